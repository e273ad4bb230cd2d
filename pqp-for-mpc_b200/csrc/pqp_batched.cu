/*
 * pqp_batched.cu -- B problems sharing one Hessian Qd, fp32 SIMT version (sm_100a).
 *
 * Same arithmetic as PQP_CPU.c:603-618 + 590-596 applied to B right-hand sides at once:
 *   NUM = (Q^- + theta) Y + F^-,  DEN = (Q^+ + theta) Y + F^+,  Y <- NUM/DEN o Y
 * with Y [N x B].  Problems are independent, so a CTA owns a tile of 32 problems for the WHOLE
 * solve: its Y tile (double-buffered) and Fd tile live in shared memory, the iteration loop needs
 * only __syncthreads (no grid barrier, no host round trip, no HBM traffic inside the loop), and
 * the two pre-split operand matrices (x-independent, built once at setup, k-major so a slab is a
 * straight 16-byte cp.async copy) stream from L2 through a double-buffered shared-memory stage.
 *
 * This is the exact-fp32 batched path (opts.use_tensor_cores = 0) and the cross-check of the
 * tcgen05 path.  Summation order per (row, problem): k ascending, fused multiply-add.
 *
 * Work per iteration: 4*N^2*B flop.
 */
#include "pqp_internal.h"

#define BT_NB 32   /* problems per CTA */
#define BT_TI 128  /* rows per row tile */
#define BT_KS 16   /* k per staged slab */
#define BT_THREADS 256

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
	const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

/*
 * QpT/QnT: [Kpad x Ipad] fp32, element [k][i] = max(0,+-Qd[i][k]) + theta_i*(i==k), zero padded
 * (Kpad multiple of 16, Ipad multiple of 128).  Built by pqp_launch_build_split_t.
 */
__global__ void __launch_bounds__(BT_THREADS, 1)
batched_simt_kernel(const float *__restrict__ QpT, const float *__restrict__ QnT, int Kpad, int Ipad, int N, int B,
		    const float *__restrict__ Fd, float *__restrict__ Y, int iters)
{
	extern __shared__ __align__(16) float smem[];
	float *Ys0 = smem;                          /* [Kpad][32] */
	float *Ys1 = Ys0 + (size_t)Kpad * BT_NB;    /* [Kpad][32] */
	float *Fs = Ys1 + (size_t)Kpad * BT_NB;     /* [Kpad][32] */
	float *As = Fs + (size_t)Kpad * BT_NB;      /* [2 stages][2 signs][BT_KS][BT_TI] */

	const int tid = threadIdx.x;
	const int ti = tid / 8, tb = tid % 8;
	const int b0 = blockIdx.x * BT_NB;

	/* load this CTA's problems: Y0 and Fd, transposed to k-major; padding problems get y=1, Fd=1 */
	for (int e = tid; e < Kpad * BT_NB; e += BT_THREADS) {
		const int b = e / Kpad, k = e % Kpad; /* consecutive threads walk one problem's vector */
		float y = 0.0f, f = 0.0f;
		if (k < N) {
			if (b0 + b < B) {
				y = Y[(size_t)(b0 + b) * N + k];
				f = Fd[(size_t)(b0 + b) * N + k];
			} else {
				y = 1.0f;
				f = 1.0f;
			}
		}
		Ys0[k * BT_NB + b] = y;
		Ys1[k * BT_NB + b] = 0.0f;
		Fs[k * BT_NB + b] = f;
	}

	const int row_tiles = Ipad / BT_TI, kslabs = Kpad / BT_KS;
	const int slabs_per_iter = row_tiles * kslabs;
	const long long total = (long long)slabs_per_iter * iters;

	auto issue = [&](long long s) {
		const int w = (int)(s % slabs_per_iter);
		const int rt = w / kslabs, ks = w % kslabs;
		float *dst = As + (size_t)(s & 1) * 2 * BT_KS * BT_TI;
		/* 16 k-rows x 128 floats per sign = 512 float4 per sign; 256 threads -> 2 + 2 copies */
#pragma unroll
		for (int u = 0; u < 2; u++) {
			const int f4 = tid + u * BT_THREADS; /* 0..511 */
			const int kk = f4 / (BT_TI / 4), c4 = f4 % (BT_TI / 4);
			const size_t g = (size_t)(ks * BT_KS + kk) * Ipad + rt * BT_TI + c4 * 4;
			cp_async16(dst + kk * BT_TI + c4 * 4, QpT + g);
			cp_async16(dst + BT_KS * BT_TI + kk * BT_TI + c4 * 4, QnT + g);
		}
		cp_async_commit();
	};

	float num[4][4], den[4][4];
#pragma unroll
	for (int r = 0; r < 4; r++)
#pragma unroll
		for (int c = 0; c < 4; c++) num[r][c] = den[r][c] = 0.0f;

	float *Ycur = Ys0, *Ynext = Ys1;
	if (total > 0) issue(0);
	for (long long s = 0; s < total; s++) {
		cp_async_wait_all();
		__syncthreads();
		if (s + 1 < total) issue(s + 1);
		const int w = (int)(s % slabs_per_iter);
		const int rt = w / kslabs, ks = w % kslabs;
		const float *ap = As + (size_t)(s & 1) * 2 * BT_KS * BT_TI;
		const float *an = ap + BT_KS * BT_TI;
		const float *yk = Ycur + (size_t)ks * BT_KS * BT_NB + tb * 4;
#pragma unroll
		for (int k = 0; k < BT_KS; k++) {
			const float4 p4 = *reinterpret_cast<const float4 *>(ap + k * BT_TI + ti * 4);
			const float4 n4 = *reinterpret_cast<const float4 *>(an + k * BT_TI + ti * 4);
			const float4 y4 = *reinterpret_cast<const float4 *>(yk + k * BT_NB);
			const float pv[4] = { p4.x, p4.y, p4.z, p4.w }, nv[4] = { n4.x, n4.y, n4.z, n4.w };
			const float yv[4] = { y4.x, y4.y, y4.z, y4.w };
#pragma unroll
			for (int r = 0; r < 4; r++)
#pragma unroll
				for (int c = 0; c < 4; c++) {
					den[r][c] = fmaf(pv[r], yv[c], den[r][c]);
					num[r][c] = fmaf(nv[r], yv[c], num[r][c]);
				}
		}
		if (ks == kslabs - 1) {
			/* row tile finished: multiplicative update for its 4 rows x 4 problems */
#pragma unroll
			for (int r = 0; r < 4; r++) {
				const int i = rt * BT_TI + ti * 4 + r;
				if (i < N) {
					const float4 y4 = *reinterpret_cast<const float4 *>(Ycur + (size_t)i * BT_NB + tb * 4);
					const float4 f4 = *reinterpret_cast<const float4 *>(Fs + (size_t)i * BT_NB + tb * 4);
					const float yv[4] = { y4.x, y4.y, y4.z, y4.w }, fv[4] = { f4.x, f4.y, f4.z, f4.w };
					float o[4];
#pragma unroll
					for (int c = 0; c < 4; c++) {
						const float nn = num[r][c] + fmaxf(-fv[c], 0.0f);
						const float dd = den[r][c] + fmaxf(fv[c], 0.0f);
						o[c] = __fdiv_rn(nn, dd) * yv[c];
					}
					*reinterpret_cast<float4 *>(Ynext + (size_t)i * BT_NB + tb * 4) = make_float4(o[0], o[1], o[2], o[3]);
				}
#pragma unroll
				for (int c = 0; c < 4; c++) num[r][c] = den[r][c] = 0.0f;
			}
			if (rt == row_tiles - 1) {
				float *t = Ycur;
				Ycur = Ynext;
				Ynext = t;
			}
		}
	}
	__syncthreads();
	for (int e = tid; e < N * BT_NB; e += BT_THREADS) {
		const int b = e / N, k = e % N;
		if (b0 + b < B) Y[(size_t)(b0 + b) * N + k] = Ycur[k * BT_NB + b];
	}
}

/* builds the k-major pre-split operands from the signed Qd and theta */
__global__ void build_split_t_kernel(float *__restrict__ QpT, float *__restrict__ QnT, int Kpad, int Ipad,
				     const float *__restrict__ Q, int ldq, const float *__restrict__ theta, int N)
{
	__shared__ float t[32][33];
	/* tile of Q: rows i0.., cols k0..  ->  out rows k, cols i */
	const int i0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
	for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
		const int i = i0 + dy, k = k0 + threadIdx.x;
		t[dy][threadIdx.x] = (i < N && k < N) ? Q[(size_t)i * ldq + k] : 0.0f;
	}
	__syncthreads();
	for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
		const int k = k0 + dy, i = i0 + threadIdx.x;
		if (k < Kpad && i < Ipad) {
			const float q = t[threadIdx.x][dy];
			float qp = fmaxf(q, 0.0f), qn = fmaxf(-q, 0.0f);
			if (i == k && i < N) {
				/* theta on the diagonal of both, as PQP_CPU.c:527,536 does */
				qp += theta[i];
				qn += theta[i];
			}
			QpT[(size_t)k * Ipad + i] = qp;
			QnT[(size_t)k * Ipad + i] = qn;
		}
	}
}

cudaError_t pqp_launch_build_split_t(float *QpT, float *QnT, int Kpad, int Ipad, const float *Q, int ldq,
						const float *theta, int N, cudaStream_t s)
{
	const int m = Kpad > Ipad ? Kpad : Ipad;
	dim3 block(32, 8), grid((m + 31) / 32, (m + 31) / 32);
	build_split_t_kernel<<<grid, block, 0, s>>>(QpT, QnT, Kpad, Ipad, Q, ldq, theta, N);
	return cudaGetLastError();
}

static size_t batched_simt_smem(int Kpad)
{
	return sizeof(float) * (3 * (size_t)Kpad * BT_NB + 2 * 2 * BT_KS * BT_TI);
}

int pqp_batched_simt_supported(int N)
{
	return batched_simt_smem(pqp_round_up(N, BT_KS)) <= 227 * 1024;
}

cudaError_t pqp_launch_batched_simt_split(const float *QpT, const float *QnT, int Kpad, int Ipad, int N, int B,
						     const float *Fd, float *Y, int iters, cudaStream_t s)
{
	const size_t smem = batched_simt_smem(Kpad);
	cudaError_t e = cudaFuncSetAttribute(batched_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	const int grid = (B + BT_NB - 1) / BT_NB;
	batched_simt_kernel<<<grid, BT_THREADS, smem, s>>>(QpT, QnT, Kpad, Ipad, N, B, Fd, Y, iters);
	return cudaGetLastError();
}
