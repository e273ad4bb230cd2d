/*
 * pqp_setup_kernels.cu -- the small kernels around the iteration loop (sm_100a).
 *
 *   matmul_strict / matmul_simt   convertToDual's products      PQP_CPU.c:489-498 (a3/a4 in SURVEY 8a)
 *   theta                         computeTheta + diagonalAdd    PQP_CPU.c:503-519, 235-242 (a5)
 *   fp / fd / md                  computeFp, computeFd, computeMd PQP_CPU.c:373-382, 456-460, 472-479 (a2/a3)
 *   recover                       computeUfromY                 PQP_CPU.c:352-360 (a11)
 *   status                        the quantities terminate() tests, PQP_CPU.c:673-687, from g = Qd y + Fd (a10)
 *
 * "strict" variants keep PQP_CPU.c's summation order (k ascending, separately rounded multiply
 * and add via __fmul_rn/__fadd_rn, which nvcc never contracts) and are bit-identical to it.
 */
#include "pqp_internal.h"

#define STRICT_TILE 32

/* ---------------------------------------------------------------------------------------------
 * C = A * op(B), reference order.  Classic 32x32 shared-memory tiling: tiles advance k ascending
 * and each thread adds its tile's 32 products k ascending, so every output sees exactly the
 * sequence  acc = fl(acc + fl(a_ik * b_kj)),  k = 0..b-1  of PQP_CPU.c:90-99.
 * ------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(STRICT_TILE *STRICT_TILE)
matmul_strict_kernel(float *__restrict__ C, int ldc, const float *__restrict__ A, int lda,
		     const float *__restrict__ B, int ldb, int transB, int a, int b, int c)
{
	__shared__ float As[STRICT_TILE][STRICT_TILE + 1];
	__shared__ float Bs[STRICT_TILE][STRICT_TILE + 1]; /* Bs[k][j] */
	const int tx = threadIdx.x, ty = threadIdx.y;
	const int i = blockIdx.y * STRICT_TILE + ty;
	const int j = blockIdx.x * STRICT_TILE + tx;
	float acc = 0.0f;
	for (int k0 = 0; k0 < b; k0 += STRICT_TILE) {
		const int ka = k0 + tx;
		As[ty][tx] = (i < a && ka < b) ? A[(size_t)i * lda + ka] : 0.0f;
		if (transB) {
			/* B stored [c x b]: read row (blockIdx.x*32 + ty), column k0+tx, store transposed */
			const int jr = blockIdx.x * STRICT_TILE + ty;
			Bs[tx][ty] = (jr < c && ka < b) ? B[(size_t)jr * ldb + ka] : 0.0f;
		} else {
			const int kb = k0 + ty;
			Bs[ty][tx] = (kb < b && j < c) ? B[(size_t)kb * ldb + j] : 0.0f;
		}
		__syncthreads();
		const int kmax = min(STRICT_TILE, b - k0);
		for (int k = 0; k < kmax; k++) acc = __fadd_rn(acc, __fmul_rn(As[ty][k], Bs[k][tx]));
		__syncthreads();
	}
	if (i < a && j < c) C[(size_t)i * ldc + j] = acc;
}

cudaError_t pqp_launch_matmul_strict(float *C, int ldc, const float *A, int lda, const float *B, int ldb, int transB,
				     int a, int b, int c, cudaStream_t s)
{
	dim3 block(STRICT_TILE, STRICT_TILE), grid((c + STRICT_TILE - 1) / STRICT_TILE, (a + STRICT_TILE - 1) / STRICT_TILE);
	matmul_strict_kernel<<<grid, block, 0, s>>>(C, ldc, A, lda, B, ldb, transB, a, b, c);
	return cudaGetLastError();
}

/* ---------------------------------------------------------------------------------------------
 * fp32 SIMT GEMM, 64x64 tile, 16-deep k slab, 4x4 outputs per thread.  FAST mode without tensor
 * cores, and the cross-check of the tcgen05 path.
 * ------------------------------------------------------------------------------------------- */
#define SG_BM 64
#define SG_BN 64
#define SG_BK 16
__global__ void __launch_bounds__(256)
matmul_simt_kernel(float *__restrict__ C, int ldc, const float *__restrict__ A, int lda, const float *__restrict__ B,
		   int ldb, int transB, int a, int b, int c)
{
	__shared__ float As[SG_BK][SG_BM + 4]; /* As[k][i] */
	__shared__ float Bs[SG_BK][SG_BN + 4]; /* Bs[k][j] */
	const int tid = threadIdx.x;
	const int tx = tid % 16, ty = tid / 16;
	const int i0 = blockIdx.y * SG_BM, j0 = blockIdx.x * SG_BN;
	float acc[4][4] = {};
	for (int k0 = 0; k0 < b; k0 += SG_BK) {
		/* A tile: 64 rows x 16 k; thread -> (row = tid/4, 4 consecutive k) */
		{
			const int r = tid / 4, kk = (tid % 4) * 4;
#pragma unroll
			for (int u = 0; u < 4; u++) {
				const int gi = i0 + r, gk = k0 + kk + u;
				As[kk + u][r] = (gi < a && gk < b) ? A[(size_t)gi * lda + gk] : 0.0f;
			}
		}
		if (transB) {
			const int r = tid / 4, kk = (tid % 4) * 4;
#pragma unroll
			for (int u = 0; u < 4; u++) {
				const int gj = j0 + r, gk = k0 + kk + u;
				Bs[kk + u][r] = (gj < c && gk < b) ? B[(size_t)gj * ldb + gk] : 0.0f;
			}
		} else {
			const int kk = tid / 16, jj = (tid % 16) * 4;
#pragma unroll
			for (int u = 0; u < 4; u++) {
				const int gk = k0 + kk, gj = j0 + jj + u;
				Bs[kk][jj + u] = (gk < b && gj < c) ? B[(size_t)gk * ldb + gj] : 0.0f;
			}
		}
		__syncthreads();
#pragma unroll
		for (int k = 0; k < SG_BK; k++) {
			float av[4], bv[4];
#pragma unroll
			for (int u = 0; u < 4; u++) {
				av[u] = As[k][ty * 4 + u];
				bv[u] = Bs[k][tx * 4 + u];
			}
#pragma unroll
			for (int u = 0; u < 4; u++)
#pragma unroll
				for (int v = 0; v < 4; v++) acc[u][v] = fmaf(av[u], bv[v], acc[u][v]);
		}
		__syncthreads();
	}
#pragma unroll
	for (int u = 0; u < 4; u++)
#pragma unroll
		for (int v = 0; v < 4; v++) {
			const int gi = i0 + ty * 4 + u, gj = j0 + tx * 4 + v;
			if (gi < a && gj < c) C[(size_t)gi * ldc + gj] = acc[u][v];
		}
}

cudaError_t pqp_launch_matmul_simt(float *C, int ldc, const float *A, int lda, const float *B, int ldb, int transB,
				   int a, int b, int c, cudaStream_t s)
{
	dim3 grid((c + SG_BN - 1) / SG_BN, (a + SG_BM - 1) / SG_BM);
	matmul_simt_kernel<<<grid, 256, 0, s>>>(C, ldc, A, lda, B, ldb, transB, a, b, c);
	return cudaGetLastError();
}

/* out[c x r] = in[r x c]' through a padded 32x32 tile */
__global__ void transpose_kernel(float *__restrict__ out, int ldo, const float *__restrict__ in, int ldi, int r, int c)
{
	__shared__ float t[32][33];
	int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
	for (int dy = threadIdx.y; dy < 32; dy += blockDim.y)
		if (y0 + dy < r && x < c) t[dy][threadIdx.x] = in[(size_t)(y0 + dy) * ldi + x];
	__syncthreads();
	x = blockIdx.y * 32 + threadIdx.x; /* input row -> output column */
	y0 = blockIdx.x * 32;             /* input column -> output row */
	for (int dy = threadIdx.y; dy < 32; dy += blockDim.y)
		if (y0 + dy < c && x < r) out[(size_t)(y0 + dy) * ldo + x] = t[threadIdx.x][dy];
}

cudaError_t pqp_launch_transpose(float *out, int ldo, const float *in, int ldi, int r, int c, cudaStream_t s)
{
	dim3 block(32, 8), grid((c + 31) / 32, (r + 31) / 32);
	transpose_kernel<<<grid, block, 0, s>>>(out, ldo, in, ldi, r, c);
	return cudaGetLastError();
}

/* ---------------------------------------------------------------------------------------------
 * theta.  strict: `Q` is the TRANSPOSE (so that thread i walking j ascending reads coalesced);
 * fast: warp per row of Q with a shuffle tree.
 * ------------------------------------------------------------------------------------------- */
__global__ void theta_strict_kernel(float *__restrict__ theta, const float *__restrict__ QT, int ldq, int N, float floor_)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= N) return;
	float s = 0.0f;
	for (int j = 0; j < N; j++) {
		const float q = QT[(size_t)j * ldq + i];
		const float qn = (0.0f > -q) ? 0.0f : -q; /* max(0.0,-q) of PQP_CPU.c:209 */
		s = __fadd_rn(s, __fmul_rn(qn, 1.0f));
	}
	theta[i] = (s > floor_) ? s : floor_;
}

__global__ void theta_fast_kernel(float *__restrict__ theta, const float *__restrict__ Q, int ldq, int N, float floor_)
{
	const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x % 32;
	if (row >= N) return;
	const float4 *r4 = reinterpret_cast<const float4 *>(Q + (size_t)row * ldq);
	float s = 0.0f;
	for (int c = lane; c < ldq / 4; c += 32) {
		const float4 q = r4[c];
		s += fmaxf(-q.x, 0.0f) + fmaxf(-q.y, 0.0f) + fmaxf(-q.z, 0.0f) + fmaxf(-q.w, 0.0f);
	}
#pragma unroll
	for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
	if (lane == 0) theta[row] = fmaxf(s, floor_);
}

cudaError_t pqp_launch_theta(float *theta, const float *Q, int ldq, int N, float floor_, int strict, cudaStream_t s)
{
	if (strict)
		theta_strict_kernel<<<(N + 127) / 128, 128, 0, s>>>(theta, Q, ldq, N, floor_);
	else
		theta_fast_kernel<<<(N + 7) / 8, 256, 0, s>>>(theta, Q, ldq, N, floor_);
	return cudaGetLastError();
}

/* ---------------------------------------------------------------------------------------------
 * Fp = Fp1*D + Fp2*x - Fp3 in the reference's order:  t1 = sum_k Fp1[i,k] D[k];  t2 = sum_k
 * Fp2[i,k] x[k];  Fp = (t1 + 1*t2) + (-1*Fp3)   (PQP_CPU.c:375-379).  Always in this order:
 * it is tiny, so FAST and STRICT share it and Fp is bit-identical to the reference in both.
 * ------------------------------------------------------------------------------------------- */
__global__ void fp_kernel(float *__restrict__ Fp, const float *__restrict__ Fp1, const float *__restrict__ Fp2,
			  const float *__restrict__ Fp3, const float *__restrict__ Fp_const, const float *__restrict__ D,
			  int D_stride, const float *__restrict__ X, int B, int M, int nd, int nState)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (long long)B * M) return;
	const int b = (int)(t / M), i = (int)(t % M);
	if (nState == 0) {
		Fp[t] = Fp_const[i];
		return;
	}
	float t1 = 0.0f, t2 = 0.0f;
	const float *d = D + (size_t)b * D_stride;
	for (int k = 0; k < nd; k++) t1 = __fadd_rn(t1, __fmul_rn(Fp1[(size_t)i * nd + k], d[k]));
	const float *x = X + (size_t)b * nState;
	for (int k = 0; k < nState; k++) t2 = __fadd_rn(t2, __fmul_rn(Fp2[(size_t)i * nState + k], x[k]));
	float f = __fadd_rn(t1, __fmul_rn(1.0f, t2));
	f = __fadd_rn(f, __fmul_rn(-1.0f, Fp3[i]));
	Fp[t] = f;
}

cudaError_t pqp_launch_fp(float *Fp, const float *Fp1, const float *Fp2, const float *Fp3, const float *Fp_const,
			  const float *D, int D_stride, const float *X, int B, int M, int nd, int nState, cudaStream_t s)
{
	const long long n = (long long)B * M;
	fp_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(Fp, Fp1, Fp2, Fp3, Fp_const, D, D_stride, X, B, M, nd, nState);
	return cudaGetLastError();
}

/*
 * Fd = GQ*Fp + Kp.  One output per thread, k ascending, separately rounded mul and add: the reference's order
 * (PQP_CPU.c:458-459), so Fd is bit-identical to PQP_CPU.c's in every mode.  A block forms a 32-row x 32-problem tile of
 * outputs from shared-memory copies of the two operand tiles (coalesced loads; the k loop then reads a broadcast GQ value and a
 * conflict-free Fp column), instead of every thread walking its own strided row of GQ in global memory.
 */
#define FD_TR 32
#define FD_TB 32
#define FD_KC 64
__global__ void __launch_bounds__(FD_TR * FD_TB / 4)
fd_seq_kernel(float *__restrict__ Fd, const float *__restrict__ GQ, const float *__restrict__ Fp, const float *__restrict__ Kp, int B, int N,
	      int M)
{
	__shared__ float g_s[FD_TR][FD_KC + 1], f_s[FD_TB][FD_KC + 1];
	const int i0 = blockIdx.x * FD_TR, b0 = blockIdx.y * FD_TB;
	const int tid = threadIdx.x;          /* 256 threads: thread = (row r, problem group of 4) */
	const int r = tid % FD_TR, pg = tid / FD_TR; /* pg in 0..7 -> problems pg, pg+8, pg+16, pg+24 */
	float acc[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
	for (int k0 = 0; k0 < M; k0 += FD_KC) {
		const int kc = min(FD_KC, M - k0);
		for (int e = tid; e < FD_TR * FD_KC; e += blockDim.x) {
			const int rr = e / FD_KC, kk = e % FD_KC;
			g_s[rr][kk] = (i0 + rr < N && kk < kc) ? GQ[(size_t)(i0 + rr) * M + k0 + kk] : 0.0f;
			f_s[rr][kk] = (b0 + rr < B && kk < kc) ? Fp[(size_t)(b0 + rr) * M + k0 + kk] : 0.0f;
		}
		__syncthreads();
		for (int k = 0; k < kc; k++) {
			const float g = g_s[r][k];
#pragma unroll
			for (int u = 0; u < 4; u++) acc[u] = __fadd_rn(acc[u], __fmul_rn(g, f_s[pg + 8 * u][k]));
		}
		__syncthreads();
	}
	if (i0 + r < N) {
		const float kp = __fmul_rn(1.0f, Kp[i0 + r]);
#pragma unroll
		for (int u = 0; u < 4; u++) {
			const int b = b0 + pg + 8 * u;
			if (b < B) Fd[(size_t)b * N + i0 + r] = __fadd_rn(acc[u], kp);
		}
	}
}

/* single problem, FAST: warp per row, coalesced, shuffle tree */
__global__ void fd_warp_kernel(float *__restrict__ Fd, const float *__restrict__ GQ, const float *__restrict__ Fp,
			       const float *__restrict__ Kp, int N, int M)
{
	const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x % 32;
	if (row >= N) return;
	const float *g = GQ + (size_t)row * M;
	float acc = 0.0f;
	for (int k = lane; k < M; k += 32) acc = fmaf(g[k], Fp[k], acc);
#pragma unroll
	for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
	if (lane == 0) Fd[row] = acc + Kp[row];
}

cudaError_t pqp_launch_fd(float *Fd, const float *GQ, const float *Fp, const float *Kp, int B, int N, int M, int strict,
			  cudaStream_t s)
{
	if (!strict && B == 1 && M >= 256) {
		fd_warp_kernel<<<(N + 7) / 8, 256, 0, s>>>(Fd, GQ, Fp, Kp, N, M);
	} else {
		fd_seq_kernel<<<dim3((N + FD_TR - 1) / FD_TR, (B + FD_TB - 1) / FD_TB), FD_TR * FD_TB / 4, 0, s>>>(Fd, GQ, Fp, Kp, B, N, M);
	}
	return cudaGetLastError();
}

/* ---------------------------------------------------------------------------------------------
 * Md = Fp' Qp_inv Fp - Mp(x), one block per problem.  Only shifts the reported dual cost, so the
 * block-tree summation order (not the reference's) is within the status tolerance.
 * ------------------------------------------------------------------------------------------- */
__device__ __forceinline__ float block_sum_256(float v, float *red)
{
#pragma unroll
	for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	__syncthreads();
	if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = v;
	__syncthreads();
	float t = 0.0f;
	for (int w = 0; w < (int)blockDim.x / 32; w++) t += red[w];
	return t;
}

/* t[b][j] = sum_k Fp[b][k] Qp_inv[k][j]: a block takes 32 columns j (lane = column: 128-byte row segments of Qp_inv) of one problem,
 * its 8 warps take k = w, w+8, ... and are folded through shared memory in warp order (deterministic).  Round 1 walked all M rows
 * per thread column inside ONE block per problem: M = 2048 meant 16 k dependent loads per thread on a single SM. */
__global__ void __launch_bounds__(256) md_rowvec_kernel(float *__restrict__ t, const float *__restrict__ Fp, const float *__restrict__ Qp_inv, int M)
{
	__shared__ float part[8][33];
	const int b = blockIdx.y, j = blockIdx.x * 32 + threadIdx.x % 32, w = threadIdx.x / 32;
	const float *f = Fp + (size_t)b * M;
	float acc = 0.0f;
	if (j < M)
		for (int k = w; k < M; k += 8) acc = fmaf(f[k], Qp_inv[(size_t)k * M + j], acc);
	part[w][threadIdx.x % 32] = acc;
	__syncthreads();
	if (w == 0 && j < M) {
		float s = part[0][threadIdx.x];
		for (int u = 1; u < 8; u++) s += part[u][threadIdx.x];
		t[(size_t)b * M + j] = s;
	}
}

__global__ void __launch_bounds__(256)
md_kernel(float *__restrict__ Md, const float *__restrict__ Fp, const float *__restrict__ t,
	  const float *__restrict__ Mp1, const float *__restrict__ Mp2, const float *__restrict__ Mp3,
	  const float *__restrict__ Mp4, const float *__restrict__ Mp5, const float *__restrict__ Mp6, float Mp0,
	  const float *__restrict__ D, int D_stride, const float *__restrict__ X, int M, int nd, int nState)
{
	__shared__ float red[8];
	const int b = blockIdx.x;
	const float *f = Fp + (size_t)b * M, *tb = t + (size_t)b * M;
	float part = 0.0f;
	for (int j = threadIdx.x; j < M; j += blockDim.x) part = fmaf(tb[j], f[j], part);
	const float quad = block_sum_256(part, red);
	float mp = Mp0;
	if (Mp1) {
		const float *x = X + (size_t)b * nState, *d = D + (size_t)b * D_stride;
		float p = 0.0f;
		for (int j = threadIdx.x; j < nState; j += blockDim.x) {
			float t1 = 0.0f, t2 = 0.0f;
			for (int k = 0; k < nState; k++) t1 = fmaf(x[k], Mp1[(size_t)k * nState + j], t1);
			for (int k = 0; k < nd; k++) t2 = fmaf(d[k], Mp2[(size_t)k * nState + j], t2);
			p += 0.5f * (t1 + t2 + Mp4[j]) * x[j];
		}
		for (int j = threadIdx.x; j < nd; j += blockDim.x) {
			float t3 = 0.0f;
			for (int k = 0; k < nd; k++) t3 = fmaf(d[k], Mp3[(size_t)k * nd + j], t3);
			p += 0.5f * (t3 + Mp5[j]) * d[j];
		}
		mp = block_sum_256(p, red) + 0.5f * Mp6[0];
	}
	if (threadIdx.x == 0) Md[b] = quad - mp;
}

/* tmp: [B x M] scratch */
cudaError_t pqp_launch_md(float *Md, float *tmp, const float *Fp, const float *Qp_inv, const float *Mp1, const float *Mp2,
			  const float *Mp3, const float *Mp4, const float *Mp5, const float *Mp6, float Mp0, const float *D,
			  int D_stride, const float *X, int B, int M, int nd, int nState, cudaStream_t s)
{
	for (int b0 = 0; b0 < B; b0 += 65535) { /* gridDim.y limit */
		const int nb = B - b0 < 65535 ? B - b0 : 65535;
		md_rowvec_kernel<<<dim3((M + 31) / 32, nb), 256, 0, s>>>(tmp + (size_t)b0 * M, Fp + (size_t)b0 * M, Qp_inv, M);
	}
	md_kernel<<<B, 256, 0, s>>>(Md, Fp, tmp, Mp1, Mp2, Mp3, Mp4, Mp5, Mp6, Mp0, D, D_stride, X, M, nd, nState);
	return cudaGetLastError();
}

/* ---------------------------------------------------------------------------------------------
 * Primal recovery U = -(Qp_inv (Gp' y + Fp)), reference order (PQP_CPU.c:355-358).
 * Stage 1: tmp[b][i] = (sum_k Gp[k,i] y_k) + 1*Fp_i   (thread per (b,i): k ascending, coalesced over i)
 * Stage 2: U[b][i]   = -(sum_k Qp_inv[i,k] tmp_k)
 * ------------------------------------------------------------------------------------------- */
__global__ void recover_stage1(float *__restrict__ tmp, const float *__restrict__ Y, int ldy, const float *__restrict__ Fp,
			       const float *__restrict__ Gp, int B, int N, int M)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (long long)B * M) return;
	const int b = (int)(t / M), i = (int)(t % M);
	const float *y = Y + (size_t)b * ldy;
	float acc = 0.0f;
	for (int k = 0; k < N; k++) acc = __fadd_rn(acc, __fmul_rn(Gp[(size_t)k * M + i], y[k]));
	tmp[t] = __fadd_rn(acc, __fmul_rn(1.0f, Fp[t]));
}
__global__ void recover_stage2(float *__restrict__ U, const float *__restrict__ tmp, const float *__restrict__ Qp_inv, int B,
			       int M)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (long long)B * M) return;
	const int b = (int)(t / M), i = (int)(t % M);
	const float *v = tmp + (size_t)b * M, *q = Qp_inv + (size_t)i * M;
	float acc = 0.0f;
	for (int k = 0; k < M; k++) acc = __fadd_rn(acc, __fmul_rn(q[k], v[k]));
	U[t] = -acc;
}

cudaError_t pqp_launch_recover(float *U, float *tmp, const float *Y, int ldy, const float *Fp, const float *Gp,
			       const float *Qp_inv, int B, int N, int M, int strict, cudaStream_t s)
{
	(void)strict; /* both modes use the reference order: the recovery is O(NM) once per solve */
	const long long n = (long long)B * M;
	const unsigned blocks = (unsigned)((n + 127) / 128);
	recover_stage1<<<blocks, 128, 0, s>>>(tmp, Y, ldy, Fp, Gp, B, N, M);
	recover_stage2<<<blocks, 128, 0, s>>>(U, tmp, Qp_inv, B, M);
	return cudaGetLastError();
}

__global__ void fill_kernel(float *p, float v, size_t n)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
cudaError_t pqp_launch_fill(float *p, float v, size_t n, cudaStream_t s)
{
	if (n == 0) return cudaSuccess;
	size_t blocks = (n + 255) / 256;
	if (blocks > 148 * 8) blocks = 148 * 8;
	fill_kernel<<<(unsigned)blocks, 256, 0, s>>>(p, v, n);
	return cudaGetLastError();
}

/* ---------------------------------------------------------------------------------------------
 * Status of a given y: g = Qd y + Fd, then min g, y'g, Jd = sum y (g + Fd)/2 + Md/2, ||min(y,g)||inf.
 * One block per problem; warps take rows round-robin.  (these replace the five GEMVs
 * and four dot products of terminate(), PQP_CPU.c:673-687.)
 * ------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(256)
status_kernel(pqp_status *__restrict__ st, const float *__restrict__ Q, int ldq, int N, const float *__restrict__ Y, int ldy,
	      const float *__restrict__ Fd, const float *__restrict__ Md, const float *__restrict__ Kp, float erc, float eac,
	      int iters, float *__restrict__ viol_out)
{
	__shared__ float r_min[8], r_gap[8], r_jd[8], r_kkt[8], r_viol[8];
	const int b = blockIdx.x, warp = threadIdx.x / 32, lane = threadIdx.x % 32;
	const float *y = Y + (size_t)b * ldy, *fd = Fd + (size_t)b * N;
	float vmin = INFINITY, gap = 0.0f, jd = 0.0f, kkt = 0.0f, viol = -INFINITY;
	for (int i = warp; i < N; i += 8) {
		const float *q = Q + (size_t)i * ldq;
		float acc = 0.0f;
		for (int k = lane; k < N; k += 32) acc = fmaf(q[k], y[k], acc);
#pragma unroll
		for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
		const float g = acc + fd[i], yi = y[i];
		vmin = fminf(vmin, g);
		gap = fmaf(yi, g, gap);
		jd = fmaf(yi, 0.5f * (g + fd[i]), jd);
		kkt = fmaxf(kkt, fabsf(fminf(yi, g)));
		const float tol = Kp ? fmaxf(erc * Kp[i], eac) : eac;
		viol = fmaxf(viol, -g - tol);
	}
	if (lane == 0) {
		r_min[warp] = vmin; r_gap[warp] = gap; r_jd[warp] = jd; r_kkt[warp] = kkt; r_viol[warp] = viol;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int w = 1; w < 8; w++) {
			vmin = fminf(vmin, r_min[w]); gap += r_gap[w]; jd += r_jd[w];
			kkt = fmaxf(kkt, r_kkt[w]); viol = fmaxf(viol, r_viol[w]);
		}
		pqp_status o;
		o.iters = iters;
		o.converged = 0;
		o.min_slack = vmin;
		o.gap = gap;
		o.Jd = jd + (Md ? 0.5f * Md[b] : 0.0f);
		o.kkt = kkt;
		st[b] = o;
		/* max_i(-g_i - max(erc*Kp_i, eac)): <= 0 is checkFeas/compare passing (PQP_CPU.c:632-641, :334-343) */
		if (viol_out) viol_out[b] = viol;
	}
}

cudaError_t pqp_launch_status(pqp_status *st, const float *Q, int ldq, int N, const float *Y, int ldy, const float *Fd,
			      const float *Md, const float *Kp, float erc, float eac, int B, int iters, float *viol_out, cudaStream_t s)
{
	status_kernel<<<B, 256, 0, s>>>(st, Q, ldq, N, Y, ldy, Fd, Md, Kp, erc, eac, iters, viol_out);
	return cudaGetLastError();
}

/* receding-horizon shift of the duals: four blocks of pH steps x nI rows; step k takes step k+1, the last step is held */
__global__ void shift_duals_kernel(float *__restrict__ out, const float *__restrict__ in, int B, int pH, int nI, float y_floor)
{
	const int N = 4 * pH * nI;
	const size_t total = (size_t)B * N;
	for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
		const int i = (int)(e % N);
		const int blk = i / (pH * nI), within = i % (pH * nI), k = within / nI, j = within % nI;
		const int ks = k + 1 < pH ? k + 1 : k;
		out[e] = fmaxf(in[e - i + (size_t)blk * pH * nI + (size_t)ks * nI + j], y_floor);
	}
}

cudaError_t pqp_launch_shift_duals(float *out, const float *in, int B, int pH, int nI, float y_floor, cudaStream_t s)
{
	const size_t total = (size_t)B * 4 * pH * nI;
	size_t blocks = (total + 255) / 256;
	if (blocks > 148 * 8) blocks = 148 * 8;
	if (blocks < 1) blocks = 1;
	shift_duals_kernel<<<(unsigned)blocks, 256, 0, s>>>(out, in, B, pH, nI, y_floor);
	return cudaGetLastError();
}

/* ---- state-dependent constraint offsets: Fd[b][i] += sum_k Kx[i][k] x_b[k] + sum_k Kd[i][k] D_b[k] -------------- */
__global__ void fd_offsets_kernel(float *Fd, const float *Kx, const float *X, int nS, const float *Kd, const float *D, int D_stride, int nd,
				  int B, int N)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
	if (i >= N || b >= B) return;
	float acc = Fd[(size_t)b * N + i];
	if (Kx) {
		float t = 0.0f;
		for (int k = 0; k < nS; k++) t = __fadd_rn(t, __fmul_rn(Kx[(size_t)i * nS + k], X[(size_t)b * nS + k]));
		acc = __fadd_rn(acc, t);
	}
	if (Kd) {
		float t = 0.0f;
		for (int k = 0; k < nd; k++) t = __fadd_rn(t, __fmul_rn(Kd[(size_t)i * nd + k], D[(size_t)b * D_stride + k]));
		acc = __fadd_rn(acc, t);
	}
	Fd[(size_t)b * N + i] = acc;
}

cudaError_t pqp_launch_fd_offsets(float *Fd, const float *Kx, const float *X, int nState, const float *Kd, const float *D, int D_stride,
				  int nd, int B, int N, cudaStream_t s)
{
	if (B <= 0 || N <= 0) return cudaSuccess;
	for (int b0 = 0; b0 < B; b0 += 65535) { /* gridDim.y limit */
		const int nb = B - b0 < 65535 ? B - b0 : 65535;
		fd_offsets_kernel<<<dim3((N + 127) / 128, nb), 128, 0, s>>>(Fd + (size_t)b0 * N, Kx, X ? X + (size_t)b0 * nState : X, nState, Kd,
									    D ? D + (size_t)b0 * D_stride : D, D_stride, nd, nb, N);
	}
	return cudaGetLastError();
}

/* ---- updateY2 + updY on the reference's dense split operands (single-step parity instrument, PQP_CPU.c:603-618, 590-596) ---- */
__global__ void update_y2_dense_kernel(float *Yn, const float *Y, const float *Qp, const float *Qn, const float *Fdp, const float *Fdn, int N)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= N) return;
	const float *rn = Qn + (size_t)i * N, *rp = Qp + (size_t)i * N;
	float num = 0.0f, den = 0.0f;
	for (int k = 0; k < N; k++) num = __fadd_rn(num, __fmul_rn(rn[k], Y[k]));
	for (int k = 0; k < N; k++) den = __fadd_rn(den, __fmul_rn(rp[k], Y[k]));
	num = __fadd_rn(num, Fdn[i]);
	den = __fadd_rn(den, Fdp[i]);
	Yn[i] = __fmul_rn(__fdiv_rn(num, den), Y[i]);
}

cudaError_t pqp_launch_update_y2_dense(float *Yn, const float *Y, const float *Qp, const float *Qn, const float *Fdp, const float *Fdn, int N,
				       cudaStream_t s)
{
	update_y2_dense_kernel<<<(N + 63) / 64, 64, 0, s>>>(Yn, Y, Qp, Qn, Fdp, Fdn, N);
	return cudaGetLastError();
}

/* ---- computeCost (PQP_CPU.c:648-666) and computeMd (:472-479) in the reference's order ------------------------------------------
 * Both start with the row vector t = z' A (matrixMultiply(tmp, Z, 1, Q, 0, 1, n, n): t_j = sum_k z_k A[k][j], k ascending from
 * zero, separately rounded multiply and add) followed by the dot product t z (j ascending).  Thread j owns t_j (coalesced over
 * j); one thread then walks the two dot products.  O(n^2) once per call; bit-identical to the reference. */
__global__ void rowvec_mat_strict_kernel(float *__restrict__ t, const float *__restrict__ z, const float *__restrict__ A, int n)
{
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n) return;
	float acc = 0.0f;
	for (int k = 0; k < n; k++) acc = __fadd_rn(acc, __fmul_rn(z[k], A[(size_t)k * n + j]));
	t[j] = acc;
}
/* mode 0: J = 1/2 t z + F z + m/2 with the reference's promotions (0.5*tmp[0] is a double product rounded into the float J,
 * M[0]/2 a float division); mode 1: Md = t z - m */
__global__ void quad_finish_kernel(float *__restrict__ out, const float *__restrict__ t, const float *__restrict__ z,
				   const float *__restrict__ F, const float *__restrict__ m, int n, int mode)
{
	if (blockIdx.x || threadIdx.x) return;
	float q = 0.0f;
	for (int j = 0; j < n; j++) q = __fadd_rn(q, __fmul_rn(t[j], z[j]));
	if (mode == 1) {
		out[0] = __fsub_rn(q, m ? m[0] : 0.0f);
		return;
	}
	float J = 0.0f;
	J = (float)((double)J + 0.5 * (double)q);
	float l = 0.0f;
	for (int j = 0; j < n; j++) l = __fadd_rn(l, __fmul_rn(F[j], z[j]));
	J = __fadd_rn(J, l);
	J = __fadd_rn(J, __fdiv_rn(m ? m[0] : 0.0f, 2.0f));
	out[0] = J;
}
cudaError_t pqp_launch_quad_form(float *out, float *tmp, const float *z, const float *A, const float *F, const float *m, int n, int mode,
				 cudaStream_t s)
{
	rowvec_mat_strict_kernel<<<(n + 127) / 128, 128, 0, s>>>(tmp, z, A, n);
	quad_finish_kernel<<<1, 32, 0, s>>>(out, tmp, z, F, m, n, mode);
	return cudaGetLastError();
}

/* ---- Brand's acceleration / line-search step (the dead branch of PQP_CPU.c:721-735: computeph :625-630, computealphaY :545-575,
 * updateY1 :579-588; computeph's `matrixAdd(ph, ph, ...)` read as `+= Fd`, the evident fix) -- opt-in, pqp_opts.accelerate ----
 *     ph = max(0, -(Qd y + Fd));   alpha = -((y'Qd + Fd') ph) / (ph'Qd ph) if ph'Qd ph > 0 else 0;   y <- y + alpha ph
 * Every sum in matrixMultiply's order (k ascending from zero, separately rounded multiply and add), so the step is bit-identical to
 * the reference's arithmetic (with that fix) in both orders of the library.  Three passes over Qd per step (row form for ph, column form for
 * ph'Qd and y'Qd, as the reference multiplies them), each as the tiled sequential-k product of fd_seq_kernel.
 *   seqdot_kernel<COL, EPI>: out[b][i] = epi(sum_k A(i,k) V[b][k]),  A(i,k) = Q[i][k] (COL = 0) or Q[k][i] (COL = 1)
 *   EPI 0: the sum;  1: sum + 1*Fd[b][i];  2: max(0, -(sum + 1*Fd[b][i])) */
template <int COL, int EPI>
__global__ void __launch_bounds__(256) seqdot_kernel(float *__restrict__ out, const float *__restrict__ Q, int ldq, const float *__restrict__ V, int ldv,
						      const float *__restrict__ Fd, int B, int N)
{
	__shared__ float a_s[32][65], v_s[32][65];
	const int i0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
	const int tid = threadIdx.x, r = tid % 32, pg = tid / 32; /* thread = (row r, problems pg, pg+8, pg+16, pg+24) */
	float acc[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
	for (int k0 = 0; k0 < N; k0 += 64) {
		const int kc = min(64, N - k0);
		for (int e = tid; e < 32 * 64; e += 256) {
			if (COL) { /* A(i,k) = Q[k][i]: consecutive threads along i */
				const int kk = e / 32, rr = e % 32;
				a_s[rr][kk] = (i0 + rr < N && kk < kc) ? Q[(size_t)(k0 + kk) * ldq + i0 + rr] : 0.0f;
			} else {
				const int rr = e / 64, kk = e % 64;
				a_s[rr][kk] = (i0 + rr < N && kk < kc) ? Q[(size_t)(i0 + rr) * ldq + k0 + kk] : 0.0f;
			}
			const int rb = e / 64, kb = e % 64;
			v_s[rb][kb] = (b0 + rb < B && kb < kc) ? V[(size_t)(b0 + rb) * ldv + k0 + kb] : 0.0f;
		}
		__syncthreads();
		for (int k = 0; k < kc; k++) {
			const float a = a_s[r][k];
#pragma unroll
			for (int u = 0; u < 4; u++) acc[u] = __fadd_rn(acc[u], __fmul_rn(a, v_s[pg + 8 * u][k]));
		}
		__syncthreads();
	}
	if (i0 + r < N) {
#pragma unroll
		for (int u = 0; u < 4; u++) {
			const int b = b0 + pg + 8 * u;
			if (b >= B) continue;
			float v = acc[u];
			if (EPI >= 1) v = __fadd_rn(v, __fmul_rn(1.0f, Fd[(size_t)b * N + i0 + r]));
			if (EPI == 2) v = (0.0f > -v) ? 0.0f : -v;
			out[(size_t)b * N + i0 + r] = v;
		}
	}
}

/* one block per problem: thread 0 walks the two dot products in order, then the block applies y += alpha*ph */
__global__ void __launch_bounds__(256) accel_finish_kernel(float *__restrict__ Y, int ldy, const float *__restrict__ ph, const float *__restrict__ t,
							    const float *__restrict__ t2, int N)
{
	__shared__ float alpha_s;
	const int b = blockIdx.x;
	const float *p = ph + (size_t)b * N, *tt = t + (size_t)b * N, *tt2 = t2 + (size_t)b * N;
	if (threadIdx.x == 0) {
		float s0 = 0.0f, alpha = 0.0f;
		for (int j = 0; j < N; j++) s0 = __fadd_rn(s0, __fmul_rn(tt[j], p[j]));
		if (s0 > 0.0f) {
			float s2 = 0.0f;
			for (int j = 0; j < N; j++) s2 = __fadd_rn(s2, __fmul_rn(tt2[j], p[j]));
			alpha = __fdiv_rn(-s2, s0);
		}
		alpha_s = alpha;
	}
	__syncthreads();
	const float alpha = alpha_s;
	float *y = Y + (size_t)b * ldy;
	for (int i = threadIdx.x; i < N; i += blockDim.x) y[i] = __fadd_rn(y[i], __fmul_rn(alpha, p[i]));
}

/* ws: [3][B][N] scratch (ph, ph'Qd, y'Qd + Fd') */
cudaError_t pqp_launch_accel_step(float *Y, int ldy, const float *Q, int ldq, const float *Fd, float *ws, int B, int N, cudaStream_t s)
{
	float *ph = ws, *t = ws + (size_t)B * N, *t2 = t + (size_t)B * N;
	for (int b0 = 0; b0 < B; b0 += 32 * 65535) { /* gridDim.y limit */
		const int nb = min(B - b0, 32 * 65535);
		const dim3 grid((N + 31) / 32, (nb + 31) / 32);
		const size_t o = (size_t)b0 * N;
		seqdot_kernel<0, 2><<<grid, 256, 0, s>>>(ph + o, Q, ldq, Y + (size_t)b0 * ldy, ldy, Fd + o, nb, N);
		seqdot_kernel<1, 0><<<grid, 256, 0, s>>>(t + o, Q, ldq, ph + o, N, NULL, nb, N);
		seqdot_kernel<1, 1><<<grid, 256, 0, s>>>(t2 + o, Q, ldq, Y + (size_t)b0 * ldy, ldy, Fd + o, nb, N);
	}
	accel_finish_kernel<<<B, 256, 0, s>>>(Y, ldy, ph, t, t2, N);
	return cudaGetLastError();
}
