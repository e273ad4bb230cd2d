/*
 * pqp_gemv.cu -- single-problem PQP iteration loop for sm_100a (B200): the hot path.
 *
 * Replaces the loop of solveQuadraticDual (PQP_CPU.c:718-740): updateY2 (:603-618) + updY
 * (:590-596) + copyMatrix (:737) + terminate (:673-687), which the reference GPU build runs as
 * 45 launches, 20 cudaMalloc/Free pairs and 2 blocking D2H copies PER ITERATION
 * (PQP_GPU_optimized.cu:799-818).  Here the whole loop is ONE cooperative launch:
 *
 *   - one CTA per SM, each owning a contiguous slab of rows of the signed Qd (stored once,
 *     [N x ldq] fp32; the Q+/Q- split of PQP_CPU.c:524-537 is two FMNMX in registers);
 *   - per iteration every CTA streams its slab with 128-bit coalesced loads, 16 independent
 *     loads in flight per thread (4 rows x 4 column chunks), y staged in shared memory;
 *   - the first `resident_rows` rows of each slab live in shared memory for the whole launch,
 *     so for small N (whole Q on chip: N=1024 is 28 KB per SM) no iteration touches HBM or L2
 *     for Q at all, and for large N they simply shave that fraction off the HBM stream;
 *   - num/den partial sums: warp shuffle tree, then a fixed-order sum over the 16 warps, then
 *     the multiplicative update y_i <- (num_i/den_i)*y_i with IEEE division, all in-kernel;
 *   - y ping-pongs between two global vectors (L2-resident); a release/acquire grid barrier
 *     separates iterations; no host round trip until the loop is finished;
 *   - the stop-test quantities (min g, y'g, Jd, ||min(y,g)||inf with g = den - num = Qd y + Fd,
 *     SURVEY 3.3) are reduced from values already in registers on the passes that need them.
 *
 * Summation order (FAST): for row i, lane l of warp w adds columns {4*(w*32+l+512*u)+e} in
 * ascending u,e; lanes are combined by xor-shuffle (16,8,4,2,1), warps in ascending w, then
 * theta_i*y_i and F are added.  It depends only on N, never on the SM count or the grid, so
 * results are reproducible across devices and across 1..8 GPU runs.
 *
 * Algorithmic bytes per iteration: 4*N*ldq (Q) + 16*N (y in, y out, F+ F- as Fd, theta).
 */
#include "pqp_internal.h"

#include <cooperative_groups.h>

#define RG 4 /* rows in flight per warp */
#define UN 4 /* column chunks (of 512 float4) in flight per warp */

__device__ __forceinline__ float4 ldg_stream(const float4 *p)
{
	float4 r;
	asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.f32 {%0,%1,%2,%3}, [%4];"
		     : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
		     : "l"(p));
	return r;
}

__device__ __forceinline__ void grid_barrier(unsigned *counter, unsigned &target, unsigned nblocks)
{
	__syncthreads();
	if (threadIdx.x == 0) {
		target += nblocks;
		__threadfence();
		asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
		unsigned v;
		do {
			asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
		} while ((int)(v - target) < 0);
		__threadfence();
	}
	__syncthreads();
}

__device__ __forceinline__ void acc4(float &num, float &den, const float4 q, const float4 y)
{
	den = fmaf(fmaxf(q.x, 0.0f), y.x, den);
	num = fmaf(fmaxf(-q.x, 0.0f), y.x, num);
	den = fmaf(fmaxf(q.y, 0.0f), y.y, den);
	num = fmaf(fmaxf(-q.y, 0.0f), y.y, num);
	den = fmaf(fmaxf(q.z, 0.0f), y.z, den);
	num = fmaf(fmaxf(-q.z, 0.0f), y.z, num);
	den = fmaf(fmaxf(q.w, 0.0f), y.w, den);
	num = fmaf(fmaxf(-q.w, 0.0f), y.w, num);
}

/* quantities of the stop test, combined in a fixed order */
struct Eval {
	float vmin, gap, jd, kkt, viol;
};
__device__ __forceinline__ Eval eval_identity()
{
	Eval e;
	e.vmin = INFINITY; e.gap = 0.0f; e.jd = 0.0f; e.kkt = 0.0f; e.viol = -INFINITY;
	return e;
}
__device__ __forceinline__ Eval eval_combine(Eval a, Eval b)
{
	a.vmin = fminf(a.vmin, b.vmin); a.gap += b.gap; a.jd += b.jd;
	a.kkt = fmaxf(a.kkt, b.kkt); a.viol = fmaxf(a.viol, b.viol);
	return a;
}
__device__ __forceinline__ Eval eval_shfl_xor(Eval e, int o)
{
	Eval r;
	r.vmin = __shfl_xor_sync(0xffffffffu, e.vmin, o);
	r.gap = __shfl_xor_sync(0xffffffffu, e.gap, o);
	r.jd = __shfl_xor_sync(0xffffffffu, e.jd, o);
	r.kkt = __shfl_xor_sync(0xffffffffu, e.kkt, o);
	r.viol = __shfl_xor_sync(0xffffffffu, e.viol, o);
	return r;
}

/*
 * Shared memory: y_s[ldq] | part[2][16][rows_max] (num, den partials per warp) | red[16*5] |
 * resident slab [resident_rows x ldq]
 */
__global__ void __launch_bounds__(PQP_GEMV_THREADS, 1) gemv_persistent_kernel(const pqp_gemv_args a, int rows_max)
{
	extern __shared__ __align__(16) float smem[];
	const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
	const int N = a.N, ldq = a.ldq, n4 = ldq / 4;
	const unsigned G = gridDim.x;
	const int r0 = (int)((long long)N * blockIdx.x / G), r1 = (int)((long long)N * (blockIdx.x + 1) / G);
	const int nrows = r1 - r0;

	float *y_s = smem;
	float *part = y_s + ldq;                          /* [2][WARPS][rows_max] */
	float *red = part + 2 * PQP_GEMV_WARPS * rows_max; /* [WARPS][5] */
	float4 *res4 = reinterpret_cast<float4 *>(red + PQP_GEMV_WARPS * 8);
	const int res_rows = min(a.resident_rows, nrows);

	/* one-time: park the resident rows of this CTA's slab in shared memory */
	for (int r = 0; r < res_rows; r++) {
		const float4 *src = reinterpret_cast<const float4 *>(a.Q + (size_t)(r0 + r) * ldq);
		for (int c = tid; c < n4; c += PQP_GEMV_THREADS) res4[(size_t)r * n4 + c] = ldg_stream(src + c);
	}

	const bool fixed = a.iters > 0;
	const int last_pass = fixed ? a.iters : a.max_iters;
	unsigned bar_target = 0;
	int evals = 0;
	const float4 *y_s4 = reinterpret_cast<const float4 *>(y_s);

	for (int p = 0;; p++) {
		const float *y_in = (p & 1) ? a.ybuf1 : a.ybuf0;
		float *y_out = (p & 1) ? a.ybuf0 : a.ybuf1;
		const bool is_last = (p == last_pass);
		const bool do_eval = is_last || (!fixed && (p % a.check_every == 0));

		/* stage y (written by all CTAs in the previous pass; L2 is the point of coherence) */
		for (int c = tid; c < n4; c += PQP_GEMV_THREADS)
			reinterpret_cast<float4 *>(y_s)[c] = __ldcg(reinterpret_cast<const float4 *>(y_in) + c);
		__syncthreads();

		/* ---- stream the slab: RG rows x UN chunks per warp-step -------------------------- */
		for (int rg = 0; rg < nrows; rg += RG) {
			float num[RG], den[RG];
#pragma unroll
			for (int r = 0; r < RG; r++) num[r] = den[r] = 0.0f;

			if (rg + RG <= res_rows) {
				/* fully resident group: shared memory only */
				for (int c = tid; c < n4; c += PQP_GEMV_THREADS) {
					const float4 y4 = y_s4[c];
#pragma unroll
					for (int r = 0; r < RG; r++) acc4(num[r], den[r], res4[(size_t)(rg + r) * n4 + c], y4);
				}
			} else {
				const float4 *row4[RG];
				bool rv[RG];
#pragma unroll
				for (int r = 0; r < RG; r++) {
					rv[r] = (rg + r) < nrows;
					row4[r] = reinterpret_cast<const float4 *>(a.Q + (size_t)(r0 + (rv[r] ? rg + r : 0)) * ldq);
				}
				for (int cb = 0; cb < n4; cb += PQP_GEMV_THREADS * UN) {
					float4 q[UN][RG];
#pragma unroll
					for (int u = 0; u < UN; u++) {
						const int c = cb + u * PQP_GEMV_THREADS + tid;
#pragma unroll
						for (int r = 0; r < RG; r++) {
							if (c < n4 && rv[r]) {
								if (rg + r < res_rows)
									q[u][r] = res4[(size_t)(rg + r) * n4 + c];
								else
									q[u][r] = ldg_stream(row4[r] + c);
							} else {
								q[u][r] = make_float4(0.f, 0.f, 0.f, 0.f);
							}
						}
					}
#pragma unroll
					for (int u = 0; u < UN; u++) {
						const int c = cb + u * PQP_GEMV_THREADS + tid;
						const float4 y4 = (c < n4) ? y_s4[c] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
						for (int r = 0; r < RG; r++) acc4(num[r], den[r], q[u][r], y4);
					}
				}
			}
#pragma unroll
			for (int r = 0; r < RG; r++) {
#pragma unroll
				for (int o = 16; o; o >>= 1) {
					num[r] += __shfl_xor_sync(0xffffffffu, num[r], o);
					den[r] += __shfl_xor_sync(0xffffffffu, den[r], o);
				}
			}
			if (lane == 0) {
#pragma unroll
				for (int r = 0; r < RG; r++)
					if (rg + r < nrows) {
						part[(0 * PQP_GEMV_WARPS + warp) * rows_max + rg + r] = num[r];
						part[(1 * PQP_GEMV_WARPS + warp) * rows_max + rg + r] = den[r];
					}
			}
		}
		__syncthreads();

		/* ---- finish rows: fixed-order warp sum, theta, F, update, stop-test terms ----------- */
		Eval ev = eval_identity();
		for (int t = tid; t < nrows; t += PQP_GEMV_THREADS) {
			const int i = r0 + t;
			float num = 0.0f, den = 0.0f;
#pragma unroll
			for (int w = 0; w < PQP_GEMV_WARPS; w++) {
				num += part[(0 * PQP_GEMV_WARPS + w) * rows_max + t];
				den += part[(1 * PQP_GEMV_WARPS + w) * rows_max + t];
			}
			const float yi = y_s[i], th = a.theta[i], fd = a.Fd[i];
			num = fmaf(th, yi, num) + fmaxf(-fd, 0.0f);
			den = fmaf(th, yi, den) + fmaxf(fd, 0.0f);
			if (!is_last) y_out[i] = __fdiv_rn(num, den) * yi;
			if (do_eval) {
				const float g = den - num; /* theta*y cancels: g = Qd y + Fd */
				Eval e;
				e.vmin = g;
				e.gap = yi * g;
				e.jd = yi * (0.5f * (g + fd));
				e.kkt = fabsf(fminf(yi, g));
				e.viol = -g - (a.Kp ? fmaxf(a.erc * a.Kp[i], a.eac) : a.eac);
				ev = eval_combine(ev, e);
			}
		}
		if (do_eval) {
#pragma unroll
			for (int o = 16; o; o >>= 1) ev = eval_combine(ev, eval_shfl_xor(ev, o));
			if (lane == 0) {
				red[warp * 8 + 0] = ev.vmin; red[warp * 8 + 1] = ev.gap; red[warp * 8 + 2] = ev.jd;
				red[warp * 8 + 3] = ev.kkt; red[warp * 8 + 4] = ev.viol;
			}
			__syncthreads();
			if (tid == 0) {
				Eval t = eval_identity();
				for (int w = 0; w < PQP_GEMV_WARPS; w++) {
					Eval e;
					e.vmin = red[w * 8 + 0]; e.gap = red[w * 8 + 1]; e.jd = red[w * 8 + 2];
					e.kkt = red[w * 8 + 3]; e.viol = red[w * 8 + 4];
					t = eval_combine(t, e);
				}
				float *slot = a.partials + ((size_t)(evals & 1) * G + blockIdx.x) * 8;
				slot[0] = t.vmin; slot[1] = t.gap; slot[2] = t.jd; slot[3] = t.kkt; slot[4] = t.viol;
			}
		}

		grid_barrier(a.barrier, bar_target, G);

		if (do_eval) {
			/* every CTA folds all CTAs' partials in the same order -> the same decision everywhere */
			Eval t = eval_identity();
			if (warp == 0) {
				const float *base = a.partials + (size_t)(evals & 1) * G * 8;
				for (unsigned c = lane; c < G; c += 32) {
					Eval e;
					e.vmin = __ldcg(base + c * 8 + 0); e.gap = __ldcg(base + c * 8 + 1);
					e.jd = __ldcg(base + c * 8 + 2); e.kkt = __ldcg(base + c * 8 + 3);
					e.viol = __ldcg(base + c * 8 + 4);
					t = eval_combine(t, e);
				}
#pragma unroll
				for (int o = 16; o; o >>= 1) t = eval_combine(t, eval_shfl_xor(t, o));
				if (lane == 0) {
					const float Jd = t.jd + (a.Md ? 0.5f * a.Md[0] : 0.0f);
					const bool conv = !fixed && t.viol <= 0.0f && fabsf(t.gap) <= a.eaj &&
							  fabsf(t.gap) <= a.erj * fabsf(Jd);
					red[0] = (conv || is_last) ? 1.0f : 0.0f;
					if ((conv || is_last) && blockIdx.x == 0) {
						pqp_status o;
						o.iters = p;
						o.converged = conv ? 1 : 0;
						o.min_slack = t.vmin;
						o.gap = t.gap;
						o.Jd = Jd;
						o.kkt = t.kkt;
						*a.status = o;
						*a.result_buf = p & 1;
					}
				}
			}
			evals++;
			__syncthreads();
			if (red[0] != 0.0f) break;
			__syncthreads(); /* red[] is rewritten by the next eval pass */
		}
	}
}

cudaError_t pqp_gemv_smem_bytes(int N, int ldq, int grid, int resident_rows, size_t *bytes)
{
	const int rows_max = (N + grid - 1) / grid + 1;
	*bytes = sizeof(float) * ((size_t)ldq + 2 * PQP_GEMV_WARPS * (size_t)rows_max + PQP_GEMV_WARPS * 8 +
				  (size_t)resident_rows * ldq);
	return cudaSuccess;
}

cudaError_t pqp_launch_gemv_persistent(const pqp_gemv_args *a, cudaStream_t s)
{
	size_t smem;
	pqp_gemv_smem_bytes(a->N, a->ldq, a->grid, a->resident_rows, &smem);
	cudaError_t e = cudaFuncSetAttribute(gemv_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	e = cudaMemsetAsync(a->barrier, 0, sizeof(unsigned), s);
	if (e != cudaSuccess) return e;
	int rows_max = (a->N + a->grid - 1) / a->grid + 1;
	pqp_gemv_args args = *a;
	void *params[] = { (void *)&args, (void *)&rows_max };
	return cudaLaunchCooperativeKernel((const void *)gemv_persistent_kernel, dim3(a->grid), dim3(PQP_GEMV_THREADS), params,
					   smem, s);
}

/* ---------------------------------------------------------------------------------------------
 * STRICT step: thread i owns row i and walks k ascending over the TRANSPOSE (coalesced across
 * threads), reproducing PQP_CPU.c:608-614 bit for bit:
 *   num = sum_k fl( (max(0,-q_ik) [+theta_i if k==i]) * y_k ),  den likewise with max(0,q_ik);
 *   num += 1*F-_i; den += 1*F+_i;  y+_i = (num/den)*y_i.
 * Explicit zero terms of the reference's dense split matrices are added too (they are exact).
 * ------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(128) gemv_strict_kernel(const float *__restrict__ QT, int ldq, int N,
							   const float *__restrict__ theta, const float *__restrict__ Fd,
							   const float *__restrict__ y_in, float *__restrict__ y_out)
{
	extern __shared__ float ys[];
	for (int k = threadIdx.x; k < N; k += blockDim.x) ys[k] = y_in[k];
	__syncthreads();
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= N) return;
	const float th = theta[i];
	float num = 0.0f, den = 0.0f;
	for (int k = 0; k < N; k++) {
		const float q = QT[(size_t)k * ldq + i];
		float qn = (0.0f > -q) ? 0.0f : -q; /* matrixNeg, PQP_CPU.c:209 */
		float qp = (0.0f > q) ? 0.0f : q;   /* matrixPos, PQP_CPU.c:192 */
		if (k == i) {                       /* + theta on the diagonal, PQP_CPU.c:527,536 */
			qn = __fadd_rn(qn, th);
			qp = __fadd_rn(qp, th);
		}
		const float yk = ys[k];
		num = __fadd_rn(num, __fmul_rn(qn, yk));
		den = __fadd_rn(den, __fmul_rn(qp, yk));
	}
	const float fd = Fd[i];
	const float fdn = (0.0f > -fd) ? 0.0f : -fd, fdp = (0.0f > fd) ? 0.0f : fd; /* PQP_CPU.c:703-704 */
	num = __fadd_rn(num, __fmul_rn(1.0f, fdn));
	den = __fadd_rn(den, __fmul_rn(1.0f, fdp));
	y_out[i] = __fmul_rn(__fdiv_rn(num, den), ys[i]);
}

cudaError_t pqp_launch_gemv_strict_step(const pqp_gemv_args *a, const float *y_in, float *y_out, cudaStream_t s)
{
	const size_t smem = sizeof(float) * (size_t)a->N;
	if (smem > 48 * 1024) { /* per device and context, so set on every launch that needs it (a cheap host-side call) */
		cudaError_t e = cudaFuncSetAttribute(gemv_strict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
		if (e != cudaSuccess) return e;
	}
	gemv_strict_kernel<<<(a->N + 127) / 128, 128, smem, s>>>(a->QT, a->ldq, a->N, a->theta, a->Fd, y_in, y_out);
	return cudaGetLastError();
}
