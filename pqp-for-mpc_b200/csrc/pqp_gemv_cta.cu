/*
 * pqp_gemv_cta.cu -- the PQP loop for a problem that fits ONE thread block (N <= 128; the shipped example has N = 28).
 *
 * With a handful of rows there is nothing to spread over SMs: the multi-CTA kernels pay ~1 us per update for the exchange of y
 * through L2, the CPU reference 2 us per update at N = 28.  Here the whole solve (PQP_CPU.c:718-740) runs in one CTA: four
 * threads per row, each holding its quarter of the row in registers already split into max(q,0) and max(-q,0), y double
 * buffered in shared memory, one __syncthreads per update.  The stop test of terminate() (PQP_CPU.c:673-687, on g = Qd y + Fd)
 * is a block reduction every check_every updates (TOL), and the status block of a fixed-count solve one more pass at the end.
 * Summation order: thread part p of row i sums columns p*CPT .. p*CPT+CPT-1 ascending, the four parts combine by xor-shuffle
 * 1, 2 -- fixed, so results are reproducible bit for bit (FAST order; the STRICT kernel keeps the reference's own order).
 */
#include "pqp_internal.h"

#define CT_TPR 4     /* threads per row */
#define CT_MAXN 128

template <int CPT, bool TOL> __global__ void __launch_bounds__(CT_TPR *CT_MAXN, 1) gemv_cta_kernel(const pqp_gemv_args a0, int fd_stride, int y_stride)
{
	/* one block per problem: block b takes Fd, y, Md and the status block of problem b (strides 0: a single problem) */
	pqp_gemv_args a = a0;
	a.Fd += (size_t)blockIdx.x * fd_stride;
	a.ybuf0 += (size_t)blockIdx.x * y_stride;
	a.ybuf1 += (size_t)blockIdx.x * y_stride;
	a.status += blockIdx.x;
	if (a.Md && fd_stride) a.Md += blockIdx.x;
	__shared__ float ys[2][CT_MAXN + 32];
	__shared__ float red[16][8];
	__shared__ int stop_s;
	const int tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
	const int N = a.N, row = tid / CT_TPR, part = tid % CT_TPR;
	const bool active = row < N;
	const int c0 = part * CPT;

	/* this thread's slice of its row, split once */
	float qp[CPT], qn[CPT];
#pragma unroll
	for (int j = 0; j < CPT; j++) {
		const int c = c0 + j;
		const float q = (active && c < N) ? a.Q[(size_t)row * a.ldq + c] : 0.0f;
		qp[j] = fmaxf(q, 0.0f);
		qn[j] = fmaxf(-q, 0.0f);
	}
	float th = 0.0f, fd = 0.0f, kp_tol = a.eac;
	if (active) {
		th = a.theta[row];
		fd = a.Fd[row];
		if (a.Kp) kp_tol = fmaxf(a.erc * a.Kp[row], a.eac);
	}
	const float fdp = fmaxf(fd, 0.0f), fdn = fmaxf(-fd, 0.0f);
	for (int i = tid; i < CT_MAXN + 32; i += blockDim.x) {
		ys[0][i] = i < N ? a.ybuf0[i] : 0.0f;
		ys[1][i] = 0.0f;
	}
	if (tid == 0) stop_s = 0;
	__syncthreads();

	const int updates = TOL ? a.max_iters : a.iters;
	int done = 0, converged = 0;
	int next_chk = 0;
	float s_min = 0.0f, s_gap = 0.0f, s_jd = 0.0f, s_kkt = 0.0f;
	for (int p = 0;; p++) {
		const float *y = ys[p & 1];
		float num = 0.0f, den = 0.0f;
#pragma unroll
		for (int j = 0; j < CPT; j += 4) {
			const float4 yv = *reinterpret_cast<const float4 *>(y + c0 + j);
			den = fmaf(qp[j], yv.x, den); num = fmaf(qn[j], yv.x, num);
			if (j + 1 < CPT) { den = fmaf(qp[j + 1], yv.y, den); num = fmaf(qn[j + 1], yv.y, num); }
			if (j + 2 < CPT) { den = fmaf(qp[j + 2], yv.z, den); num = fmaf(qn[j + 2], yv.z, num); }
			if (j + 3 < CPT) { den = fmaf(qp[j + 3], yv.w, den); num = fmaf(qn[j + 3], yv.w, num); }
		}
		num += __shfl_xor_sync(0xffffffffu, num, 1);
		den += __shfl_xor_sync(0xffffffffu, den, 1);
		num += __shfl_xor_sync(0xffffffffu, num, 2);
		den += __shfl_xor_sync(0xffffffffu, den, 2);
		const float y_mine = active ? y[row] : 0.0f;
		num = fmaf(th, y_mine, num) + fdn;
		den = fmaf(th, y_mine, den) + fdp;

		const bool last = (p == updates);
		const bool chk = last || (TOL && p == next_chk);
		if (chk) {
			/* the terms of the stop test / status block on y_p */
			float e_min = INFINITY, e_gap = 0.0f, e_jd = 0.0f, e_kkt = 0.0f, e_viol = -INFINITY;
			if (active && part == 0) {
				const float gq = den - num;
				e_min = gq;
				e_gap = y_mine * gq;
				e_jd = y_mine * (0.5f * (gq + fd));
				e_kkt = fabsf(fminf(y_mine, gq));
				e_viol = -gq - kp_tol;
			}
#pragma unroll
			for (int o = 16; o; o >>= 1) {
				e_min = fminf(e_min, __shfl_xor_sync(0xffffffffu, e_min, o));
				e_gap += __shfl_xor_sync(0xffffffffu, e_gap, o);
				e_jd += __shfl_xor_sync(0xffffffffu, e_jd, o);
				e_kkt = fmaxf(e_kkt, __shfl_xor_sync(0xffffffffu, e_kkt, o));
				e_viol = fmaxf(e_viol, __shfl_xor_sync(0xffffffffu, e_viol, o));
			}
			if (lane == 0) {
				red[warp][0] = e_min; red[warp][1] = e_gap; red[warp][2] = e_jd; red[warp][3] = e_kkt; red[warp][4] = e_viol;
			}
			__syncthreads();
			if (tid == 0) {
				const int nw = (blockDim.x + 31) / 32;
				for (int w = 1; w < nw; w++) {
					e_min = fminf(e_min, red[w][0]); e_gap += red[w][1]; e_jd += red[w][2];
					e_kkt = fmaxf(e_kkt, red[w][3]); e_viol = fmaxf(e_viol, red[w][4]);
				}
				const float Jd = e_jd + (a.Md ? 0.5f * a.Md[0] : 0.0f);
				const int conv = TOL && e_viol <= 0.0f && fabsf(e_gap) <= a.eaj && fabsf(e_gap) <= a.erj * fabsf(Jd);
				red[0][0] = e_min; red[0][1] = e_gap; red[0][2] = Jd; red[0][3] = e_kkt;
				stop_s = (conv || last) ? (conv ? 2 : 1) : 0;
			}
			__syncthreads();
			if (stop_s) {
				done = p;
				converged = stop_s == 2;
				s_min = red[0][0]; s_gap = red[0][1]; s_jd = red[0][2]; s_kkt = red[0][3];
				if (active && part == 0) a.ybuf1[row] = y_mine;
				break;
			}
			if (TOL && !last) next_chk += a.check_every;
		}
		if (active && part == 0) ys[(p + 1) & 1][row] = __fdiv_rn(num, den) * y_mine;
		__syncthreads();
	}
	if (tid == 0) {
		pqp_status o;
		o.iters = done;
		o.converged = converged;
		o.min_slack = s_min;
		o.gap = s_gap;
		o.Jd = s_jd;
		o.kkt = s_kkt;
		*a.status = o;
		*a.result_buf = 1;
	}
}

int pqp_gemv_cta_supported(int N) { return N >= 1 && N <= CT_MAXN; }

template <bool TOL> static const void *cta_fn(int cpt)
{
	switch (cpt) {
	case 8: return (const void *)gemv_cta_kernel<8, TOL>;
	case 16: return (const void *)gemv_cta_kernel<16, TOL>;
	default: return (const void *)gemv_cta_kernel<32, TOL>;
	}
}

/* B problems sharing Q, one block each (B = 1: the single-problem call): Fd, y_0 (ybuf0), the result (ybuf1, may be ybuf0: a block reads
 * all of its y_0 before it writes), Md and the status blocks of problem b at b * stride */
cudaError_t pqp_launch_gemv_cta_batch(const pqp_gemv_args *a, int B, int fd_stride, int y_stride, cudaStream_t s)
{
	const int N = a->N;
	const int cpt = N <= 32 ? 8 : (N <= 64 ? 16 : 32);
	const void *fn = a->iters > 0 ? cta_fn<false>(cpt) : cta_fn<true>(cpt);
	int threads = (CT_TPR * N + 31) / 32 * 32;
	if (threads < 32) threads = 32;
	pqp_gemv_args args = *a;
	void *params[] = { (void *)&args, (void *)&fd_stride, (void *)&y_stride };
	return cudaLaunchKernel(fn, dim3(B), dim3(threads), params, 0, s);
}

/* result left in ybuf1, status written by the kernel */
cudaError_t pqp_launch_gemv_cta(const pqp_gemv_args *a, cudaStream_t s) { return pqp_launch_gemv_cta_batch(a, 1, 0, 0, s); }
