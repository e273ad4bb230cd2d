/*
 * pqp_gemv_small.cu -- single-problem PQP loop for problems whose whole Hessian fits ON CHIP IN REGISTERS
 * (N <= ~2k: config C1 = the shipped example, config C2 = generator N=1024).  sm_100a.
 *
 * Same loop as pqp_gemv_tma.cu (PQP_CPU.c:718-740 as one cooperative launch), re-shaped for the regime where the
 * iteration is pure latency, not bandwidth:
 *   - each CTA owns <= 16 rows of the signed Qd, ALL of them processed concurrently: 16/wpr rows at a time, wpr
 *     warps per row, every thread keeping its slice of the row in REGISTERS for the whole solve (no shared-memory
 *     or L2 traffic for Q inside the loop at all);
 *   - y travels as {value, epoch} packets (flag-in-data): one 64-bit store per dual, polling loads on the reader
 *     side, so the grid barrier and the reload of y are one L2 round trip; each CTA fetches y once per iteration
 *     into (double-buffered) shared memory and its 16 warps read it from there;
 *   - per iteration: one cooperative poll-load of y, one __syncthreads (two when several warps share a row),
 *     8*CPT FMNMX+FFMA pairs per thread, one shuffle tree, one 64-bit store per row.
 *
 * Summation order: lane l of part w of a row adds float4 columns w*32 + l + 32*wpr*u (u ascending, even and odd u
 * in two interleaved chains that are added at the end), xor-shuffle 16..1, parts ascending.  Deterministic and
 * independent of the SM count, but NOT the order of the streaming kernels (last-bit differences).
 */
#include "pqp_internal.h"

#include <type_traits>

#ifndef SMALL_TOL_DBG
#define SMALL_TOL_DBG 0 /* experiment switch: 1 no decision read, 2 no check publish */
#endif
#define SM_THREADS 512
#define SM_WARPS 16

__device__ __forceinline__ void sm_st_packet(uint2 *dst, float v, uint32_t epoch)
{
	asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(__float_as_uint(v)), "r"(epoch) : "memory");
}
__device__ __forceinline__ float sm_ld_packet(const uint2 *src, uint32_t epoch)
{
	uint32_t v, e;
	for (;;) {
		asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(e) : "l"(src) : "memory");
		if (e == epoch) break;
	}
	return __uint_as_float(v);
}
__device__ __forceinline__ float4 sm_ld_packet4(const uint2 *src, uint32_t epoch)
{
	uint32_t a0, e0, a1, e1, a2, e2, a3, e3;
	for (;;) {
		asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(e0), "=r"(a1), "=r"(e1) : "l"(src) : "memory");
		asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a2), "=r"(e2), "=r"(a3), "=r"(e3) : "l"(src + 2) : "memory");
		if (e0 == epoch && e1 == epoch && e2 == epoch && e3 == epoch) break;
	}
	return make_float4(__uint_as_float(a0), __uint_as_float(a1), __uint_as_float(a2), __uint_as_float(a3));
}
/* two values and their epochs in one 16-byte store / polling load */
__device__ __forceinline__ void sm_st_packet2(uint4 *dst, float v0, float v1, uint32_t epoch)
{
	asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(__float_as_uint(v0)), "r"(epoch), "r"(__float_as_uint(v1)), "r"(epoch)
		     : "memory");
}
__device__ __forceinline__ void sm_ld_packet2(const uint4 *src, uint32_t epoch, float &v0, float &v1)
{
	uint32_t a0, e0, a1, e1;
	for (;;) {
		asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(e0), "=r"(a1), "=r"(e1) : "l"(src) : "memory");
		if (e0 == epoch && e1 == epoch) break;
	}
	v0 = __uint_as_float(a0);
	v1 = __uint_as_float(a1);
}
__device__ __forceinline__ void sm_acc4(float &num, float &den, const float4 q, const float4 y)
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
__device__ __forceinline__ void sm_grid_barrier(unsigned *counter, unsigned &target, unsigned nblocks)
{
	__syncthreads();
	if (threadIdx.x == 0) {
		target += nblocks;
		asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
		unsigned v;
		do {
			asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
		} while ((int)(v - target) < 0);
	}
	__syncthreads();
}

/*
 * TOL = true: run to the stop test (a.iters <= 0).  Every a.check_every passes (and at a.max_iters) the pass first evaluates
 * the stop test of terminate() on g = den - num = Qd y + Fd BEFORE applying its update: per-CTA partial results
 * travel as {value, epoch} packets like y does, every CTA folds all of them in the same order and so takes the same decision;
 * a converged problem leaves with exactly the y that passed.  One extra L2 round trip per check, none per update.
 */
template <int CPT, bool TOL>
__global__ void __launch_bounds__(SM_THREADS, 1) gemv_small_kernel(const pqp_gemv_args a, int wpr, uint2 *pk0, uint2 *pk1)
{
	extern __shared__ __align__(16) float y_s[]; /* [2][ldq]: y of the current / next pass */
	__shared__ float part_s[2][SM_WARPS][2]; /* [parity][warp][num, den] */
	__shared__ float ev_s[SM_WARPS][8];
	const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
	const int N = a.N, ldq = a.ldq, n4 = ldq / 4;
	const unsigned G = gridDim.x;
	const int r0 = (int)((long long)N * blockIdx.x / G), r1 = (int)((long long)N * (blockIdx.x + 1) / G);
	const int nrows = r1 - r0;
	const int slot = warp / wpr, part = warp % wpr;
	const bool active = slot < nrows;
	const int row = r0 + (active ? slot : 0);
	const int stride = 32 * wpr;
	const bool finisher = active && part == 0 && lane == 0;

	/* this thread's slice of its row, for the whole solve: the first CREG float4 in registers, the rest (16-column
	 * instantiation only) in shared memory -- 64 registers of q plus the loop state spilled 46 registers into the loop
	 * (measured 4.0 us/update at N=1152 against 1.33 at N=1024) */
	constexpr int CREG = CPT > 8 ? 8 : CPT;
	float4 q[CREG];
	float4 *q_s = reinterpret_cast<float4 *>(y_s + 2 * (size_t)ldq); /* [CPT - CREG][SM_THREADS] */
#pragma unroll
	for (int u = 0; u < CPT; u++) {
		const int c = part * 32 + lane + stride * u;
		const float4 v = (active && c < n4) ? __ldg(reinterpret_cast<const float4 *>(a.Q + (size_t)row * ldq) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
		if (u < CREG) q[u < CREG ? u : 0] = v;
		else q_s[(u - CREG) * SM_THREADS + tid] = v;
	}
	float th_r = 0.0f, fd_r = 0.0f, kp_tol = a.eac;
	if (active) {
		th_r = a.theta[row];
		fd_r = a.Fd[row];
		if (a.Kp) kp_tol = fmaxf(a.erc * a.Kp[row], a.eac);
	}
	const int passes = TOL ? a.max_iters + 1 : a.iters + 1;
	unsigned bar_target = 0;
	uint4 *cpk4 = reinterpret_cast<uint4 *>(a.partials); /* TOL: [2][G][3] check packets of 16 bytes */
	float *fin = a.partials + (TOL ? (size_t)2 * G * 12 : 0); /* final per-CTA slots, clear of packets a slow CTA may still be polling */
	__shared__ float chk_s[TOL ? 148 : 1][6];

	/* one pass; CHK / DEC are compile-time so that the plain pass carries none of the stop-test code (with runtime flags the
	 * skipped blocks still cost 0.7 us per pass).  Returns true when the kernel is done. */
	auto pass = [&](const int p, auto chk_c, auto dec_c) -> bool {
		constexpr bool CHK = decltype(chk_c)::value, DEC = decltype(dec_c)::value;
		const bool is_last = (p == passes - 1);
		const uint2 *pk_in = (p & 1) ? pk1 : pk0;
		uint2 *pk_out = (p & 1) ? pk0 : pk1;
		/* all threads: fetch y of this pass (packets -> plain floats in shared memory), once per CTA */
		float4 *ys4 = reinterpret_cast<float4 *>(y_s + (size_t)(p & 1) * ldq);
		for (int c = tid; c < n4; c += SM_THREADS) {
			float4 y4 = make_float4(0.f, 0.f, 0.f, 0.f);
			if (p == 0) {
				y4 = __ldcg(reinterpret_cast<const float4 *>(a.ybuf0) + c);
			} else if (c * 4 + 3 < N) {
				y4 = sm_ld_packet4(pk_in + 4 * c, (uint32_t)p);
			} else {
				float t[4] = { 0.f, 0.f, 0.f, 0.f };
				for (int e = 0; e < 4; e++)
					if (c * 4 + e < N) t[e] = sm_ld_packet(pk_in + 4 * c + e, (uint32_t)p);
				y4 = make_float4(t[0], t[1], t[2], t[3]);
			}
			ys4[c] = y4;
		}
		constexpr bool decide = TOL && DEC;
		if (decide && tid < 3 * (int)G) {
			/* the stop test of pass p-1: one 16-byte polling load per thread, all in flight at once */
			const unsigned chk = (unsigned)((p - 1) / a.check_every) + 1u;
			const int c = tid / 3, k = tid % 3;
			float v0, v1;
			sm_ld_packet2(cpk4 + ((size_t)(chk & 1u) * G + c) * 3 + k, chk, v0, v1);
			chk_s[c][2 * k] = v0;
			chk_s[c][2 * k + 1] = v1;
		}
		__syncthreads();
		if (decide) {
			/* every warp of every CTA folds all CTAs' values in the same order, so all take the same decision */
			float v_min = INFINITY, v_gap = 0.0f, v_jd = 0.0f, v_kkt = 0.0f, v_viol = -INFINITY;
			for (unsigned c = lane; c < G; c += 32) {
				v_min = fminf(v_min, chk_s[c][0]); v_gap += chk_s[c][1]; v_jd += chk_s[c][2]; v_kkt = fmaxf(v_kkt, chk_s[c][3]);
				v_viol = fmaxf(v_viol, chk_s[c][4]);
			}
#pragma unroll
			for (int o = 16; o; o >>= 1) {
				v_min = fminf(v_min, __shfl_xor_sync(0xffffffffu, v_min, o));
				v_gap += __shfl_xor_sync(0xffffffffu, v_gap, o);
				v_jd += __shfl_xor_sync(0xffffffffu, v_jd, o);
				v_kkt = fmaxf(v_kkt, __shfl_xor_sync(0xffffffffu, v_kkt, o));
				v_viol = fmaxf(v_viol, __shfl_xor_sync(0xffffffffu, v_viol, o));
			}
			const float Jd = v_jd + (a.Md ? 0.5f * a.Md[0] : 0.0f);
			if (v_viol <= 0.0f && fabsf(v_gap) <= a.eaj && fabsf(v_gap) <= a.erj * fabsf(Jd)) {
				/* converged at y_{p-1}, which still sits in the other half of the double-buffered shared copy */
				if (finisher) a.ybuf1[row] = y_s[(size_t)((p - 1) & 1) * ldq + row];
				if (blockIdx.x == 0 && tid == 0) {
					pqp_status o;
					o.iters = p - 1; o.converged = 1; o.min_slack = v_min; o.gap = v_gap; o.Jd = Jd; o.kkt = v_kkt;
					*a.status = o;
					*a.result_buf = 1;
				}
				return true; /* uniform across the grid */
			}
		}
		float num0 = 0.0f, den0 = 0.0f, num1 = 0.0f, den1 = 0.0f;
#pragma unroll
		for (int u = 0; u < CPT; u++) {
			const int c = part * 32 + lane + stride * u;
			const float4 y4 = (c < n4) ? ys4[c] : make_float4(0.f, 0.f, 0.f, 0.f);
			const float4 qv = u < CREG ? q[u < CREG ? u : 0] : q_s[(u - CREG) * SM_THREADS + tid];
			if (u & 1) sm_acc4(num1, den1, qv, y4);
			else sm_acc4(num0, den0, qv, y4);
		}
		float num = num0 + num1, den = den0 + den1;
#pragma unroll
		for (int o = 16; o; o >>= 1) {
			num += __shfl_xor_sync(0xffffffffu, num, o);
			den += __shfl_xor_sync(0xffffffffu, den, o);
		}
		if (wpr > 1) {
			if (lane == 0) {
				part_s[p & 1][warp][0] = num;
				part_s[p & 1][warp][1] = den;
			}
			__syncthreads();
			if (finisher) {
				num = 0.0f;
				den = 0.0f;
				for (int w = 0; w < wpr; w++) {
					num += part_s[p & 1][slot * wpr + w][0];
					den += part_s[p & 1][slot * wpr + w][1];
				}
			}
		}
		float e_min = INFINITY, e_gap = 0.0f, e_jd = 0.0f, e_kkt = 0.0f, e_viol = -INFINITY;
		if (finisher) {
			const float y_mine = y_s[(size_t)(p & 1) * ldq + row];
			num = fmaf(th_r, y_mine, num) + fmaxf(-fd_r, 0.0f);
			den = fmaf(th_r, y_mine, den) + fmaxf(fd_r, 0.0f);
		}
		if (TOL && CHK && !is_last) {
			/* ---- publish this CTA's share of the stop test on y_p; the decision is read at the start of the NEXT pass, when
			 * the packets have long arrived (no exchange latency added), and y_{p+1} is published below as always ---- */
			const unsigned chk = (unsigned)(p / a.check_every) + 1u;
			if (finisher) {
				const float y_mine = y_s[(size_t)(p & 1) * ldq + row];
				const float gq = den - num;
				e_min = gq; e_gap = y_mine * gq; e_jd = y_mine * (0.5f * (gq + fd_r)); e_kkt = fabsf(fminf(y_mine, gq)); e_viol = -gq - kp_tol;
			}
			if (lane == 0) {
				ev_s[warp][0] = e_min; ev_s[warp][1] = e_gap; ev_s[warp][2] = e_jd; ev_s[warp][3] = e_kkt; ev_s[warp][4] = e_viol;
			}
			__syncthreads();
			if (tid == 0) {
				float m = ev_s[0][0], ga = ev_s[0][1], jd = ev_s[0][2], kk = ev_s[0][3], vi = ev_s[0][4];
				for (int w = 1; w < SM_WARPS; w++) {
					m = fminf(m, ev_s[w][0]); ga += ev_s[w][1]; jd += ev_s[w][2]; kk = fmaxf(kk, ev_s[w][3]); vi = fmaxf(vi, ev_s[w][4]);
				}
				uint4 *mine = cpk4 + ((size_t)(chk & 1u) * G + blockIdx.x) * 3; /* {value, epoch, value, epoch} x 3, one 16-byte store each */
				sm_st_packet2(mine + 0, m, ga, chk);
				sm_st_packet2(mine + 1, jd, kk, chk);
				sm_st_packet2(mine + 2, vi, 0.0f, chk);
			}
			e_min = INFINITY; e_gap = 0.0f; e_jd = 0.0f; e_kkt = 0.0f; e_viol = -INFINITY;
		}
		if (finisher) {
			const float y_mine = y_s[(size_t)(p & 1) * ldq + row];
			if (!is_last) {
				sm_st_packet(pk_out + row, __fdiv_rn(num, den) * y_mine, (uint32_t)(p + 1));
			} else {
				a.ybuf1[row] = y_mine;
				const float gq = den - num;
				e_min = gq;
				e_gap = y_mine * gq;
				e_jd = y_mine * (0.5f * (gq + fd_r));
				e_kkt = fabsf(fminf(y_mine, gq));
				e_viol = -gq - kp_tol;
			}
		}
		if (is_last) {
			if (lane == 0) {
				ev_s[warp][0] = e_min; ev_s[warp][1] = e_gap; ev_s[warp][2] = e_jd; ev_s[warp][3] = e_kkt; ev_s[warp][4] = e_viol;
			}
			__syncthreads();
			if (tid == 0) {
				for (int w = 1; w < SM_WARPS; w++) {
					e_min = fminf(e_min, ev_s[w][0]); e_gap += ev_s[w][1]; e_jd += ev_s[w][2];
					e_kkt = fmaxf(e_kkt, ev_s[w][3]); e_viol = fmaxf(e_viol, ev_s[w][4]);
				}
				float *sl = fin + (size_t)blockIdx.x * 8;
				sl[0] = e_min; sl[1] = e_gap; sl[2] = e_jd; sl[3] = e_kkt; sl[4] = e_viol;
				__threadfence();
			}
			sm_grid_barrier(a.barrier, bar_target, G);
			if (blockIdx.x == 0 && warp == 0) {
				float v_min = INFINITY, v_gap = 0.0f, v_jd = 0.0f, v_kkt = 0.0f;
				for (unsigned c = lane; c < G; c += 32) {
					const float *sl = fin + (size_t)c * 8;
					v_min = fminf(v_min, __ldcg(sl + 0)); v_gap += __ldcg(sl + 1); v_jd += __ldcg(sl + 2);
					v_kkt = fmaxf(v_kkt, __ldcg(sl + 3));
				}
#pragma unroll
				for (int o = 16; o; o >>= 1) {
					v_min = fminf(v_min, __shfl_xor_sync(0xffffffffu, v_min, o));
					v_gap += __shfl_xor_sync(0xffffffffu, v_gap, o);
					v_jd += __shfl_xor_sync(0xffffffffu, v_jd, o);
					v_kkt = fmaxf(v_kkt, __shfl_xor_sync(0xffffffffu, v_kkt, o));
				}
				if (lane == 0) {
					pqp_status o;
					o.iters = TOL ? p : a.iters;
					o.converged = 0; /* TOL: this block is only reached at the cap; a converged run left at the decision above */
					o.min_slack = v_min;
					o.gap = v_gap;
					o.Jd = v_jd + (a.Md ? 0.5f * a.Md[0] : 0.0f);
					o.kkt = v_kkt;
					*a.status = o;
					*a.result_buf = 1;
				}
			}
			return true;
		}
		return false;
	};

	if (!TOL) {
		for (int p = 0; p < passes; p++)
			if (pass(p, std::false_type{}, std::false_type{})) break;
	} else {
		/* check passes at p = 0, check_every, 2*check_every, ...; the decision of a check is read at the start of the next pass */
		int next_chk = 0;
		bool pending = false;
		for (int p = 0; p < passes; p++) {
			const bool chk = (p == next_chk);
			bool done;
			if (chk && pending) done = pass(p, std::true_type{}, std::true_type{});
			else if (chk) done = pass(p, std::true_type{}, std::false_type{});
			else if (pending) done = pass(p, std::false_type{}, std::true_type{});
			else done = pass(p, std::false_type{}, std::false_type{});
			if (done) break;
			pending = chk;
			if (chk) next_chk += a.check_every;
		}
	}
}

/* wpr (warps per row, power of two) and columns per thread; returns 0 when the shape does not fit this kernel */
int pqp_gemv_small_plan(int N, int ldq, int grid, int *wpr_out, int *cpt_out)
{
	const int rows_max = (N + grid - 1) / grid;
	if (rows_max > SM_WARPS) return 0;
	int slots = 1;
	while (slots < rows_max) slots *= 2;
	const int wpr = SM_WARPS / slots;
	const int n4 = ldq / 4;
	const int need = (n4 + 32 * wpr - 1) / (32 * wpr);
	/* columns (float4) per thread: 8 of them live in registers, the rest in shared memory (up to 12 x 512 x 16 B = 96 KB) */
	static const int choices[] = { 1, 2, 4, 8, 12, 16, 20 };
	int cpt = 0;
	for (unsigned i = 0; i < sizeof choices / sizeof choices[0]; i++)
		if (choices[i] >= need) {
			cpt = choices[i];
			break;
		}
	if (!cpt) return 0;
	*wpr_out = wpr;
	*cpt_out = cpt;
	return 1;
}

template <bool TOL> static const void *small_fn(int cpt)
{
	switch (cpt) {
	case 1: return (const void *)gemv_small_kernel<1, TOL>;
	case 2: return (const void *)gemv_small_kernel<2, TOL>;
	case 4: return (const void *)gemv_small_kernel<4, TOL>;
	case 8: return (const void *)gemv_small_kernel<8, TOL>;
	case 12: return (const void *)gemv_small_kernel<12, TOL>;
	case 16: return (const void *)gemv_small_kernel<16, TOL>;
	case 20: return (const void *)gemv_small_kernel<20, TOL>;
	default: return NULL;
	}
}

cudaError_t pqp_launch_gemv_small(const pqp_gemv_args *a, int wpr, int cpt, void *pk0, void *pk1, cudaStream_t s)
{
	const void *fn = a->iters > 0 ? small_fn<false>(cpt) : small_fn<true>(cpt);
	if (!fn) return cudaErrorInvalidValue;
	cudaError_t e = cudaMemsetAsync(a->barrier, 0, sizeof(unsigned), s);
	if (e == cudaSuccess && a->iters <= 0) e = cudaMemsetAsync(a->partials, 0, (size_t)2 * a->grid * 3 * sizeof(uint4), s); /* check packets: epoch 0 = none */
	if (e == cudaSuccess) e = cudaMemsetAsync(pk0, 0xFF, (size_t)a->ldq * sizeof(uint2), s);
	if (e == cudaSuccess) e = cudaMemsetAsync(pk1, 0xFF, (size_t)a->ldq * sizeof(uint2), s);
	if (e != cudaSuccess) return e;
	pqp_gemv_args args = *a;
	uint2 *p0 = reinterpret_cast<uint2 *>(pk0), *p1 = reinterpret_cast<uint2 *>(pk1);
	void *params[] = { (void *)&args, (void *)&wpr, (void *)&p0, (void *)&p1 };
	/* y (two passes) + the columns of q that do not live in registers (16-column instantiation: 8 float4 per thread) */
	const size_t smem = 2 * (size_t)a->ldq * sizeof(float) + (cpt > 8 ? (size_t)(cpt - 8) * SM_THREADS * sizeof(float4) : 0);
	if (smem > 48 * 1024) {
		e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess) return e;
	}
	return cudaLaunchCooperativeKernel(fn, dim3(a->grid), dim3(SM_THREADS), params, smem, s);
}
