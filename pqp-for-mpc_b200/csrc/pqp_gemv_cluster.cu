/*
 * pqp_gemv_cluster.cu -- the PQP loop of ONE mid-size problem (64 < N <= 512: a single condensed-MPC QP, config C4's shape as one
 * problem) on ONE thread-block cluster, y exchanged through distributed shared memory.  sm_100a.
 *
 * The multi-CTA kernels (pqp_gemv_small.cu) pay two L2 round trips per update for the exchange of y: ~1.0 us per update from N = 128
 * to N = 512 whatever the arithmetic (and the one-block kernel slows to 1.0-1.25 us above N = 64).  A problem of this size fits the registers of 16 SMs, and 16 SMs are one cluster (non-portable
 * size; 8 where that is refused): solveQuadraticDual's loop (PQP_CPU.c:718-740) then never leaves the cluster.
 *   - CTA c of the cluster owns rows [cR, cR + R), R = ceil(N / CS); warp w of its 16 takes rows w and w + 16 of them, lane l the
 *     float4 column groups l, l + 32, ... -- the whole matrix sits in registers for the solve, already split into max(q,0) /
 *     max(-q,0) when that fits (<= 32 elements per thread), signed otherwise;
 *   - per update: y from shared memory, the row sums, an xor-shuffle tree, the update with IEEE division (updateY2 / updY,
 *     PQP_CPU.c:603-618, :590-596), the CTA's R new duals into a staging row, ONE block barrier, then warp p sends the staging row to
 *     CTA p of the cluster (itself included) with 16-byte st.shared::cluster.  There is no fence, no mbarrier and no cluster barrier
 *     in the loop: duals are never negative, so the SIGN BIT of every element carries its arrival flag (it alternates with every use
 *     of a buffer); a reader polls its own column groups in its own shared memory until all four flags of a group have turned, and
 *     strips them.  y is double buffered: a peer can only send y_{t+2} after it has received our share of y_{t+1}, which we send
 *     behind the block barrier that follows every warp's last read of y_t;
 *   - y in shared memory is laid out in the order it arrives: chunk c at c * Rpad (Rpad = R rounded up to a multiple of 4, the
 *     16-byte store); the matrix columns are permuted to match when they are loaded, padding holds zeros on both sides (and is
 *     sent, flagged, like everything else);
 *   - run to tolerance (TOL): every check_every updates each CTA sends the five terms of terminate()'s test on y_t for its rows
 *     (PQP_CPU.c:673-687 on g = Qd y + Fd) to every CTA as 8-byte {value, epoch} packets; every warp of every CTA folds
 *     all CTAs' terms in the same xor tree at the start of the next update and takes the same decision; a converged run leaves with
 *     y_t, which is still in the other half of the double buffer -- bit-identical to the fixed-count solve at the reported count.
 * Summation order: lane l adds its column groups ascending in two chains per row (p / n), xor-shuffle 16..1.  Fixed, so results are
 * reproducible bit for bit; FAST order (the STRICT kernel keeps the reference's own).
 */
#include "pqp_internal.h"
#include "pqp_umma.cuh"
#include "pqp_imma.cuh"

#include <stdlib.h>

#define GC_THREADS 512
#define GC_WARPS 16
#define GC_MAXCS 16
#ifndef GC_SMALL_N
#define GC_SMALL_N 128 /* problems up to this size take a cluster of 8: measured, one problem runs as fast (N = 96: 0.50 against 0.52 us per update) and twice as many clusters run at a time (16 problems: 0.98 against 1.5-1.6 ms per 1000 updates); at N = 144-256 a single problem is 10 % slower on 8 */
#endif
#define GC_MAXNP 512 /* padded length of y: 4 column groups of 128 */

namespace {

__device__ __forceinline__ uint32_t gc_cluster_size()
{
	uint32_t r;
	asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
	return r;
}

struct GcSmem {
	uint32_t y[2][GC_MAXNP];       /* bits of y with the arrival flag in the sign bit (y >= 0: the bit is free) */
	uint32_t stage[2][64];         /* the CTA's new duals, flagged, as they are sent */
	float part_tx[8];              /* TOL: the CTA's terms of the stop test */
	uint2 part_rx[2][GC_MAXCS][8]; /* TOL: all CTAs' terms as {value, epoch} packets */
	float part_fin[GC_MAXCS][8];   /* CTA 0: all CTAs' terms of the final status block */
	float red[GC_WARPS][8];
};

__device__ __forceinline__ uint4 gc_lds_volatile(const uint32_t *p)
{
	uint4 v;
	asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(umma::smem_addr(p)) : "memory");
	return v;
}
__device__ __forceinline__ uint32_t gc_lds_volatile1(const uint32_t *p)
{
	uint32_t v;
	asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(umma::smem_addr(p)) : "memory");
	return v;
}
__device__ __forceinline__ uint2 gc_lds_volatile2(const uint2 *p)
{
	uint2 v;
	asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(umma::smem_addr(p)) : "memory");
	return v;
}
/* bounded spin (a protocol bug must not hang the device): 2^28 polls of shared memory within one update are seconds */
__device__ __forceinline__ void gc_spin(unsigned &n)
{
	if (++n > (1u << 28)) __trap();
}

template <int RW, int U, bool TOL> __global__ void __launch_bounds__(GC_THREADS, 1) gemv_cluster_kernel(const pqp_gemv_args a0, int R, int Rpad, int fd_stride,
												     int y_stride)
{
	/* one cluster per problem: cluster b of the grid takes Fd, y, Md and the status block of problem b (strides 0: a single problem) */
	pqp_gemv_args a = a0;
	{
		const int prob = (int)(blockIdx.x / gc_cluster_size());
		a.Fd += (size_t)prob * fd_stride;
		a.ybuf0 += (size_t)prob * y_stride;
		a.ybuf1 += (size_t)prob * y_stride;
		a.status += prob;
		if (a.Md && fd_stride) a.Md += prob;
	}
	constexpr bool SPLIT = RW * U * 4 <= 32;
	__shared__ __align__(128) GcSmem sm;
	const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
	const uint32_t rank = cluster_ctarank(), CS = gc_cluster_size();
	const int N = a.N;
	const int NP = (int)CS * Rpad; /* positions of y that are ever sent; the rest of the buffers stays zero */

	/* y_0 straight from global memory, in arrival order, flag 0; everything else (padding, buffer 1, packets) zero */
	for (int i = tid; i < 2 * GC_MAXNP; i += GC_THREADS) {
		const int pos = i % GC_MAXNP, c = pos / Rpad, r = pos - c * Rpad, col = c * R + r;
		(&sm.y[0][0])[i] = (i < GC_MAXNP && c < (int)CS && r < R && col < N) ? (__float_as_uint(a.ybuf0[col]) & 0x7fffffffu) : 0u;
	}
	for (int i = tid; i < 2 * 64; i += GC_THREADS) (&sm.stage[0][0])[i] = 0u;
	for (int i = tid; i < 2 * GC_MAXCS * 8; i += GC_THREADS) (&sm.part_rx[0][0][0])[i] = make_uint2(0u, 0u);

	/* this thread's share of the matrix: rows warp + 16 rw of the CTA's R, column groups lane + 32 u */
	float qa[RW][U][4], qb[SPLIT ? RW : 1][SPLIT ? U : 1][4];
	float th[RW], fdp[RW], fdn[RW], fd[RW], tolr[RW];
	int rowl[RW];
	bool rok[RW];
#pragma unroll
	for (int rw = 0; rw < RW; rw++) {
		rowl[rw] = warp + GC_WARPS * rw;
		const int grow = (int)rank * R + rowl[rw];
		rok[rw] = rowl[rw] < R && grow < N;
#pragma unroll
		for (int u = 0; u < U; u++)
#pragma unroll
			for (int e = 0; e < 4; e++) {
				const int pos = 4 * (lane + 32 * u) + e, c = pos / Rpad, r = pos - c * Rpad, col = c * R + r;
				const float q = (rok[rw] && c < (int)CS && r < R && col < N) ? __ldg(a.Q + (size_t)grow * a.ldq + col) : 0.0f;
				if (SPLIT) {
					qa[rw][u][e] = fmaxf(q, 0.0f);
					qb[rw][u][e] = fmaxf(-q, 0.0f);
				} else {
					qa[rw][u][e] = q;
				}
			}
		th[rw] = rok[rw] ? a.theta[grow] : 1.0f;
		fd[rw] = rok[rw] ? a.Fd[grow] : 0.0f;
		fdp[rw] = fmaxf(fd[rw], 0.0f);
		fdn[rw] = fmaxf(-fd[rw], 0.0f);
		tolr[rw] = (rok[rw] && a.Kp) ? fmaxf(a.erc * a.Kp[grow], a.eac) : a.eac;
	}
	__syncthreads();
	cluster_sync_all(); /* every CTA's buffers are initialised before anyone sends */

	const int updates = TOL ? a.max_iters : a.iters;
	int next_chk = 0;     /* TOL: check passes at t = 0, check_every, ... */
	bool pending = false; /* TOL: the previous pass sent stop-test terms; decide before this pass's sums */
	int done = updates, converged = 0;
	float s_min = 0.0f, s_gap = 0.0f, s_jd = 0.0f, s_kkt = 0.0f;
	unsigned spins = 0;
	/* the row this LANE finishes (two rows per warp: lanes 16-31 the second) and its constants; its dual stays in a register: this warp
	 * computed it, it need not come back through the exchange */
	const int sel = (RW == 2 && lane >= 16) ? RW - 1 : 0;
	const float th_l = th[sel], fdn_l = fdn[sel], fdp_l = fdp[sel], fd_l = fd[sel], tol_l = tolr[sel];
	const bool rok_l = rok[sel];
	const int rowl_l = rowl[sel];
	float ym_l = rok_l ? __uint_as_float(sm.y[0][(int)rank * Rpad + rowl_l]) : 0.0f;

	for (int t = 0;; t++) {
		const int cur = t & 1, nxt = cur ^ 1;
		const bool last = (t == updates);
		spins = 0; /* the bound is on the polls of ONE update, however long the solve */
		const bool chk = TOL && !last && t == next_chk;
		/* arrival flag of y_t in buffer cur: its k-th use (k = (t-1)/2) carries flag (k+1)&1 -- the first use 1, against the zeros the
		 * buffer starts with; y_0 is there from the start */
		const uint32_t want = t > 0 ? ((uint32_t)(((t - 1) >> 1) + 1) & 1u) << 31 : 0u;
		const uint32_t want_nxt = ((uint32_t)((t >> 1) + 1) & 1u) << 31; /* flag of y_{t+1} in buffer nxt */

		if (TOL && pending) {
			/* the stop test of pass t-1: lane c of every warp waits for CTA c's five terms; a fixed xor tree folds them, so every warp of
			 * every CTA takes the same decision */
			float v_min = INFINITY, v_gap = 0.0f, v_jd = 0.0f, v_kkt = 0.0f, v_viol = -INFINITY;
			if (lane < (int)CS) {
				uint2 pk[5];
#pragma unroll
				for (int k = 0; k < 5; k++) {
					pk[k] = gc_lds_volatile2(&sm.part_rx[cur][lane][k]);
					while (pk[k].y != (uint32_t)t) {
						gc_spin(spins);
						pk[k] = gc_lds_volatile2(&sm.part_rx[cur][lane][k]);
					}
				}
				v_min = __uint_as_float(pk[0].x); v_gap = __uint_as_float(pk[1].x); v_jd = __uint_as_float(pk[2].x);
				v_kkt = __uint_as_float(pk[3].x); v_viol = __uint_as_float(pk[4].x);
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
				/* converged at y_{t-1}: it is still in the other half of the double buffer (nobody sends y_{t+1} into it: all stop here) */
				done = t - 1; converged = 1; s_min = v_min; s_gap = v_gap; s_jd = Jd; s_kkt = v_kkt;
				for (int r = tid; r < R; r += GC_THREADS) {
					const int grow = (int)rank * R + r;
					if (grow < N) a.ybuf1[grow] = __uint_as_float(sm.y[nxt][(int)rank * Rpad + r] & 0x7fffffffu);
				}
				break;
			}
			pending = false;
		}

		/* y_t: poll this lane's column groups until every element carries the flag of this update, then strip it */
		float4 yv[U];
#pragma unroll
		for (int u = 0; u < U; u++) {
			const uint32_t *src = &sm.y[cur][4 * (lane + 32 * u)];
			uint4 v = gc_lds_volatile(src);
			if (t > 0 && 4 * (lane + 32 * u) < NP) {
				while (((v.x ^ want) | (v.y ^ want) | (v.z ^ want) | (v.w ^ want)) >> 31) {
					gc_spin(spins);
					v = gc_lds_volatile(src);
				}
			}
			yv[u] = make_float4(__uint_as_float(v.x & 0x7fffffffu), __uint_as_float(v.y & 0x7fffffffu), __uint_as_float(v.z & 0x7fffffffu),
					    __uint_as_float(v.w & 0x7fffffffu));
		}
		float num[RW], den[RW];
#pragma unroll
		for (int rw = 0; rw < RW; rw++) {
			float n = 0.0f, d = 0.0f;
#pragma unroll
			for (int u = 0; u < U; u++) {
				const float ye[4] = { yv[u].x, yv[u].y, yv[u].z, yv[u].w };
#pragma unroll
				for (int e = 0; e < 4; e++) {
					if (SPLIT) {
						d = fmaf(qa[rw][u][e], ye[e], d);
						n = fmaf(qb[rw][u][e], ye[e], n);
					} else {
						d = fmaf(fmaxf(qa[rw][u][e], 0.0f), ye[e], d);
						n = fmaf(fmaxf(-qa[rw][u][e], 0.0f), ye[e], n);
					}
				}
			}
			num[rw] = n;
			den[rw] = d;
		}
		/* row totals over the 32 lanes.  Two rows per warp: the first round also sorts them onto the two half-warps (lanes 0-15 end up
		 * with row 0, lanes 16-31 with row 1), so the tree, the update and the division are issued once for both rows */
		float n_t, d_t;
		if (RW == 2) {
			const bool hi = lane >= 16;
			n_t = (hi ? num[RW - 1] : num[0]) + __shfl_xor_sync(0xffffffffu, hi ? num[0] : num[RW - 1], 16);
			d_t = (hi ? den[RW - 1] : den[0]) + __shfl_xor_sync(0xffffffffu, hi ? den[0] : den[RW - 1], 16);
#pragma unroll
			for (int o = 8; o; o >>= 1) {
				n_t += __shfl_xor_sync(0xffffffffu, n_t, o);
				d_t += __shfl_xor_sync(0xffffffffu, d_t, o);
			}
		} else {
			n_t = num[0];
			d_t = den[0];
#pragma unroll
			for (int o = 16; o; o >>= 1) {
				n_t += __shfl_xor_sync(0xffffffffu, n_t, o);
				d_t += __shfl_xor_sync(0xffffffffu, d_t, o);
			}
		}
		float e_min = INFINITY, e_gap = 0.0f, e_jd = 0.0f, e_kkt = 0.0f, e_viol = -INFINITY;
		{
			const float y_mine = ym_l;
			const float nn = fmaf(th_l, y_mine, n_t) + fdn_l;
			const float dd = fmaf(th_l, y_mine, d_t) + fdp_l;
			if ((last || chk) && rok_l) {
				const float gq = dd - nn;
				e_min = gq;
				e_gap = y_mine * gq;
				e_jd = y_mine * (0.5f * (gq + fd_l));
				e_kkt = fabsf(fminf(y_mine, gq));
				e_viol = -gq - tol_l;
			}
			if (RW == 2 && (last || chk)) {
				/* lane 0 folds its row, then the row of lane 16 */
				const float m2 = __shfl_sync(0xffffffffu, e_min, 16), g2 = __shfl_sync(0xffffffffu, e_gap, 16), j2 = __shfl_sync(0xffffffffu, e_jd, 16),
					    k2 = __shfl_sync(0xffffffffu, e_kkt, 16), v2 = __shfl_sync(0xffffffffu, e_viol, 16);
				e_min = fminf(e_min, m2); e_gap += g2; e_jd += j2; e_kkt = fmaxf(e_kkt, k2); e_viol = fmaxf(e_viol, v2);
			}
			const bool writer = RW == 2 ? (lane & 15) == 0 : lane == 0;
			if (last) {
				if (writer && rok_l) a.ybuf1[(int)rank * R + rowl_l] = y_mine; /* the answer, as a plain vector */
			} else {
				/* (a NaN dual keeps its payload but not its sign: the bit carries the arrival flag) */
				const float yn = rok_l ? __fdiv_rn(nn, dd) * y_mine : 0.0f;
				ym_l = __uint_as_float(__float_as_uint(yn) & 0x7fffffffu);
				if (writer && rowl_l < Rpad) sm.stage[nxt][rowl_l] = (__float_as_uint(yn) & 0x7fffffffu) | want_nxt; /* padding rows: zeros, flagged */
			}
		}
		if (last || chk) {
			if (lane == 0) {
				sm.red[warp][0] = e_min; sm.red[warp][1] = e_gap; sm.red[warp][2] = e_jd; sm.red[warp][3] = e_kkt; sm.red[warp][4] = e_viol;
			}
			__syncthreads();
			if (tid == 0) {
				for (int w = 1; w < GC_WARPS; w++) {
					e_min = fminf(e_min, sm.red[w][0]); e_gap += sm.red[w][1]; e_jd += sm.red[w][2];
					e_kkt = fmaxf(e_kkt, sm.red[w][3]); e_viol = fmaxf(e_viol, sm.red[w][4]);
				}
				sm.part_tx[0] = e_min; sm.part_tx[1] = e_gap; sm.part_tx[2] = e_jd; sm.part_tx[3] = e_kkt; sm.part_tx[4] = e_viol;
			}
		}
		if (last) {
			/* status block of a fixed-count solve (or of a capped run): every CTA's terms to CTA 0 through its shared memory */
			__syncthreads();
			if (tid < 5) {
				const uint32_t dst = map_peer(umma::smem_addr(&sm.part_fin[rank][tid]), 0u);
				asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(dst), "f"(sm.part_tx[tid]) : "memory");
			}
			cluster_sync_all();
			if (rank == 0 && warp == 0) {
				/* the same xor tree as the stop test's fold: a run that stops at the cap reports the bits a check would have seen */
				float v_min = INFINITY, v_gap = 0.0f, v_jd = 0.0f, v_kkt = 0.0f;
				if (lane < (int)CS) {
					const float *pr = sm.part_fin[lane];
					v_min = pr[0]; v_gap = pr[1]; v_jd = pr[2]; v_kkt = pr[3];
				}
#pragma unroll
				for (int o = 16; o; o >>= 1) {
					v_min = fminf(v_min, __shfl_xor_sync(0xffffffffu, v_min, o));
					v_gap += __shfl_xor_sync(0xffffffffu, v_gap, o);
					v_jd += __shfl_xor_sync(0xffffffffu, v_jd, o);
					v_kkt = fmaxf(v_kkt, __shfl_xor_sync(0xffffffffu, v_kkt, o));
				}
				s_min = v_min; s_gap = v_gap; s_jd = v_jd + (a.Md ? 0.5f * a.Md[0] : 0.0f); s_kkt = v_kkt;
			}
			done = updates;
			break;
		}
		/* ONE block barrier per update, then warp p sends the staging row to CTA p of the cluster -- itself included -- with 16-byte
		 * remote stores.  No fence and no mbarrier: every element carries its own arrival flag in the sign bit (y >= 0; a 16-byte store
		 * need not land as one piece), the terms of the stop test travel as 8-byte {value, epoch} packets.  The barrier also keeps every
		 * warp of the CTA -- including warps that own no row and that nobody would wait for -- within one update of the others: a warp
		 * that fell behind could be overtaken by y_{t+2} landing in the buffer it still polls for y_t.
		 * Measured alternatives (N = 144..256, cycles per update): bulk copies with complete_tx on the receiver's mbarrier ~1700 (the
		 * copy engine's start-up); 16-byte stores + a release-arrive at cluster scope ~2100 (the cluster-scope release / acquire);
		 * 4-byte stores straight from the producing warps, no barrier, ~3800 (the SM-to-SM network wants few, large packets);
		 * this form ~1400.  The flag-in-data exchange through L2 of pqp_gemv_small.cu: ~1900. */
		__syncthreads();
		if (warp < (int)CS) {
			const uint32_t peer = (uint32_t)warp;
			if (lane < Rpad / 4) {
				const uint4 v = reinterpret_cast<const uint4 *>(sm.stage[nxt])[lane];
				st_cluster_v4(map_peer(umma::smem_addr(&sm.y[nxt][(int)rank * Rpad + 4 * lane]), peer), v);
			} else if (chk && lane >= 16 && lane < 21) {
				st_cluster_v2(map_peer(umma::smem_addr(&sm.part_rx[nxt][rank][lane - 16]), peer), __float_as_uint(sm.part_tx[lane - 16]), (uint32_t)(t + 1));
			}
		}
		if (TOL && chk) {
			pending = true;
			next_chk += a.check_every;
		}
	}
	if (converged) {
		/* every CTA took the same decision; CTA 0 reports it */
		if (rank == 0 && tid == 0) {
			pqp_status o;
			o.iters = done; o.converged = 1; o.min_slack = s_min; o.gap = s_gap; o.Jd = s_jd; o.kkt = s_kkt;
			*a.status = o;
			*a.result_buf = 1;
		}
		cluster_sync_all(); /* nobody exits while a peer's last stores may still be in flight towards it */
		return;
	}
	if (rank == 0 && tid == 0) {
		pqp_status o;
		o.iters = done; o.converged = 0; o.min_slack = s_min; o.gap = s_gap; o.Jd = s_jd; o.kkt = s_kkt;
		*a.status = o;
		*a.result_buf = 1;
	}
}

int gc_max_cluster = -1; /* 16, 8 or 0 (no cluster launch possible), decided once per process */

template <int RW, bool TOL> const void *gc_fn_u(int U)
{
	switch (U) {
	case 1: return (const void *)gemv_cluster_kernel<RW, 1, TOL>;
	case 2: return (const void *)gemv_cluster_kernel<RW, 2, TOL>;
	case 3: return (const void *)gemv_cluster_kernel<RW, 3, TOL>;
	default: return (const void *)gemv_cluster_kernel<RW, 4, TOL>;
	}
}
template <bool TOL> const void *gc_fn(int RW, int U) { return RW == 1 ? gc_fn_u<1, TOL>(U) : gc_fn_u<2, TOL>(U); }

void gc_geometry(int N, int CS, int *R, int *Rpad, int *RW, int *U)
{
	*R = (N + CS - 1) / CS;
	*Rpad = (*R + 3) / 4 * 4;
	*RW = (*R + GC_WARPS - 1) / GC_WARPS;
	*U = (CS * *Rpad + 127) / 128;
}

/* the largest cluster the device schedules for this kernel: 16 (non-portable), else 8, else none */
int gc_probe(void)
{
	if (gc_max_cluster >= 0) return gc_max_cluster;
	gc_max_cluster = 0;
	const void *fn = (const void *)gemv_cluster_kernel<1, 1, false>;
	for (int cs = 16; cs >= 8; cs -= 8) {
		if (cs > 8 && cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
			cudaGetLastError();
			continue;
		}
		cudaLaunchConfig_t cfg;
		memset(&cfg, 0, sizeof cfg);
		cfg.gridDim = dim3(cs);
		cfg.blockDim = dim3(GC_THREADS);
		cudaLaunchAttribute attr[1];
		attr[0].id = cudaLaunchAttributeClusterDimension;
		attr[0].val.clusterDim.x = cs;
		attr[0].val.clusterDim.y = 1;
		attr[0].val.clusterDim.z = 1;
		cfg.attrs = attr;
		cfg.numAttrs = 1;
		int nclusters = 0;
		if (cudaOccupancyMaxActiveClusters(&nclusters, fn, &cfg) == cudaSuccess && nclusters >= 1) {
			gc_max_cluster = cs;
			break;
		}
		cudaGetLastError();
	}
	return gc_max_cluster;
}

/* cluster size for a problem of N duals: the largest the device schedules, but 8 where 8 CTAs hold the matrix as comfortably (N <= 256:
 * at most two rows per warp either way) -- twice as many clusters then run at a time, which is what a small batch wants */
int gc_size_for(int N)
{
	const int cs = gc_probe();
	const char *e = pqp_env("PQP_CLUSTER_SIZE");
	if (e && (atoi(e) == 8 || atoi(e) == 16) && atoi(e) <= cs) return atoi(e);
	if (cs == 16 && N <= GC_SMALL_N) return 8;
	return cs;
}

} /* namespace */

/* 64 < N: up to there one thread block does it without any exchange in 0.18-0.30 us per update (pqp_gemv_cta.cu; measured 0.97 at
 * N = 96 and 1.25 at N = 128, against 0.52 here).  Up to 32 rows per CTA (two per warp) and 512
 * padded columns, i.e. N <= 512 on a cluster of 16: beyond that the 16 SMs' arithmetic costs more than the exchange saves (measured:
 * N = 144..256 0.55-0.58 us per update against 0.95-1.00 for the multi-CTA kernel, N = 480 0.83 against 1.05, N = 640 slower). */
int pqp_gemv_cluster_supported(int N)
{
	const char *e = pqp_env("PQP_GEMV_CLUSTER");
	if (e && atoi(e) == 0) return 0;
	const char *m = pqp_env("PQP_GEMV_CLUSTER_MIN"); /* experiments: the largest N the cluster leaves alone */
	if (N <= (m ? atoi(m) : 64)) return 0;
	if (gc_probe() == 0) return 0;
	const int cs = gc_size_for(N);
	int R, Rpad, RW, U;
	gc_geometry(N, cs, &R, &Rpad, &RW, &U);
	return RW <= 2 && U <= 4 && cs * Rpad <= GC_MAXNP;
}

/* B problems sharing Q, one cluster each (B = 1: the single-problem call): Fd, y_0 (ybuf0), the result (ybuf1, may be ybuf0), Md and the
 * status blocks of problem b at b * stride.  Clusters are independent; the device runs as many at a time as it has GPCs with 16 SMs. */
cudaError_t pqp_launch_gemv_cluster_batch(const pqp_gemv_args *a, int B, int fd_stride, int y_stride, cudaStream_t s)
{
	if (gc_probe() == 0 || B < 1) return cudaErrorNotSupported;
	const int cs = gc_size_for(a->N);
	int R, Rpad, RW, U;
	gc_geometry(a->N, cs, &R, &Rpad, &RW, &U);
	const void *fn = a->iters > 0 ? gc_fn<false>(RW, U) : gc_fn<true>(RW, U);
	cudaError_t e = cudaSuccess;
	if (cs > 8) e = cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
	if (e != cudaSuccess) return e;
	cudaLaunchConfig_t cfg;
	memset(&cfg, 0, sizeof cfg);
	cfg.gridDim = dim3(cs * B);
	cfg.blockDim = dim3(GC_THREADS);
	cfg.stream = s;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = cs;
	attr[0].val.clusterDim.y = 1;
	attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr;
	cfg.numAttrs = 1;
	pqp_gemv_args args = *a;
	void *params[] = { (void *)&args, (void *)&R, (void *)&Rpad, (void *)&fd_stride, (void *)&y_stride };
	return cudaLaunchKernelExC(&cfg, fn, params);
}

/* result left in ybuf1, status written by the kernel */
cudaError_t pqp_launch_gemv_cluster(const pqp_gemv_args *a, cudaStream_t s) { return pqp_launch_gemv_cluster_batch(a, 1, 0, 0, s); }
