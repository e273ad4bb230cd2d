/*
 * pqp_gemv_tma.cu -- single-problem PQP loop, TMA-staged (sm_100a).  The production path for
 * fixed-count solves; pqp_gemv.cu's LDG kernel remains the general fallback (tolerance mode,
 * rows too long for the ring).
 *
 * Same loop as pqp_gemv.cu (PQP_CPU.c:718-740 in one cooperative launch, one CTA per SM owning a
 * slab of rows of the signed Qd), restructured around the TMA engine so that the HBM stream never
 * stops -- not for the grid barrier, not for the y exchange, not for the update:
 *
 *   producer   warp 16, one elected thread: for every pass and every non-resident row of the slab,
 *              waits for a free ring stage (mbarrier "empty"), arms the stage's "full" mbarrier with
 *              the row's byte count and issues ONE cp.async.bulk (UBLKCP) global->shared copy of the
 *              whole row (N*4 bytes) with an L2 cache-policy hint.  Q does not depend on y, so the
 *              producer runs ahead across iteration boundaries: while the consumers sit in the grid
 *              barrier the ring (S rows per SM, ~19 MB chip-wide) keeps HBM busy.
 *   consumers  warps 0-15: each thread keeps ITS columns of y in registers (float4 x YC, reloaded
 *              from L2 once per iteration), waits on "full", reads the row from shared memory with
 *              conflict-free 128-bit loads, forms num += max(-q,0)*y, den += max(q,0)*y, releases
 *              the stage ("empty"), warp-shuffles, and parks per-warp partials in shared memory.
 *   finish     threads t < rows: fixed-order sum over the 16 warps, + theta_i*y_i + F-/F+, the
 *              multiplicative update with IEEE division, y+ to the ping-pong vector, stop-test terms.
 *   residency  the first R rows of every slab stay in shared memory for the whole launch (never
 *              re-fetched); when the whole slab fits (N <= ~2.6k) the loop never touches HBM/L2 for Q.
 *   L2 pinning the next P rows of every slab are fetched with an evict_last policy and the rest with
 *              evict_first, so a fixed ~P/rows fraction of Q is served from the 126 MB L2 on every
 *              iteration instead of HBM (Q at N=8192 is 268 MB: it cannot all stay).
 *
 * Summation order is IDENTICAL to pqp_gemv.cu (thread tid owns float4 columns tid+512u, u ascending;
 * xor-shuffle 16..1; warps ascending), so both kernels return bit-identical y.
 */
#include "pqp_internal.h"

#include <stdlib.h>

#define CONSUMERS PQP_GEMV_THREADS /* 512 */
#define TMA_THREADS (CONSUMERS + 32)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
/* bounded (10 s of %globaltimer): a protocol bug traps instead of hanging the device; the timer is only read on the failing path */
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	asm volatile(
		"{\n\t"
		".reg .pred p;\n\t"
		".reg .u64 t0, t1;\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@p bra WAIT_DONE;\n\t"
		"mov.u64 t0, %%globaltimer;\n\t"
		"WAIT_LOOP:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@p bra WAIT_DONE;\n\t"
		"mov.u64 t1, %%globaltimer;\n\t"
		"sub.u64 t1, t1, t0;\n\t"
		"setp.lt.u64 p, t1, 10000000000;\n\t"
		"@p bra WAIT_LOOP;\n\t"
		"trap;\n\t"
		"WAIT_DONE:\n\t"
		"}" ::"r"(smem_u32(bar)),
		"r"(parity)
		: "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
			     smem_u32(dst)),
		     "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
		     : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_last()
{
	uint64_t p;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
	return p;
}
__device__ __forceinline__ uint64_t policy_evict_first()
{
	uint64_t p;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
	return p;
}
__device__ __forceinline__ uint64_t make_policy(int kind)
{
	uint64_t p;
	switch (kind) {
	case 1: asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); break;
	case 2: asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); break;
	case 3: asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;" : "=l"(p)); break;
	default: asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p)); break;
	}
	return p;
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory"); }

__device__ __forceinline__ void grid_barrier_consumers(unsigned *counter, unsigned &target, unsigned nblocks)
{
	consumer_sync();
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
	consumer_sync();
}

/*
 * Flag-in-data exchange of y ("LL" protocol): every dual travels as an 8-byte packet {value, epoch} written with
 * one 64-bit store and read with polling loads; a packet is valid for pass p when its epoch equals p.  This fuses
 * the grid barrier and the y reload into a single L2 round trip per iteration.  Ping-pong safety: a CTA writes
 * epoch p+2 into buffer p&1 only after it has read ALL of epoch p+1, which exists only once every CTA has finished
 * reading epoch p (each CTA writes its rows after a CTA-wide sync that follows its reads).
 */
__device__ __forceinline__ void st_packet(uint2 *dst, float v, uint32_t epoch)
{
	asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(__float_as_uint(v)), "r"(epoch) : "memory");
}
__device__ __forceinline__ float ld_packet(const uint2 *src, uint32_t epoch)
{
	uint32_t v, e;
	for (;;) {
		asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(e) : "l"(src) : "memory");
		if (e == epoch) break;
		__nanosleep(20);
	}
	return __uint_as_float(v);
}
/* four consecutive packets -> float4 */
__device__ __forceinline__ float4 ld_packet4(const uint2 *src, uint32_t epoch)
{
	uint32_t a0, e0, a1, e1, a2, e2, a3, e3;
	for (;;) {
		asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(e0), "=r"(a1), "=r"(e1) : "l"(src) : "memory");
		asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a2), "=r"(e2), "=r"(a3), "=r"(e3) : "l"(src + 2) : "memory");
		if (e0 == epoch && e1 == epoch && e2 == epoch && e3 == epoch) break;
		__nanosleep(20);
	}
	return make_float4(__uint_as_float(a0), __uint_as_float(a1), __uint_as_float(a2), __uint_as_float(a3));
}

__device__ __forceinline__ void acc4t(float &num, float &den, const float4 q, const float4 y)
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

/*
 * Tensor memory as a second on-chip store for rows (long-row instantiation, YC = 4: N <= 8192).  The loop is HBM-bound and has no
 * tensor-core work, so the SM's 256 KB of TMEM would sit idle: it holds eight more rows per SM for the whole launch (37 MB over 148
 * SMs that HBM/L2 do not deliver again on every update).  Warp w reaches lanes 32*(w%4)..+31; a row takes 64 columns: consumer warps
 * w, w+4, w+8, w+12 share a lane quarter and own columns 16*(w/4)..+15 of the row's slot -- a thread's 16 columns are its four
 * float4 of the row.  The parked rows are spread evenly through the slab so the TMA ring refills while they are worked on.
 */
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float4 &q0, float4 &q1, float4 &q2, float4 &q3)
{
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
		     : "=f"(q0.x), "=f"(q0.y), "=f"(q0.z), "=f"(q0.w), "=f"(q1.x), "=f"(q1.y), "=f"(q1.z), "=f"(q1.w), "=f"(q2.x), "=f"(q2.y), "=f"(q2.z),
		       "=f"(q2.w), "=f"(q3.x), "=f"(q3.y), "=f"(q3.z), "=f"(q3.w)
		     : "r"(taddr));
	/* the wait names the registers it makes valid, so that no use of them can be scheduled ahead of it */
	asm volatile("tcgen05.wait::ld.sync.aligned;"
		     : "+f"(q0.x), "+f"(q0.y), "+f"(q0.z), "+f"(q0.w), "+f"(q1.x), "+f"(q1.y), "+f"(q1.z), "+f"(q1.w), "+f"(q2.x), "+f"(q2.y), "+f"(q2.z),
		       "+f"(q2.w), "+f"(q3.x), "+f"(q3.y), "+f"(q3.z), "+f"(q3.w));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float4 q0, const float4 q1, const float4 q2, const float4 q3)
{
	asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
		     "f"(q0.x), "f"(q0.y), "f"(q0.z), "f"(q0.w), "f"(q1.x), "f"(q1.y), "f"(q1.z), "f"(q1.w), "f"(q2.x), "f"(q2.y), "f"(q2.z), "f"(q2.w),
		     "f"(q3.x), "f"(q3.y), "f"(q3.z), "f"(q3.w)
		     : "memory");
}
#define TMA_TM_ROWS_MAX 128 /* rows per slab the tensor-memory table covers */

struct TmaGeom {
	int stages;     /* ring depth S */
	int resident;   /* R rows per slab kept in shared memory */
	int pinned;     /* P streamed rows per slab fetched evict_last */
	int rows_max;
	int tmem_rows;  /* rows per slab parked in tensor memory (TMU instantiation), 0..8 */
	int pol_keep, pol_stream; /* 0 normal, 1 evict_first, 2 evict_last, 3 evict_unchanged */
	uint2 *pk0, *pk1;         /* packet vectors [ldq] (epochs pre-set to 0xFFFFFFFF), NULL = counter barrier + plain y */
};

/*
 * Shared memory: ring [S][ldq] | resident [R][ldq] | part [2][16][rows_max] | red [16*8] | full[S], empty[S]
 */
template <int YC, bool TMU>
__global__ void __launch_bounds__(TMA_THREADS, 1) gemv_tma_kernel(const pqp_gemv_args a, const TmaGeom g)
{
	static_assert(!TMU || YC == 4, "tensor-memory rows: 16 columns per thread");
	__shared__ uint32_t tmem_base_s;
	__shared__ short tm_col[TMU ? TMA_TM_ROWS_MAX : 1]; /* column of row t of the slab in tensor memory, or -1 */
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
	const int N = a.N, ldq = a.ldq, n4 = ldq / 4;
	const unsigned G = gridDim.x;
	const int r0 = (int)((long long)N * blockIdx.x / G), r1 = (int)((long long)N * (blockIdx.x + 1) / G);
	const int nrows = r1 - r0;
	const int S = g.stages, R = min(g.resident, nrows), rows_max = g.rows_max;
	const uint32_t row_bytes = (uint32_t)ldq * 4u;

	float *ring = reinterpret_cast<float *>(smem_raw);
	float *resid = ring + (size_t)S * ldq;
	float *part = resid + (size_t)g.resident * ldq;
	float *red = part + 2 * PQP_GEMV_WARPS * rows_max;
	uint64_t *full = reinterpret_cast<uint64_t *>(red + PQP_GEMV_WARPS * 8);
	uint64_t *empty = full + S;

	const int passes = a.iters + 1; /* iters updates + one evaluation pass */

	if (tid == 0) {
		for (int s = 0; s < S; s++) {
			mbar_init(&full[s], 1);
			mbar_init(&empty[s], PQP_GEMV_WARPS);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (TMU) {
		const int L = nrows - R, TM = min(g.tmem_rows, L);
		for (int t = tid; t < nrows; t += TMA_THREADS) {
			short c = -1;
			if (t >= R && TM > 0) {
				const int x = t - R, before = (int)(((long long)x * TM) / L);
				if ((int)(((long long)(x + 1) * TM) / L) != before) c = (short)(64 * before);
			}
			tm_col[t] = c;
		}
		if (warp == 0) {
			asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
			asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
		}
		asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	}
	__syncthreads();
	if (TMU) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

	if (warp == PQP_GEMV_WARPS) {
		/* ================= producer ================= */
		if (lane == 0) {
			const uint64_t pol_keep = make_policy(g.pol_keep), pol_stream = make_policy(g.pol_stream);
			/* resident rows: one-time bulk copies, all on full[0]'s first phase */
			/* ring position and phase are stepped, not derived from a row counter: a 64-bit idx % S, idx / S per row and warp was a
			 * 35-instruction dependent chain in front of every barrier wait (ncu source view: 1/5 of the row loop's instructions) */
			int ps = 0;
			uint32_t pph = 0;
			if (R > 0) {
				mbar_arrive_expect_tx(&full[0], row_bytes * (uint32_t)R);
				for (int t = 0; t < R; t++)
					bulk_g2s(resid + (size_t)t * ldq, a.Q + (size_t)(r0 + t) * ldq, row_bytes, &full[0], pol_stream);
				if (++ps == S) { ps = 0; pph ^= 1u; } /* stage 0 / phase 0 is consumed by the residency handshake */
			}
			const int T = max(nrows - R - (TMU ? min(g.tmem_rows, nrows - R) : 0), 1), P = min(g.pinned, T); /* streamed rows per pass */
			for (int p = 0; p < passes; p++) {
				for (int t = R, k = 0; t < nrows; t++) {
					if (TMU && tm_col[t] >= 0) continue; /* lives in tensor memory */
					const int s = ps;
					const uint32_t ph = pph;
					if (++ps == S) { ps = 0; pph ^= 1u; }
					mbar_wait(&empty[s], ph ^ 1u);
					mbar_arrive_expect_tx(&full[s], row_bytes);
					bulk_g2s(ring + (size_t)s * ldq, a.Q + (size_t)(r0 + t) * ldq, row_bytes, &full[s],
						 /* pinned rows are spread evenly through the slab so L2 hits and HBM misses overlap in time */
						 (((long long)(k + 1) * P) / T != ((long long)k * P) / T) ? pol_keep : pol_stream); /* k: among the streamed rows */
					k++;
				}
			}
		}
		return;
	}

	/* ================= consumers ================= */
	int cs = 0; /* ring stage and phase of the next streamed row */
	uint32_t cph = 0;
	if (R > 0) {
		mbar_wait(&full[0], 0);
		__syncwarp();
		if (lane == 0) mbar_arrive(&empty[0]);
		if (++cs == S) { cs = 0; cph ^= 1u; }
	}
	/* per-row constants of the rows this thread finishes */
	float th_r = 0.0f, fd_r = 0.0f, kp_tol = a.eac;
	if (tid < nrows) {
		th_r = a.theta[r0 + tid];
		fd_r = a.Fd[r0 + tid];
		if (a.Kp) kp_tol = fmaxf(a.erc * a.Kp[r0 + tid], a.eac);
	}
	unsigned bar_target = 0;
	uint32_t tm_mine = 0;
	if (TMU) {
		/* park the rows: each consumer thread loads its four float4 of the row and stores them into its lane */
		tm_mine = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 16);
		for (int t = R; t < nrows; t++) {
			const int col = tm_col[t];
			if (col < 0) continue;
			const float4 *src = reinterpret_cast<const float4 *>(a.Q + (size_t)(r0 + t) * ldq);
			float4 q[4];
#pragma unroll
			for (int u = 0; u < 4; u++) {
				const int c = tid + u * CONSUMERS;
				q[u] = (c < n4) ? __ldcs(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
			}
			tmem_st16(tm_mine + (uint32_t)col, q[0], q[1], q[2], q[3]);
		}
		asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
	}

	for (int p = 0; p < passes; p++) {
		const float *y_in = (p & 1) ? a.ybuf1 : a.ybuf0;
		float *y_out = (p & 1) ? a.ybuf0 : a.ybuf1;
		const bool is_last = (p == passes - 1);
		const bool ll = g.pk0 != nullptr;
		const uint2 *pk_in = (p & 1) ? g.pk1 : g.pk0;
		uint2 *pk_out = (p & 1) ? g.pk0 : g.pk1;

		float4 yv[YC];
		float y_mine = 0.0f;
		if (ll && p > 0) {
			/* columns beyond N carry no packets: the padding of Q is zero, so any finite value works */
#pragma unroll
			for (int u = 0; u < YC; u++) {
				const int c = tid + u * CONSUMERS;
				yv[u] = (c * 4 + 3 < N) ? ld_packet4(pk_in + 4 * c, (uint32_t)p) : make_float4(0.f, 0.f, 0.f, 0.f);
				if (c * 4 + 3 >= N && c * 4 < N) {
					float t[4] = { 0.f, 0.f, 0.f, 0.f };
					for (int e = 0; e < 4; e++)
						if (c * 4 + e < N) t[e] = ld_packet(pk_in + 4 * c + e, (uint32_t)p);
					yv[u] = make_float4(t[0], t[1], t[2], t[3]);
				}
			}
			if (tid < nrows) y_mine = ld_packet(pk_in + r0 + tid, (uint32_t)p);
		} else {
#pragma unroll
			for (int u = 0; u < YC; u++) {
				const int c = tid + u * CONSUMERS;
				yv[u] = (c < n4) ? __ldcg(reinterpret_cast<const float4 *>(y_in) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
			}
			if (tid < nrows) y_mine = __ldcg(y_in + r0 + tid);
		}

		if constexpr (YC <= 2) { /* short rows (N <= 4096): the loop is issue-bound, not bandwidth-bound (ncu: L2 hit 99 %, issue slots 53 %) */
			/*
			 * Rows are taken two at a time: both rows' operands are in registers before the arithmetic starts (twice the independent
			 * work per wait), and the two xor-shuffle trees are folded into one -- after the first exchange (xor 16) lanes 0-15 carry
			 * row t and lanes 16-31 row t+1, so the remaining four steps serve both rows.  The pairing of partial sums is exactly that
			 * of the one-row tree (16, 8, 4, 2, 1), hence the same bits as pqp_gemv.cu.
			 */
			for (int t = 0; t < nrows; t += 2) {
				const bool two = (t + 1 < nrows);
				const float4 *src[2];
				int sidx[2] = { 0, 0 };
	#pragma unroll
				for (int k = 0; k < 2; k++) {
					const int tt = t + k;
					if (k == 1 && !two) {
						src[k] = src[0];
					} else if (tt < R) {
						src[k] = reinterpret_cast<const float4 *>(resid + (size_t)tt * ldq);
					} else {
						sidx[k] = cs;
						mbar_wait(&full[cs], cph);
						src[k] = reinterpret_cast<const float4 *>(ring + (size_t)cs * ldq);
						if (++cs == S) { cs = 0; cph ^= 1u; }
					}
				}
				float4 q0[YC], q1[YC];
	#pragma unroll
				for (int u = 0; u < YC; u++) {
					const int c = tid + u * CONSUMERS;
					q0[u] = (c < n4) ? src[0][c] : make_float4(0.f, 0.f, 0.f, 0.f);
					q1[u] = (two && c < n4) ? src[1][c] : make_float4(0.f, 0.f, 0.f, 0.f);
				}
				__syncwarp();
				if (lane == 0) { /* the rows are in registers: hand the stages back */
					if (t >= R) mbar_arrive(&empty[sidx[0]]);
					if (two && t + 1 >= R) mbar_arrive(&empty[sidx[1]]);
				}
				float num0 = 0.0f, den0 = 0.0f, num1 = 0.0f, den1 = 0.0f;
	#pragma unroll
				for (int u = 0; u < YC; u++) {
					acc4t(num0, den0, q0[u], yv[u]);
					acc4t(num1, den1, q1[u], yv[u]);
				}
				/* xor 16: lower half keeps row t, upper half row t+1 */
				const bool up = lane >= 16;
				const float sn = up ? num0 : num1, sd = up ? den0 : den1; /* what this lane gives away */
				const float rn = __shfl_xor_sync(0xffffffffu, sn, 16), rd = __shfl_xor_sync(0xffffffffu, sd, 16);
				float num = (up ? num1 : num0) + rn, den = (up ? den1 : den0) + rd;
	#pragma unroll
				for (int o = 8; o; o >>= 1) {
					num += __shfl_xor_sync(0xffffffffu, num, o);
					den += __shfl_xor_sync(0xffffffffu, den, o);
				}
				if ((lane & 15) == 0 && (!up || two)) {
					const int tt = t + (up ? 1 : 0);
					part[(0 * PQP_GEMV_WARPS + warp) * rows_max + tt] = num;
					part[(1 * PQP_GEMV_WARPS + warp) * rows_max + tt] = den;
				}
			}
		} else { /* long rows: one row per step (two would need two 32-48 KB stages at once and 2*YC float4 of operands in registers) */
			for (int t = 0; t < nrows; t++) {
				const float4 *src = nullptr;
				int s = 0;
				const int col = TMU ? (int)tm_col[t] : -1;
				const bool ringrow = t >= R && col < 0;
				if (t < R) {
					src = reinterpret_cast<const float4 *>(resid + (size_t)t * ldq);
				} else if (ringrow) {
					s = cs;
					mbar_wait(&full[s], cph);
					src = reinterpret_cast<const float4 *>(ring + (size_t)s * ldq);
				}
				float num = 0.0f, den = 0.0f;
				float4 q[YC];
				if (TMU && col >= 0) {
					if constexpr (YC == 4) tmem_ld16(tm_mine + (uint32_t)col, q[0], q[1], q[2], q[3]);
				} else {
#pragma unroll
					for (int u = 0; u < YC; u++) {
						const int c = tid + u * CONSUMERS;
						q[u] = (c < n4) ? src[c] : make_float4(0.f, 0.f, 0.f, 0.f);
					}
				}
				if (ringrow) {
					__syncwarp();
					if (lane == 0) mbar_arrive(&empty[s]); /* the row is in registers: hand the stage back */
					if (++cs == S) { cs = 0; cph ^= 1u; }
				}
#pragma unroll
				for (int u = 0; u < YC; u++) acc4t(num, den, q[u], yv[u]);
#pragma unroll
				for (int o = 16; o; o >>= 1) {
					num += __shfl_xor_sync(0xffffffffu, num, o);
					den += __shfl_xor_sync(0xffffffffu, den, o);
				}
				if (lane == 0) {
					part[(0 * PQP_GEMV_WARPS + warp) * rows_max + t] = num;
					part[(1 * PQP_GEMV_WARPS + warp) * rows_max + t] = den;
				}
			}
		}
		consumer_sync();

		float e_min = INFINITY, e_gap = 0.0f, e_jd = 0.0f, e_kkt = 0.0f, e_viol = -INFINITY;
		if (tid < nrows) {
			float num = 0.0f, den = 0.0f;
#pragma unroll
			for (int w = 0; w < PQP_GEMV_WARPS; w++) {
				num += part[(0 * PQP_GEMV_WARPS + w) * rows_max + tid];
				den += part[(1 * PQP_GEMV_WARPS + w) * rows_max + tid];
			}
			num = fmaf(th_r, y_mine, num) + fmaxf(-fd_r, 0.0f);
			den = fmaf(th_r, y_mine, den) + fmaxf(fd_r, 0.0f);
			if (!is_last) {
				const float yn = __fdiv_rn(num, den) * y_mine;
				if (ll) st_packet(pk_out + r0 + tid, yn, (uint32_t)(p + 1));
				else y_out[r0 + tid] = yn;
			} else if (ll) {
				a.ybuf1[r0 + tid] = y_mine; /* the answer, as a plain vector */
			}
			if (is_last) {
				const float gq = den - num;
				e_min = gq;
				e_gap = y_mine * gq;
				e_jd = y_mine * (0.5f * (gq + fd_r));
				e_kkt = fabsf(fminf(y_mine, gq));
				e_viol = -gq - kp_tol;
			}
		}
		if (is_last) {
#pragma unroll
			for (int o = 16; o; o >>= 1) {
				e_min = fminf(e_min, __shfl_xor_sync(0xffffffffu, e_min, o));
				e_gap += __shfl_xor_sync(0xffffffffu, e_gap, o);
				e_jd += __shfl_xor_sync(0xffffffffu, e_jd, o);
				e_kkt = fmaxf(e_kkt, __shfl_xor_sync(0xffffffffu, e_kkt, o));
				e_viol = fmaxf(e_viol, __shfl_xor_sync(0xffffffffu, e_viol, o));
			}
			if (lane == 0) {
				red[warp * 8 + 0] = e_min; red[warp * 8 + 1] = e_gap; red[warp * 8 + 2] = e_jd;
				red[warp * 8 + 3] = e_kkt; red[warp * 8 + 4] = e_viol;
			}
			consumer_sync();
			if (tid == 0) {
				for (int w = 1; w < PQP_GEMV_WARPS; w++) {
					e_min = fminf(e_min, red[w * 8 + 0]); e_gap += red[w * 8 + 1]; e_jd += red[w * 8 + 2];
					e_kkt = fmaxf(e_kkt, red[w * 8 + 3]); e_viol = fmaxf(e_viol, red[w * 8 + 4]);
				}
				float *slot = a.partials + (size_t)blockIdx.x * 8;
				slot[0] = e_min; slot[1] = e_gap; slot[2] = e_jd; slot[3] = e_kkt; slot[4] = e_viol;
			}
		}

		if (!ll || is_last) grid_barrier_consumers(a.barrier, bar_target, G);
		else consumer_sync(); /* part[] and the row constants are reused by the next pass */

		if (is_last && blockIdx.x == 0 && warp == 0) {
			float v_min = INFINITY, v_gap = 0.0f, v_jd = 0.0f, v_kkt = 0.0f;
			for (unsigned c = lane; c < G; c += 32) {
				const float *sl = a.partials + (size_t)c * 8;
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
				o.iters = a.iters;
				o.converged = 0;
				o.min_slack = v_min;
				o.gap = v_gap;
				o.Jd = v_jd + (a.Md ? 0.5f * a.Md[0] : 0.0f);
				o.kkt = v_kkt;
				*a.status = o;
				*a.result_buf = ll ? 1 : (p & 1);
			}
		}
	}
	if (TMU) {
		asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
		consumer_sync();
		if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base_s) : "memory");
	}
}

static size_t tma_smem_bytes(int ldq, const TmaGeom &g)
{
	return sizeof(float) * ((size_t)(g.stages + g.resident) * ldq + 2 * PQP_GEMV_WARPS * (size_t)g.rows_max + PQP_GEMV_WARPS * 8) +
	       sizeof(uint64_t) * 2 * (size_t)g.stages;
}

/* picks ring depth / residency for the shared-memory budget; returns 0 when the shape does not fit this kernel */
int pqp_gemv_tma_plan(int N, int ldq, int grid, size_t smem_budget, int *stages, int *resident, int *yc)
{
	const int n4 = ldq / 4;
	int YC = (n4 + CONSUMERS - 1) / CONSUMERS;
	if (YC > 8) return 0;
	YC = YC <= 1 ? 1 : (YC <= 2 ? 2 : (YC <= 4 ? 4 : 8));
	const int rows_max = (N + grid - 1) / grid + 1;
	if (rows_max > CONSUMERS) return 0;
	TmaGeom g;
	g.rows_max = rows_max;
	g.pinned = 0;
	/* whole slab resident if it fits (then a token ring of 1), else the deepest ring up to 6 with the rest resident */
	g.stages = 1;
	g.resident = rows_max;
	if (tma_smem_bytes(ldq, g) <= smem_budget) {
		*stages = 1; *resident = rows_max; *yc = YC;
		return 1;
	}
	const size_t row = (size_t)ldq * sizeof(float);
	g.resident = 0;
	g.stages = 0;
	const size_t fixed = tma_smem_bytes(ldq, g);
	if (fixed + 2 * row + 64 > smem_budget) return 0;
	int total_rows = (int)((smem_budget - fixed - 128) / (row + 16));
	/* ring depth by BYTES in flight: one SM needs ~100 KB outstanding to cover the HBM latency at its share of
	 * the bandwidth (measured at N=8192 with L2 pinning on: 4 x 32 KB rows beat 3 and 6) */
	int S = (int)((128 * 1024 + row - 1) / row);
	if (S < 3) S = 3;
	if (S > total_rows) S = total_rows;
	if (S < 2) return 0;
	*stages = S;
	*resident = total_rows - S;
	*yc = YC;
	return 1;
}

cudaError_t pqp_launch_gemv_tma(const pqp_gemv_args *a, int stages, int resident, int pinned, int yc, void *pk0, void *pk1,
				cudaStream_t s)
{
	TmaGeom g;
	g.stages = stages;
	g.resident = resident;
	g.pinned = pinned;
	g.rows_max = (a->N + a->grid - 1) / a->grid + 1;
	/* eight rows per slab parked in tensor memory: long-row instantiation only, and only when something is streamed at all */
	g.tmem_rows = (yc == 4 && g.rows_max <= TMA_TM_ROWS_MAX && resident < g.rows_max) ? 8 : 0;
	if (pqp_env("PQP_TMA_TMEM")) {
		const int v = atoi(pqp_env("PQP_TMA_TMEM"));
		if (v >= 0 && v <= 8 && g.tmem_rows) g.tmem_rows = v;
	}
	g.pol_keep = 2;
	g.pol_stream = 1;
	g.pk0 = reinterpret_cast<uint2 *>(pk0);
	g.pk1 = reinterpret_cast<uint2 *>(pk1);
	if (g.pk0) {
		cudaError_t e0 = cudaMemsetAsync(pk0, 0xFF, (size_t)a->ldq * sizeof(uint2), s);
		if (e0 == cudaSuccess) e0 = cudaMemsetAsync(pk1, 0xFF, (size_t)a->ldq * sizeof(uint2), s);
		if (e0 != cudaSuccess) return e0;
	}
	if (pqp_env("PQP_POL_KEEP")) g.pol_keep = atoi(pqp_env("PQP_POL_KEEP"));
	if (pqp_env("PQP_POL_STREAM")) g.pol_stream = atoi(pqp_env("PQP_POL_STREAM"));
	const size_t smem = tma_smem_bytes(a->ldq, g);
	const void *fn = NULL;
	switch (yc) {
	case 1: fn = (const void *)gemv_tma_kernel<1, false>; break;
	case 2: fn = (const void *)gemv_tma_kernel<2, false>; break;
	case 4: fn = g.tmem_rows > 0 ? (const void *)gemv_tma_kernel<4, true> : (const void *)gemv_tma_kernel<4, false>; break;
	case 8: fn = (const void *)gemv_tma_kernel<8, false>; break;
	default: return cudaErrorInvalidValue;
	}
	cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	e = cudaMemsetAsync(a->barrier, 0, sizeof(unsigned), s);
	if (e != cudaSuccess) return e;
	pqp_gemv_args args = *a;
	void *params[] = { (void *)&args, (void *)&g };
	return cudaLaunchCooperativeKernel(fn, dim3(a->grid), dim3(TMA_THREADS), params, smem, s);
}
