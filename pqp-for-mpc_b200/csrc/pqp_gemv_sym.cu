/*
 * pqp_gemv_sym.cu -- single-problem PQP loop that reads only the UPPER TRIANGLE of a symmetric Qd (sm_100a).
 *
 * Qd = Gp*Qp_inv*Gp' (computeQd, PQP_CPU.c:440-443) is symmetric in exact arithmetic; whenever the fp32 matrix the
 * handle holds is symmetric element for element (checked once on the device: the generator instances of
 * testing/test_generator.c are, their Qp_inv being diagonal), the update of updateY2 (PQP_CPU.c:603-618)
 *
 *     num_i = sum_j max(-q_ij,0) y_j      den_i = sum_j max(q_ij,0) y_j
 *
 * can be formed from the strictly upper triangle alone: an element q_ij (i<j) feeds row i with y_j AND row j with y_i.
 * That halves the bytes per update (2N^2 instead of 4N^2: 134 MB at N=8192, most of which the 126 MB L2 holds), so the
 * loop leaves the HBM roofline of pqp_gemv_tma.cu behind.  SURVEY.md section 8(f)4.
 *
 * Layout (built once per handle): the upper triangle as 128x128 tiles, strictly-lower part, diagonal and padding zeroed,
 * each tile stored as two UNITS of 64 rows x 128 columns (32 KB, elements in the order the lanes read them), units ordered
 * by column strip J, then tile row I, then half h.  CTA c owns a contiguous range of units -- one contiguous piece of
 * memory -- cut so that every CTA carries the same cost (units + column flushes).
 *
 *   copies     every warp runs its own cp.async pipeline over its 4 KB slab of each unit (SY_D units deep; the first R units
 *              of the range stay in shared memory for the whole launch).  Q does not depend on y, so the pipelines run ahead
 *              across iteration boundaries.  No block-wide synchronisation in the unit loop.
 *   units      8 warps x 8 rows; lane (a,b) owns rows a, 4+a of its warp's eight and columns 16b..16b+15:
 *                row direction    rn/rd[row] += max(-+q,0) * y_J[col]   -> three shuffle rounds over b, then one 16-byte packet
 *                                 {num, epoch, den, epoch} per (tile, row) straight from registers (deferred by one unit so the
 *                                 shuffle chain overlaps the next unit's arithmetic)
 *                column direction cn/cd[col] += max(-+q,0) * y_I[row]   -> stays in registers for the whole strip; at a
 *                                 strip change the sums are folded over a, added over the 8 warps in fixed order through
 *                                 shared memory and leave as one packet per (CTA, strip, column)
 *              arithmetic in packed pairs (fma.rn.f32x2): 64 FFMA2 + 64 FMNMX per lane and unit.
 *   owners     CTA c also owns rows [N*c/G, N*(c+1)/G): lanes along consecutive rows (their packets are contiguous), four
 *              thread groups over the terms of a row (row direction: one per tile of its tile row; column direction: one per
 *              CTA that touched its strip), summed in fixed order; then (theta_i + max(-+q_ii,0)) y_i and F-+, the update
 *              with IEEE division, and y_i+ leaves as a {value, epoch} packet.
 *   y          every CTA collects the blocks of y its units use (flag-in-data exchange as in pqp_gemv_tma.cu, restricted to
 *              those blocks; see the comment at need[] for why that keeps the ping-pong buffers safe).
 * All sums are in a fixed order: results are reproducible bit for bit from run to run.
 */
#include "pqp_internal.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define SY_BS 128                   /* tile edge */
#define SY_UR 64                    /* rows per unit */
#define SY_UNIT (SY_UR * SY_BS)     /* floats per unit (32 KB) */
#define SY_CONS 256
#define SY_WARPS 8
#define SY_THREADS SY_CONS
#define SY_FLUSH_COST 1.7
#define SY_RC 6                     /* per owned row: theta+q-, theta+q+, Fd, slack tolerance, y, previous y */
#define SY_D 4                      /* units in flight per warp (cp.async groups) */
#define SY_YB 4                     /* float4 of y in flight per thread */
#define SY_UMAX 512                 /* most units per CTA the tensor-memory tables cover (N = 16384 on 148 CTAs: 113) */

#ifdef PQP_SYM_DEBUG
#define SY_DBG g.dbg_
#else
#define SY_DBG 0
#endif

typedef unsigned long long u64;

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
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
/* back-off between two polls of a packet that is not there yet.  A wait in this kernel lasts microseconds; one that lasts seconds
 * means a packet will never come (a defect): trap, so that the launch fails instead of hanging the device. */
#define SY_SPIN_LIMIT (1u << 23)
__device__ __forceinline__ void spin_pause(unsigned &spins)
{
	__nanosleep(20);
	if (++spins > SY_SPIN_LIMIT) __trap();
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(SY_CONS) : "memory"); }

__device__ __forceinline__ void grid_barrier_consumers(unsigned *counter, unsigned &target, unsigned nblocks)
{
	consumer_sync();
	if (threadIdx.x == 0) {
		target += nblocks;
		__threadfence();
		asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
		unsigned v, spins = 0;
		do {
			asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
			if (++spins > (SY_SPIN_LIMIT << 2)) __trap();
		} while ((int)(v - target) < 0);
		__threadfence();
	}
	consumer_sync();
}

/* {value, epoch} packets of y (8 bytes) and {num, epoch, den, epoch} packets of partial sums (16 bytes): one store, polling loads */
__device__ __forceinline__ void st_packet(uint2 *dst, float v, uint32_t epoch)
{
	asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(__float_as_uint(v)), "r"(epoch) : "memory");
}
__device__ __forceinline__ float ld_packet(const uint2 *src, uint32_t epoch)
{
	uint32_t v, e;
	unsigned spins = 0;
	for (;;) {
		asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(e) : "l"(src) : "memory");
		if (e == epoch) break;
		spin_pause(spins);
	}
	return __uint_as_float(v);
}
__device__ __forceinline__ void st_pair(uint4 *dst, float num, float den, uint32_t epoch)
{
	asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(__float_as_uint(num)), "r"(epoch),
		     "r"(__float_as_uint(den)), "r"(epoch)
		     : "memory");
}
__device__ __forceinline__ uint4 ld_pair_raw(const uint4 *src)
{
	uint4 v;
	asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src) : "memory");
	return v;
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint64_t policy)
{
	asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(policy) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

/*
 * Tensor memory as a second on-chip store for units.  The loop has no tensor-core work, so the SM's 256 KB of TMEM (128 lanes x
 * 512 columns x 32 bit) would sit idle; it holds eight more 32 KB units per SM for the whole launch (37 MB over 148 SMs that
 * neither HBM nor L2 has to deliver again on every update).  A warp reaches lanes 32*(warp%4)..+31 only: warps 0-3 use columns
 * [64u, 64u+32) of unit slot u, warps 4-7 columns [64u+32, 64u+64); a thread's 32 columns are its eight float4 of the unit.
 */
/* sixteen columns -> four float4; the registers are valid only after tmem_ld_wait() */
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float4 &q0, float4 &q1, float4 &q2, float4 &q3)
{
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
		     : "=f"(q0.x), "=f"(q0.y), "=f"(q0.z), "=f"(q0.w), "=f"(q1.x), "=f"(q1.y), "=f"(q1.z), "=f"(q1.w), "=f"(q2.x), "=f"(q2.y), "=f"(q2.z),
		       "=f"(q2.w), "=f"(q3.x), "=f"(q3.y), "=f"(q3.z), "=f"(q3.w)
		     : "r"(taddr));
}
/* the wait names the registers it makes valid, so that no use of them can be scheduled ahead of it */
__device__ __forceinline__ void tmem_ld_wait(float4 &q0, float4 &q1, float4 &q2, float4 &q3)
{
	asm volatile("tcgen05.wait::ld.sync.aligned;"
		     : "+f"(q0.x), "+f"(q0.y), "+f"(q0.z), "+f"(q0.w), "+f"(q1.x), "+f"(q1.y), "+f"(q1.z), "+f"(q1.w), "+f"(q2.x), "+f"(q2.y), "+f"(q2.z),
		       "+f"(q2.w), "+f"(q3.x), "+f"(q3.y), "+f"(q3.z), "+f"(q3.w));
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float4 (&q)[8])
{
	uint32_t r[32];
	const int order[8] = { 0, 1, 4, 5, 2, 3, 6, 7 }; /* the two halves the unit loop consumes one after the other */
#pragma unroll
	for (int k = 0; k < 8; k++) {
		const float4 v = q[order[k]];
		r[4 * k] = __float_as_uint(v.x); r[4 * k + 1] = __float_as_uint(v.y);
		r[4 * k + 2] = __float_as_uint(v.z); r[4 * k + 3] = __float_as_uint(v.w);
	}
	asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
		     "%29,%30,%31,%32};" ::"r"(taddr),
		     "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
		     "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
		     "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
		     : "memory");
}

/* packed fp32 pairs (FFMA2) */
__device__ __forceinline__ u64 pk2(float lo, float hi)
{
	u64 r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
	return r;
}
__device__ __forceinline__ void up2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
	u64 d;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
	return d;
}

__device__ __forceinline__ u64 add2(u64 a, u64 b)
{
	float al, ah, bl, bh;
	up2(a, al, ah);
	up2(b, bl, bh);
	return pk2(al + bl, ah + bh);
}
__device__ __forceinline__ u64 shfl2(u64 v, int o)
{
	float l, h;
	up2(v, l, h);
	return pk2(__shfl_xor_sync(0xffffffffu, l, o), __shfl_xor_sync(0xffffffffu, h, o));
}

/* one float4 of one row: p = max(q,0), n = max(-q,0); row sums against y_J, column sums against y_I */
__device__ __forceinline__ void sym_row(const float4 q, const u64 yj01, const u64 yj23, const float yi, const u64 neg1, u64 &rn, u64 &rd,
					u64 &cn01, u64 &cn23, u64 &cd01, u64 &cd23)
{
	const u64 yi2 = pk2(yi, yi);
	const u64 p01 = pk2(fmaxf(q.x, 0.0f), fmaxf(q.y, 0.0f)), p23 = pk2(fmaxf(q.z, 0.0f), fmaxf(q.w, 0.0f));
	/* max(-q,0) on the ALU pipe (p - q through the FMA pipe is the same value, measured 1% slower: the FMA pipe is the busier one) */
	const u64 n01 = pk2(fmaxf(-q.x, 0.0f), fmaxf(-q.y, 0.0f)), n23 = pk2(fmaxf(-q.z, 0.0f), fmaxf(-q.w, 0.0f));
	(void)neg1;
	rd = fma2(p01, yj01, rd);
	rn = fma2(n01, yj01, rn);
	cd01 = fma2(p01, yi2, cd01);
	cn01 = fma2(n01, yi2, cn01);
	rd = fma2(p23, yj23, rd);
	rn = fma2(n23, yj23, rn);
	cd23 = fma2(p23, yi2, cd23);
	cn23 = fma2(n23, yi2, cn23);
}

/* row sums of one unit over the eight column lanes: the first round also sorts the lane's two rows onto the two half-groups;
 * lanes 0 and 4 of every group of eight then hold {num, den} of rows a and 4+a and store the packet */
__device__ __forceinline__ void row_finish(float n0, float n1, float d0, float d1, uint4 *dst, uint32_t ep, int lane)
{
	const bool hi4 = (lane & 4) != 0;
	float num = (hi4 ? n1 : n0) + __shfl_xor_sync(0xffffffffu, hi4 ? n0 : n1, 4);
	float den = (hi4 ? d1 : d0) + __shfl_xor_sync(0xffffffffu, hi4 ? d0 : d1, 4);
	num += __shfl_xor_sync(0xffffffffu, num, 2);
	den += __shfl_xor_sync(0xffffffffu, den, 2);
	num += __shfl_xor_sync(0xffffffffu, num, 1);
	den += __shfl_xor_sync(0xffffffffu, den, 1);
	if ((lane & 3) == 0) st_pair(dst, num, den, ep);
}

/* y packets of pass p -> plain floats in shared memory, for the blocks listed in need[].  Not inlined: the polling buffers must
 * not take part in the register allocation of the unit loop. */
__device__ __noinline__ void sym_fetch_y(float *y_s, const uint2 *pk_in, const int *need, int nneed4, int n4, int N, int nb, uint32_t p, int tid)
{
	unsigned spins = 0;
	/* every thread spins on ONE float4 (many loads in flight per thread while the packets are not there yet saturate the
	 * L2 with polls: measured +9 us per update at N=8192), then takes the rest of its share in one batch */
	for (int xb = tid; xb < nneed4; xb += SY_YB * SY_CONS) {
		uint4 lo[SY_YB], hi[SY_YB];
		int cc[SY_YB];
#pragma unroll
		for (int q = 0; q < SY_YB; q++) {
			const int x = xb + q * SY_CONS;
			cc[q] = x < nneed4 ? need[1 + x / (SY_BS / 4)] * (SY_BS / 4) + x % (SY_BS / 4) : n4;
		}
		if (cc[0] < n4) {
			lo[0] = ld_pair_raw(reinterpret_cast<const uint4 *>(pk_in + 4 * cc[0]));
			hi[0] = ld_pair_raw(reinterpret_cast<const uint4 *>(pk_in + 4 * cc[0] + 2));
			while (lo[0].y != p || lo[0].w != p || hi[0].y != p || hi[0].w != p) {
				spin_pause(spins);
				lo[0] = ld_pair_raw(reinterpret_cast<const uint4 *>(pk_in + 4 * cc[0]));
				hi[0] = ld_pair_raw(reinterpret_cast<const uint4 *>(pk_in + 4 * cc[0] + 2));
			}
		}
#pragma unroll
		for (int q = 1; q < SY_YB; q++) {
			if (cc[q] < n4) {
				lo[q] = ld_pair_raw(reinterpret_cast<const uint4 *>(pk_in + 4 * cc[q]));
				hi[q] = ld_pair_raw(reinterpret_cast<const uint4 *>(pk_in + 4 * cc[q] + 2));
			}
		}
#pragma unroll
		for (int q = 0; q < SY_YB; q++) {
			const int c = cc[q];
			if (c < n4) {
				while (lo[q].y != p || lo[q].w != p || hi[q].y != p || hi[q].w != p) {
					spin_pause(spins);
					lo[q] = ld_pair_raw(reinterpret_cast<const uint4 *>(pk_in + 4 * c));
					hi[q] = ld_pair_raw(reinterpret_cast<const uint4 *>(pk_in + 4 * c + 2));
				}
				reinterpret_cast<float4 *>(y_s)[c] = make_float4(__uint_as_float(lo[q].x), __uint_as_float(lo[q].z),
										 __uint_as_float(hi[q].x), __uint_as_float(hi[q].z));
			}
		}
	}
	if (need[need[0]] == nb - 1) /* the last, partial float4 of y */
		for (int i = 4 * n4 + tid; i < N; i += SY_CONS) y_s[i] = ld_packet(pk_in + i, p);
}

/* the packets of row i: thread group k of SY_TPR sums terms k, k+SY_TPR, ... in ascending order, up to SY_OB polling loads in
 * flight.  Not inlined (see above).  Two shapes: 4 groups x 64 rows per round (N >= 6144: one round for the ~56 rows of a CTA) and
 * 8 groups x 32 rows (smaller N: half the terms per thread; measured 6-12% faster at N = 2560..4096, 4% slower at 8192). */
template <int SY_TPR, int SY_OB> __device__ __noinline__ float2 sym_row_terms(const uint4 *rp_in, const uint4 *cp_in, const int *tab_c0, const int *tab_c1, const int *tab_j0,
					      int nb, int maxseg, int i, int k, uint32_t ep)
{
	float num = 0.0f, den = 0.0f;
	unsigned spins = 0;
	const int Ib = i / SY_BS, li = i % SY_BS;
	const int nr = nb - Ib, c0 = tab_c0[Ib], nt = nr + tab_c1[Ib] - c0 + 1;
	auto term = [&](int m) -> const uint4 * {
		if (m < nr) {
			const int J = Ib + m;
			return rp_in + ((size_t)J * (J + 1) / 2 + Ib) * SY_BS + li;
		}
		const int cc = c0 + (m - nr);
		return cp_in + ((size_t)cc * maxseg + (Ib - tab_j0[cc])) * SY_BS + li;
	};
	/* terms k, k+SY_TPR, ... in ascending order: spin on the first, then SY_OB polling loads in flight */
	if (k < nt) {
		const uint4 *s0 = term(k);
		uint4 v0 = ld_pair_raw(s0);
		while (v0.y != ep || v0.w != ep) {
			spin_pause(spins);
			v0 = ld_pair_raw(s0);
		}
		num = __uint_as_float(v0.x);
		den = __uint_as_float(v0.z);
	}
	for (int m0 = k + SY_TPR; m0 < nt; m0 += SY_TPR * SY_OB) {
		const uint4 *src[SY_OB];
		uint4 v[SY_OB];
#pragma unroll
		for (int q = 0; q < SY_OB; q++) {
			const int m = m0 + SY_TPR * q;
			src[q] = m < nt ? term(m) : nullptr;
		}
#pragma unroll
		for (int q = 0; q < SY_OB; q++)
			if (src[q]) v[q] = ld_pair_raw(src[q]);
#pragma unroll
		for (int q = 0; q < SY_OB; q++)
			if (src[q]) {
				while (v[q].y != ep || v[q].w != ep) {
					spin_pause(spins);
					v[q] = ld_pair_raw(src[q]);
				}
				num += __uint_as_float(v[q].x);
				den += __uint_as_float(v[q].z);
			}
	}
	return make_float2(num, den);
}

} // namespace

struct SymGeom {
	const float *units;    /* [U][SY_UNIT], lane-major inside a unit (sym_build_units_kernel) */
	const int *cta_u0;     /* [G+1] first unit of every CTA */
	const int *cta_j0;     /* [G] column strip of that unit */
	const int *strip_c0;   /* [nb] first / last CTA touching strip J */
	const int *strip_c1;
	int nb, U, maxseg;
	int resident, pinned, rows_max;
	int pol_keep, pol_stream;
	uint2 *pk0, *pk1;      /* y packets [ldq], epochs preset to 0xFFFFFFFF */
	uint4 *rowpart;        /* [2][nT][128] row-direction packets, zeroed before launch (epoch 0 = none) */
	uint4 *colpart;        /* [2][G][maxseg][128] column-direction packets */
	int tmem_units;        /* units per CTA parked in tensor memory for the whole launch (0: tensor memory not allocated) */
	int dbg_;              /* timing experiments only (PQP_SYM_DBG): 1 no arithmetic, 2 no row packets, 4 no unit loads, 8 no column flush; results invalid */
	long long *prof;       /* debug (PQP_SYM_PROF=1): [G][4] cycles in y fetch / units / owner phase, else NULL */
};

/*
 * Shared memory: ring [SY_D][SY_UNIT] | resident [R][SY_UNIT] | y [nb*128] | scr [2][8][128] float2 | osum [4][64] float2 | rowc [rows_max][6] | red [8*8]
 *
 * Lane mapping inside a unit (64 rows x 128 columns): warp w owns rows 8w..8w+7; lane (a = lane>>3, b = lane&7) owns rows
 * 8w+a and 8w+4+a and columns 16b..16b+15 -- eight float4 per unit, stored lane-major so every shared-memory access is
 * conflict-free and every lane reads back exactly the 16-byte pieces it copied itself.  Each warp therefore runs its own
 * cp.async pipeline over its 4 KB slab of every unit (SY_D units deep, running ahead across iteration boundaries): there is no
 * block-wide synchronisation in the unit loop at all.  Row sums need three shuffle rounds over b; the sixteen column sums of a
 * lane stay in its registers for the whole strip (reduced over a and over the warps only when the strip changes).
 */
/*
 * TOL (run to tolerance, iters <= 0): every check_every passes the owners also form the stop-test terms of terminate()
 * (PQP_CPU.c:673-687, on g = Qd y + Fd) for y_p and publish them as three 16-byte packets per CTA; y_{p+1} is published as
 * always and the decision is read by every CTA at the start of the NEXT pass, all folding the 148 x 3 values in the same
 * order, so all decide alike.  A converged run leaves with y_p (each owner keeps the previous value of its rows): bit for bit
 * what the fixed-count solve returns at the reported count.
 */
template <bool TOL, int SY_TPR, bool TMU> __global__ void __launch_bounds__(SY_THREADS, 1) gemv_sym_kernel(const pqp_gemv_args a, const SymGeom g)
{
	__shared__ float chk_s[TOL ? 160 : 1][6];
	__shared__ uint32_t tmem_base_s;
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
	const int N = a.N, nb = g.nb;
	const unsigned G = gridDim.x;
	const int cta = blockIdx.x;
	const int r0 = (int)((long long)N * cta / G), r1 = (int)((long long)N * (cta + 1) / G);
	const int nrows = r1 - r0;
	const int u0 = g.cta_u0[cta], nU = g.cta_u0[cta + 1] - u0;
	const int R = min(g.resident, nU);
	const int L = nU - R;                /* units of the range outside shared memory ... */
	const int TM = TMU ? min(g.tmem_units, L) : 0; /* ... of which TM, spread evenly (so that the copy pipeline keeps running while they are worked on), live in tensor memory */
	const size_t nT = (size_t)nb * (nb + 1) / 2;

	float *ring = reinterpret_cast<float *>(smem_raw);
	float *resid = ring + (size_t)SY_D * SY_UNIT;
	float *y_s = resid + (size_t)g.resident * SY_UNIT;
	float2 *scr = reinterpret_cast<float2 *>(y_s + (size_t)nb * SY_BS); /* [2][8][128] column sums of the warps */
	float2 *osum = scr + 2 * SY_WARPS * SY_BS; /* [SY_TPR][64] partial sums of the owner phase */
	float *rowc = reinterpret_cast<float *>(osum + SY_CONS);
	float *red = rowc + SY_RC * (size_t)g.rows_max;
	int *tab_c0 = reinterpret_cast<int *>(red + SY_WARPS * 8); /* strip_c0 [nb], strip_c1 [nb], cta_j0 [G]: the owner phase reads them every pass */
	int *tab_c1 = tab_c0 + nb;
	int *tab_j0 = tab_c1 + nb;
	int *need = tab_j0 + G; /* [1 + nb]: count, then the 128-wide blocks of y this CTA reads: the column strips and tile rows of its units, nothing else.
	 * A CTA that reads block b holds a tile touching block b, so the owners of those rows wait for its packets before they publish
	 * the next y: the ping-pong packet buffers stay safe without a global barrier (a CTA polling rows it does not feed could be
	 * overtaken by two passes and spin on an epoch that is gone).  The y of its OWN rows never leaves the CTA (rowc[5*rr+4]). */

	/* TMU: ukind[t] = column of unit t in tensor memory or -1; snext[i] = the i-th unit of the range that is streamed */
	short *ukind = reinterpret_cast<short *>(need + 1 + nb);
	short *snext = ukind + SY_UMAX;

	const int passes = (TOL ? a.max_iters : a.iters) + 1; /* updates + one evaluation pass */
	uint4 *cpk4 = reinterpret_cast<uint4 *>(a.partials); /* TOL: [2][G][3] check packets */
	float *fin = a.partials + (TOL ? (size_t)2 * G * 12 : 0); /* final per-CTA slots, clear of the check packets */

	/* ---- this warp's copy pipeline ---- */
	const int T = nU - R - TM; /* streamed units per pass */
	const float *mine = g.units + (size_t)u0 * SY_UNIT + (warp * 256 + lane) * 4;
	const uint32_t ring_w = smem_u32(ring) + (uint32_t)(warp * 256 + lane) * 16u;
	const uint64_t pol_keep = make_policy(g.pol_keep), pol_stream = make_policy(g.pol_stream);
	const int Pn = min(g.pinned, max(T, 1));
	long long left = (long long)passes * T; /* copies still to request */
	const float *rsrc = mine + (size_t)R * SY_UNIT; /* next streamed unit to request, its ring slot, its place in the range */
	uint32_t rdst_ring = ring_w;
	int iu = 0, islot = 0, pacc = 0;
	auto request = [&]() {
		if (left > 0) {
			if (TMU) rsrc = mine + (size_t)snext[iu] * SY_UNIT; /* the units in tensor memory are not in this list */
			/* Pn of every T units are fetched evict_last, spread evenly through the range so L2 hits and HBM misses overlap in time */
			pacc += Pn;
			const bool keep = pacc >= T;
			if (keep) pacc -= T;
			const uint64_t pol = keep ? pol_keep : pol_stream;
#pragma unroll
			for (int j = 0; j < 8; j++) cp_async16(rdst_ring + j * 512u, rsrc + j * 128, pol);
			left--;
			rsrc += SY_UNIT;
			rdst_ring += SY_UNIT * 4u;
			if (++iu == T) {
				iu = 0;
				pacc = 0;
				rsrc = mine + (size_t)R * SY_UNIT;
			}
			if (++islot == SY_D) {
				islot = 0;
				rdst_ring = ring_w;
			}
		}
		cp_async_commit(); /* one group per call, empty or not: the group arithmetic below stays uniform */
	};
	if (TMU) {
		for (int x = tid; x < L; x += SY_CONS) {
			const int before = (int)(((long long)x * TM) / L); /* units in tensor memory among the first x */
			const bool here = (int)(((long long)(x + 1) * TM) / L) != before;
			ukind[R + x] = here ? (short)(64 * before) : (short)-1;
			if (!here) snext[x - before] = (short)(R + x);
		}
		consumer_sync();
	}
	{
		const uint32_t rdst = smem_u32(resid) + (uint32_t)(warp * 256 + lane) * 16u;
		for (int t = 0; t < R; t++)
#pragma unroll
			for (int j = 0; j < 8; j++) cp_async16(rdst + (uint32_t)t * (SY_UNIT * 4u) + j * 512u, mine + (size_t)t * SY_UNIT + j * 128, pol_stream);
		cp_async_commit();
		for (int m = 0; m < SY_D - 1; m++) request();
		cp_async_wait<SY_D - 1>(); /* the resident units have landed */
	}
	int cslot = 0; /* ring slot of the next streamed unit to consume */
	uint32_t tm_mine = 0; /* this thread's lane / column origin in tensor memory */
	if (TMU) {
		if (warp == 0) {
			asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
			asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
		}
		asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
		consumer_sync();
		asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
		tm_mine = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 32);
		for (int x = 0; x < L; x++) {
			const int col = ukind[R + x];
			if (col < 0) continue;
			const float4 *src = reinterpret_cast<const float4 *>(mine + (size_t)(R + x) * SY_UNIT);
			float4 q[8];
#pragma unroll
			for (int k = 0; k < 8; k++) q[k] = __ldcs(src + 32 * k);
			tmem_st32(tm_mine + (uint32_t)col, q);
		}
		asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
	}

	/* per-row constants of the rows this CTA finishes: theta_i + max(-+q_ii, 0) as the reference forms Q-+ + theta in fp32
	 * (computeQdn_theta / computeQdp_theta, PQP_CPU.c:524-537), F_i, and the slack tolerance */
	for (int rr = tid; rr < nrows; rr += SY_CONS) {
		const int i = r0 + rr;
		const float th = a.theta[i], qii = a.Q[(size_t)i * a.ldq + i];
		rowc[SY_RC * rr + 0] = th + fmaxf(-qii, 0.0f);
		rowc[SY_RC * rr + 1] = th + fmaxf(qii, 0.0f);
		rowc[SY_RC * rr + 2] = a.Fd[i];
		rowc[SY_RC * rr + 3] = a.Kp ? fmaxf(a.erc * a.Kp[i], a.eac) : a.eac;
		rowc[SY_RC * rr + 4] = __ldcg(a.ybuf0 + i); /* y of the rows this CTA owns */
	}
	for (int i = N + tid; i < nb * SY_BS; i += SY_CONS) y_s[i] = 0.0f; /* the padding of Q is zero; keep 0 * y finite */
	const int Jstart = g.cta_j0[cta];
	const int Istart = (u0 >> 1) - Jstart * (Jstart + 1) / 2, hstart = u0 & 1;
	for (int i = tid; i < nb; i += SY_CONS) {
		tab_c0[i] = g.strip_c0[i];
		tab_c1[i] = g.strip_c1[i];
	}
	for (int i = tid; i < (int)G; i += SY_CONS) tab_j0[i] = g.cta_j0[i];
	if (tid == 0) {
		/* y blocks needed: mark in need[1..nb], then compact in place */
		for (int i = 0; i < nb; i++) need[1 + i] = 0;
		const int ntiles = ((u0 + nU - 1) >> 1) - (u0 >> 1) + 1;
		int J = Jstart, I = Istart;
		for (int k = 0; k < ntiles; k++) {
			need[1 + J] = 1;
			need[1 + I] = 1;
			if (++I > J) {
				J++;
				I = 0;
			}
		}
		int cnt = 0;
		for (int i = 0; i < nb; i++)
			if (need[1 + i]) need[1 + cnt++] = i;
		need[0] = cnt;
	}
	consumer_sync();
	const int nneed4 = need[0] * (SY_BS / 4); /* float4 of y this CTA fetches per pass */
	const u64 neg1 = pk2(-1.0f, -1.0f);
	const int n4 = N / 4; /* whole float4 of y */
	const int la = lane >> 3, lb = lane & 7;
	unsigned bar_target = 0, flushes = 0;
	float e_min = INFINITY, e_gap = 0.0f, e_jd = 0.0f, e_kkt = 0.0f, e_viol = -INFINITY;

	int next_chk = 0;     /* TOL: check passes at p = 0, check_every, 2*check_every, ... */
	bool pending = false; /* TOL: the previous pass published a check; decide before this pass's units */
	for (int p = 0; p < passes; p++) {
		const bool is_last = (p == passes - 1);
		const bool chk = TOL && !is_last && p == next_chk;
		const uint2 *pk_in = (p & 1) ? g.pk1 : g.pk0;
		uint2 *pk_out = (p & 1) ? g.pk0 : g.pk1;
		const uint32_t ep = (uint32_t)p + 1u; /* epoch of this pass's partial packets */
		uint4 *rp_out = g.rowpart + (size_t)(p & 1) * nT * SY_BS;
		uint4 *cp_out = g.colpart + (size_t)(p & 1) * G * g.maxseg * SY_BS;

		long long tA = 0, tB = 0, tC = 0;
		if (g.prof) tA = clock64();
		if (TOL && pending) {
			/* the stop test of pass p-1: one 16-byte polling load per packet; published before that pass's y, so it is there */
			const unsigned cq = (unsigned)((p - 1) / a.check_every) + 1u;
			unsigned spins = 0;
			for (int x = tid; x < 3 * (int)G; x += SY_CONS) {
				const uint4 *src = cpk4 + (size_t)(cq & 1u) * G * 3 + x;
				uint4 v = ld_pair_raw(src);
				while (v.y != cq || v.w != cq) {
					spin_pause(spins);
					v = ld_pair_raw(src);
				}
				chk_s[x / 3][2 * (x % 3)] = __uint_as_float(v.x);
				chk_s[x / 3][2 * (x % 3) + 1] = __uint_as_float(v.z);
			}
		}
		/* ---- y of this pass -> shared memory ---- */
		if (p == 0) {
			for (int i = tid; i < N; i += SY_CONS) y_s[i] = __ldcg(a.ybuf0 + i);
		} else {
			sym_fetch_y(y_s, pk_in, need, nneed4, n4, N, nb, (uint32_t)p, tid);
		}
		consumer_sync();
		if (TOL && pending) {
			/* every warp of every CTA folds all CTAs' values in the same order: one decision for the whole grid */
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
				/* converged at y_{p-1}: every owner still holds it */
				for (int rr = tid; rr < nrows; rr += SY_CONS) a.ybuf1[r0 + rr] = rowc[SY_RC * rr + 5];
				if (cta == 0 && tid == 0) {
					pqp_status o;
					o.iters = p - 1; o.converged = 1; o.min_slack = v_min; o.gap = v_gap; o.Jd = Jd; o.kkt = v_kkt;
					*a.status = o;
					*a.result_buf = 1;
				}
				break; /* uniform across the grid */
			}
		}
		if (g.prof) tB = clock64();

		/* ---- the units of this CTA ---- */
		{
			int J = Jstart, I = Istart, h = hstart, seg = 0;
			bool fresh = true;
			u64 yj[8], cn[8], cd[8]; /* this lane's 16 columns as 8 pairs */
			float pn0 = 0.f, pn1 = 0.f, pd0 = 0.f, pd1 = 0.f; /* row sums of the previous unit, not yet reduced */
			uint4 *pdst = nullptr;
			bool pend = false;
			for (int t = 0; t < nU; t++) {
				if (fresh) {
					const float4 *yp = reinterpret_cast<const float4 *>(y_s + J * SY_BS + 16 * lb);
#pragma unroll
					for (int k = 0; k < 4; k++) {
						const float4 v = yp[k];
						yj[2 * k] = pk2(v.x, v.y);
						yj[2 * k + 1] = pk2(v.z, v.w);
					}
#pragma unroll
					for (int k = 0; k < 8; k++) cn[k] = cd[k] = 0ull; /* +0.0f pairs */
					fresh = false;
				}
				const float4 *up = nullptr;
				bool in_tm = false;
				uint32_t tm_col = 0;
				if (t < R) {
					up = reinterpret_cast<const float4 *>(resid + (size_t)t * SY_UNIT) + warp * 256 + lane;
				} else {
					const int col = TMU ? (int)ukind[t] : -1;
					in_tm = TMU && col >= 0;
					if (in_tm) {
						tm_col = (uint32_t)col;
					} else {
						cp_async_wait<SY_D - 2>(); /* all but the newest SY_D-2 groups: this unit's slab is in */
						up = reinterpret_cast<const float4 *>(ring + (size_t)cslot * SY_UNIT) + warp * 256 + lane;
						if (++cslot == SY_D) cslot = 0;
					}
				}
				const float *yip = y_s + I * SY_BS + h * SY_UR + 8 * warp + la;
				const float yi0 = yip[0], yi1 = yip[4];
				u64 rn0 = 0, rd0 = 0, rn1 = 0, rd1 = 0;
				/* two complete copies of the unit's arithmetic: a unit from tensor memory waits for its second half in the middle,
				 * and a wait inside the block the other units run through would cut that block's instruction schedule in two */
				if (TMU && in_tm) {
					float4 q[8];
					tmem_ld16(tm_mine + tm_col, q[0], q[1], q[4], q[5]);
					tmem_ld_wait(q[0], q[1], q[4], q[5]);
					tmem_ld16(tm_mine + tm_col + 16, q[2], q[3], q[6], q[7]); /* in flight behind the first half's arithmetic */
#pragma unroll
					for (int k = 0; k < 2; k++) {
						sym_row(q[k], yj[2 * k], yj[2 * k + 1], yi0, neg1, rn0, rd0, cn[2 * k], cn[2 * k + 1], cd[2 * k], cd[2 * k + 1]);
						sym_row(q[4 + k], yj[2 * k], yj[2 * k + 1], yi1, neg1, rn1, rd1, cn[2 * k], cn[2 * k + 1], cd[2 * k], cd[2 * k + 1]);
					}
					tmem_ld_wait(q[2], q[3], q[6], q[7]);
#pragma unroll
					for (int k = 2; k < 4; k++) {
						sym_row(q[k], yj[2 * k], yj[2 * k + 1], yi0, neg1, rn0, rd0, cn[2 * k], cn[2 * k + 1], cd[2 * k], cd[2 * k + 1]);
						sym_row(q[4 + k], yj[2 * k], yj[2 * k + 1], yi1, neg1, rn1, rd1, cn[2 * k], cn[2 * k + 1], cd[2 * k], cd[2 * k + 1]);
					}
				} else {
					float4 q[8];
					if (!(SY_DBG & 4)) {
#pragma unroll
						for (int k = 0; k < 8; k++) q[k] = up[32 * k];
					} else {
#pragma unroll
						for (int k = 0; k < 8; k++) q[k] = make_float4(1.f, -1.f, 2.f, -2.f);
					}
					if (SY_DBG & 1) {
						rn0 = pk2(q[0].x + q[1].y + q[2].z + q[3].w, yi0);
						rn1 = pk2(q[4].x + q[5].y + q[6].z + q[7].w, yi1);
					} else {
#pragma unroll
						for (int k = 0; k < 4; k++) {
							sym_row(q[k], yj[2 * k], yj[2 * k + 1], yi0, neg1, rn0, rd0, cn[2 * k], cn[2 * k + 1], cd[2 * k], cd[2 * k + 1]);
							sym_row(q[4 + k], yj[2 * k], yj[2 * k + 1], yi1, neg1, rn1, rd1, cn[2 * k], cn[2 * k + 1], cd[2 * k], cd[2 * k + 1]);
						}
					}
				}
				/* the slab is consumed (its values feed the sums above): request the one SY_D-1 units ahead into the freed slot */
				if (t >= R && !in_tm) request();
				/* the row sums of the PREVIOUS unit are reduced and stored here, next to this unit's arithmetic, so that a warp
				 * never sits on its own shuffle/store chain */
				if (pend && !(SY_DBG & 2)) row_finish(pn0, pn1, pd0, pd1, pdst, ep, lane);
				{
					float l0, h0, l1, h1;
					up2(rn0, l0, h0); up2(rn1, l1, h1);
					pn0 = l0 + h0; pn1 = l1 + h1;
					up2(rd0, l0, h0); up2(rd1, l1, h1);
					pd0 = l0 + h0; pd1 = l1 + h1;
					pdst = rp_out + (size_t)(u0 + t) * SY_UR + 8 * warp + ((lane & 4) ? 4 : 0) + la; /* tile * 128 + h * 64: units are two to a tile */
					pend = true;
				}
				/* next unit; a strip change (or the end of the range) flushes the column sums */
				h ^= 1;
				bool flush = (t == nU - 1);
				if (h == 0) {
					I++;
					if (I > J) flush = true;
				}
				if (flush && (SY_DBG & 8)) {
					seg++;
					fresh = true;
				} else if (flush) {
					/* over the four row lanes (a): after two folding rounds lane (a,b) holds columns 16b+4a..16b+4a+3 */
					const bool a1 = (lane & 16) != 0, a0 = (lane & 8) != 0;
					u64 yn[4], yd[4];
#pragma unroll
					for (int j = 0; j < 4; j++) {
						yn[j] = add2(a1 ? cn[4 + j] : cn[j], shfl2(a1 ? cn[j] : cn[4 + j], 16));
						yd[j] = add2(a1 ? cd[4 + j] : cd[j], shfl2(a1 ? cd[j] : cd[4 + j], 16));
					}
					u64 zn[2], zd[2];
#pragma unroll
					for (int j = 0; j < 2; j++) {
						zn[j] = add2(a0 ? yn[2 + j] : yn[j], shfl2(a0 ? yn[j] : yn[2 + j], 8));
						zd[j] = add2(a0 ? yd[2 + j] : yd[j], shfl2(a0 ? yd[j] : yd[2 + j], 8));
					}
					/* the eight warps' sums are added in fixed order through shared memory (two buffers: one barrier per flush) into one
					 * packet per (CTA, strip, column); per-warp packets would spare the barrier but cost the owners eight times the
					 * column terms -- measured slower overall */
					float c0, c1, c2, c3, e0, e1, e2, e3;
					up2(zn[0], c0, c1); up2(zn[1], c2, c3); up2(zd[0], e0, e1); up2(zd[1], e2, e3);
					float2 *scb = scr + (size_t)(flushes & 1) * SY_WARPS * SY_BS;
					flushes++;
					float4 *sc = reinterpret_cast<float4 *>(scb + warp * SY_BS + 16 * lb + 4 * la);
					sc[0] = make_float4(c0, e0, c1, e1);
					sc[1] = make_float4(c2, e2, c3, e3);
					consumer_sync();
					if (tid < SY_BS) {
						float sn = 0.0f, sd = 0.0f;
#pragma unroll
						for (int w = 0; w < SY_WARPS; w++) {
							const float2 v = scb[w * SY_BS + tid];
							sn += v.x;
							sd += v.y;
						}
						st_pair(cp_out + ((size_t)cta * g.maxseg + seg) * SY_BS + tid, sn, sd, ep);
					}
					seg++;
					fresh = true;
				}
				if (h == 0 && I > J) {
					J++;
					I = 0;
				}
			}
			if (pend && !(SY_DBG & 2)) row_finish(pn0, pn1, pd0, pd1, pdst, ep, lane);
		}

		if (g.prof) tC = clock64();
		/* ---- the rows this CTA owns: collect the packets, update, publish ---- */
		const uint4 *rp_in = rp_out, *cp_in = cp_out;
		for (int gb = 0; SY_DBG && gb < nrows; gb += SY_CONS) /* experiments: packets may be missing, publish anything */
			if (gb + tid < nrows && !is_last) st_packet(pk_out + r0 + gb + tid, 1.0f, (uint32_t)(p + 1));
		if (!SY_DBG)
		/* lanes along consecutive rows (their packets are contiguous: a warp-wide load touches 4-5 lines, not 32 -- with a thread
		 * group per row the phase cost one cycle per packet, 5000 cycles), SY_TPR thread groups over the terms of a row */
		for (int gb = 0; gb < nrows; gb += SY_CONS / SY_TPR) {
			const int rr = gb + tid % (SY_CONS / SY_TPR), k = tid / (SY_CONS / SY_TPR);
			const bool valid = rr < nrows;
			float num = 0.0f, den = 0.0f;
			if (valid) {
				const float2 nd = sym_row_terms<SY_TPR, (SY_TPR == 8 ? 9 : 18)>(rp_in, cp_in, tab_c0, tab_c1, tab_j0, nb, g.maxseg, r0 + rr, k, ep);
				num = nd.x;
				den = nd.y;
			}
			osum[k * (SY_CONS / SY_TPR) + tid % (SY_CONS / SY_TPR)] = make_float2(num, den);
			consumer_sync();
			if (k == 0) {
				num = den = 0.0f;
#pragma unroll
				for (int kk = 0; kk < SY_TPR; kk++) {
					const float2 v = osum[kk * (SY_CONS / SY_TPR) + tid];
					num += v.x;
					den += v.y;
				}
			}
			if (gb + SY_CONS / SY_TPR < nrows) consumer_sync(); /* osum is reused by the next round */
			if (valid && k == 0) {
				const int i = r0 + rr;
				const float y_mine = rowc[SY_RC * rr + 4], fd_r = rowc[SY_RC * rr + 2];
				num = fmaf(rowc[SY_RC * rr + 0], y_mine, num) + fmaxf(-fd_r, 0.0f);
				den = fmaf(rowc[SY_RC * rr + 1], y_mine, den) + fmaxf(fd_r, 0.0f);
				if (!is_last) {
					const float yn = __fdiv_rn(num, den) * y_mine;
					st_packet(pk_out + i, yn, (uint32_t)(p + 1));
					rowc[SY_RC * rr + 4] = yn;
					if (TOL) rowc[SY_RC * rr + 5] = y_mine;
				} else {
					a.ybuf1[i] = y_mine; /* the answer, as a plain vector */
				}
				if (is_last || chk) {
					const float gq = den - num;
					e_min = fminf(e_min, gq);
					e_gap += y_mine * gq;
					e_jd += y_mine * (0.5f * (gq + fd_r));
					e_kkt = fmaxf(e_kkt, fabsf(fminf(y_mine, gq)));
					e_viol = fmaxf(e_viol, -gq - rowc[SY_RC * rr + 3]);
				}
			}
		}

		if (g.prof && tid == 0 && p > 0) {
			const long long tD = clock64();
			g.prof[cta * 4 + 0] += tB - tA;
			g.prof[cta * 4 + 1] += tC - tB;
			g.prof[cta * 4 + 2] += tD - tC;
		}
		if (is_last || chk) {
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
				for (int w = 1; w < SY_WARPS; w++) {
					e_min = fminf(e_min, red[w * 8 + 0]); e_gap += red[w * 8 + 1]; e_jd += red[w * 8 + 2];
					e_kkt = fmaxf(e_kkt, red[w * 8 + 3]); e_viol = fmaxf(e_viol, red[w * 8 + 4]);
				}
				if (is_last) {
					float *slot = fin + (size_t)cta * 8;
					slot[0] = e_min; slot[1] = e_gap; slot[2] = e_jd; slot[3] = e_kkt; slot[4] = e_viol;
				} else {
					/* this CTA's share of the stop test on y_p, epoch = number of the check; read at the start of the next pass */
					const unsigned cq = (unsigned)(p / a.check_every) + 1u;
					uint4 *mine_c = cpk4 + ((size_t)(cq & 1u) * G + cta) * 3;
					st_pair(mine_c + 0, e_min, e_gap, cq);
					st_pair(mine_c + 1, e_jd, e_kkt, cq);
					st_pair(mine_c + 2, e_viol, 0.0f, cq);
				}
			}
			e_min = INFINITY; e_gap = 0.0f; e_jd = 0.0f; e_kkt = 0.0f; e_viol = -INFINITY;
		}
		if (is_last) {
			grid_barrier_consumers(a.barrier, bar_target, G);
			if (cta == 0 && warp == 0) {
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
					o.iters = p; /* TOL: the cap; a converged run left at the decision above */
					o.converged = 0;
					o.min_slack = v_min;
					o.gap = v_gap;
					o.Jd = v_jd + (a.Md ? 0.5f * a.Md[0] : 0.0f);
					o.kkt = v_kkt;
					*a.status = o;
					*a.result_buf = 1;
				}
			}
		} else {
			consumer_sync(); /* y_s is rewritten by the next pass */
		}
		if (TOL) {
			pending = chk;
			if (chk) next_chk += a.check_every;
		}
	}
	cp_async_wait<0>();
	if (TMU) {
		asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
		consumer_sync();
		if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base_s) : "memory");
	}
}

/* ---- one-time kernels: symmetry test and the unit array ---------------------------------------------------------------- */

/* counts pairs (i<j) with Q_ij != Q_ji (as floats: -0 == +0, NaN != NaN); 32x32 tiles, both read along rows */
__global__ void sym_check_kernel(const float *Q, int ldq, int N, unsigned *mismatch)
{
	__shared__ float tA[32][33], tB[32][33];
	const int bi = blockIdx.y, bj = blockIdx.x;
	if (bi > bj) return;
	const int tx = threadIdx.x, ty = threadIdx.y; /* 32 x 8 */
	for (int r = ty; r < 32; r += 8) {
		const int i = bi * 32 + r, j = bj * 32 + tx;
		tA[r][tx] = (i < N && j < N) ? Q[(size_t)i * ldq + j] : 0.0f;
		const int i2 = bj * 32 + r, j2 = bi * 32 + tx;
		tB[r][tx] = (i2 < N && j2 < N) ? Q[(size_t)i2 * ldq + j2] : 0.0f;
	}
	__syncthreads();
	unsigned bad = 0;
	for (int r = ty; r < 32; r += 8)
		if (!(tA[r][tx] == tB[tx][r])) bad++;
	bad = __reduce_add_sync(0xffffffffu, bad);
	if (tx == 0 && bad) atomicAdd(mismatch, bad);
}

/*
 * Qd = Gp Qp_inv Gp' built by pqp_setup is symmetric in exact arithmetic whenever Qp_inv is; in fp32 its two copies of an
 * element are two different sums and differ in the last bits (they do not for the generator's diagonal Qp_inv).  FAST order
 * already takes Qd from a GEMM whose sums are not the reference's, so it may as well take ONE value per pair: when every pair
 * agrees to rounding -- |Q_ij - Q_ji| <= tol * sqrt(|Q_ii Q_jj|), the scale of a rounding error of such a product -- both
 * copies are replaced by their mean.  Pass 1 counts the pairs that do not agree (or are not finite); pass 2 does nothing
 * unless that count is zero, so an unsymmetric Qp_inv keeps the matrix it produced.  Never applied to a user-supplied Qd.
 */
__global__ void sym_near_check_kernel(const float *Q, int ldq, int N, float tol, unsigned *mismatch)
{
	__shared__ float tA[32][33], tB[32][33];
	const int bi = blockIdx.y, bj = blockIdx.x;
	if (bi > bj) return;
	const int tx = threadIdx.x, ty = threadIdx.y; /* 32 x 8 */
	for (int r = ty; r < 32; r += 8) {
		const int i = bi * 32 + r, j = bj * 32 + tx;
		tA[r][tx] = (i < N && j < N) ? Q[(size_t)i * ldq + j] : 0.0f;
		const int i2 = bj * 32 + r, j2 = bi * 32 + tx;
		tB[r][tx] = (i2 < N && j2 < N) ? Q[(size_t)i2 * ldq + j2] : 0.0f;
	}
	__syncthreads();
	unsigned bad = 0;
	for (int r = ty; r < 32; r += 8) {
		const int i = bi * 32 + r, j = bj * 32 + tx;
		if (i < N && j < N && i < j) {
			const float a = tA[r][tx], b = tB[tx][r];
			const float scale = sqrtf(fabsf(Q[(size_t)i * ldq + i]) * fabsf(Q[(size_t)j * ldq + j]));
			if (!(fabsf(a - b) <= tol * scale)) bad++; /* also counts NaN / Inf */
		}
	}
	bad = __reduce_add_sync(0xffffffffu, bad);
	if (tx == 0 && bad) atomicAdd(mismatch, bad);
}

__global__ void sym_mean_kernel(float *Q, int ldq, int N, const unsigned *mismatch)
{
	__shared__ float tA[32][33], tB[32][33];
	if (*mismatch) return;
	const int bi = blockIdx.y, bj = blockIdx.x;
	if (bi > bj) return;
	const int tx = threadIdx.x, ty = threadIdx.y;
	for (int r = ty; r < 32; r += 8) {
		const int i = bi * 32 + r, j = bj * 32 + tx;
		tA[r][tx] = (i < N && j < N) ? Q[(size_t)i * ldq + j] : 0.0f;
		const int i2 = bj * 32 + r, j2 = bi * 32 + tx;
		tB[r][tx] = (i2 < N && j2 < N) ? Q[(size_t)i2 * ldq + j2] : 0.0f;
	}
	__syncthreads();
	for (int r = ty; r < 32; r += 8) {
		const int i = bi * 32 + r, j = bj * 32 + tx; /* upper copy (i, j), coalesced along j */
		if (i < N && j < N && i < j) Q[(size_t)i * ldq + j] = 0.5f * (tA[r][tx] + tB[tx][r]);
		const int i2 = bj * 32 + r, j2 = bi * 32 + tx; /* lower copy (i2, j2) = mean of (j2, i2) and (i2, j2) */
		if (i2 < N && j2 < N && j2 < i2) Q[(size_t)i2 * ldq + j2] = 0.5f * (tA[tx][r] + tB[r][tx]);
	}
}

cudaError_t pqp_launch_sym_mean(float *Q, int ldq, int N, float tol, unsigned *mismatch, cudaStream_t s)
{
	cudaError_t e = cudaMemsetAsync(mismatch, 0, sizeof(unsigned), s);
	if (e != cudaSuccess) return e;
	const int nbk = (N + 31) / 32;
	sym_near_check_kernel<<<dim3(nbk, nbk), dim3(32, 8), 0, s>>>(Q, ldq, N, tol, mismatch);
	sym_mean_kernel<<<dim3(nbk, nbk), dim3(32, 8), 0, s>>>(Q, ldq, N, mismatch);
	return cudaGetLastError();
}

__global__ void sym_build_units_kernel(float *units, const float *Q, int ldq, int N, int nb)
{
	const int u = blockIdx.x; /* unit */
	const int tile = u >> 1, h = u & 1;
	int J = (int)((sqrtf(8.0f * (float)tile + 1.0f) - 1.0f) * 0.5f);
	while (J * (J + 1) / 2 > tile) J--;
	while ((J + 1) * (J + 2) / 2 <= tile) J++;
	const int I = tile - J * (J + 1) / 2;
	float *dst = units + (size_t)u * SY_UNIT;
	/* storage order: [warp 8][half-row-group t 2][column group k 4][lane 32][4]; lane (a = lane>>3, b = lane&7) */
	for (int e = threadIdx.x; e < SY_UNIT; e += blockDim.x) {
		const int ee = e & 3, lane = (e >> 2) & 31, k = (e >> 7) & 3, t = (e >> 9) & 1, w = e >> 10;
		const int r = 8 * w + 4 * t + (lane >> 3), c = 16 * (lane & 7) + 4 * k + ee;
		const int gi = I * SY_BS + h * SY_UR + r, gj = J * SY_BS + c;
		dst[e] = (gi < gj && gj < N) ? Q[(size_t)gi * ldq + gj] : 0.0f;
	}
}

cudaError_t pqp_launch_sym_check(const float *Q, int ldq, int N, unsigned *mismatch, cudaStream_t s)
{
	cudaError_t e = cudaMemsetAsync(mismatch, 0, sizeof(unsigned), s);
	if (e != cudaSuccess) return e;
	const int nbk = (N + 31) / 32;
	sym_check_kernel<<<dim3(nbk, nbk), dim3(32, 8), 0, s>>>(Q, ldq, N, mismatch);
	return cudaGetLastError();
}

void pqp_gemv_sym_counts(int N, int *nb, int *U)
{
	*nb = (N + SY_BS - 1) / SY_BS;
	*U = *nb * (*nb + 1); /* two units per tile */
}

size_t pqp_gemv_sym_units_bytes(int N)
{
	int nb, U;
	pqp_gemv_sym_counts(N, &nb, &U);
	return (size_t)U * SY_UNIT * sizeof(float);
}

cudaError_t pqp_launch_build_sym_units(float *units, const float *Q, int ldq, int N, cudaStream_t s)
{
	int nb, U;
	pqp_gemv_sym_counts(N, &nb, &U);
	sym_build_units_kernel<<<U, 256, 0, s>>>(units, Q, ldq, N, nb);
	return cudaGetLastError();
}

/* host tables for a grid of G CTAs: cta_u0 [G+1], cta_j0 [G], strip_c0 / strip_c1 [nb]; returns the largest number of strips a CTA touches */
int pqp_gemv_sym_tables(int N, int G, int *cta_u0, int *cta_j0, int *strip_c0, int *strip_c1)
{
	int nb, U;
	pqp_gemv_sym_counts(N, &nb, &U);
	/* contiguous ranges of equal COST: a unit counts 1, the column flush at the end of a strip SY_FLUSH_COST (measured: the barrier in it makes every warp wait
	 * for the slowest); the short strips at the start of the triangle would otherwise make the first CTAs the slowest of every pass */
	{
		const double fc = SY_FLUSH_COST;
		const double total = (double)U + fc * nb;
		int c = 1, J = 0;
		double acc = 0.0;
		cta_u0[0] = 0;
		for (int u = 0; u < U && c < G; u++) {
			while ((J + 1) * (J + 2) <= u) J++;
			acc += 1.0 + ((u == (J + 1) * (J + 2) - 1) ? fc : 0.0);
			/* cut after unit u when the running cost reaches c/G of the total, keeping one unit for every CTA still to come */
			while (c < G && (acc >= total * c / G || U - (u + 1) <= G - c) && u + 1 > cta_u0[c - 1]) {
				cta_u0[c++] = u + 1;
				if (U - (u + 1) > G - c) break;
			}
		}
		while (c <= G) cta_u0[c++] = U;
		cta_u0[G] = U;
	}
	/* strip J holds units [J(J+1), (J+1)(J+2)) */
	int maxseg = 1;
	for (int c = 0, J = 0; c < G; c++) {
		while ((J + 1) * (J + 2) <= cta_u0[c]) J++;
		cta_j0[c] = J;
		int Jl = J;
		while ((Jl + 1) * (Jl + 2) <= cta_u0[c + 1] - 1) Jl++;
		if (cta_u0[c + 1] > cta_u0[c] && Jl - J + 1 > maxseg) maxseg = Jl - J + 1;
	}
	for (int J = 0, c = 0; J < nb; J++) {
		const int first = J * (J + 1), last = (J + 1) * (J + 2) - 1;
		while (cta_u0[c + 1] <= first) c++;
		strip_c0[J] = c;
		int cl = c;
		while (cta_u0[cl + 1] <= last) cl++;
		strip_c1[J] = cl;
	}
	return maxseg;
}

/* the host tables, for the CPU-side tests (not part of include/pqp.h); returns maxseg and the unit / block counts */
extern "C" int pqp_internal_sym_tables(int N, int G, int *cta_u0, int *cta_j0, int *strip_c0, int *strip_c1, int *nb, int *U)
{
	pqp_gemv_sym_counts(N, nb, U);
	if (*U < G) return 0;
	return pqp_gemv_sym_tables(N, G, cta_u0, cta_j0, strip_c0, strip_c1);
}

static size_t sym_smem_bytes(int nb, int rows_max, int resident)
{
	return sizeof(float) * ((size_t)(SY_D + resident) * SY_UNIT + (size_t)nb * SY_BS + 4 * SY_WARPS * SY_BS + 2 * SY_CONS + SY_RC * (size_t)rows_max +
				SY_WARPS * 8) + sizeof(int) * (3 * (size_t)nb + 1 + 512) + 2 * SY_UMAX * sizeof(short) + 128; /* + the tables (grid <= 512) */
}

/* residency for the shared-memory budget (the ring is SY_D units deep); 0 when the shape does not fit */
int pqp_gemv_sym_plan(int N, int grid, size_t smem_budget, int *stages, int *resident)
{
	int nb, U;
	pqp_gemv_sym_counts(N, &nb, &U);
	if (U < grid || N < 4 * SY_BS) return 0;
	const int rows_max = (N + grid - 1) / grid + 1;
	const int umax = (U + grid - 1) / grid;
	const size_t fixed = sym_smem_bytes(nb, rows_max, 0) + 4096; /* + the kernel's static shared memory (stop-test values, TMEM base) */
	if (fixed > smem_budget) return 0;
	int res = (int)((smem_budget - fixed) / ((size_t)SY_UNIT * 4));
	if (res > umax) res = umax;
	*stages = SY_D;
	*resident = res;
	return 1;
}

cudaError_t pqp_launch_gemv_sym(const pqp_gemv_args *a, const pqp_sym_plan *pl, void *pk0, void *pk1, cudaStream_t s)
{
	SymGeom g;
	g.units = pl->units;
	g.cta_u0 = pl->cta_u0; g.cta_j0 = pl->cta_j0; g.strip_c0 = pl->strip_c0; g.strip_c1 = pl->strip_c1;
	g.nb = pl->nb; g.U = pl->U; g.maxseg = pl->maxseg;
	g.resident = pl->resident; g.pinned = pl->pinned; g.tmem_units = pl->tmem;
	g.rows_max = (a->N + a->grid - 1) / a->grid + 1;
	g.pol_keep = 2;
	g.pol_stream = 1;
	if (pqp_env("PQP_POL_KEEP")) g.pol_keep = atoi(pqp_env("PQP_POL_KEEP"));
	if (pqp_env("PQP_POL_STREAM")) g.pol_stream = atoi(pqp_env("PQP_POL_STREAM"));
	g.pk0 = reinterpret_cast<uint2 *>(pk0);
	g.pk1 = reinterpret_cast<uint2 *>(pk1);
	g.rowpart = reinterpret_cast<uint4 *>(pl->rowpart);
	g.colpart = reinterpret_cast<uint4 *>(pl->colpart);
	g.dbg_ = pqp_env("PQP_SYM_DBG") ? atoi(pqp_env("PQP_SYM_DBG")) : 0; /* honoured only when built with -DPQP_SYM_DEBUG */
	g.prof = NULL;
	static long long *prof_dev = NULL;
	const int prof = pqp_env("PQP_SYM_PROF") && atoi(pqp_env("PQP_SYM_PROF"));
	if (prof) {
		if (!prof_dev) cudaMalloc((void **)&prof_dev, sizeof(long long) * 4 * 256);
		cudaMemsetAsync(prof_dev, 0, sizeof(long long) * 4 * 256, s);
		g.prof = prof_dev;
	}
	cudaError_t e = cudaMemsetAsync(pk0, 0xFF, (size_t)a->ldq * sizeof(uint2), s);
	if (e == cudaSuccess) e = cudaMemsetAsync(pk1, 0xFF, (size_t)a->ldq * sizeof(uint2), s);
	if (e == cudaSuccess) e = cudaMemsetAsync(pl->rowpart, 0, pl->rowpart_bytes, s);
	if (e == cudaSuccess) e = cudaMemsetAsync(pl->colpart, 0, pl->colpart_bytes, s);
	if (e == cudaSuccess) e = cudaMemsetAsync(a->barrier, 0, sizeof(unsigned), s);
	if (e != cudaSuccess) return e;
	const size_t smem = sym_smem_bytes(g.nb, g.rows_max, g.resident);
	const bool wide = a->N >= 6144; /* owner phase: 4 thread groups x 64 rows per round, else 8 x 32 */
	const int umax_cta = (pl->U + a->grid - 1) / a->grid + 2;
	if (2 * umax_cta > SY_UMAX) g.tmem_units = 0; /* the cost-balanced ranges differ by a few units from the mean */
	const bool tmu = g.tmem_units > 0;
	const void *fn = a->iters > 0 ? (wide ? (tmu ? (const void *)gemv_sym_kernel<false, 4, true> : (const void *)gemv_sym_kernel<false, 4, false>)
					      : (tmu ? (const void *)gemv_sym_kernel<false, 8, true> : (const void *)gemv_sym_kernel<false, 8, false>))
				      : (wide ? (tmu ? (const void *)gemv_sym_kernel<true, 4, true> : (const void *)gemv_sym_kernel<true, 4, false>)
					      : (tmu ? (const void *)gemv_sym_kernel<true, 8, true> : (const void *)gemv_sym_kernel<true, 8, false>));
	if (a->iters <= 0) e = cudaMemsetAsync(a->partials, 0, (size_t)2 * a->grid * 3 * sizeof(uint4), s); /* check packets: epoch 0 = none */
	if (e != cudaSuccess) return e;
	e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	pqp_gemv_args args = *a;
	void *params[] = { (void *)&args, (void *)&g };
	e = cudaLaunchCooperativeKernel(fn, dim3(a->grid), dim3(SY_THREADS), params, smem, s);
	if (prof && e == cudaSuccess) {
		long long hp[4 * 256];
		cudaStreamSynchronize(s);
		cudaMemcpy(hp, prof_dev, sizeof hp, cudaMemcpyDeviceToHost);
		double sum[3] = { 0, 0, 0 }, mx[3] = { 0, 0, 0 }, mn[3] = { 1e30, 1e30, 1e30 };
		for (int c = 0; c < a->grid; c++)
			for (int k = 0; k < 3; k++) {
				const double v = (double)hp[c * 4 + k] / (double)(a->iters > 0 ? a->iters : 1);
				sum[k] += v;
				if (v > mx[k]) mx[k] = v;
				if (v < mn[k]) mn[k] = v;
			}
		if (atoi(pqp_env("PQP_SYM_PROF")) > 1)
			for (int c = 0; c < a->grid; c++)
				fprintf(stderr, "cta %3d: y %6.0f units %6.0f owner %6.0f\n", c, (double)hp[c * 4] / (a->iters > 0 ? a->iters : 1), (double)hp[c * 4 + 1] / (a->iters > 0 ? a->iters : 1),
					(double)hp[c * 4 + 2] / (a->iters > 0 ? a->iters : 1));
		fprintf(stderr, "pqp: gemv_sym cycles per pass (min/mean/max over CTAs): y fetch %.0f/%.0f/%.0f  units %.0f/%.0f/%.0f  owner %.0f/%.0f/%.0f\n",
			mn[0], sum[0] / a->grid, mx[0], mn[1], sum[1] / a->grid, mx[1], mn[2], sum[2] / a->grid, mx[2]);
	}
	return e;
}
