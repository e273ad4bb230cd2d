/*
 * pqp_imma.cuh -- shared pieces of the int8 digit-plane batched kernels (pqp_batched_imma.cu: one CTA per 32 problems;
 * pqp_batched_imma_pair.cu: a CTA pair per 64 problems): layout constants, inline-PTX wrappers, kernel parameters.
 */
#ifndef PQP_IMMA_CUH
#define PQP_IMMA_CUH

#include "pqp_internal.h"
#include "pqp_umma.cuh"

#define BI_SLICE 4096u            /* one 128 x 32 u8 tile */
#define BI_CHUNK (3u * BI_SLICE)  /* three digit planes of one (matrix, M tile, K step) */
#define BI_A_LBO 2048u            /* K-major A: byte stride between 16-element k groups */
#define BI_A_SBO 128u             /*            byte stride between 8-row groups */
#define BI_B_LBO 128u             /* MN-major B: byte stride between 8-k groups (16 problems x 8 k = 128 B core matrix) */
#define BI_YBITS 22
#define BI_MAX_MT 4
/* wait-time profile slots (experiment aid, PQP_IMMA_DBG=8) */
enum { PROF_MMA_TOTAL = 0, PROF_MMA_WAIT_BREADY, PROF_MMA_WAIT_TMEM, PROF_MMA_WAIT_FULL, PROF_EPI_TOTAL, PROF_EPI_WAIT_TMEM, PROF_EPI_REQUANT,
       PROF_PROD_WAIT_EMPTY };
#define PROF_T(var) const long long var = prof_on ? clock64() : 0
#define PROF_ADD(slot, t0) \
	if (prof_on) prof_acc[slot] += clock64() - (t0)

namespace {

__device__ __forceinline__ uint32_t cluster_ctarank()
{
	uint32_t r;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
	return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank()
{
	uint32_t r;
	asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
	return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
	asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
	asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s_plain(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(umma::smem_addr(dst)),
		     "l"(src), "r"(bytes), "r"(umma::smem_addr(bar))
		     : "memory");
}
__device__ __forceinline__ void bulk_g2s_mcast(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint16_t mask)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
			     umma::smem_addr(dst)),
		     "l"(src), "r"(bytes), "r"(umma::smem_addr(bar)), "h"(mask)
		     : "memory");
}
__device__ __forceinline__ void mma_commit_mcast(uint64_t *bar, uint16_t mask)
{
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
			     umma::smem_addr(bar)),
		     "h"(mask)
		     : "memory");
}
/* D[tmem] (+)= A[smem, u8, K-major] * B[smem, s8, MN-major]; int32 accumulate (exact) */
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
	asm volatile(
		"{\n\t"
		".reg .pred p;\n\t"
		"setp.ne.b32 p, %4, 0;\n\t"
		"tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
		"}" ::"r"(d_tmem),
		"l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
		: "memory");
}
/* one lane of a converged warp (the tensor-core / bulk-copy instructions take uniform operands: issuing them from
 * warp-uniform code under elect.sync lets ptxas keep the descriptors in uniform registers) */
__device__ __forceinline__ bool elect_one()
{
	uint32_t pred;
	asm volatile(
		"{\n\t"
		".reg .pred p;\n\t"
		"elect.sync _|p, 0xffffffff;\n\t"
		"selp.u32 %0, 1, 0, p;\n\t"
		"}"
		: "=r"(pred));
	return pred != 0;
}
/* the three MMAs of one K step: planes A0, A1, A2 (4 KB apart) against [Y0|Y1|Y2], [Y0|Y1], [Y0] */
__device__ __forceinline__ void mma_i8_step(uint32_t d, uint32_t d1, uint32_t d2, uint64_t a_desc, uint64_t b_desc, uint32_t id3, uint32_t id2,
					    uint32_t id1, uint32_t accumulate)
{
	asm volatile(
		"{\n\t"
		".reg .pred p, pt;\n\t"
		".reg .b64 a1, a2;\n\t"
		"setp.ne.b32 p, %8, 0;\n\t"
		"setp.eq.b32 pt, 0, 0;\n\t"
		"add.s64 a1, %3, 256;\n\t"
		"add.s64 a2, %3, 512;\n\t"
		"tcgen05.mma.cta_group::1.kind::i8 [%0], %3, %4, %5, p;\n\t"
		"tcgen05.mma.cta_group::1.kind::i8 [%1], a1, %4, %6, pt;\n\t"
		"tcgen05.mma.cta_group::1.kind::i8 [%2], a2, %4, %7, pt;\n\t"
		"}" ::"r"(d),
		"r"(d1), "r"(d2), "l"(a_desc), "l"(b_desc), "r"(id3), "r"(id2), "r"(id1), "r"(accumulate)
		: "memory");
}
/* three consecutive K steps (a 36 KB ring stage) in one go: 9 MMAs, descriptors derived inside the asm block so the
 * issuing thread moves its operands to uniform registers once per stage instead of once per K step */
__device__ __forceinline__ void mma_i8_step3(uint32_t d, uint32_t d1, uint32_t d2, uint64_t a_desc, uint64_t b_desc, uint32_t id3, uint32_t id2,
					     uint32_t id1, uint32_t accumulate)
{
	asm volatile(
		"{\n\t"
		".reg .pred p, pt;\n\t"
		".reg .b64 a01, a02, a10, a11, a12, a20, a21, a22, b1, b2;\n\t"
		"setp.ne.b32 p, %8, 0;\n\t"
		"setp.eq.b32 pt, 0, 0;\n\t"
		"add.s64 a01, %3, 256;\n\t"
		"add.s64 a02, %3, 512;\n\t"
		"add.s64 a10, %3, 768;\n\t"
		"add.s64 a11, %3, 1024;\n\t"
		"add.s64 a12, %3, 1280;\n\t"
		"add.s64 a20, %3, 1536;\n\t"
		"add.s64 a21, %3, 1792;\n\t"
		"add.s64 a22, %3, 2048;\n\t"
		"add.s64 b1, %4, 32;\n\t"
		"add.s64 b2, %4, 64;\n\t"
		"tcgen05.mma.cta_group::1.kind::i8 [%0], %3, %4, %5, p;\n\t"
		"tcgen05.mma.cta_group::1.kind::i8 [%1], a01, %4, %6, pt;\n\t"
		"tcgen05.mma.cta_group::1.kind::i8 [%2], a02, %4, %7, pt;\n\t"
		"tcgen05.mma.cta_group::1.kind::i8 [%0], a10, b1, %5, pt;\n\t"
		"tcgen05.mma.cta_group::1.kind::i8 [%1], a11, b1, %6, pt;\n\t"
		"tcgen05.mma.cta_group::1.kind::i8 [%2], a12, b1, %7, pt;\n\t"
		"tcgen05.mma.cta_group::1.kind::i8 [%0], a20, b2, %5, pt;\n\t"
		"tcgen05.mma.cta_group::1.kind::i8 [%1], a21, b2, %6, pt;\n\t"
		"tcgen05.mma.cta_group::1.kind::i8 [%2], a22, b2, %7, pt;\n\t"
		"}" ::"r"(d),
		"r"(d1), "r"(d2), "l"(a_desc), "l"(b_desc), "r"(id3), "r"(id2), "r"(id1), "r"(accumulate)
		: "memory");
}
/* instruction descriptor: kind::i8, D = s32, A = u8 K-major, B = s8 MN-major */
__host__ __device__ constexpr uint32_t idesc_i8(int M, int N)
{
	return (2u << 4)                      /* c_format = S32 */
	       | (0u << 7)                    /* a_format = unsigned 8 bit */
	       | (1u << 10)                   /* b_format = signed 8 bit */
	       | (0u << 15)                   /* A K-major */
	       | (1u << 16)                   /* B MN-major */
	       | ((uint32_t)(N >> 3) << 17)   /* n_dim */
	       | ((uint32_t)(M >> 4) << 24);  /* m_dim */
}
__device__ __forceinline__ void tmem_ld16_i32(uint32_t taddr, int *v)
{
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
		     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
		       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
		     : "r"(taddr)
		     : "memory");
}
__device__ __forceinline__ void tmem_ld8_i32(uint32_t taddr, int *v)
{
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
		     : "r"(taddr)
		     : "memory");
}
template <int PW> __device__ __forceinline__ void tmem_ld_i32(uint32_t taddr, int *v)
{
	if (PW == 16) tmem_ld16_i32(taddr, v);
	else tmem_ld8_i32(taddr, v);
}
/* 32 lanes x 16 consecutive 32-bit columns from registers into TMEM (thread l of the warp writes lane base_lane + l) */
__device__ __forceinline__ void tmem_st16_f32(uint32_t taddr, const float *v)
{
	asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
		     "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
		     "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
		     "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
		     "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
		     : "memory");
	asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st8_f32(uint32_t taddr, const float *v)
{
	asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
		     "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
		     "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
		     : "memory");
	asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
/* 32 lanes x 4 columns, no wait (pair with tmem_st_wait) */
__device__ __forceinline__ void tmem_st4_f32_nowait(uint32_t taddr, const float *v)
{
	asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])),
		     "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3]))
		     : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

/* power-of-two scales of a problem whose largest dual has biased exponent field ex (pmax < 2^(ex-126)):
 * quantise with 2^(22-f), f = ex-126;  undo with 2^(f-22).  ex clamped so both stay normal floats. */
__device__ __forceinline__ void problem_scales(uint32_t pmax_bits, float &sc, float &isc)
{
	uint32_t ex = pmax_bits >> 23;
	ex = ex < 22u ? 22u : (ex > 254u ? 254u : ex);
	sc = __uint_as_float((275u - ex) << 23);  /* 2^(148-ex) */
	isc = __uint_as_float((ex - 21u) << 23);  /* 2^(ex-148) */
}

/*
 * Signed base-256 digits of four quantised duals b0..b3 (0 <= b <= 2^22 + headroom), packed one plane per word:
 *   b = Y0*65536 + Y1*256 + Y2,  Y1, Y2 in [-128, 127].  With c = b + 128 and c1 = (c >> 8) + 128:
 *   Y2 = (c & 255) - 128, Y1 = (c1 & 255) - 128, Y0 = c1 >> 8; as bytes x - 128 == x ^ 0x80, applied to the packed word.
 */
__device__ __forceinline__ void digits4(int b0, int b1, int b2, int b3, uint32_t &w0, uint32_t &w1, uint32_t &w2)
{
	const uint32_t c_0 = (uint32_t)(b0 + 128), c_1 = (uint32_t)(b1 + 128), c_2 = (uint32_t)(b2 + 128), c_3 = (uint32_t)(b3 + 128);
	const uint32_t d_0 = (uint32_t)(((int)c_0 >> 8) + 128), d_1 = (uint32_t)(((int)c_1 >> 8) + 128), d_2 = (uint32_t)(((int)c_2 >> 8) + 128),
		       d_3 = (uint32_t)(((int)c_3 >> 8) + 128);
	/* byte 0 of four words -> one word; byte 1 of four words -> one word */
	w2 = __byte_perm(__byte_perm(c_0, c_1, 0x0040), __byte_perm(c_2, c_3, 0x0040), 0x5410) ^ 0x80808080u;
	w1 = __byte_perm(__byte_perm(d_0, d_1, 0x0040), __byte_perm(d_2, d_3, 0x0040), 0x5410) ^ 0x80808080u;
	w0 = __byte_perm(__byte_perm(d_0, d_1, 0x0051), __byte_perm(d_2, d_3, 0x0051), 0x5410);
}

/* ---- cluster (CTA pair) helpers: DSMEM stores / reductions / mbarrier arrivals on the peer ---- */
__device__ __forceinline__ uint32_t map_peer(uint32_t local_addr, uint32_t rank)
{
	uint32_t r;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
	return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v)
{
	asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void red_max_cluster(uint32_t addr, uint32_t v)
{
	asm volatile("red.relaxed.cluster.shared::cluster.max.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
/* arrive on an mbarrier of any CTA of the cluster (address from map_peer, or the own CTA's mapped address) */
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t addr)
{
	asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
/* relaxed arrival; the caller has issued fence_release_cluster() after its writes */
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t addr)
{
	asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void fence_release_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity)
{
	asm volatile(
		"{\n\t"
		".reg .pred p;\n\t"
		".reg .u64 t0, t1;\n\t"
		"mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
		"@p bra PAIR_DONE;\n\t"
		"mov.u64 t0, %%globaltimer;\n\t"
		"PAIR_WAIT:\n\t"
		"mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
		"@p bra PAIR_DONE;\n\t"
		"mov.u64 t1, %%globaltimer;\n\t"
		"sub.u64 t1, t1, t0;\n\t"
		"setp.lt.u64 p, t1, 10000000000;\n\t"
		"@p bra PAIR_WAIT;\n\t"
		"trap;\n\t" /* bounded like umma::mbar_wait */
		"PAIR_DONE:\n\t"
		"}" ::"r"(umma::smem_addr(bar)),
		"r"(parity)
		: "memory");
}
/* all state spaces: the remote (shared::cluster) digit stores must reach the peer's tensor core too */
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void st_cluster_v2(uint32_t addr, uint32_t a, uint32_t b)
{
	asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void tmem_ld4_i32(uint32_t taddr, int *v)
{
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}

} /* namespace */

struct BiParams {
	const unsigned char *Atiles; /* [MT][2 (neg,pos)][NKS][3 planes][4096 B] */
	const float4 *rowc;          /* [MT*128] {dn, dp, rsn, rsp}: diagonal terms and row scales (2^(e-8)) */
	const float *Fd;             /* [B x N] */
	float *Y;                    /* [B x N] in: y0, out: y_K */
	int N, B, iters;
	int MT, NKS;                 /* M tiles of 128 rows, K steps of 32 (padded to a multiple of ksc) */
	int ksc;                     /* K steps per ring stage: 3 (36 KB stages) or 1 */
	int dbuf;                    /* 1: two plane buffers -> new digits are stored speculatively while the MMAs still read the old ones */
	int stages;                  /* ring depth */
	uint32_t b_sbo;              /* B operand: byte stride between 16-problem groups = Kpad*16 */
	/* run-to-tolerance (iters <= 0): every check_every updates one EVALUATION pass (same MMAs, no update) forms g = den - num =
	 * Qd y + Fd per row and reduces the stop test of terminate() (PQP_CPU.c:673-687) per problem; a problem that
	 * passes is frozen at exactly that y.  All passes of a problem before it freezes are the fixed-count passes. */
	int max_iters, check_every;
	float erc, eac, eaj, erj;
	const float *Kp;             /* [N] or NULL */
	const float *Md;             /* [B] or NULL */
	pqp_status *status;          /* [B], written in tolerance mode */
	/* pqp_batched_imma_paired.cu only -- the h(x) refresh and the primal recovery inside the loop kernel (0 / NULL: not fused):
	 *   fx_GQ != NULL: Fp_b = Fp1 D_b + Fp2 x_b - Fp3 (computeFp, PQP_CPU.c:373-382) and Fd_b = GQ Fp_b + Kp (computeFd, :456-460) are
	 *   formed by the kernel's prologue in the reference's order and also written to fx_Fp_out / fx_Fd_out; Fd (above) is not read.
	 *   rc_U != NULL: U_b = -Qp_inv (Gp' y_b + Fp_b) (computeUfromY, :352-360) by the kernel's epilogue, reference order.
	 *   y0_const: Y holds no start vector, every dual starts at y_init (PQP_CPU.c:710). */
	const float *fx_X, *fx_D, *fx_Fp1, *fx_Fp2, *fx_Fp3, *fx_Fpc, *fx_GQ, *fx_Kp;
	int fx_nS, fx_nd, fx_Dstride, fx_M;
	float *fx_Fp_out, *fx_Fd_out;
	const float *rc_Gp, *rc_Qp_inv, *rc_Fp;
	float *rc_U;
	int y0_const;
	float y_init;
	/* pqp_batched_imma_paired.cu only -- run to tolerance in chunks (the host drives; pqp_api.cu tol_chunked):
	 *   m_resume [B]: max_k y[k] of the iterate BEFORE the start vector (left in m_out by the previous chunk): the start digits are then
	 *   quantised with the bound-based scale the uninterrupted loop would have used, so a chunked run equals the one-shot run bit for bit;
	 *   eval_part != NULL: after the `iters` updates one EVALUATION pass (same MMAs, no update) forms g = den - num = Qd y + Fd per row and
	 *   leaves, per warp and problem, {max(-g - tol), min g, sum y g, sum y (g + Fd)/2, max |min(y, g)|} in eval_part
	 *   [pair][rank][warp][group][8][5] -- the terms of terminate() (PQP_CPU.c:673-687), folded in a fixed order */
	const float *m_resume;
	float *m_out;
	float *eval_part;
	int dbg;                     /* experiment switches (PQP_IMMA_DBG): 2 skip all MMAs, 4 skip epilogue math, 8 print wait-time profile */
	long long *prof;             /* dbg & 8: [8] cycle counters of CTA 0 (see PROF_*) */
};

#endif
