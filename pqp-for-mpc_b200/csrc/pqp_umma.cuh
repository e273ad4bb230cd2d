/*
 * pqp_umma.cuh -- inline-PTX wrappers for the sm_100a tensor path (tcgen05 + TMEM + mbarrier +
 * bulk async copies) and the 3xTF32 operand split.  Hand-written: no CUTLASS/CuTe in the product.
 *
 * Shared-memory operand layout used everywhere here: the canonical K-major, no-swizzle UMMA layout.
 * A "core matrix" is 8 rows x 16 bytes (4 tf32) stored contiguously (128 B, row r at r*16).  For an
 * operand with R rows and K columns
 *      addr(r, k) = (k/4)*LBO + (r/8)*SBO + (r%8)*16 + (k%4)*4
 * SBO = byte stride between 8-row groups, LBO = byte stride between 4-column groups (both multiples
 * of 16).  One tcgen05.mma.kind::tf32 consumes K = 8 (two column groups) per instruction.
 */
#ifndef PQP_UMMA_CUH
#define PQP_UMMA_CUH

#include <cuda_runtime.h>
#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

/* ---- mbarrier ------------------------------------------------------------------------------ */
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
/* Bounded: a wait that has failed for 10 s of %globaltimer (every legitimate wait in these kernels is microseconds) traps -- a
 * protocol bug then surfaces as a CUDA error from the call instead of a hung device.  The timer is only read on the failing path. */
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	asm volatile(
		"{\n\t"
		".reg .pred p;\n\t"
		".reg .u64 t0, t1;\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@p bra UMMA_DONE;\n\t"
		"mov.u64 t0, %%globaltimer;\n\t"
		"UMMA_WAIT:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@p bra UMMA_DONE;\n\t"
		"mov.u64 t1, %%globaltimer;\n\t"
		"sub.u64 t1, t1, t0;\n\t"
		"setp.lt.u64 p, t1, 10000000000;\n\t"
		"@p bra UMMA_WAIT;\n\t"
		"trap;\n\t"
		"UMMA_DONE:\n\t"
		"}" ::"r"(smem_addr(bar)),
		"r"(parity)
		: "memory");
}

/* ---- proxies and fences ----------------------------------------------------------------------- */
/* generic-proxy writes to shared memory -> visible to the async proxy (tensor core / TMA reads) */
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

/* ---- TMEM ------------------------------------------------------------------------------------- */
/* one full warp; writes the TMEM base address to *slot (shared memory); ncols power of two >= 32 */
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t ncols)
{
	asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(slot)), "r"(ncols) : "memory");
	asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
	asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
/* 32 lanes x 16 consecutive 32-bit columns: thread l of the warp receives lane (base_lane + l) */
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v)
{
	uint32_t r[16];
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
		     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
		       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
		     : "r"(taddr)
		     : "memory");
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
	for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

/* ---- descriptors ---------------------------------------------------------------------------------- */
/* shared-memory matrix descriptor, K-major, SWIZZLE_NONE, version 1 (Blackwell) */
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
	uint64_t d = 0;
	d |= (uint64_t)((saddr >> 4) & 0x3FFF);            /* [0,14)  start address >> 4 */
	d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;  /* [16,30) leading (K-direction) byte offset >> 4 */
	d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;  /* [32,46) stride (8-row group) byte offset >> 4 */
	d |= (uint64_t)1 << 46;                            /* [46,48) descriptor version = 1 */
	return d;                                          /* base_offset 0, lbo_mode 0, layout_type 0 (no swizzle) */
}
/* instruction descriptor for kind::tf32, fp32 accumulate, both operands K-major, dense */
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N)
{
	return (1u << 4)              /* c_format  = F32  */
	       | (2u << 7)            /* a_format  = TF32 */
	       | (2u << 10)           /* b_format  = TF32 */
	       | ((uint32_t)(N >> 3) << 17) /* n_dim */
	       | ((uint32_t)(M >> 4) << 24); /* m_dim */
}

/* D[tmem] (+)= A[smem] * B[smem]; accumulate = 0 overwrites D */
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
	asm volatile(
		"{\n\t"
		".reg .pred p;\n\t"
		"setp.ne.b32 p, %4, 0;\n\t"
		"tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
		"}" ::"r"(d_tmem),
		"l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
		: "memory");
}
/* arrive on an mbarrier when every tcgen05.mma issued so far by this thread has completed */
__device__ __forceinline__ void mma_commit(uint64_t *bar)
{
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}

/* ---- 3xTF32 split ------------------------------------------------------------------------------------ */
/*
 * x = hi + lo + O(2^-24 |x|), both exactly representable in tf32 (10 explicit mantissa bits), so the
 * tensor core's truncation of the low 13 bits is a no-op.  hi is x rounded to nearest (cvt.rna.tf32.f32); lo is the exact
 * remainder rounded to nearest.  a*b ~= hi_a*hi_b + hi_a*lo_b + lo_a*hi_b  (error <= ~2^-22 |ab|).
 */
__device__ __forceinline__ float tf32_rn(float x)
{
	/* one instruction (round to nearest, ties away from zero; low 13 bits zero) instead of four integer operations: the setup GEMM's
	 * converter warps were the bound of that kernel (round 2) */
	uint32_t u;
	asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
	return __uint_as_float(u);
}
__device__ __forceinline__ void tf32_split(float x, float &hi, float &lo)
{
	hi = tf32_rn(x);
	lo = tf32_rn(x - hi);
}

} /* namespace umma */
#endif
