/*
 * pqp_gemm_umma.cu -- the one-time dual construction on the 5th-gen tensor cores (sm_100a):
 *     C[a x c] = A[a x b] * Bt[c x b]'        (both operands K-major: row r holds its K values contiguously)
 * used by pqp_setup for  GQ = Gp*Qp_inv  (Bt = Qp_inv')  and  Qd = GQ*Gp'  (Bt = Gp), i.e. convertToDual's two
 * products (PQP_CPU.c:492, 442), which ARE dense contractions (2NM^2 + 2N^2M flop: 344 GFLOP at N=8192, M=2048).
 *
 * fp32 in, fp32 out, 3xTF32 on tcgen05.mma.kind::tf32 with fp32 accumulation in TMEM:
 *     a*b ~= hi_a*hi_b + hi_a*lo_b + lo_a*hi_b        (umma::tf32_split; error ~2^-22 per product)
 *
 * One CTA computes a 128 x 128 tile of C.  Per 32-deep K chunk all 256 threads load the fp32 operands from
 * global memory, split them into hi/lo, and store them into shared memory in the canonical K-major no-swizzle
 * UMMA layout (pqp_umma.cuh); one thread then issues 4 K-steps x 3 tcgen05.mma (M=128, N=128, K=8) and commits
 * them to the stage's mbarrier, so the tensor pipe works on chunk k while the threads load chunk k+1 (two
 * stages).  The accumulator lives in 128 TMEM columns; the epilogue reads it back with tcgen05.ld (32 lanes x
 * 16 columns per instruction) and stores fp32 rows.
 */
#include "pqp_internal.h"
#include "pqp_umma.cuh"

#include <stdlib.h>

#define GT_M 128
#define GT_N 128
#define GT_K 32
#define GT_THREADS 256
#define GT_DRAIN 2 /* chunks (of K = 32) per in-TMEM accumulation chain */
#define GT_SBO 128u
#define GT_LBO (16u * 128u + 16u) /* +16 B: the 8 column groups of one row land in different bank groups */
#define GT_TILE_BYTES (8u * GT_LBO)  /* 8 column groups of 4 */

struct GemmSmem {
	/* [stage][operand: A_hi, A_lo, B_hi, B_lo] */
	unsigned char tiles[2][4][GT_TILE_BYTES];
	uint64_t mma_done[2];
	uint32_t tmem_base;
};

/* loads 128 rows x 32 k of a K-major fp32 matrix, splits, stores hi/lo tiles */
__device__ __forceinline__ void load_split_tile(unsigned char *hi_tile, unsigned char *lo_tile, const float *__restrict__ G, int ld,
						int row0, int nrows, int k0, int nk, bool vec_ok)
{
	/* 128 rows x 8 column groups = 1024 float4; thread t takes groups t%8 of rows t/8 + 32u */
	const int cg = threadIdx.x % 8;
#pragma unroll
	for (int u = 0; u < 4; u++) {
		const int r = threadIdx.x / 8 + 32 * u;
		const int gr = row0 + r, gk = k0 + cg * 4;
		float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
		if (gr < nrows) {
			const float *src = G + (size_t)gr * ld + gk;
			if (vec_ok && gk + 3 < nk) {
				v = *reinterpret_cast<const float4 *>(src);
			} else {
				if (gk + 0 < nk) v.x = src[0];
				if (gk + 1 < nk) v.y = src[1];
				if (gk + 2 < nk) v.z = src[2];
				if (gk + 3 < nk) v.w = src[3];
			}
		}
		float4 h, l;
		umma::tf32_split(v.x, h.x, l.x);
		umma::tf32_split(v.y, h.y, l.y);
		umma::tf32_split(v.z, h.z, l.z);
		umma::tf32_split(v.w, h.w, l.w);
		const uint32_t off = (uint32_t)cg * GT_LBO + (uint32_t)(r / 8) * GT_SBO + (uint32_t)(r % 8) * 16u;
		*reinterpret_cast<float4 *>(hi_tile + off) = h;
		*reinterpret_cast<float4 *>(lo_tile + off) = l;
	}
}

__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_3xtf32_kernel(float *__restrict__ C, int ldc, const float *__restrict__ A, int lda, const float *__restrict__ Bt, int ldb, int a,
		   int b, int c)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	GemmSmem &sm = *reinterpret_cast<GemmSmem *>(smem_raw);
	const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
	const int i0 = blockIdx.y * GT_M, j0 = blockIdx.x * GT_N;

	if (tid == 0) {
		umma::mbar_init(&sm.mma_done[0], 1);
		umma::mbar_init(&sm.mma_done[1], 1);
		umma::mbar_fence_init();
	}
	if (warp == 0) umma::tmem_alloc(&sm.tmem_base, 128);
	umma::tc_fence_before();
	__syncthreads();
	umma::tc_fence_after();
	const uint32_t tmem = sm.tmem_base;

	const bool vecA = (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
	const bool vecB = (ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(Bt) & 15) == 0);
	const uint32_t idesc = umma::idesc_tf32(GT_M, GT_N);
	const int nchunks = (b + GT_K - 1) / GT_K;

	/*
	 * The tensor core adds into its fp32 accumulator with truncation (measured on B200: a chain of n accumulating
	 * MMAs over non-negative data ends ~n * 2.8e-8 low, tools/tc_bias_probe.py), so long K chains are NOT left in
	 * TMEM: every GT_DRAIN chunks (K = 64: 24 accumulating MMAs) the partial tile is read back and added to fp32
	 * registers with round-to-nearest, and the next chain starts from zero (accumulate = 0).
	 */
	const int lane_base = 32 * (warp % 4), col_base = 64 * (warp / 4);
	float acc[64];
#pragma unroll
	for (int e = 0; e < 64; e++) acc[e] = 0.0f;
	int commits[2] = { 0, 0 }; /* commits issued per stage barrier (uniform across threads) */

	for (int kc = 0; kc < nchunks; kc++) {
		const int st = kc & 1;
		/* the MMAs that last read this stage (chunk kc-2) have completed: either drained below or waited here */
		if (kc >= 2 && ((kc - 2) % GT_DRAIN) != GT_DRAIN - 1) umma::mbar_wait(&sm.mma_done[st], (uint32_t)((commits[st] - 1) & 1));
		load_split_tile(sm.tiles[st][0], sm.tiles[st][1], A, lda, i0, a, kc * GT_K, b, vecA);
		load_split_tile(sm.tiles[st][2], sm.tiles[st][3], Bt, ldb, j0, c, kc * GT_K, b, vecB);
		umma::fence_proxy_async();
		__syncthreads();
		const bool chain_start = (kc % GT_DRAIN) == 0;
		if (tid == 0) {
			umma::tc_fence_after();
			const uint32_t a_hi = umma::smem_addr(sm.tiles[st][0]), a_lo = umma::smem_addr(sm.tiles[st][1]);
			const uint32_t b_hi = umma::smem_addr(sm.tiles[st][2]), b_lo = umma::smem_addr(sm.tiles[st][3]);
#pragma unroll
			for (int ks = 0; ks < GT_K / 8; ks++) {
				const uint32_t o = (uint32_t)ks * 2u * GT_LBO;
				const uint64_t dah = umma::smem_desc(a_hi + o, GT_LBO, GT_SBO), dal = umma::smem_desc(a_lo + o, GT_LBO, GT_SBO);
				const uint64_t dbh = umma::smem_desc(b_hi + o, GT_LBO, GT_SBO), dbl = umma::smem_desc(b_lo + o, GT_LBO, GT_SBO);
				umma::mma_tf32(tmem, dal, dbh, idesc, (!chain_start || ks) ? 1u : 0u); /* small terms first */
				umma::mma_tf32(tmem, dah, dbl, idesc, 1u);
				umma::mma_tf32(tmem, dah, dbh, idesc, 1u);
			}
			umma::mma_commit(&sm.mma_done[st]);
		}
		commits[st]++;
		if ((kc % GT_DRAIN) == GT_DRAIN - 1 || kc == nchunks - 1) {
			/* end of a chain: wait for it, fold it into the registers */
			umma::mbar_wait(&sm.mma_done[st], (uint32_t)((commits[st] - 1) & 1));
			umma::tc_fence_after();
#pragma unroll
			for (int cc = 0; cc < 64; cc += 16) {
				float v[16];
				umma::tmem_ld16(tmem + ((uint32_t)lane_base << 16) + (uint32_t)(col_base + cc), v);
#pragma unroll
				for (int e = 0; e < 16; e++) acc[cc + e] += v[e];
			}
			umma::tc_fence_before();
			__syncthreads(); /* everyone has read the tile before the next chain overwrites it */
		}
	}
	/* epilogue: thread = one row of C, 64 consecutive columns */
	const int gi = i0 + lane_base + lane;
	if (gi < a) {
		float *dst = C + (size_t)gi * ldc + j0 + col_base;
#pragma unroll
		for (int e = 0; e < 64; e++)
			if (j0 + col_base + e < c) dst[e] = acc[e];
	}
	umma::tc_fence_before();
	__syncthreads();
	if (warp == 0) umma::tmem_dealloc(tmem, 128);
}

cudaError_t pqp_launch_gemm_umma(float *C, int ldc, const float *A, int lda, const float *Bt, int ldb, int a, int b, int c,
				 cudaStream_t s)
{
	/* the warp-specialised pipeline (pqp_gemm_umma_ws.cu) for anything but small products; PQP_GEMM_WS=0 keeps this kernel */
	if (pqp_gemm_umma_ws_wanted(a, b, c)) return pqp_launch_gemm_umma_ws(C, ldc, A, lda, Bt, ldb, a, b, c, NULL, s);
	const size_t smem = sizeof(GemmSmem) + 128;
	cudaError_t e = cudaFuncSetAttribute(gemm_3xtf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	dim3 grid((c + GT_N - 1) / GT_N, (a + GT_M - 1) / GT_M);
	gemm_3xtf32_kernel<<<grid, GT_THREADS, smem, s>>>(C, ldc, A, lda, Bt, ldb, a, b, c);
	return cudaGetLastError();
}
