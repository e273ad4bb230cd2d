/*
 * pqp_gemm_umma_ws.cu -- the one-time dual construction (convertToDual's two products, PQP_CPU.c:492, :442) on the 5th-gen tensor
 * cores as a warp-specialised pipeline:      C[a x c] = A[a x b] * Bt[c x b]'      fp32 in / out, 3xTF32, fp32 accumulate.
 *
 * Round 1's kernel (pqp_gemm_umma.cu, kept as the fallback for small shapes) had all 256 threads load, split and store a chunk,
 * then one thread issue its MMAs, a block barrier per 32-deep chunk and an exposed global-load latency per chunk: 12 ms for the
 * 344 GFLOP of config C3 (28 TFLOP/s fp32-equivalent, 9 % of the measured tf32 rate).  Here the roles are separate warps that
 * only meet at mbarriers:
 *   warps 0-7   loaders/converters: fp32 rows from global memory (float4, the loads of the next TWO 16-deep chunks in flight while the
 *               current one is split), umma::tf32_split into hi/lo, stores in the K-major no-swizzle UMMA layout of a 4-stage ring;
 *   warp 8      MMA issuer: per chunk 2 K-steps x 3 tcgen05.mma.kind::tf32 (lo*hi, hi*lo, hi*hi; M = 128, N = 192, K = 8) into one of
 *               TWO TMEM accumulators; tcgen05.commit frees the stage / publishes the accumulator;
 *   warps 9-20  drain/epilogue (4 TMEM lane quarters x 3 column blocks of 64): the tensor core adds into its fp32 accumulator with
 *               truncation (~2.8e-8 relative per accumulating MMA, always downwards on non-negative data), so a chain is cut after
 *               4 chunks (24 MMAs), read back with tcgen05.ld and added to fp32 registers with round-to-nearest -- while the tensor
 *               pipe already runs the next chain into the other accumulator.
 * CTA tile 128 x 192: N = 192 MMAs run at 94 % of the N = 256 rate (profiles/tensor_peaks_r2_shapes.txt) and leave the drain warps 64
 * accumulator registers per thread.  Roofline: tensor (tf32): 3 MMA flop per algorithmic flop.
 */
#include "pqp_internal.h"

#include <stdlib.h>
#include "pqp_umma.cuh"

#define GW_M 128
#define GW_N 192
#define GW_K 16
#define GW_LOADERS 256
#define GW_DRAINERS 384
#define GW_THREADS (GW_LOADERS + 32 + GW_DRAINERS)
#define GW_DRAIN 4 /* chunks (of K = 16) per in-TMEM accumulation chain: 24 accumulating MMAs */
#define GW_SBO 128u
#define GW_LBO_A (16u * 128u + 16u) /* 128 rows; +16 B: the 8 column groups of a row land in different bank groups */
#define GW_LBO_B (24u * 128u + 16u) /* 192 rows */
#define GW_CG (GW_K / 4)             /* column groups (float4) per row of a chunk */
#define GW_TILE_A (GW_CG * GW_LBO_A)
#define GW_TILE_B (GW_CG * GW_LBO_B)
#define GW_STAGE (2u * GW_TILE_A + 2u * GW_TILE_B)
#define GW_STAGES 4

namespace {
__device__ __forceinline__ bool gw_elect()
{
	uint32_t pred;
	asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
	return pred != 0;
}
/* float4 number f (0 .. rows*GW_CG-1) of a [rows x GW_K] chunk: row f/GW_CG, column group f%GW_CG */
__device__ __forceinline__ float4 gw_load(const float *__restrict__ G, int ld, int row0, int nrows, int k0, int nk, bool vec, int f)
{
	const int r = f / GW_CG, cg = f % GW_CG, gr = row0 + r, gk = k0 + cg * 4;
	float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
	if (gr < nrows) {
		const float *src = G + (size_t)gr * ld + gk;
		if (vec && gk + 3 < nk) {
			v = __ldcg(reinterpret_cast<const float4 *>(src)); /* L2 only: a 16-deep chunk uses half of each 128-byte line, the tiny L1 left beside 165 KB of shared memory would evict the other half before the next chunk asks for it */
		} else {
			if (gk + 0 < nk) v.x = src[0];
			if (gk + 1 < nk) v.y = src[1];
			if (gk + 2 < nk) v.z = src[2];
			if (gk + 3 < nk) v.w = src[3];
		}
	}
	return v;
}
__device__ __forceinline__ void gw_split_store(unsigned char *hi_tile, unsigned char *lo_tile, uint32_t lbo, int f, float4 v)
{
	const int r = f / GW_CG, cg = f % GW_CG;
	float4 h, l;
	umma::tf32_split(v.x, h.x, l.x);
	umma::tf32_split(v.y, h.y, l.y);
	umma::tf32_split(v.z, h.z, l.z);
	umma::tf32_split(v.w, h.w, l.w);
	const uint32_t off = (uint32_t)cg * lbo + (uint32_t)(r >> 3) * GW_SBO + (uint32_t)(r & 7) * 16u;
	*reinterpret_cast<float4 *>(hi_tile + off) = h;
	*reinterpret_cast<float4 *>(lo_tile + off) = l;
}
} /* namespace */

/* (registers are handed out per 4 warps: the 21 warps count as 24, so 80 registers per thread is the most a launch accepts) */
__global__ void __launch_bounds__(GW_THREADS, 1)
gemm_3xtf32_ws_kernel(float *__restrict__ C, int ldc, const float *__restrict__ A, int lda, const float *__restrict__ Bt, int ldb, int a, int b,
		      int c, const unsigned *__restrict__ sym_bad)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	unsigned char *ring = smem_raw; /* [stage][A_hi | A_lo | B_hi | B_lo] */
	uint64_t *full = reinterpret_cast<uint64_t *>(ring + GW_STAGES * GW_STAGE);
	uint64_t *empty = full + GW_STAGES;
	uint64_t *acc_full = empty + GW_STAGES; /* [2] */
	uint64_t *acc_empty = acc_full + 2;     /* [2] */
	uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
	const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
	/* tile order: groups of 8 row blocks, column tiles outermost inside a group -- the ~148 CTAs in flight then share 8 row blocks of A
	 * and ~18 column tiles of B (30 MB at C3) instead of 4 row blocks and ALL of B (67 MB per wave, re-read from HBM 8x: ncu) */
	int ti, tj;
	{
		const int tx = gridDim.x, ty = gridDim.y, L = blockIdx.y * tx + blockIdx.x;
		const int grp = L / (8 * tx), within = L - grp * 8 * tx;
		const int rows = min(8, ty - grp * 8);
		ti = grp * 8 + within % rows;
		tj = within / rows;
	}
	const int i0 = ti * GW_M, j0 = tj * GW_N;
	/* symmetric product (C = X S X' with S symmetric: Qd = Gp Qp_inv Gp'): only the tiles that touch the upper triangle are multiplied;
	 * their elements above the diagonal are stored twice, as C_ij and as C_ji.  *sym_bad is the device-side count of unequal pairs
	 * of S (pqp_launch_sym_check), so the decision needs no host round trip. */
	const bool sym = sym_bad != nullptr && a == c && __ldg(sym_bad) == 0u;
	if (sym && j0 + GW_N - 1 < i0) return; /* whole tile below the diagonal: its mirror image writes it */
	const int nchunks = (b + GW_K - 1) / GW_K;
	const int nchains = (nchunks + GW_DRAIN - 1) / GW_DRAIN;

	if (tid == 0) {
		for (int s = 0; s < GW_STAGES; s++) {
			umma::mbar_init(&full[s], GW_LOADERS / 32); /* one arrival per loader warp */
			umma::mbar_init(&empty[s], 1);
		}
		for (int i = 0; i < 2; i++) {
			umma::mbar_init(&acc_full[i], 1);
			umma::mbar_init(&acc_empty[i], GW_DRAINERS / 32);
		}
		umma::mbar_fence_init();
	}
	if (warp == GW_LOADERS / 32) umma::tmem_alloc(tmem_slot, 512);
	umma::tc_fence_before();
	__syncthreads();
	umma::tc_fence_after();
	const uint32_t tmem = *tmem_slot;

	if (warp < GW_LOADERS / 32) {
		/* ================= loaders / converters ================= */
		const bool vecA = (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
		const bool vecB = (ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(Bt) & 15) == 0);
		/* this thread's share of a chunk: NA of the 128*GW_CG float4 of A, NB of the 192*GW_CG of B; TWO chunks of loads in flight */
		constexpr int NA = GW_M * GW_CG / GW_LOADERS, NB = GW_N * GW_CG / GW_LOADERS;
		float4 ra[2][NA], rb[2][NB];
		auto fetch = [&](int slot, int kc) {
#pragma unroll
			for (int u = 0; u < NA; u++) ra[slot][u] = gw_load(A, lda, i0, a, kc * GW_K, b, vecA, tid + GW_LOADERS * u);
#pragma unroll
			for (int u = 0; u < NB; u++) rb[slot][u] = gw_load(Bt, ldb, j0, c, kc * GW_K, b, vecB, tid + GW_LOADERS * u);
		};
		auto convert = [&](int slot, int kc) {
			const int st = kc % GW_STAGES;
			unsigned char *sa = ring + (size_t)st * GW_STAGE;
			umma::mbar_wait(&empty[st], (uint32_t)(((kc / GW_STAGES) & 1) ^ 1));
#pragma unroll
			for (int u = 0; u < NA; u++) gw_split_store(sa, sa + GW_TILE_A, GW_LBO_A, tid + GW_LOADERS * u, ra[slot][u]);
#pragma unroll
			for (int u = 0; u < NB; u++) gw_split_store(sa + 2 * GW_TILE_A, sa + 2 * GW_TILE_A + GW_TILE_B, GW_LBO_B, tid + GW_LOADERS * u, rb[slot][u]);
		};
		auto publish = [&](int kc) {
			umma::fence_proxy_async();
			__syncwarp();
			if (lane == 0) umma::mbar_arrive(&full[kc % GW_STAGES]);
		};
		fetch(0, 0);
		if (nchunks > 1) fetch(1, 1);
		for (int kc = 0; kc < nchunks; kc += 2) {
			convert(0, kc);
			if (kc + 2 < nchunks) fetch(0, kc + 2);
			publish(kc);
			if (kc + 1 < nchunks) {
				convert(1, kc + 1);
				if (kc + 3 < nchunks) fetch(1, kc + 3);
				publish(kc + 1);
			}
		}
	} else if (warp == GW_LOADERS / 32) {
		/* ================= MMA issuer ================= */
		const uint32_t idesc = umma::idesc_tf32(GW_M, GW_N);
		for (int kc = 0; kc < nchunks; kc++) {
			const int st = kc % GW_STAGES, chain = kc / GW_DRAIN, buf = chain & 1;
			const bool chain_start = (kc % GW_DRAIN) == 0;
			if (chain_start) {
				umma::mbar_wait(&acc_empty[buf], (uint32_t)(((chain >> 1) & 1) ^ 1));
				umma::tc_fence_after();
			}
			umma::mbar_wait(&full[st], (uint32_t)((kc / GW_STAGES) & 1));
			umma::tc_fence_after();
			if (gw_elect()) {
				const uint32_t base = umma::smem_addr(ring + (size_t)st * GW_STAGE);
				const uint32_t a_hi = base, a_lo = base + GW_TILE_A, b_hi = base + 2 * GW_TILE_A, b_lo = b_hi + GW_TILE_B;
				const uint32_t d = tmem + (uint32_t)buf * GW_N;
#pragma unroll
				for (int ks = 0; ks < GW_K / 8; ks++) {
					const uint32_t oa = (uint32_t)ks * 2u * GW_LBO_A, ob = (uint32_t)ks * 2u * GW_LBO_B;
					const uint64_t dah = umma::smem_desc(a_hi + oa, GW_LBO_A, GW_SBO), dal = umma::smem_desc(a_lo + oa, GW_LBO_A, GW_SBO);
					const uint64_t dbh = umma::smem_desc(b_hi + ob, GW_LBO_B, GW_SBO), dbl = umma::smem_desc(b_lo + ob, GW_LBO_B, GW_SBO);
					umma::mma_tf32(d, dal, dbh, idesc, (!chain_start || ks) ? 1u : 0u); /* small terms first */
					umma::mma_tf32(d, dah, dbl, idesc, 1u);
					umma::mma_tf32(d, dah, dbh, idesc, 1u);
				}
				umma::mma_commit(&empty[st]);
				if ((kc % GW_DRAIN) == GW_DRAIN - 1 || kc == nchunks - 1) umma::mma_commit(&acc_full[buf]);
			}
			__syncwarp();
		}
	} else {
		/* ================= drain / epilogue ================= */
		const int dw = warp - GW_LOADERS / 32 - 1; /* 0..11 */
		const int q = warp % 4, cb = dw / 4;       /* TMEM lane quarter (fixed by the warp id), column block of 64 */
		const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
		float acc[64];
#pragma unroll
		for (int e = 0; e < 64; e++) acc[e] = 0.0f;
		for (int chain = 0; chain < nchains; chain++) {
			const int buf = chain & 1;
			umma::mbar_wait(&acc_full[buf], (uint32_t)((chain >> 1) & 1));
			umma::tc_fence_after();
#pragma unroll
			for (int cc = 0; cc < 64; cc += 16) {
				float v[16];
				umma::tmem_ld16(tmem + lane_addr + (uint32_t)(buf * GW_N + cb * 64 + cc), v);
#pragma unroll
				for (int e = 0; e < 16; e++) acc[cc + e] = __fadd_rn(acc[cc + e], v[e]);
			}
			umma::tc_fence_before();
			__syncwarp();
			if (lane == 0) umma::mbar_arrive(&acc_empty[buf]);
		}
		const int gi = i0 + 32 * q + lane;
		const int jb = j0 + cb * 64;
		if (gi < a) {
			/* streaming stores: C (268 MB at C3) must not push the operands out of L2; 16 bytes per store where the row allows it */
			float *dst = C + (size_t)gi * ldc + jb;
			const bool whole = !sym || jb >= i0 + GW_M; /* no element of this block lies below the diagonal */
			if (whole && (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && jb + 64 <= c) {
#pragma unroll
				for (int e = 0; e < 64; e += 4) __stcs(reinterpret_cast<float4 *>(dst + e), make_float4(acc[e], acc[e + 1], acc[e + 2], acc[e + 3]));
			} else {
#pragma unroll
				for (int e = 0; e < 64; e++)
					if (jb + e < c && (whole || jb + e >= gi)) __stcs(dst + e, acc[e]);
			}
		}
		if (sym) {
			/* the mirror image: lane = row of the tile, so C_ji of one column j is 32 consecutive floats over the warp */
#pragma unroll
			for (int e = 0; e < 64; e++)
				if (gi < a && jb + e < c && jb + e > gi) __stcs(C + (size_t)(jb + e) * ldc + gi, acc[e]);
		}
	}
	umma::tc_fence_before();
	__syncthreads();
	if (warp == GW_LOADERS / 32) umma::tmem_dealloc(tmem, 512);
}

/* worth it from a few tiles per SM on; small products stay on the simpler kernel */
int pqp_gemm_umma_ws_wanted(int a, int b, int c)
{
	const char *e = pqp_env("PQP_GEMM_WS"); /* PQP_GEMM_WS=0 keeps round 1's kernel */
	return a >= 256 && c >= 192 && b >= 64 && !(e && atoi(e) == 0);
}

cudaError_t pqp_launch_gemm_umma_ws(float *C, int ldc, const float *A, int lda, const float *Bt, int ldb, int a, int b, int c,
				    const unsigned *sym_bad, cudaStream_t s)
{
	const size_t smem = (size_t)GW_STAGES * GW_STAGE + 128;
	cudaError_t e = cudaFuncSetAttribute(gemm_3xtf32_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	dim3 grid((c + GW_N - 1) / GW_N, (a + GW_M - 1) / GW_M);
	gemm_3xtf32_ws_kernel<<<grid, GW_THREADS, smem, s>>>(C, ldc, A, lda, Bt, ldb, a, b, c, sym_bad);
	return cudaGetLastError();
}
