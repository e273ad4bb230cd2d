/*
 * pqp_batched_umma.cu -- B problems sharing one Hessian on the 5th-gen tensor cores (sm_100a).
 *
 * The PQP update for many MPC states at once (PQP_CPU.c:603-618 + 590-596 applied to B right-hand sides)
 *      NUM = (Q^- + theta) Y + F^-,   DEN = (Q^+ + theta) Y + F^+,   Y <- NUM/DEN o Y,      Y in R^{N x B}
 * is two N x N x B contractions per iteration (4*N^2*B flop), so it runs as
 * tcgen05.mma.kind::tf32 with fp32 accumulators in TMEM and 3xTF32 operand splitting:
 *      q*y ~= q_hi*y_hi + q_hi*y_lo + q_lo*y_hi        (all operands >= 0 here: no cancellation inside the sums)
 *
 * Mapping (one CTA = 32 problems for the WHOLE solve; no grid barrier, no host round trip, no HBM traffic in the loop):
 *   M = 128 rows of Q (an "M tile"), K = the columns of Q, N = problems.
 *   A operand  = pre-split, pre-tiled (Q^- + theta | Q^+ + theta) x (hi | lo), built once at setup
 *                (x-independent, 3.9 MB at N=480, L2 resident), streamed through a 4-stage shared-memory ring
 *                by 1-D bulk async copies (UBLKCP), MULTICAST to the CTAs of a cluster so L2 is read once per
 *                cluster instead of once per CTA.
 *   B operand  = the CTA's Y tile, [y_hi | y_lo], K-major in shared memory for the whole solve; rewritten in
 *                place by the epilogue every iteration.
 *   D          = 2 x 4 accumulator tiles of 128 lanes x 64 columns = all 512 TMEM columns:
 *                columns [0,32) += q_hi*y_hi (+ q_lo*y_hi), columns [32,64) = q_hi*y_lo  (one N=64 MMA over
 *                [y_hi|y_lo] and one N=32 MMA per K step: A_hi is read from shared memory once, not twice).
 *   warp 0     producer (bulk copies + mbarriers), warp 1 MMA issuer (single thread) + TMEM allocator,
 *   warps 2-9  epilogue: tcgen05.ld the accumulators, num/den + F-/F+, IEEE division, y <- (num/den)*y,
 *              re-split into tf32 hi/lo and store straight into the B operand layout (bank-conflict free:
 *              the K-direction stride of the B tile is padded to 144 B).
 *
 * Bound: shared-memory operand bandwidth (11 KB of operands per 48 tensor-cycles at N=32 problems per SM);
 * see DESIGN.md 3.4.  Requires N <= 512.
 */
#include "pqp_internal.h"
#include "pqp_umma.cuh"

#include <stdlib.h>
#include <string.h>

#define BU_NB 32            /* problems per CTA */
#define BU_KC 16            /* k per staged chunk (2 MMA K-steps) */
#define BU_STAGES 4
#define BU_TILE 8192u       /* one 128 x 16 tf32 tile: 4 column groups x 2048 B */
#define BU_CHUNK (2u * BU_TILE) /* hi + lo */
#define BU_A_LBO 2048u
#define BU_A_SBO 128u
#define BU_B_LBO 144u       /* 128 B core matrix + 16 B pad: epilogue stores are conflict free */
#define BU_EPI_WARPS 8
#define BU_THREADS (64 + 32 * BU_EPI_WARPS)

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

struct BuParams {
	const unsigned char *Atiles; /* [2 mats][MT][NKC][hi|lo][8192 B] */
	const float *Fd;             /* [B x N] */
	float *Y;                    /* [B x N] in: y0, out: y_K */
	int N, B, iters;
	int MT, NKC;                 /* M tiles of 128 rows, K chunks of 16 */
	uint32_t b_sbo;              /* byte stride between 8-problem groups of the B tile = (Kpad/4)*144 */
	int dbg;                     /* experiment switches (PQP_UMMA_DBG): 1 skip lo MMA, 2 skip all MMAs, 4 skip epilogue math */
};

/*
 * shared memory: ring [STAGES][BU_CHUNK] | B_hi [4 groups x b_sbo] | B_lo [same] | barriers
 */
__global__ void __launch_bounds__(BU_THREADS, 1) batched_umma_kernel(const BuParams p)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
	const uint32_t CS = cluster_nctarank(), crank = cluster_ctarank();
	const uint16_t cmask = (uint16_t)((1u << CS) - 1u);

	unsigned char *ring = smem_raw;
	unsigned char *Bhi = ring + BU_STAGES * BU_CHUNK;
	const uint32_t b_tile_bytes = 4u * p.b_sbo;
	unsigned char *Blo = Bhi + b_tile_bytes;
	uint64_t *full = reinterpret_cast<uint64_t *>(Blo + b_tile_bytes);
	uint64_t *empty = full + BU_STAGES;
	uint64_t *tmem_full = empty + BU_STAGES;
	uint64_t *b_ready = tmem_full + 1;
	uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(b_ready + 1);

	const int N = p.N, MT = p.MT, NKC = p.NKC;
	const int chunks_per_iter = 2 * MT * NKC;
	const long long total_chunks = (long long)chunks_per_iter * p.iters;
	const int b0 = blockIdx.x * BU_NB;
	const uint32_t tmem_cols = (2 * MT * 64 <= 128) ? 128u : ((2 * MT * 64 <= 256) ? 256u : 512u);

	if (tid == 0) {
		for (int s = 0; s < BU_STAGES; s++) {
			umma::mbar_init(&full[s], 1);
			umma::mbar_init(&empty[s], CS);
		}
		umma::mbar_init(tmem_full, 1);
		umma::mbar_init(b_ready, 32 * BU_EPI_WARPS);
		umma::mbar_fence_init();
	}
	if (warp == 1) umma::tmem_alloc(tmem_slot, tmem_cols);
	umma::tc_fence_before();
	__syncthreads();
	if (CS > 1) cluster_sync_all(); /* every CTA's barriers exist before anyone multicasts into them */
	umma::tc_fence_after();
	const uint32_t tmem = *tmem_slot;

	if (warp == 0) {
		/* ================= producer: A chunks through the ring ================= */
		if (lane == 0) {
			for (long long c = 0; c < total_chunks; c++) {
				const int s = (int)(c % BU_STAGES);
				const uint32_t ph = (uint32_t)((c / BU_STAGES) & 1);
				umma::mbar_wait(&empty[s], ph ^ 1u);
				umma::mbar_arrive_expect_tx(&full[s], BU_CHUNK);
				const unsigned char *src = p.Atiles + (size_t)(c % chunks_per_iter) * BU_CHUNK;
				if (CS == 1)
					bulk_g2s_plain(ring + (size_t)s * BU_CHUNK, src, BU_CHUNK, &full[s]);
				else if ((uint32_t)(c % CS) == crank)
					bulk_g2s_mcast(ring + (size_t)s * BU_CHUNK, src, BU_CHUNK, &full[s], cmask);
			}
		}
	} else if (warp == 1) {
		/* ================= MMA issuer ================= */
		if (lane == 0) {
			const uint32_t idesc64 = umma::idesc_tf32(128, 64), idesc32 = umma::idesc_tf32(128, 32);
			const uint32_t bhi_addr = umma::smem_addr(Bhi);
			long long c = 0;
			for (int it = 0; it < p.iters; it++) {
				umma::mbar_wait(b_ready, (uint32_t)(it & 1)); /* Y tile of this iteration is in place */
				umma::tc_fence_after();
				for (int mat = 0; mat < 2; mat++)
					for (int mt = 0; mt < MT; mt++) {
						const uint32_t d = tmem + (uint32_t)((mat * MT + mt) * 64);
						for (int kc = 0; kc < NKC; kc++, c++) {
							const int s = (int)(c % BU_STAGES);
							umma::mbar_wait(&full[s], (uint32_t)((c / BU_STAGES) & 1));
							umma::tc_fence_after();
							const uint32_t a_hi = umma::smem_addr(ring + (size_t)s * BU_CHUNK), a_lo = a_hi + BU_TILE;
#pragma unroll
							for (int ks = 0; ks < 2; ks++) {
								const uint32_t ao = (uint32_t)ks * 2u * BU_A_LBO;
								const uint32_t bo = (uint32_t)(kc * 2 + ks) * 2u * BU_B_LBO;
								const uint64_t db = umma::smem_desc(bhi_addr + bo, BU_B_LBO, p.b_sbo);
								/* [q_hi*y_hi | q_hi*y_lo] then + q_lo*y_hi into the first 32 columns */
								if (!(p.dbg & 2)) umma::mma_tf32(d, umma::smem_desc(a_hi + ao, BU_A_LBO, BU_A_SBO), db, idesc64, (kc | ks) ? 1u : 0u);
								if (!(p.dbg & 3)) umma::mma_tf32(d, umma::smem_desc(a_lo + ao, BU_A_LBO, BU_A_SBO), db, idesc32, 1u);
							}
							if (CS == 1) umma::mma_commit(&empty[s]);
							else mma_commit_mcast(&empty[s], cmask); /* the stage is free once EVERY CTA of the cluster has read it */
						}
					}
				umma::mma_commit(tmem_full);
			}
		}
	} else {
		/* ================= epilogue warps ================= */
		const int ew = warp - 2;
		const int q = warp % 4;       /* TMEM lane quarter this warp may touch */
		const int half = ew / 4;      /* which 16 of the 32 problems */
		const int pb = half * 16;
		const uint32_t lane_addr = (uint32_t)(32 * q) << 16;

		/* iteration 0: load y0, split, fill the B operand (zero the K padding) */
		{
			const int kpad = NKC * BU_KC;
			for (int e = tid - 64; e < BU_NB * kpad; e += 32 * BU_EPI_WARPS) {
				const int b = e / kpad, k = e % kpad;
				float y = 0.0f;
				if (k < N) y = (b0 + b < p.B) ? p.Y[(size_t)(b0 + b) * N + k] : 1.0f;
				float hi, lo;
				umma::tf32_split(y, hi, lo);
				const uint32_t off = (uint32_t)(b / 8) * p.b_sbo + (uint32_t)(k / 4) * BU_B_LBO + (uint32_t)(b % 8) * 16u + (uint32_t)(k % 4) * 4u;
				*reinterpret_cast<float *>(Bhi + off) = hi;
				*reinterpret_cast<float *>(Blo + off) = lo;
			}
			umma::fence_proxy_async();
			umma::mbar_arrive(b_ready);
		}

		for (int it = 0; it < p.iters; it++) {
			umma::mbar_wait(tmem_full, (uint32_t)(it & 1));
			umma::tc_fence_after();
			for (int mt = 0; mt < MT; mt++) {
				const int i = mt * 128 + 32 * q + lane; /* the row of Q (= dual index) this thread finishes */
				float nhh[16], nhl[16], dhh[16], dhl[16];
				const uint32_t cn = tmem + lane_addr + (uint32_t)((0 * MT + mt) * 64 + pb);
				const uint32_t cd = tmem + lane_addr + (uint32_t)((1 * MT + mt) * 64 + pb);
				umma::tmem_ld16(cn, nhh);
				umma::tmem_ld16(cn + 32, nhl);
				umma::tmem_ld16(cd, dhh);
				umma::tmem_ld16(cd + 32, dhl);
				if (i < N && !(p.dbg & 4)) {
#pragma unroll
					for (int j = 0; j < 16; j++) {
						const int b = pb + j;
						const uint32_t off = (uint32_t)(b / 8) * p.b_sbo + (uint32_t)(i / 4) * BU_B_LBO + (uint32_t)(b % 8) * 16u + (uint32_t)(i % 4) * 4u;
						const float yo = *reinterpret_cast<const float *>(Bhi + off) + *reinterpret_cast<const float *>(Blo + off);
						const float fd = (b0 + b < p.B) ? __ldg(p.Fd + (size_t)(b0 + b) * N + i) : 1.0f;
						const float num = (nhh[j] + nhl[j]) + fmaxf(-fd, 0.0f);
						const float den = (dhh[j] + dhl[j]) + fmaxf(fd, 0.0f);
						const float yn = __fdiv_rn(num, den) * yo;
						float hi, lo;
						umma::tf32_split(yn, hi, lo);
						*reinterpret_cast<float *>(Bhi + off) = hi;
						*reinterpret_cast<float *>(Blo + off) = lo;
					}
				}
			}
			umma::fence_proxy_async();  /* the new Y tile must be visible to the tensor core's reads */
			umma::tc_fence_before();    /* and our TMEM reads ordered before the next MMAs overwrite D */
			umma::mbar_arrive(b_ready);
		}
		/* all epilogue threads are past their last B writes once b_ready's final phase completes */
		umma::mbar_wait(b_ready, (uint32_t)(p.iters & 1));
		for (int e = tid - 64; e < BU_NB * N; e += 32 * BU_EPI_WARPS) {
			const int b = e / N, k = e % N;
			if (b0 + b < p.B) {
				const uint32_t off = (uint32_t)(b / 8) * p.b_sbo + (uint32_t)(k / 4) * BU_B_LBO + (uint32_t)(b % 8) * 16u + (uint32_t)(k % 4) * 4u;
				p.Y[(size_t)(b0 + b) * N + k] = *reinterpret_cast<const float *>(Bhi + off) + *reinterpret_cast<const float *>(Blo + off);
			}
		}
	}
	umma::tc_fence_before();
	__syncthreads();
	if (CS > 1) cluster_sync_all(); /* nobody leaves while a peer may still multicast into / arrive on this CTA */
	if (warp == 1) umma::tmem_dealloc(tmem, tmem_cols);
}

/* builds the pre-split, pre-tiled A operand from the signed Qd and theta (x-independent, once per handle) */
__global__ void build_umma_tiles_kernel(unsigned char *__restrict__ tiles, const float *__restrict__ Q, int ldq,
					 const float *__restrict__ theta, int N, int MT, int NKC)
{
	/* one thread per (mat, row i of the padded 128*MT, column group of 4) */
	const int kg_per_row = NKC * 4;
	const long long total = 2LL * MT * 128 * kg_per_row;
	for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
		const int kg = (int)(t % kg_per_row);
		const int i = (int)((t / kg_per_row) % (MT * 128));
		const int mat = (int)(t / ((long long)kg_per_row * MT * 128));
		float v[4];
#pragma unroll
		for (int e = 0; e < 4; e++) {
			const int k = kg * 4 + e;
			float x = 0.0f;
			if (i < N && k < N) {
				const float qv = Q[(size_t)i * ldq + k];
				x = mat == 0 ? fmaxf(-qv, 0.0f) : fmaxf(qv, 0.0f); /* mat 0: Q^- (numerator), mat 1: Q^+ (denominator) */
				if (i == k) x += theta[i];                         /* theta on the diagonal of both, PQP_CPU.c:527,536 */
			}
			v[e] = x;
		}
		float4 h, l;
		umma::tf32_split(v[0], h.x, l.x);
		umma::tf32_split(v[1], h.y, l.y);
		umma::tf32_split(v[2], h.z, l.z);
		umma::tf32_split(v[3], h.w, l.w);
		const int mt = i / 128, r = i % 128, kc = kg / 4, g = kg % 4;
		unsigned char *blk = tiles + ((size_t)(mat * MT + mt) * NKC + kc) * BU_CHUNK;
		const uint32_t off = (uint32_t)g * BU_A_LBO + (uint32_t)(r / 8) * BU_A_SBO + (uint32_t)(r % 8) * 16u;
		*reinterpret_cast<float4 *>(blk + off) = h;
		*reinterpret_cast<float4 *>(blk + BU_TILE + off) = l;
	}
}

int pqp_batched_umma_supported(int N) { return N >= 16 && N <= 512; }

size_t pqp_batched_umma_tiles_bytes(int N)
{
	const int MT = (N + 127) / 128, NKC = (N + BU_KC - 1) / BU_KC;
	return (size_t)2 * MT * NKC * BU_CHUNK;
}

cudaError_t pqp_launch_build_umma_tiles(void *tiles, const float *Q, int ldq, const float *theta, int N, cudaStream_t s)
{
	const int MT = (N + 127) / 128, NKC = (N + BU_KC - 1) / BU_KC;
	build_umma_tiles_kernel<<<296, 256, 0, s>>>(reinterpret_cast<unsigned char *>(tiles), Q, ldq, theta, N, MT, NKC);
	return cudaGetLastError();
}

cudaError_t pqp_launch_batched_umma(const void *tiles, int N, int B, const float *Fd, float *Y, int iters, int cluster, cudaStream_t s)
{
	BuParams p;
	p.Atiles = reinterpret_cast<const unsigned char *>(tiles);
	p.Fd = Fd;
	p.Y = Y;
	p.N = N;
	p.B = B;
	p.iters = iters;
	p.MT = (N + 127) / 128;
	p.NKC = (N + BU_KC - 1) / BU_KC;
	p.b_sbo = (uint32_t)(p.NKC * BU_KC / 4) * BU_B_LBO;
	p.dbg = pqp_env("PQP_UMMA_DBG") ? atoi(pqp_env("PQP_UMMA_DBG")) : 0;
	const size_t smem = (size_t)BU_STAGES * BU_CHUNK + 2 * 4 * (size_t)p.b_sbo + (2 * BU_STAGES + 2) * sizeof(uint64_t) + 16;
	cudaError_t e = cudaFuncSetAttribute(batched_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	if (cluster < 1) cluster = 1;
	int tiles_n = (B + BU_NB - 1) / BU_NB;
	tiles_n = (tiles_n + cluster - 1) / cluster * cluster;
	cudaLaunchConfig_t cfg;
	memset(&cfg, 0, sizeof cfg);
	cfg.gridDim = dim3(tiles_n);
	cfg.blockDim = dim3(BU_THREADS);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = s;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = cluster;
	attr[0].val.clusterDim.y = 1;
	attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr;
	cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, batched_umma_kernel, p);
}
