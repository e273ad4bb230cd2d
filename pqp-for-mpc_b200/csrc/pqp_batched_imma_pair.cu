/*
 * pqp_batched_imma_pair.cu -- the int8 digit-plane batched PQP loop (see pqp_batched_imma.cu for the arithmetic) with the rows
 * of Q split over a PAIR of CTAs (a cluster of two SMs) that share 64 problems.
 *
 * Why.  The single-CTA kernel is bound by shared-memory bandwidth, not by the tensor pipe: per K step it moves 12 KB of A
 * digit planes in (bulk copy), 12 KB of them out again (MMA operand fetch) and 6 KB of Y planes, 30 KB at 128 B/clk = 234
 * cycles against 96 cycles of int8 math (in-kernel profile: 243 cycles per K step).  The A traffic does not depend on how many
 * problems ride on it, so a tile of 64 problems costs 36 KB per K step instead of 2 x 30 KB -- but 64 problems per CTA would
 * leave half the SMs idle at the batch size of config C4 (4096 states = 64 tiles).  Hence the pair: both CTAs keep the digit
 * planes of all 64 problems, each streams and multiplies only ITS M tiles (rank 0: tiles 0, 2; rank 1: tiles 1, 3), i.e. half
 * the K steps per update, and after every update each epilogue thread stores the new digits of its rows into both CTAs'
 * plane buffers (local store + st.shared::cluster).  Arithmetic, digit planes and result are those of the single-CTA kernel,
 * bit for bit (same numpy model, tests/imma_model.py).
 *
 * Cross-CTA protocol per update (all mbarriers live in each CTA's own shared memory; remote arrivals use mapa addresses):
 *   1. units (matrix, own M tile): tensor pipe -> TMEM -> epilogue, exactly as in the single-CTA kernel (local barriers);
 *      each finished tile publishes its per-problem maxima into BOTH CTAs' max slots (atomicMax + red.shared::cluster);
 *   2. every epilogue thread arrives on both CTAs' `allmax` barrier (2 x 512 arrivals).  Its completion means: the exact maxima of
 *      all rows are known, AND both CTAs' MMAs of this update have completed (a thread arrives only after its last tmem_full),
 *      so the single plane buffer of either CTA may be overwritten;
 *   3. scales from the exact maxima, digits of the thread's rows -> both plane buffers, fence.proxy.async, arrive on both CTAs'
 *      `b_ready` (2 x 512 arrivals), on which each CTA's MMA issuer waits before the next update.
 */
#include "pqp_imma.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define BP_NB 64        /* problems per CTA pair */
#define BP_PW 16        /* problems per epilogue thread */
#define BP_EW 16        /* epilogue warps: 4 TMEM lane quarters x 4 problem groups */
#define BP_ETHREADS (32 * BP_EW)
#define BP_THREADS (64 + BP_ETHREADS)
#define BP_UNIT_COLS (3u * BP_NB)
#define BP_MAX_MY 2     /* M tiles per CTA (N <= 512) */


/*
 * shared memory: ring [stages][ksc*BI_CHUNK] | planes [3][4][Kpad/8][8][16 B] | smax[2][64] | iscale[64] | barriers
 */
__global__ void __launch_bounds__(BP_THREADS, 1) batched_imma_pair_kernel(const BiParams p)
{
	constexpr int NB = BP_NB, PW = BP_PW;
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
	const uint32_t rank = cluster_ctarank(), peer = rank ^ 1u;

	const int N = p.N, MT = p.MT, NKS = p.NKS;
	const int Kpad = NKS * 32;
	unsigned char *ring = smem_raw;
	const uint32_t stage_bytes = (uint32_t)p.ksc * BI_CHUNK;
	unsigned char *Bpl = ring + (size_t)p.stages * stage_bytes;
	const uint32_t plane_bytes = (uint32_t)(NB / 16) * p.b_sbo;
	const uint32_t pbuf_bytes = 3u * plane_bytes;
	uint32_t *smax = reinterpret_cast<uint32_t *>(Bpl + pbuf_bytes);
	float *iscale = reinterpret_cast<float *>(smax + 2 * NB);
	uint64_t *full = reinterpret_cast<uint64_t *>(iscale + NB);
	uint64_t *empty = full + p.stages;
	uint64_t *tmem_full = empty + p.stages; /* [2] */
	uint64_t *tmem_empty = tmem_full + 2;   /* [2] */
	uint64_t *b_ready = tmem_empty + 2;     /* 2 x 16 arrivals: the epilogue warps of both CTAs */
	uint64_t *allmax = b_ready + 1;         /* 2 x 16 arrivals */
	uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(allmax + 1);

	const int n_my = (MT - (int)rank + 1) / 2; /* own M tiles: rank, rank + 2 */
	const int units_per_iter = 2 * n_my;
	const int chunks_per_unit = NKS / p.ksc;
	const int b0 = (int)(blockIdx.x / 2) * NB;
	const bool prof_on = (p.dbg & 8) && p.prof && blockIdx.x == 0;
	long long prof_acc[12] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };

	if (tid == 0) {
		for (int s = 0; s < p.stages; s++) {
			umma::mbar_init(&full[s], 1);
			umma::mbar_init(&empty[s], 1);
		}
		for (int i = 0; i < 2; i++) {
			umma::mbar_init(&tmem_full[i], 1);
			umma::mbar_init(&tmem_empty[i], BP_EW); /* one arrival per epilogue warp */
		}
		umma::mbar_init(b_ready, 2 * BP_EW); /* one arrival per epilogue warp of either CTA */
		umma::mbar_init(allmax, 2 * BP_EW);
		umma::mbar_fence_init();
	}
	if (tid < 2 * NB) smax[tid] = 0u;
	for (uint32_t i = tid; i < pbuf_bytes / 16u; i += blockDim.x) reinterpret_cast<uint4 *>(Bpl)[i] = make_uint4(0u, 0u, 0u, 0u);
	umma::fence_proxy_async();
	if (warp == 1) umma::tmem_alloc(tmem_slot, 512);
	umma::tc_fence_before();
	__syncthreads();
	cluster_sync_all(); /* both CTAs' barriers, max slots and zeroed planes exist before anyone touches the peer's */
	umma::tc_fence_after();
	const uint32_t tmem = *tmem_slot;

	if (warp == 0) {
		/* ================= producer: the A chunks of the own M tiles ================= */
		int st = 0;
		uint32_t ph = 0;
		for (int it = 0; it < p.iters; it++) {
			for (int lt = 0; lt < n_my; lt++) {
				const int mt = (int)rank + 2 * lt;
				for (int c = 0; c < 2 * chunks_per_unit; c++) { /* matrix 0 then matrix 1: contiguous in the tile array */
					PROF_T(tw);
					umma::mbar_wait(&empty[st], ph ^ 1u);
					PROF_ADD(PROF_PROD_WAIT_EMPTY, tw);
					if (elect_one()) {
						umma::mbar_arrive_expect_tx(&full[st], stage_bytes);
						const unsigned char *src = p.Atiles + ((size_t)mt * 2 * NKS) * BI_CHUNK + (size_t)c * stage_bytes;
						bulk_g2s_plain(ring + (size_t)st * stage_bytes, src, stage_bytes, &full[st]);
					}
					__syncwarp();
					if (++st == p.stages) { st = 0; ph ^= 1u; }
				}
			}
		}
		if (prof_on && lane == 0) p.prof[PROF_PROD_WAIT_EMPTY] = prof_acc[PROF_PROD_WAIT_EMPTY];
	} else if (warp == 1) {
		/* ================= MMA issuer ================= */
		const uint32_t id3 = idesc_i8(128, 3 * NB), id2 = idesc_i8(128, 2 * NB), id1 = idesc_i8(128, NB);
		const uint64_t a_desc0 = umma::smem_desc(umma::smem_addr(ring), BI_A_LBO, BI_A_SBO);
		const uint64_t b_desc0 = umma::smem_desc(umma::smem_addr(Bpl), BI_B_LBO, p.b_sbo);
		int st = 0;
		uint32_t ph = 0;
		long long unit = 0;
		PROF_T(tm0);
		for (int it = 0; it < p.iters; it++) {
			PROF_T(tb);
			mbar_wait_cluster(b_ready, (uint32_t)(it & 1)); /* both CTAs' digits of this update are in place */
			PROF_ADD(PROF_MMA_WAIT_BREADY, tb);
			umma::tc_fence_after();
			for (int u = 0; u < units_per_iter; u++, unit++) {
				const int buf = u & 1;
				PROF_T(te);
				umma::mbar_wait(&tmem_empty[buf], (uint32_t)(((unit >> 1) & 1) ^ 1));
				PROF_ADD(PROF_MMA_WAIT_TMEM, te);
				umma::tc_fence_after();
				const uint32_t d = tmem + (uint32_t)buf * BP_UNIT_COLS;
				for (int ch = 0; ch < chunks_per_unit; ch++) {
					PROF_T(tf);
					umma::mbar_wait(&full[st], ph);
					PROF_ADD(PROF_MMA_WAIT_FULL, tf);
					umma::tc_fence_after();
					if (elect_one()) {
						const uint64_t da = a_desc0 + (uint64_t)((uint32_t)st * (stage_bytes >> 4));
						const uint64_t db = b_desc0 + (uint64_t)((uint32_t)(ch * p.ksc) * (4u * BI_B_LBO >> 4));
						if (!(p.dbg & 2)) {
							if (p.ksc == 3) mma_i8_step3(d, d + NB, d + 2 * NB, da, db, id3, id2, id1, ch ? 1u : 0u);
							else mma_i8_step(d, d + NB, d + 2 * NB, da, db, id3, id2, id1, ch ? 1u : 0u);
						}
						umma::mma_commit(&empty[st]);
					}
					__syncwarp();
					if (++st == p.stages) { st = 0; ph ^= 1u; }
				}
				if (elect_one()) umma::mma_commit(&tmem_full[buf]);
				__syncwarp();
			}
		}
		PROF_ADD(PROF_MMA_TOTAL, tm0);
		if (prof_on && lane == 0)
			for (int i = PROF_MMA_TOTAL; i <= PROF_MMA_WAIT_FULL; i++) p.prof[i] = prof_acc[i];
	} else {
		/* ================= epilogue warps ================= */
		const int et = tid - 64;
		const int ew = warp - 2;
		const int q = warp % 4;  /* TMEM lane quarter this warp may touch */
		const int cg = ew / 4;   /* which 16 of the 64 problems */
		const int pb = cg * PW;
		const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
		const int r = 32 * q + lane;

		/* the peer's copies of the exchange targets */
		const uint32_t planes_l = umma::smem_addr(Bpl), planes_r = map_peer(planes_l, peer);
		const uint32_t smax_r = map_peer(umma::smem_addr(smax), peer);
		const uint32_t allmax_l = map_peer(umma::smem_addr(allmax), rank), allmax_r = map_peer(umma::smem_addr(allmax), peer);
		const uint32_t bready_l = map_peer(umma::smem_addr(b_ready), rank), bready_r = map_peer(umma::smem_addr(b_ready), peer);

		float y[BP_MAX_MY][PW]; /* fp32 master copy: rows (rank + 2*lt)*128 + r, problems pb..pb+15 */
		float4 rc[BP_MAX_MY];

		/* per-problem maxima of one tile -> both CTAs' slots */
		auto publish_tile_max = [&](const float(&yv)[PW], int parity) {
#pragma unroll
			for (int j = 0; j < PW; j++) {
				const uint32_t wm = __reduce_max_sync(0xffffffffu, __float_as_uint(yv[j]) & 0x7fffffffu);
				if (lane == j) {
					atomicMax(smax + parity * NB + pb + j, wm);
					red_max_cluster(smax_r + (uint32_t)(parity * NB + pb + j) * 4u, wm);
				}
			}
		};
		/* exchange: exact maxima -> scales -> digits of the own rows into both plane buffers */
		/* exchange: exact maxima -> scales -> digits of the own rows into both plane buffers */
		auto exchange = [&](int parity, uint32_t phase) {
			__syncwarp(); /* every lane's max publications precede lane 0's (release) arrivals */
			if (lane == 0) { /* one release fence for both arrivals */
				fence_release_cluster();
				mbar_arrive_cluster_relaxed(allmax_l);
				mbar_arrive_cluster_relaxed(allmax_r);
			}
			PROF_T(tx0);
			mbar_wait_cluster(allmax, phase);
			PROF_ADD(8, tx0);
			PROF_T(tx1);
			float sc[PW];
#pragma unroll
			for (int j = 0; j < PW; j++) {
				const uint32_t mx = smax[parity * NB + pb + j];
				float isc;
				problem_scales(mx, sc[j], isc);
				if (q == 0 && lane == j) iscale[pb + j] = isc;
				if ((mx >> 23) >= 255u) {
#pragma unroll
					for (int lt = 0; lt < BP_MAX_MY; lt++) y[lt][j] = __uint_as_float(0x7fc00000u);
				}
			}
#pragma unroll
			for (int lt = 0; lt < BP_MAX_MY; lt++) {
				const int i = ((int)rank + 2 * lt) * 128 + r;
				if (lt < n_my && i < Kpad) {
					uint32_t w0[4], w1[4], w2[4];
#pragma unroll
					for (int j = 0; j < PW; j += 4)
						digits4(__float2int_rn(y[lt][j] * sc[j]), __float2int_rn(y[lt][j + 1] * sc[j + 1]),
							__float2int_rn(y[lt][j + 2] * sc[j + 2]), __float2int_rn(y[lt][j + 3] * sc[j + 3]), w0[j >> 2], w1[j >> 2], w2[j >> 2]);
					const uint32_t off = (uint32_t)cg * p.b_sbo + (uint32_t)(i >> 3) * BI_B_LBO + (uint32_t)(i & 7) * 16u;
					const uint4 v0 = make_uint4(w0[0], w0[1], w0[2], w0[3]), v1 = make_uint4(w1[0], w1[1], w1[2], w1[3]),
						    v2 = make_uint4(w2[0], w2[1], w2[2], w2[3]);
					*reinterpret_cast<uint4 *>(Bpl + off) = v0;
					*reinterpret_cast<uint4 *>(Bpl + off + plane_bytes) = v1;
					*reinterpret_cast<uint4 *>(Bpl + off + 2u * plane_bytes) = v2;
					if (!(p.dbg & 16)) { /* experiment switch: timing without the remote stores (results are then wrong) */
						st_cluster_v4(planes_r + off, v0);
						st_cluster_v4(planes_r + off + plane_bytes, v1);
						st_cluster_v4(planes_r + off + 2u * plane_bytes, v2);
					}
				}
			}
			PROF_ADD(9, tx1);
			PROF_T(tx2);
			fence_proxy_async_all(); /* local and remote digits must be visible to both tensor cores */
			PROF_ADD(10, tx2);
			__syncwarp();
			if (lane == 0) {
				fence_release_cluster();
				mbar_arrive_cluster_relaxed(bready_l);
				mbar_arrive_cluster_relaxed(bready_r);
			}
		};

		/* update 0: y0 and the row constants of the own tiles */
#pragma unroll
		for (int lt = 0; lt < BP_MAX_MY; lt++) {
			const int i = ((int)rank + 2 * lt) * 128 + r;
			rc[lt] = make_float4(0.f, 0.f, 0.f, 0.f);
			if (lt < n_my) rc[lt] = __ldg(p.rowc + i);
#pragma unroll
			for (int j = 0; j < PW; j++) {
				float v = 0.0f;
				if (lt < n_my && i < N && b0 + pb + j < p.B) v = p.Y[(size_t)(b0 + pb + j) * N + i];
				y[lt][j] = v;
			}
		}
		/* Fd of the own rows never changes: park it in the 128 TMEM columns the two accumulator buffers leave free (columns
		 * 384 + 64*lt + problem), so the loop reads it with tcgen05.ld instead of an L2 round trip per update */
#pragma unroll
		for (int lt = 0; lt < BP_MAX_MY; lt++) {
			const int i = ((int)rank + 2 * lt) * 128 + r;
			float fv[PW];
#pragma unroll
			for (int j = 0; j < PW; j++) fv[j] = (lt < n_my && i < N && b0 + pb + j < p.B) ? __ldg(p.Fd + (size_t)(b0 + pb + j) * N + i) : 1.0f;
			tmem_st16_f32(tmem + lane_addr + 2u * BP_UNIT_COLS + (uint32_t)(lt * NB + pb), fv);
		}
		umma::tc_fence_before();
#pragma unroll
		for (int lt = 0; lt < BP_MAX_MY; lt++)
			if (lt < n_my) publish_tile_max(y[lt], 1);
		exchange(1, 0u);

		long long pair = 0;
		PROF_T(te0);
		for (int it = 0; it < p.iters; it++) {
			const int par_out = it & 1, par_in = par_out ^ 1;
#pragma unroll
			for (int lt = 0; lt < BP_MAX_MY; lt++) {
				if (lt < n_my) {
					const int i = ((int)rank + 2 * lt) * 128 + r;
					float sn[PW];
#pragma unroll
					for (int mat = 0; mat < 2; mat++) {
						PROF_T(tw);
						umma::mbar_wait(&tmem_full[mat], (uint32_t)(pair & 1));
						PROF_ADD(PROF_EPI_WAIT_TMEM, tw);
						umma::tc_fence_after();
						if (lt == 0 && mat == 0) {
							/* every local thread is past its reads of the maxima of the previous exchange: clear those slots for the
							 * update after this one (the peer cannot publish into them before this CTA has arrived on b_ready again) */
							named_bar_sync(2, BP_ETHREADS);
							if (et < NB) smax[par_in * NB + et] = 0u;
						}
						const float rs = mat == 0 ? rc[lt].z : rc[lt].w;
						const float dg = mat == 0 ? rc[lt].x : rc[lt].y;
#pragma unroll
						for (int h = 0; h < 2; h++) { /* two halves of 8 problems: keeps the TMEM staging registers at 24 */
							const uint32_t col = tmem + lane_addr + (uint32_t)mat * BP_UNIT_COLS + (uint32_t)(pb + 8 * h);
							int w0[8], w1[8], w2[8], fdb[8];
							tmem_ld8_i32(col, w0);
							tmem_ld8_i32(col + NB, w1);
							tmem_ld8_i32(col + 2 * NB, w2);
							tmem_ld8_i32(tmem + lane_addr + 2u * BP_UNIT_COLS + (uint32_t)(lt * NB + pb + 8 * h), fdb);
							tmem_ld_wait();
							if (h == 1) {
								umma::tc_fence_before();
								__syncwarp();
								if (lane == 0) umma::mbar_arrive(&tmem_empty[mat]);
							}
							if (!(p.dbg & 4)) {
#pragma unroll
								for (int jj = 0; jj < 8; jj++) {
									const int j = 8 * h + jj;
									const float fd = __int_as_float(fdb[jj]);
									const float t = fmaf((float)w0[jj], 65536.0f, fmaf((float)w1[jj], 256.0f, (float)w2[jj]));
									const float S = __fmul_rn(__fmul_rn(t, rs), iscale[pb + j]);
									const float dy = __fmul_rn(dg, y[lt][j]);
									if (mat == 0) {
										sn[j] = __fadd_rn(__fadd_rn(S, dy), fmaxf(-fd, 0.0f));
									} else {
										const float den = __fadd_rn(__fadd_rn(S, dy), fmaxf(fd, 0.0f));
										if (i < N) y[lt][j] = __fmul_rn(__fdiv_rn(sn[j], den), y[lt][j]);
									}
								}
							}
						}
					}
					pair++;
					publish_tile_max(y[lt], par_out);
				}
			}
			PROF_T(tq);
			exchange(par_out, (uint32_t)((it + 1) & 1));
			PROF_ADD(PROF_EPI_REQUANT, tq);
		}
		PROF_ADD(PROF_EPI_TOTAL, te0);
		if (prof_on && et == 0) {
			for (int i = PROF_EPI_TOTAL; i <= PROF_EPI_REQUANT; i++) p.prof[i] = prof_acc[i];
			for (int i = 8; i < 12; i++) p.prof[i] = prof_acc[i];
		}
#pragma unroll
		for (int lt = 0; lt < BP_MAX_MY; lt++) {
			const int i = ((int)rank + 2 * lt) * 128 + r;
			if (lt < n_my && i < N) {
#pragma unroll
				for (int j = 0; j < PW; j++)
					if (b0 + pb + j < p.B) p.Y[(size_t)(b0 + pb + j) * N + i] = y[lt][j];
			}
		}
	}
	umma::tc_fence_before();
	__syncthreads();
	cluster_sync_all(); /* nobody leaves while the peer may still store into / arrive on this CTA */
	if (warp == 1) umma::tmem_dealloc(tmem, 512);
}

int pqp_batched_imma_pair_supported(int N) { return N > 128 && N <= 512; }

cudaError_t pqp_launch_batched_imma_pair(const void *tiles, const void *rowc, int N, int B, const float *Fd, float *Y, int iters,
					 size_t smem_optin, cudaStream_t s)
{
	BiParams p;
	memset(&p, 0, sizeof p);
	p.Atiles = reinterpret_cast<const unsigned char *>(tiles);
	p.rowc = reinterpret_cast<const float4 *>(rowc);
	p.Fd = Fd;
	p.Y = Y;
	p.N = N;
	p.B = B;
	p.iters = iters;
	pqp_imma_geometry(N, &p.MT, &p.NKS, &p.ksc);
	p.b_sbo = (uint32_t)(p.NKS * 32) * 16u;
	p.dbg = pqp_env("PQP_IMMA_DBG") ? atoi(pqp_env("PQP_IMMA_DBG")) : 0;
	if (iters <= 0 || !pqp_batched_imma_pair_supported(N)) return cudaErrorInvalidValue;

	const size_t pbuf = 3 * (size_t)(BP_NB / 16) * p.b_sbo;
	const size_t stage_bytes = (size_t)p.ksc * BI_CHUNK;
	const size_t misc = 3 * BP_NB * sizeof(uint32_t) + 96;
	int stages = (int)((smem_optin - 1024 - pbuf - misc) / (stage_bytes + 16));
	if (stages > 16) stages = 16;
	if (pqp_env("PQP_IMMA_STAGES")) {
		const int v = atoi(pqp_env("PQP_IMMA_STAGES"));
		if (v >= 2 && v <= stages) stages = v;
	}
	if (stages < 2) return cudaErrorInvalidConfiguration;
	p.stages = stages;
	const size_t smem = (size_t)stages * stage_bytes + pbuf + misc + (size_t)stages * 16;
	cudaError_t e = cudaFuncSetAttribute(batched_imma_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;

	static long long *prof_dev = NULL;
	if (p.dbg & 8) {
		if (!prof_dev) cudaMalloc((void **)&prof_dev, 12 * sizeof(long long));
		cudaMemsetAsync(prof_dev, 0, 12 * sizeof(long long), s);
		p.prof = prof_dev;
	}
	cudaLaunchConfig_t cfg;
	memset(&cfg, 0, sizeof cfg);
	cfg.gridDim = dim3(2 * ((B + BP_NB - 1) / BP_NB));
	cfg.blockDim = dim3(BP_THREADS);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = s;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = 2;
	attr[0].val.clusterDim.y = 1;
	attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr;
	cfg.numAttrs = 1;
	e = cudaLaunchKernelEx(&cfg, batched_imma_pair_kernel, p);
	if ((p.dbg & 8) && e == cudaSuccess) {
		long long h[12];
		cudaStreamSynchronize(s);
		cudaMemcpy(h, prof_dev, sizeof h, cudaMemcpyDeviceToHost);
		const double it = (double)iters;
		fprintf(stderr, "  exchange split: wait allmax %.0f, scales+quantise+stores %.0f, proxy fence %.0f\n", h[8] / it, h[9] / it, h[10] / it);
		fprintf(stderr,
			"imma pair profile (CTA 0, cycles per update): mma warp total %.0f = wait b_ready %.0f + wait tmem_empty %.0f + wait full(stream) %.0f + issue %.0f | "
			"epilogue total %.0f = wait tmem_full %.0f + exchange %.0f + drain/math %.0f | producer wait empty %.0f\n",
			h[0] / it, h[1] / it, h[2] / it, h[3] / it, (h[0] - h[1] - h[2] - h[3]) / it, h[4] / it, h[5] / it, h[6] / it,
			(h[4] - h[5] - h[6]) / it, h[7] / it);
	}
	return e;
}
