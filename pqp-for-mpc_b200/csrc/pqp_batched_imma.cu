/*
 * pqp_batched_imma.cu -- B problems sharing one Hessian on the 5th-gen tensor cores (sm_100a), with
 * ERROR-FREE accumulation: tcgen05.mma.kind::i8 over byte slices of the operands, int32 accumulators in TMEM.
 *
 * The PQP update for many MPC states at once (PQP_CPU.c:603-618 + 590-596 applied to B right-hand sides)
 *      NUM = (Q^- + theta) Y + F^-,   DEN = (Q^+ + theta) Y + F^+,   Y <- NUM/DEN o Y,      Y in R^{N x B}
 * is two N x N x B contractions per iteration (4*N^2*B flop).
 *
 * Why integers.  The fp32 accumulator of tcgen05.mma.kind::tf32 truncates on every accumulating step (measured:
 * ~2.8e-8 relative per MMA, always downwards for the non-negative sums of this algorithm; tools/tc_bias_probe.py).
 * Over the 120 steps of an N=480 row that is a coherent -3e-6 bias on both sums, which the fixed-point iteration
 * amplifies well past the 1e-5 parity tolerance.  Integer tensor-core accumulation is exact, so the operands are
 * cut into byte slices (Ozaki-style splitting) instead:
 *   off-diagonal Q^+ / Q^- row i :  a_ik = rint(q_ik * 2^(24-e_i))  < 2^24,  a = A0*2^16 + A1*2^8 + A2   (u8 digits; x-independent, built once)
 *   y of problem b               :  y_k  = rint(y_k  * 2^(22-f_b)) <= 2^22,  y = Y0*2^16 + Y1*2^8 + Y2   (signed s8 digits, round to nearest,
 *                                   so the dropped cross terms are zero-mean), 2^f_b > max_k y_k, re-derived every iteration
 *   w0 = A0*Y0,  w1 = A0*Y1 + A1*Y0,  w2 = A0*Y2 + A1*Y1 + A2*Y0      exact in int32 (|w| < 2^26 for N <= 512)
 *   sum_k q_ik y_k = (w0*2^16 + w1*2^8 + w2) * 2^(e_i-8) * 2^(f_b-22)  up to the 2^-24-level roundings of the fixed-point conversions
 * The diagonal term (Q^+-_ii + theta_i)*y_i (the largest single term, PQP_CPU.c:527,536) and F^+- are added in fp32 in the
 * epilogue, the division is IEEE, and the master copy of y stays in fp32 registers: only the tensor-core operand is quantised.
 * Accuracy against a float64 run of the same algorithm is that of PQP_CPU.c's own fp32 arithmetic (tests/imma_model.py).
 *
 * It is also the cheaper contraction: 6 slice products at the int8 rate (4x tf32) = 1.5 tf32-equivalents instead of 3 (3xTF32),
 * 3 bytes per operand element instead of 8.
 *
 * Mapping (one CTA = NB problems for the WHOLE solve; no grid barrier, no host round trip, no HBM traffic in the loop):
 *   M = 128 rows of Q (an "M tile"), K = columns of Q (32 per MMA), N = problems x digit planes.
 *   A operand  = pre-sliced, pre-tiled Q^- / Q^+ digit planes (3 x 4 KB per (matrix, M tile, K step); 1.5 MB at N=480, L2
 *                resident), streamed through a shared-memory ring by 1-D bulk async copies (UBLKCP), multicast to the CTAs
 *                of a cluster.
 *   B operand  = the CTA's Y digit planes [Y0 | Y1 | Y2], MN-major in shared memory for the whole solve (so an epilogue thread,
 *                which owns one row k and 16 problems, writes its 16 digits with ONE 128-bit store per plane).
 *   D          = one "unit" (matrix, M tile) = 3*NB int32 columns [w0 | w1 | w2], double buffered in TMEM:
 *                  MMA(A0, [Y0|Y1|Y2]) -> [w0 w1 w2];  MMA(A1, [Y0|Y1]) -> += [w1 w2];  MMA(A2, [Y0]) -> += [w2]
 *                so the tensor pipe works on unit u+1 while the epilogue warps drain unit u.
 *   warp 0     producer (bulk copies + mbarriers), warp 1 MMA issuer (single thread) + TMEM allocator,
 *   warps 2..  epilogue (4 lane quarters x NB/16 problem groups): tcgen05.ld, int -> fp32, +diag, +F, divide, update the fp32
 *              master y in registers, per-problem max (redux + shared atomics), requantise, store the digit planes.
 */
#include "pqp_internal.h"
#include "pqp_umma.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pqp_imma.cuh"



/*
 * shared memory: ring [stages][ksc*BI_CHUNK] | B planes [3][NB/16][Kpad/8][8][16 B] | smax[2][NB] | iscale[2][NB] | barriers
 */
template <int NB, int PW, bool TOL>
__global__ void __launch_bounds__(64 + 128 * (NB / PW), 1) batched_imma_kernel(const BiParams p)
{
	constexpr int EW = 4 * (NB / PW);    /* epilogue warps: 4 lane quarters x NB/PW problem groups of PW problems per thread */
	constexpr int ETHREADS = 32 * EW;
	constexpr uint32_t UNIT_COLS = 3 * NB;
	/* Fd of the CTA's rows never changes: with 32 problems per CTA it fits in the TMEM columns the two accumulator buffers leave
	 * free (column 2*UNIT_COLS + 32*mt + problem) and is read back with tcgen05.ld instead of an L2 round trip per update */
	constexpr bool FD_TMEM = (NB == 32) && (PW == 8);
	constexpr uint32_t FD_COL0 = 2 * UNIT_COLS;
	constexpr uint32_t TMEM_NEED = 2 * UNIT_COLS + (FD_TMEM ? (uint32_t)(BI_MAX_MT * NB) : 0u);
	constexpr uint32_t TMEM_COLS = TMEM_NEED <= 128 ? 128u : (TMEM_NEED <= 256 ? 256u : 512u);

	extern __shared__ __align__(128) unsigned char smem_raw[];
	const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
	const uint32_t CS = cluster_nctarank(), crank = cluster_ctarank();
	const uint16_t cmask = (uint16_t)((1u << CS) - 1u);

	const int N = p.N, MT = p.MT, NKS = p.NKS;
	const int Kpad = NKS * 32;
	unsigned char *ring = smem_raw;
	const uint32_t stage_bytes = (uint32_t)p.ksc * BI_CHUNK;
	unsigned char *Bpl = ring + (size_t)p.stages * stage_bytes;
	const uint32_t plane_bytes = (uint32_t)(NB / 16) * p.b_sbo;
	const uint32_t pbuf_bytes = 3u * plane_bytes;                 /* one buffer = the three digit planes */
	const uint32_t nbuf = p.dbuf ? 2u : 1u;
	uint32_t *smax = reinterpret_cast<uint32_t *>(Bpl + nbuf * pbuf_bytes);
	float *iscale = reinterpret_cast<float *>(smax + 2 * NB);
	uint64_t *full = reinterpret_cast<uint64_t *>(iscale + 2 * NB);
	uint64_t *empty = full + p.stages;
	uint64_t *tmem_full = empty + p.stages;   /* [2] */
	uint64_t *tmem_empty = tmem_full + 2;     /* [2] */
	uint64_t *b_ready = tmem_empty + 2;
	uint64_t *decided = b_ready + 1;          /* tolerance mode: the stop decision of an evaluation pass is in *stop_flag */
	uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(decided + 1);
	uint32_t *stop_flag = tmem_slot + 1;
	uint32_t *frozen = stop_flag + 1;         /* [NB] */
	float *part = reinterpret_cast<float *>(frozen + NB); /* [EW][PW][5] per-warp partial sums of an evaluation pass */
	constexpr bool tol = TOL; /* run-to-tolerance is a separate instantiation: the fixed-count kernel carries none of its registers */

	const int units_per_iter = 2 * MT;
	const int chunks_per_unit = NKS / p.ksc;
	const int chunks_per_iter = units_per_iter * chunks_per_unit;
	const int b0 = blockIdx.x * NB;
	const bool prof_on = (p.dbg & 8) && p.prof && blockIdx.x == 0;
	long long prof_acc[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };

	if (tid == 0) {
		for (int s = 0; s < p.stages; s++) {
			umma::mbar_init(&full[s], 1);
			umma::mbar_init(&empty[s], CS);
		}
		for (int i = 0; i < 2; i++) {
			umma::mbar_init(&tmem_full[i], 1);
			umma::mbar_init(&tmem_empty[i], EW); /* one arrival per epilogue warp */
		}
		umma::mbar_init(b_ready, EW);
		umma::mbar_init(decided, ETHREADS);
		*stop_flag = 0u;
		umma::mbar_fence_init();
	}
	if (tid < 2 * NB) smax[tid] = 0u;
	if (tid < NB) frozen[tid] = (b0 + tid < p.B) ? 0u : 1u; /* padding problems never hold the CTA back */
	for (uint32_t i = tid; i < nbuf * pbuf_bytes / 16u; i += blockDim.x) reinterpret_cast<uint4 *>(Bpl)[i] = make_uint4(0u, 0u, 0u, 0u);
	umma::fence_proxy_async(); /* K padding of the planes is never rewritten: make the zeros visible to the tensor core's reads */
	if (warp == 1) umma::tmem_alloc(tmem_slot, TMEM_COLS);
	umma::tc_fence_before();
	__syncthreads();
	if (CS > 1) cluster_sync_all(); /* every CTA's barriers exist before anyone multicasts into them */
	umma::tc_fence_after();
	const uint32_t tmem = *tmem_slot;

	if (warp == 0) {
		/* ================= producer: A chunks through the ring (whole warp, one elected lane issues) ================= */
		int st = 0;
		uint32_t ph = 0;
		long long c = 0;
		int upd = 0, since = 0;
		uint32_t n_eval = 0;
		for (;;) {
			if (!tol && upd == p.iters) break;
			const bool ev = tol && (since == p.check_every || upd >= p.max_iters);
			for (int within = 0; within < chunks_per_iter; within++, c++) {
				PROF_T(tw);
				umma::mbar_wait(&empty[st], ph ^ 1u);
				PROF_ADD(PROF_PROD_WAIT_EMPTY, tw);
				if (elect_one()) {
					umma::mbar_arrive_expect_tx(&full[st], stage_bytes);
					const unsigned char *src = p.Atiles + (size_t)within * stage_bytes;
					if (CS == 1)
						bulk_g2s_plain(ring + (size_t)st * stage_bytes, src, stage_bytes, &full[st]);
					else if ((uint32_t)(c % CS) == crank)
						bulk_g2s_mcast(ring + (size_t)st * stage_bytes, src, stage_bytes, &full[st], cmask);
				}
				__syncwarp();
				if (++st == p.stages) { st = 0; ph ^= 1u; }
			}
			if (ev) {
				/* no copies may be in flight into a CTA that is about to exit: wait for the decision before running ahead */
				since = 0;
				umma::mbar_wait(decided, n_eval & 1u);
				n_eval++;
				if (*reinterpret_cast<volatile uint32_t *>(stop_flag)) break;
			} else {
				upd++;
				since++;
			}
		}
		if (prof_on && lane == 0) p.prof[PROF_PROD_WAIT_EMPTY] = prof_acc[PROF_PROD_WAIT_EMPTY];
	} else if (warp == 1) {
		/* ================= MMA issuer (whole warp in uniform control flow, one elected lane issues) ================= */
		const uint32_t id3 = idesc_i8(128, 3 * NB), id2 = idesc_i8(128, 2 * NB), id1 = idesc_i8(128, NB);
		const uint64_t a_desc0 = umma::smem_desc(umma::smem_addr(ring), BI_A_LBO, BI_A_SBO);
		const uint64_t b_desc0 = umma::smem_desc(umma::smem_addr(Bpl), BI_B_LBO, p.b_sbo);
		int st = 0;
		uint32_t ph = 0;
		long long unit = 0;
		int upd = 0, since = 0;
		uint32_t n_eval = 0, pbi = 0; /* pbi: plane buffer the MMAs of this pass read */
		PROF_T(tm0);
		for (long long pass = 0;; pass++) {
			if (!tol && upd == p.iters) break;
			const bool ev = tol && (since == p.check_every || upd >= p.max_iters);
			PROF_T(tb);
			umma::mbar_wait(b_ready, (uint32_t)(pass & 1)); /* digit planes of this pass are in place */
			PROF_ADD(PROF_MMA_WAIT_BREADY, tb);
			umma::tc_fence_after();
			for (int u = 0; u < units_per_iter; u++, unit++) {
				const int buf = u & 1;
				PROF_T(te);
				umma::mbar_wait(&tmem_empty[buf], (uint32_t)(((unit >> 1) & 1) ^ 1)); /* epilogue drained this buffer */
				PROF_ADD(PROF_MMA_WAIT_TMEM, te);
				umma::tc_fence_after();
				const uint32_t d = tmem + (uint32_t)buf * UNIT_COLS;
				for (int ch = 0; ch < chunks_per_unit; ch++) {
					PROF_T(tf);
					umma::mbar_wait(&full[st], ph);
					PROF_ADD(PROF_MMA_WAIT_FULL, tf);
					umma::tc_fence_after();
					if (elect_one()) {
						/* only the 14-bit start-address field of the descriptors moves: +768 per 12 KB of ring, +32 per K step */
						const uint64_t da = a_desc0 + (uint64_t)((uint32_t)st * (stage_bytes >> 4));
						const uint64_t db = b_desc0 + (uint64_t)((uint32_t)(ch * p.ksc) * (4u * BI_B_LBO >> 4) + (pbi ? (pbuf_bytes >> 4) : 0u));
						if (!(p.dbg & 2)) {
							if (p.ksc == 3) mma_i8_step3(d, d + NB, d + 2 * NB, da, db, id3, id2, id1, ch ? 1u : 0u);
							else mma_i8_step(d, d + NB, d + 2 * NB, da, db, id3, id2, id1, ch ? 1u : 0u);
						}
						if (CS == 1) umma::mma_commit(&empty[st]);
						else mma_commit_mcast(&empty[st], cmask); /* the stage is free once EVERY CTA of the cluster has read it */
					}
					__syncwarp();
					if (++st == p.stages) { st = 0; ph ^= 1u; }
				}
				if (elect_one()) umma::mma_commit(&tmem_full[buf]);
				__syncwarp();
			}
			if (ev) {
				since = 0;
				umma::mbar_wait(decided, n_eval & 1u);
				n_eval++;
				if (*reinterpret_cast<volatile uint32_t *>(stop_flag)) break;
			} else {
				upd++;
				since++;
				if (p.dbuf) pbi ^= 1u;
			}
		}
		PROF_ADD(PROF_MMA_TOTAL, tm0);
		if (prof_on && lane == 0)
			for (int i = PROF_MMA_TOTAL; i <= PROF_MMA_WAIT_FULL; i++) p.prof[i] = prof_acc[i];
	} else {
		/* ================= epilogue warps ================= */
		const int et = tid - 64;
		const int ew = warp - 2;
		const int q = warp % 4;    /* TMEM lane quarter this warp may touch */
		const int cg = ew / 4;     /* which PW of the NB problems */
		const int pb = cg * PW;
		const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
		const int r = 32 * q + lane; /* row within an M tile */

		float y[BI_MAX_MT][PW];    /* fp32 master copy of this thread's duals: rows mt*128+r, problems pb..pb+PW-1 */
		float4 rc[BI_MAX_MT];

		/* digits of one M tile of this thread's rows (y[mt][0..15] scaled by sc[j]) -> plane buffer `buf` */
		auto store_tile = [&](int mt, const float(&yv)[PW], const float(&sc)[PW], uint32_t buf) {
			const int i = mt * 128 + r;
			if (i < Kpad) {
				uint32_t w0[4] = { 0, 0, 0, 0 }, w1[4] = { 0, 0, 0, 0 }, w2[4] = { 0, 0, 0, 0 };
#pragma unroll
				for (int j = 0; j < PW; j += 4)
					digits4(__float2int_rn(yv[j] * sc[j]), __float2int_rn(yv[j + 1] * sc[j + 1]), __float2int_rn(yv[j + 2] * sc[j + 2]),
						__float2int_rn(yv[j + 3] * sc[j + 3]), w0[j >> 2], w1[j >> 2], w2[j >> 2]);
				unsigned char *dst = Bpl + buf * pbuf_bytes + (uint32_t)(pb >> 4) * p.b_sbo + (uint32_t)(i >> 3) * BI_B_LBO + (uint32_t)(i & 7) * 16u +
						     (uint32_t)(pb & 15);
				if (PW == 16) {
					*reinterpret_cast<uint4 *>(dst) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
					*reinterpret_cast<uint4 *>(dst + plane_bytes) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
					*reinterpret_cast<uint4 *>(dst + 2u * plane_bytes) = make_uint4(w2[0], w2[1], w2[2], w2[3]);
				} else {
					*reinterpret_cast<uint2 *>(dst) = make_uint2(w0[0], w0[1]);
					*reinterpret_cast<uint2 *>(dst + plane_bytes) = make_uint2(w1[0], w1[1]);
					*reinterpret_cast<uint2 *>(dst + 2u * plane_bytes) = make_uint2(w2[0], w2[1]);
				}
			}
		};
		auto store_planes = [&](const float(&sc)[PW], uint32_t buf) {
#pragma unroll
			for (int mt = 0; mt < BI_MAX_MT; mt++)
				if (mt < MT) store_tile(mt, y[mt], sc, buf);
		};
		/* per-problem max of one M tile's new duals -> shared atomics (non-negative floats order like their bit patterns;
		 * NaN patterns order above everything, which is what lets a NaN poison its whole problem below) */
		auto publish_tile_max = [&](const float(&yv)[PW], uint32_t *dstmax) {
#pragma unroll
			for (int j = 0; j < PW; j++) {
				const uint32_t wm = __reduce_max_sync(0xffffffffu, __float_as_uint(yv[j]) & 0x7fffffffu);
				if (lane == j) atomicMax(dstmax + pb + j, wm);
			}
		};
		/*
		 * After everyone published: the exact per-problem maxima are known.  Derive the scales, remember the inverses for the
		 * accumulator conversion of the next iteration, and make sure plane buffer `buf` holds digits quantised with exactly these
		 * scales: `have_ex` < 0 means nothing has been stored yet; otherwise the tiles were already stored speculatively with the
		 * scales of exponent field pred_bits>>23 and only need redoing where the exponent moved.
		 * A problem whose maximum is not finite is set to NaN as a whole: the reference's dense sums do exactly that one
		 * iteration later (0*NaN = NaN in every row), whereas the quantiser would silently drop the NaN.
		 */
		auto requantise = [&](int parity, uint32_t buf, bool speculated, int par_pred) {
			named_bar_sync(1, ETHREADS);
			float sc[PW];
			bool redo = !speculated;
#pragma unroll
			for (int j = 0; j < PW; j++) {
				const uint32_t mx = smax[parity * NB + pb + j];
				float isc;
				problem_scales(mx, sc[j], isc);
				if (q == 0 && lane == j) iscale[parity * NB + pb + j] = isc;
				if ((mx >> 23) >= 255u) {
#pragma unroll
					for (int mt = 0; mt < BI_MAX_MT; mt++) y[mt][j] = __uint_as_float(0x7fc00000u);
				}
				if (speculated) {
					float scp, iscp;
					problem_scales(smax[par_pred * NB + pb + j], scp, iscp);
					redo = redo || (scp != sc[j]);
				}
			}
			if (redo) store_planes(sc, buf);
			umma::fence_proxy_async(); /* the new planes must be visible to the tensor core's reads */
		};

		/* iteration 0: load y0 and the row constants */
#pragma unroll
		for (int mt = 0; mt < BI_MAX_MT; mt++) {
			const int i = mt * 128 + r;
			rc[mt] = make_float4(0.f, 0.f, 0.f, 0.f);
			if (mt < MT) rc[mt] = __ldg(p.rowc + i);
#pragma unroll
			for (int j = 0; j < PW; j++) {
				float v = 0.0f;
				if (mt < MT && i < N && b0 + pb + j < p.B) v = p.Y[(size_t)(b0 + pb + j) * N + i];
				y[mt][j] = v;
			}
		}
		if (FD_TMEM) {
#pragma unroll
			for (int mt = 0; mt < BI_MAX_MT; mt++) {
				const int i = mt * 128 + r;
				float fv[8];
#pragma unroll
				for (int j = 0; j < 8; j++) fv[j] = (mt < MT && i < N && b0 + pb + j < p.B) ? __ldg(p.Fd + (size_t)(b0 + pb + j) * N + i) : 1.0f;
				tmem_st8_f32(tmem + lane_addr + FD_COL0 + (uint32_t)(mt * NB + pb), fv);
			}
			umma::tc_fence_before();
		}
#pragma unroll
		for (int mt = 0; mt < BI_MAX_MT; mt++)
			if (mt < MT) publish_tile_max(y[mt], smax + NB); /* parity 1 = "planes for iteration 0" (iteration it publishes into parity it&1) */
		requantise(1, 0u, false, 0);
		__syncwarp();
		if (lane == 0) umma::mbar_arrive(b_ready);

		long long pair = 0; /* (pass, M tile) counter: completion index of tmem_full[0] and [1] */
		int upd = 0, since = 0;
		uint32_t n_eval = 0, pbi = 0;
		PROF_T(te0);
		for (;;) {
			if (!tol && upd == p.iters) break;
			const bool ev = tol && (since == p.check_every || upd >= p.max_iters);
			const int par_out = upd & 1;     /* max / scale slots the update of this pass fills */
			const int par_in = par_out ^ 1;  /* scales the planes being read were quantised with */
			float isc[PW];
			uint32_t fz = 0;                 /* problems of this thread that are frozen (tolerance mode) */
			if (tol) {
#pragma unroll
				for (int j = 0; j < PW; j++) fz |= (frozen[pb + j] ? 1u : 0u) << j;
			}
#pragma unroll
			for (int mt = 0; mt < BI_MAX_MT; mt++) {
				if (mt < MT) {
					const int i = mt * 128 + r;
					float fd[PW];
					if (!FD_TMEM) {
#pragma unroll
						for (int j = 0; j < PW; j++) fd[j] = (i < N && b0 + pb + j < p.B) ? __ldg(p.Fd + (size_t)(b0 + pb + j) * N + i) : 1.0f;
					}
					float sn[PW];
					float e_viol[PW], e_min[PW], e_gap[PW], e_jd[PW], e_kkt[PW]; /* evaluation pass only */
#pragma unroll
					for (int mat = 0; mat < 2; mat++) {
						PROF_T(tw);
						umma::mbar_wait(&tmem_full[mat], (uint32_t)(pair & 1));
						PROF_ADD(PROF_EPI_WAIT_TMEM, tw);
						umma::tc_fence_after();
						if (mt == 0 && mat == 0) {
							/* first unit of the pass: every thread is past the previous requantise -> safe to read the inverse
							 * scales and (update pass) to reset the max slots the NEXT requantise will use */
#pragma unroll
							for (int j = 0; j < PW; j++) isc[j] = iscale[par_in * NB + pb + j];
							if (!ev) {
								if (et < NB) smax[par_out * NB + et] = 0u;
								named_bar_sync(2, ETHREADS);
							}
						}
						const uint32_t col = tmem + lane_addr + (uint32_t)mat * UNIT_COLS + (uint32_t)pb;
						int w0[PW], w1[PW], w2[PW];
						tmem_ld_i32<PW>(col, w0);
						tmem_ld_i32<PW>(col + NB, w1);
						tmem_ld_i32<PW>(col + 2 * NB, w2);
						if (FD_TMEM && mat == 0) tmem_ld8_i32(tmem + lane_addr + FD_COL0 + (uint32_t)(mt * NB + pb), reinterpret_cast<int *>(fd));
						tmem_ld_wait();
						umma::tc_fence_before();
						__syncwarp();
						if (lane == 0) umma::mbar_arrive(&tmem_empty[mat]); /* the tensor pipe may start the next unit in this buffer */
						const float rs = mat == 0 ? rc[mt].z : rc[mt].w;
						const float dg = mat == 0 ? rc[mt].x : rc[mt].y;
						if (!(p.dbg & 4)) {
							float tol_i = 0.0f;
							if (ev && mat == 1) tol_i = (p.Kp && i < N) ? fmaxf(p.erc * __ldg(p.Kp + i), p.eac) : p.eac;
#pragma unroll
							for (int j = 0; j < PW; j++) {
								/* (w0*2^16 + w1*2^8 + w2) * 2^(e-8) * 2^(f-22): two fused roundings, then exact scalings */
								const float t = fmaf((float)w0[j], 65536.0f, fmaf((float)w1[j], 256.0f, (float)w2[j]));
								const float S = __fmul_rn(__fmul_rn(t, rs), isc[j]);
								const float dy = __fmul_rn(dg, y[mt][j]);
								if (mat == 0) {
									sn[j] = __fadd_rn(__fadd_rn(S, dy), fmaxf(-fd[j], 0.0f));
								} else {
									const float den = __fadd_rn(__fadd_rn(S, dy), fmaxf(fd[j], 0.0f));
									if (!ev) {
										if (i < N && !((fz >> j) & 1u)) y[mt][j] = __fmul_rn(__fdiv_rn(sn[j], den), y[mt][j]);
									} else {
										/* g = den - num = Qd y + Fd: slack of constraint i at U(y) */
										const float g = __fsub_rn(den, sn[j]), yi = y[mt][j];
										const bool live = i < N;
										e_viol[j] = live ? -g - tol_i : -INFINITY;
										e_min[j] = live ? g : INFINITY;
										e_gap[j] = live ? __fmul_rn(yi, g) : 0.0f;
										e_jd[j] = live ? __fmul_rn(yi, 0.5f * (g + fd[j])) : 0.0f;
										e_kkt[j] = live ? fabsf(fminf(yi, g)) : 0.0f;
									}
								}
							}
						}
					}
					pair++;
					if (!ev) {
						/* this tile's new duals are final: publish their maxima, and (two plane buffers) store their digits right away
						 * with the scales of the previous maxima -- almost always the exponent the exact maxima will confirm */
						publish_tile_max(y[mt], smax + par_out * NB);
						if (p.dbuf) {
							float scp[PW];
#pragma unroll
							for (int j = 0; j < PW; j++) {
								float iscp;
								problem_scales(smax[par_in * NB + pb + j], scp[j], iscp);
							}
							store_tile(mt, y[mt], scp, pbi ^ 1u);
						}
					} else {
						/* fold this tile's rows: fixed xor tree over the 32 rows of the warp, tiles in ascending order */
#pragma unroll
						for (int j = 0; j < PW; j++) {
							float v = e_viol[j], m = e_min[j], ga = e_gap[j], jd = e_jd[j], kk = e_kkt[j];
#pragma unroll
							for (int o = 16; o; o >>= 1) {
								v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
								m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
								ga = __fadd_rn(ga, __shfl_xor_sync(0xffffffffu, ga, o));
								jd = __fadd_rn(jd, __shfl_xor_sync(0xffffffffu, jd, o));
								kk = fmaxf(kk, __shfl_xor_sync(0xffffffffu, kk, o));
							}
							if (lane == 0) {
								float *pp = part + ((size_t)ew * PW + j) * 5;
								if (mt == 0) {
									pp[0] = v; pp[1] = m; pp[2] = ga; pp[3] = jd; pp[4] = kk;
								} else {
									pp[0] = fmaxf(pp[0], v); pp[1] = fminf(pp[1], m); pp[2] = __fadd_rn(pp[2], ga);
									pp[3] = __fadd_rn(pp[3], jd); pp[4] = fmaxf(pp[4], kk);
								}
							}
						}
					}
				}
			}
			if (!ev) {
				PROF_T(tq);
				requantise(par_out, p.dbuf ? (pbi ^ 1u) : 0u, p.dbuf != 0, par_in);
				__syncwarp(); /* every lane's digit stores and proxy fence precede lane 0's (release) arrival */
				if (lane == 0) umma::mbar_arrive(b_ready);
				PROF_ADD(PROF_EPI_REQUANT, tq);
				upd++;
				since++;
				if (p.dbuf) pbi ^= 1u;
			} else {
				/* the stop test of terminate(), per problem, on the reductions of g (no extra matrix traffic) */
				named_bar_sync(1, ETHREADS);
				if (et < NB && !frozen[et] ) {
					const int cgp = et / PW, j = et % PW;
					float v = -INFINITY, m = INFINITY, ga = 0.0f, jd = 0.0f, kk = 0.0f;
					for (int w = 0; w < 4; w++) { /* the four lane-quarter warps of this problem group, fixed order */
						const float *pp = part + ((size_t)(cgp * 4 + w) * PW + j) * 5;
						v = fmaxf(v, pp[0]); m = fminf(m, pp[1]); ga = __fadd_rn(ga, pp[2]); jd = __fadd_rn(jd, pp[3]); kk = fmaxf(kk, pp[4]);
					}
					const float Jd = jd + (p.Md ? 0.5f * p.Md[b0 + et] : 0.0f);
					const bool conv = v <= 0.0f && fabsf(ga) <= p.eaj && fabsf(ga) <= p.erj * fabsf(Jd);
					/* a problem the reference itself drives to 0/0 (all-NaN duals) can never pass: retire it unconverged
					 * instead of holding its 31 batch mates until max_iters */
					const bool dead = !(fabsf(ga) <= 3.0e38f) || !(fabsf(jd) <= 3.0e38f);
					if (conv || dead || upd >= p.max_iters) {
						pqp_status o;
						o.iters = upd; o.converged = conv ? 1 : 0; o.min_slack = m; o.gap = ga; o.Jd = Jd; o.kkt = kk;
						p.status[b0 + et] = o;
					}
					if (conv || dead) frozen[et] = 1u;
				}
				named_bar_sync(1, ETHREADS);
				if (et == 0) {
					uint32_t all = 1u;
					for (int b = 0; b < NB; b++) all &= frozen[b];
					*stop_flag = (all || upd >= p.max_iters) ? 1u : 0u;
				}
				since = 0;
				umma::mbar_arrive(decided);
				umma::mbar_wait(decided, n_eval & 1u);
				n_eval++;
				if (*reinterpret_cast<volatile uint32_t *>(stop_flag)) break;
				__syncwarp();
				if (lane == 0) umma::mbar_arrive(b_ready); /* the planes are unchanged: the next pass may start */
			}
		}
		PROF_ADD(PROF_EPI_TOTAL, te0);
		if (prof_on && et == 0)
			for (int i = PROF_EPI_TOTAL; i <= PROF_EPI_REQUANT; i++) p.prof[i] = prof_acc[i];
		/* result: the fp32 master copy */
#pragma unroll
		for (int mt = 0; mt < BI_MAX_MT; mt++) {
			const int i = mt * 128 + r;
			if (mt < MT && i < N) {
#pragma unroll
				for (int j = 0; j < PW; j++)
					if (b0 + pb + j < p.B) p.Y[(size_t)(b0 + pb + j) * N + i] = y[mt][j];
			}
		}
	}
	umma::tc_fence_before();
	__syncthreads();
	if (CS > 1) cluster_sync_all(); /* nobody leaves while a peer may still multicast into / arrive on this CTA */
	if (warp == 1) umma::tmem_dealloc(tmem, TMEM_COLS);
}

/*
 * Builds the digit planes of the off-diagonal Q^- / Q^+ and the per-row constants (x-independent, once per handle).
 * One warp per (row, matrix): pass 1 finds the row maximum, pass 2 quantises and scatters the bytes into the tiles.
 */
__global__ void build_imma_tiles_kernel(unsigned char *__restrict__ tiles, float4 *__restrict__ rowc, const float *__restrict__ Q, int ldq,
					 const float *__restrict__ theta, int N, int MT, int NKS)
{
	const int lane = threadIdx.x % 32;
	const int gw = (blockIdx.x * blockDim.x + threadIdx.x) / 32;
	const int nw = gridDim.x * blockDim.x / 32;
	const int rows = MT * 128, Kpad = NKS * 32;
	for (int i = gw; i < rows; i += nw) {
		float dg[2] = { 0.f, 0.f }, rs[2] = { 0.f, 0.f };
		for (int mat = 0; mat < 2; mat++) {
			const float sgn = mat == 0 ? -1.0f : 1.0f;
			float m = 0.0f;
			if (i < N)
				for (int k = lane; k < N; k += 32)
					if (k != i) m = fmaxf(m, fmaxf(sgn * Q[(size_t)i * ldq + k], 0.0f));
			m = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(m)));
			/* m < 2^e with e = ex-126; entries are scaled by 2^(24-e) = 2^(150-ex); the epilogue undoes it with 2^(e-8) (2^16 folded in) */
			uint32_t ex = __float_as_uint(m) >> 23;
			float scale = 0.0f;
			if (m > 0.0f && ex >= 32u && ex <= 254u) { /* rows whose largest entry is below 2^-95 are treated as zero */
				scale = __uint_as_float((277u - ex) << 23);  /* 2^(150-ex) */
				rs[mat] = __uint_as_float((ex - 7u) << 23);  /* 2^(ex-134) = 2^(e-8) */
			}
			if (i < N) {
				const float qii = Q[(size_t)i * ldq + i];
				dg[mat] = __fadd_rn(fmaxf(sgn * qii, 0.0f), theta[i]); /* theta added on the diagonal, PQP_CPU.c:527,536 */
			}
			for (int k = lane; k < Kpad; k += 32) {
				uint32_t a = 0;
				if (i < N && k < N && k != i) a = __float2uint_rn(fmaxf(sgn * Q[(size_t)i * ldq + k], 0.0f) * scale);
				const int mt = i / 128, rr = i % 128, ks = k / 32, kk = k % 32;
				unsigned char *blk = tiles + ((size_t)(mt * 2 + mat) * NKS + ks) * BI_CHUNK;
				const uint32_t off = (uint32_t)(kk / 16) * BI_A_LBO + (uint32_t)(rr / 8) * BI_A_SBO + (uint32_t)(rr % 8) * 16u + (uint32_t)(kk % 16);
				blk[off] = (unsigned char)(a >> 16);
				blk[BI_SLICE + off] = (unsigned char)((a >> 8) & 255u);
				blk[2 * BI_SLICE + off] = (unsigned char)(a & 255u);
			}
		}
		if (lane == 0) rowc[i] = make_float4(dg[0], dg[1], rs[0], rs[1]);
	}
}

int pqp_batched_imma_supported(int N) { return N >= 16 && N <= 128 * BI_MAX_MT; }

/* K steps per ring stage and the padded K-step count: stages of 3 K steps (36 KB) keep the issuing thread off the
 * critical path; tiny problems keep single-step stages */
void pqp_imma_geometry(int N, int *MT, int *NKS, int *ksc)
{
	int nks = (N + 31) / 32;
	int k = nks >= 3 ? 3 : 1;
	if (pqp_env("PQP_IMMA_KSC") && atoi(pqp_env("PQP_IMMA_KSC")) == 1) k = 1;
	*ksc = k;
	*NKS = (nks + k - 1) / k * k;
	*MT = (N + 127) / 128;
}

size_t pqp_batched_imma_tiles_bytes(int N)
{
	int MT, NKS, ksc;
	pqp_imma_geometry(N, &MT, &NKS, &ksc);
	return (size_t)2 * MT * NKS * BI_CHUNK;
}
size_t pqp_batched_imma_rowc_bytes(int N) { return (size_t)((N + 127) / 128) * 128 * sizeof(float4); }

cudaError_t pqp_launch_build_imma_tiles(void *tiles, void *rowc, const float *Q, int ldq, const float *theta, int N, cudaStream_t s)
{
	int MT, NKS, ksc;
	pqp_imma_geometry(N, &MT, &NKS, &ksc);
	build_imma_tiles_kernel<<<148, 256, 0, s>>>(reinterpret_cast<unsigned char *>(tiles), reinterpret_cast<float4 *>(rowc), Q, ldq, theta, N,
						    MT, NKS);
	return cudaGetLastError();
}

template <int NB, int PW, bool TOL>
static cudaError_t launch_imma(const BiParams &p0, int cluster, size_t smem_optin, cudaStream_t s)
{
	BiParams p = p0;
	const size_t pbuf = 3 * (size_t)(NB / 16) * p.b_sbo;
	const size_t stage_bytes = (size_t)p.ksc * BI_CHUNK;
	const size_t misc = 4 * NB * sizeof(uint32_t) + 64 /* barriers besides the ring's */ + 32 + NB * sizeof(uint32_t) +
			    (size_t)(4 * NB) * 5 * sizeof(float) /* evaluation partials [EW][PW][5], EW*PW = 4*NB */;
	/* two plane buffers when at least three ring stages still fit beside them */
	p.dbuf = (2 * pbuf + misc + 3 * (stage_bytes + 16) + 1024 <= smem_optin) ? 1 : 0;
	if (pqp_env("PQP_IMMA_DBUF")) p.dbuf = p.dbuf && atoi(pqp_env("PQP_IMMA_DBUF")) != 0;
	const size_t fixed = (p.dbuf ? 2 : 1) * pbuf + misc;
	int stages = (int)((smem_optin - 1024 - fixed) / (stage_bytes + 16));
	if (stages > 16) stages = 16;
	if (pqp_env("PQP_IMMA_STAGES")) {
		const int v = atoi(pqp_env("PQP_IMMA_STAGES"));
		if (v >= 2 && v <= stages) stages = v;
	}
	if (stages < 2) return cudaErrorInvalidConfiguration;
	p.stages = stages;
	const size_t smem = (size_t)stages * stage_bytes + fixed + (size_t)stages * 16;
	cudaError_t e = cudaFuncSetAttribute(batched_imma_kernel<NB, PW, TOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	if (cluster < 1) cluster = 1;
	if (cluster > 8) {
		e = cudaFuncSetAttribute(batched_imma_kernel<NB, PW, TOL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
		if (e != cudaSuccess) return e;
	}
	int ctas = (p.B + NB - 1) / NB;
	ctas = (ctas + cluster - 1) / cluster * cluster;
	cudaLaunchConfig_t cfg;
	memset(&cfg, 0, sizeof cfg);
	cfg.gridDim = dim3(ctas);
	cfg.blockDim = dim3(64 + 128 * (NB / PW));
	cfg.dynamicSmemBytes = smem;
	cfg.stream = s;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = cluster;
	attr[0].val.clusterDim.y = 1;
	attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr;
	cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, batched_imma_kernel<NB, PW, TOL>, p);
}

cudaError_t pqp_launch_batched_imma(const void *tiles, const void *rowc, int N, int B, const float *Fd, float *Y, int iters, int nb,
				    int cluster, size_t smem_optin, const pqp_imma_tol *tolp, cudaStream_t s)
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
	if (iters <= 0) {
		if (!tolp || !tolp->status) return cudaErrorInvalidValue;
		p.iters = 0;
		p.max_iters = tolp->max_iters;
		p.check_every = tolp->check_every > 0 ? tolp->check_every : 1;
		p.erc = tolp->erc; p.eac = tolp->eac; p.eaj = tolp->eaj; p.erj = tolp->erj;
		p.Kp = tolp->Kp; p.Md = tolp->Md; p.status = tolp->status;
		cluster = 1; /* CTAs stop independently */
	}
	static long long *prof_dev = NULL;
	if (p.dbg & 8) {
		if (!prof_dev) cudaMalloc((void **)&prof_dev, 8 * sizeof(long long));
		cudaMemsetAsync(prof_dev, 0, 8 * sizeof(long long), s);
		p.prof = prof_dev;
	}
	/* PW = problems per epilogue thread: 8 doubles the epilogue warps of the 32-problem tile (16 instead of 8), which is what
	 * hides the latency of its dependent fp32 chain (division, conversions) behind the tensor pipe */
	const int pw = pqp_env("PQP_IMMA_PW") ? atoi(pqp_env("PQP_IMMA_PW")) : 8;
	cudaError_t e;
	if (iters <= 0) e = launch_imma<32, 8, true>(p, cluster, smem_optin, s);
	else if (nb == 64) e = launch_imma<64, 16, false>(p, cluster, smem_optin, s);
	else if (pw == 16) e = launch_imma<32, 16, false>(p, cluster, smem_optin, s);
	else e = launch_imma<32, 8, false>(p, cluster, smem_optin, s);
	if ((p.dbg & 8) && e == cudaSuccess) {
		long long h[8];
		cudaStreamSynchronize(s);
		cudaMemcpy(h, prof_dev, sizeof h, cudaMemcpyDeviceToHost);
		const double it = (double)iters;
		fprintf(stderr,
			"imma profile (CTA 0, cycles per iteration): mma warp total %.0f = wait b_ready %.0f + wait tmem_empty %.0f + wait full(stream) %.0f + issue %.0f | "
			"epilogue total %.0f = wait tmem_full %.0f + requantise %.0f + drain/math %.0f | producer wait empty %.0f\n",
			h[0] / it, h[1] / it, h[2] / it, h[3] / it, (h[0] - h[1] - h[2] - h[3]) / it, h[4] / it, h[5] / it, h[6] / it,
			(h[4] - h[5] - h[6]) / it, h[7] / it);
	}
	return e;
}
