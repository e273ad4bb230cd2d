/*
 * include/pqp.h -- C ABI of the B200-native PQP solver (libpqp_b200.so).
 *
 * The reference (yashsoni501/PQP-for-MPC) has no library API: its boundary is the set of
 * free functions every one of its programs calls from main() (PQP_CPU.c:988-999).  This
 * header is the drop-in for that path; each entry point names the reference interface it
 * replaces.  Plain pointers and sizes only; every pointer argument documented as
 * "host or device" is classified with cudaPointerGetAttributes, because the reference's
 * CPU program passes host arrays and its GPU programs pass device arrays to the same names.
 *
 * Conventions: every function returns PQP_OK (0) or a negative pqp_error (the reference
 * calls exit() on a failed cudaMalloc, PQP_GPU_optimized.cu:85-89; this library never
 * exits).  The caller owns all host buffers; a handle owns its device memory and one CUDA
 * stream, is bound to one device, and is not thread-safe.  Calls are synchronous on return
 * unless the name ends in _async.  There is NO CPU fallback: without a CUDA device every
 * compute entry point returns PQP_ERR_NO_DEVICE.
 *
 * All matrices are fp32 row-major, exactly as the reference's input() leaves them
 * (PQP_CPU.c:757-930).
 */
#ifndef PQP_B200_H
#define PQP_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef enum pqp_error {
	PQP_OK = 0,
	PQP_ERR_INVALID = -1,      /* bad argument (NULL, non-positive size, unknown option) */
	PQP_ERR_NO_DEVICE = -2,    /* no usable CUDA device / wrong architecture */
	PQP_ERR_CUDA = -3,         /* a CUDA runtime call failed; see pqp_last_cuda_error() */
	PQP_ERR_ALLOC = -4,        /* host or device allocation failed */
	PQP_ERR_IO = -5,           /* file missing or malformed */
	PQP_ERR_UNSUPPORTED = -6   /* shape outside what the kernels were built for */
} pqp_error;

/*
 * Problem sizes.  The reference fixes them with #defines (PQP_CPU.c:13-17) and derives
 * M = pHorizon*nInput, N = 4*pHorizon*nInput (PQP_CPU.c:940-941).  Here M and N are explicit
 * so the testing/ generator shapes (any M, N) fit too; pqp_dims_mpc() fills them the
 * reference's way.  nState == 0 means "Fp is a constant vector" (the testing/ file format).
 */
typedef struct pqp_dims {
	int M;        /* primal variables            (pHorizon*nInput) */
	int N;        /* constraints = dual variables (4*pHorizon*nInput) */
	int nState;   /* length of x                 */
	int nDisH;    /* length of D                 (nDis*pHorizon) */
	int pHorizon, nInput, nOutput, nDis; /* informational; used by the example loader */
} pqp_dims;

/*
 * Host-side problem data = the arrays PQP_CPU.c's main() holds after input() (PQP_CPU.c:988).
 * Fp(x) = Fp1*D + Fp2*x - Fp3 (computeFp, PQP_CPU.c:373-382).  When nState == 0 the constant
 * vector Fp (length M) is used instead and Fp1/Fp2/Fp3 may be NULL.  Mp1..Mp6 (computeMp,
 * PQP_CPU.c:395-428) only shift the reported costs; they may be NULL (then Mp(x) = Mp0).
 * Z and Theta are loaded by the reference and never used (PQP_CPU.c:889-911); kept for the loader.
 */
typedef struct pqp_host_problem {
	const float *Qp_inv; /* [M x M]       */
	const float *Gp;     /* [N x M]       */
	const float *Kp;     /* [N]           */
	const float *Fp1;    /* [M x nDisH]   */
	const float *Fp2;    /* [M x nState]  */
	const float *Fp3;    /* [M]           */
	const float *D;      /* [nDisH] default disturbance */
	const float *Mp1;    /* [nState x nState] */
	const float *Mp2;    /* [nDisH x nState]  */
	const float *Mp3;    /* [nDisH x nDisH]   */
	const float *Mp4;    /* [nState] */
	const float *Mp5;    /* [nDisH]  */
	const float *Mp6;    /* [1]      */
	const float *Fp;     /* [M]   constant Fp when nState == 0 (testing/ format) */
	float Mp0;           /*       constant Mp when Mp1 == NULL  (testing/ format) */
	const float *x;      /* [nState] the instance's own state (example/x.txt), optional */
	const float *Z;      /* unused by the algorithm */
	const float *Theta;  /* unused by the algorithm */
	/* State-dependent constraint offsets (the reference never forms them): when non-NULL,
	 *     Kp(x, D) = Kp + Kx*x + Kd*D        hence      Fd(x, D) = GQ*Fp(x, D) + Kp + Kx*x + Kd*D,
	 * added after computeFd's sum (PQP_CPU.c:456-460) in the order written, k ascending.  NULL (the default, and what every
	 * loader returns) is PQP_CPU.c's constant Kp.  pqp_output_offsets() builds them from Z / Theta.  They stay the caller's:
	 * pqp_free_problem() leaves them alone. */
	const float *Kx;     /* [N x nState] or NULL */
	const float *Kd;     /* [N x nDisH]  or NULL */
} pqp_host_problem;

typedef enum pqp_order {
	/* split-K / warp-shuffle summation; tensor cores where the shape makes a contraction.
	 * Results within the stated tolerance of PQP_CPU.c (max normwise relative error 1e-5). */
	PQP_ORDER_FAST = 0,
	/* every sum in PQP_CPU.c's own order (k ascending, separately rounded mul and add):
	 * bit-identical to the reference built without FMA contraction.  Parity mode. */
	PQP_ORDER_STRICT = 1
} pqp_order;

typedef struct pqp_opts {
	float theta_floor; /* PQP_CPU.c:240 literal 5 (testing/ harness uses 100)   default 5    */
	float y_init;      /* PQP_CPU.c:710                                         default 1000 */
	float erc, eac, eaj, erj; /* PQP_CPU.c:19-22                                 default 1e-6 */
	int order;         /* pqp_order                                             default FAST */
	int device;        /* CUDA device ordinal, -1 = current                     default -1   */
	int max_iters;     /* cap for run-to-tolerance mode                         default 100000 */
	int check_every;   /* run-to-tolerance: test every this many iterations     default 8    */
	int batch_capacity;/* problems the workspace is sized for (grows on demand) default 1    */
	int use_tensor_cores; /* FAST mode: 0 fp32 SIMT everywhere; 1 (default) tcgen05: 3xTF32 for the setup GEMMs and int8 digit
			       * planes with exact int32 accumulation for the batched loop (accuracy of PQP_CPU.c's own fp32);
			       * 2 the batched loop on 3xTF32 instead (its truncating accumulator leaves it above the 1e-5 parity
			       * tolerance after many updates: opt-in, for comparison) */
	int l2_persist;    /* 1: pin as much of Q as the device allows in L2 (GEMV regime) default 1 */
	int exploit_symmetry; /* FAST mode, one problem, fixed count: 1 (default) when the fp32 Qd of the handle is symmetric element
			       * for element (tested once on the device) the loop reads its upper triangle only -- half the bytes per
			       * update, same sums in a different order; 0 always read the full matrix.  Never applies to a Qd that is
			       * not exactly symmetric.  pqp_setup (not pqp_setup_dual), FAST order: the Qd it
			       * builds from Gp Qp_inv Gp' gets one value per pair (Q_ij, Q_ji) -- their mean -- when all pairs agree to
			       * rounding (1e-5 sqrt(Q_ii Q_jj)), so a dense symmetric Qp_inv also reaches the upper-triangle loop */
	int exploit_structure; /* FAST mode, batched loop: 1 (default) when the constraint rows come in +/- pairs as in the reference's
			       * MPC layout (Gp = [G; -G] per half, N = 4*pHorizon*nInput, PQP_CPU.c:941), i.e. the fp32 Qd of the handle
			       * satisfies Qd[i+N/4][j] == -Qd[i][j] == Qd[i][j+-N/4] element for element (tested once on the device), the
			       * tensor-core loop multiplies only the N/2 representative rows and updates each row together with its
			       * partner: the same exact integer sums, half the work.  0: always all N rows.  Never applies to a Qd
			       * without that structure */
	int accelerate;    /* 0 (default): the reference's live loop, updateY2 only.  P > 0 with a fixed count: after every P-th
			       * multiplicative update one acceleration / line-search step y <- y + alpha*ph, ph = max(0, -(Qd y + Fd)),
			       * alpha = -((y'Qd + Fd')ph)/(ph'Qd ph) -- the branch solveQuadraticDual leaves dead behind `if(1)`
			       * (PQP_CPU.c:721-735; computeph :625-630, computealphaY :545-575, updateY1 :579-588), with computeph's
			       * `matrixAdd(ph, ph, ...)` read as `+= Fd` (as written it doubles Qd y and never sees Fd).  Not counted in
			       * the update count, not applied after the last update; ignored when running to tolerance */
} pqp_opts;

/* Per-problem result of a solve (replaces the printf's of PQP_CPU.c:741,1005-1006). */
typedef struct pqp_status {
	int iters;        /* updates applied */
	int converged;    /* run-to-tolerance mode: 1 if the stop test passed */
	float min_slack;  /* min_i g_i with g = Qd y + Fd = Kp - Gp U(y): primal feasibility (checkFeas, PQP_CPU.c:632) */
	float gap;        /* y'g = Jp(U(y)) + Jd(y): duality gap (terminate, PQP_CPU.c:682-684) */
	float Jd;         /* dual cost 1/2 y'Qd y + Fd'y + Md/2 (computeCost, PQP_CPU.c:648) */
	float kkt;        /* || min(y, g) ||_inf */
} pqp_status;

typedef struct pqp_handle pqp_handle;

/* ---- defaults / helpers ------------------------------------------------------------ */
void pqp_default_opts(pqp_opts *o);
/* M = pH*nInput, N = 4*pH*nInput, nDisH = nDis*pH  (PQP_CPU.c:940-941) */
void pqp_dims_mpc(pqp_dims *d, int pHorizon, int nState, int nInput, int nOutput, int nDis);
const char *pqp_strerror(int code);
/*
 * Kx [N x nState], Kd [N x nDisH] from the output maps the reference loads and never uses (Z [nOut x nState], Theta
 * [nOut x nDisH], nOut = nOutput*pHorizon; PQP_CPU.c:889-911), read as the free response of the constrained outputs:
 * out = Z*x + Theta*D + (output rows of Gp)*U.  With the row blocks of the example (N = 4M: U <= umax, -U <= -umin,
 * out <= outmax, -out <= -outmin) the upper output bound moves by -(Z*x + Theta*D) and the lower by +(Z*x + Theta*D):
 * rows [2M, 2M+nOut) of Kx, Kd are -Z, -Theta, rows [3M, 3M+nOut) are +Z, +Theta, all others zero.  Host-only helper;
 * needs N == 4*M and nOut <= M.
 */
int pqp_output_offsets(const pqp_dims *d, const float *Z, const float *Theta, float *Kx, float *Kd);
const char *pqp_last_cuda_error(void);
/* number of usable sm_100 devices (0 without a GPU); never fails */
int pqp_device_count(void);

/* ---- loaders (host only; no GPU needed) --------------------------------------------- */
/*
 * Replaces input(), PQP_CPU.c:757-930: reads <dir>/{Qp_inv,Gp,Kp,Fp1,Fp2,Fp3,Mp1..Mp6,Theta,Z,D,x}.txt,
 * column-major token streams, into freshly malloc'ed row-major arrays.  dims must be set
 * (the files carry no sizes).  Free with pqp_free_problem.
 */
int pqp_load_example(const char *dir, const pqp_dims *dims, pqp_host_problem *out);
/*
 * Replaces the reader of testing/CPU version/PQP_CPU_test.c:936-976 for the files
 * testing/test_generator.c:936-987 writes ("M N", diag(Qp_inv), Fp, Mp, Kp, Gp rows in {0,1,-1}).
 * Reads the file literally: -1 stays -1 and Kp is the file's line 5 (the reference reader
 * overwrites Kp with rand() and maps -1 to +1; documented defects, not reproduced).
 */
int pqp_load_testfile(const char *path, pqp_dims *dims, pqp_host_problem *out);
/*
 * Synthetic instance with the distribution of testing/test_generator.c:936-987 (diag(Qp_inv),
 * Fp, Mp, Kp ~ U[0,100] rounded to 6 decimals as its "%f" does; Gp in {0,+1,-1} w.p. 1/3 each),
 * from a seeded splitmix64 stream instead of srand(time(0)) (test_generator.c:994).
 */
int pqp_generate_testproblem(unsigned long long seed, int M, int N, pqp_dims *dims, pqp_host_problem *out);
/* writes the testing/ file format (test_generator.c:940-987) */
int pqp_write_testfile(const char *path, const pqp_dims *dims, const pqp_host_problem *p);
void pqp_free_problem(pqp_host_problem *p);

/* ---- the solver ----------------------------------------------------------------------- */
/*
 * One-time, x-independent work on the device.  Replaces Gauss_Jordan + convertToDual
 * (PQP_CPU.c:989,994 -> 489-498) and the setup half of solveQuadraticDual (PQP_CPU.c:696-708):
 *   GQ = Gp*Qp_inv [N x M], Qd = GQ*Gp' [N x N], theta_i = max(sum_j max(0,-Qd_ij), floor),
 * uploads Fp1/Fp2/Fp3/Kp/Mp*, allocates every workspace the solve calls need.
 * Qd is stored ONCE, signed; the Q+/Q- split of PQP_CPU.c:524-537 is applied in registers.
 */
int pqp_setup(pqp_handle **out, const pqp_dims *dims, const pqp_host_problem *prob, const pqp_opts *opts);
/*
 * Same, starting from an already-formed dual (the arguments solveQuadraticDual receives,
 * PQP_CPU.c:694): Qd [N x N], optional Gp/Qp_inv for primal recovery (may be NULL).
 * Fd is then supplied per solve through pqp_solve_dual.  Qd: host or device.
 */
int pqp_setup_dual(pqp_handle **out, int N, const float *Qd, int M, const float *Gp, const float *Qp_inv,
		   const pqp_opts *opts);
void pqp_destroy(pqp_handle *h);

/*
 * Solves B problems that share the handle's Qd and differ in their state x (and optionally D).
 * Replaces computeFp + computeFd + solveQuadraticDual (PQP_CPU.c:991,494,996):
 *   Fp_b = Fp1*D_b + Fp2*x_b - Fp3;  Fd_b = GQ*Fp_b + Kp;  Y_b <- y_init (or Y0_b);
 *   iters > 0: exactly `iters` updates (the fixed-count loop of testing/CPU version/PQP_CPU_test.c:717);
 *   iters <= 0: until the stop test on (min_slack, gap, Jd) passes or opts.max_iters.
 * X [B x nState] (ignored when nState == 0), D [B x nDisH] or NULL (problem's D for all),
 * Y0 [B x N] or NULL, Y [B x N] out, st [B] out or NULL.  X, D, Y0, Y: host or device.
 * B == 1 runs the persistent GEMV kernel, B > 1 the batched kernel (tensor cores), which in run-to-tolerance mode evaluates
 * the stop test per problem every opts.check_every updates and freezes each problem at exactly the y that passed: its result
 * equals the fixed-count solve with iters = st[b].iters, whatever its batch mates do.
 */
int pqp_solve_batch(pqp_handle *h, const float *X, const float *D, int B, int iters, const float *Y0,
		    float *Y, pqp_status *st);
/* Same loop with the linear term given directly: Fd [B x N] (host or device). */
int pqp_solve_dual(pqp_handle *h, const float *Fd, int B, int iters, const float *Y0, float *Y, pqp_status *st);
/*
 * pqp_solve_dual with the constant of the dual cost given too: Md [B] (host or device; computeMd, PQP_CPU.c:472-479), so that
 * status.Jd = 1/2 y'Qd y + Fd'y + Md/2 is computeCost's value (PQP_CPU.c:648-666) and the relative gap test of terminate()
 * (|gap| <= erj*|Jd|, PQP_CPU.c:684) runs against the same scale as the reference's.  Md == NULL is pqp_solve_dual.
 */
int pqp_solve_dual_full(pqp_handle *h, const float *Fd, const float *Md, int B, int iters, const float *Y0, float *Y, pqp_status *st);
/*
 * Constraint bounds Kp [N] (host or device) for a handle built by pqp_setup_dual, which receives none: the feasibility test of
 * run-to-tolerance solves then uses compare()'s per-row tolerance max(erc*Kp_i, eac) (PQP_CPU.c:334-343) instead of eac alone.
 * NULL removes them again.  (pqp_setup takes Kp from the problem.)
 */
int pqp_set_constraint_bounds(pqp_handle *h, const float *Kp);
/*
 * U_b = -Qp_inv*(Gp'*Y_b + Fp_b), computeUfromY PQP_CPU.c:352-360, with Fp_b the vectors the
 * last pqp_solve_batch on this handle formed (pass Fp != NULL [B x M] to override, e.g. after
 * pqp_solve_dual).  Y: host or device, U [B x M]: host or device.
 */
int pqp_recover_primal(pqp_handle *h, const float *Y, const float *Fp, int B, float *U);
/* pqp_solve_batch followed by pqp_recover_primal on the device-resident Y (one stream, no host sync between). */
int pqp_solve_batch_primal(pqp_handle *h, const float *X, const float *D, int B, int iters, const float *Y0,
			   float *Y, float *U, pqp_status *st);

/*
 * Receding-horizon warm start (the caller the report describes, PQP_CPU.c:710 always restarts from 1000).
 * The duals of a condensed MPC QP come in four blocks of pHorizon*nInput rows (N = 4*pHorizon*nInput, PQP_CPU.c:941), each
 * ordered by horizon step: Ynext[blk][k] = max(Y[blk][k+1], y_floor) for k < pHorizon-1, and the last step is held.  The floor
 * matters: the multiplicative update can never leave y_i = 0, so a dual that went to zero while its constraint was inactive
 * could not become active again in a later period (the reference avoids the question by restarting from 1000 every time).
 * Y, Ynext [B x N], host or device; Ynext may equal Y.  Feed Ynext as Y0 to the next period's pqp_solve_batch.  Needs
 * dims.pHorizon*nInput*4 == N.  Uses the handle's workspace (call pqp_get_linear_terms before it, not after).
 */
int pqp_shift_duals(pqp_handle *h, const float *Y, int B, float y_floor, float *Ynext);

/*
 * out[a x c] = op(A)[a x b] * op(B)[b x c] on the GPU: the reference's matrixMultiply (PQP_CPU.c:84-147) with its
 * transpose flags (tA: A is stored [b x a]; tB: B is stored [c x b]).  engine: PQP_MM_STRICT = the reference's
 * summation order, bit-identical; PQP_MM_SIMT = fp32 FMA tiles; PQP_MM_TENSOR = tcgen05 3xTF32 (fp32-level accuracy).
 * All pointers host or device; device < 0 = current.
 */
enum { PQP_MM_STRICT = 0, PQP_MM_SIMT = 1, PQP_MM_TENSOR = 2 };
int pqp_matmul(float *out, const float *A, int tA, const float *B, int tB, int a, int b, int c, int engine, int device);
/*
 * One multiplicative update from the reference's own DENSE operands, for single-step tests against updateY2 + updY
 * (PQP_CPU.c:603-618, :590-596): Y_next_i = (Qdn_theta[i,:]*Y + Fdn_i) / (Qdp_theta[i,:]*Y + Fdp_i) * Y_i, every sum k ascending
 * with separately rounded multiply and add, IEEE division -- bit-identical to the reference.  Host or device pointers;
 * matrices [N x N] row-major.  (The solver proper never forms these two matrices: it keeps one signed Qd and theta.)
 */
int pqp_update_y2(float *Y_next, const float *Y, const float *Qdp_theta, const float *Qdn_theta, const float *Fdp, const float *Fdn, int N,
		  int device);

/*
 * The reference's small free functions on the device, in the reference's own summation order (bit-identical to PQP_CPU.c), for
 * callers that hold host arrays and no handle (libpqp_compat.so): all pointers host or device, device < 0 = current.
 *   pqp_compute_fp        Fp = Fp1*D + Fp2*x - Fp3                       computeFp      PQP_CPU.c:373-382
 *   pqp_compute_cost      J  = 1/2 Z'QZ + F'Z + Mc/2  (Mc NULL: 0)        computeCost    PQP_CPU.c:648-666
 *   pqp_compute_md        Md = Fp'Qp_inv Fp - Mp      (Mp NULL: 0)        computeMd      PQP_CPU.c:472-479
 *   pqp_compute_u_from_y  U_b = -Qp_inv (Gp'Y_b + Fp_b), b < B            computeUfromY  PQP_CPU.c:352-360
 */
int pqp_compute_fp(float *Fp, const float *Fp1, const float *Fp2, const float *Fp3, const float *D, const float *x, int M, int nDisH, int nState,
		   int device);
int pqp_compute_cost(float *J, const float *Z, const float *Q, const float *F, const float *Mc, int N, int device);
int pqp_compute_md(float *Md, const float *Fp, const float *Qp_inv, const float *Mp, int M, int device);
int pqp_compute_u_from_y(float *U, const float *Y, const float *Fp, const float *Gp, const float *Qp_inv, int N, int M, int B, int device);

/* ---- introspection (tests, benches) ----------------------------------------------------- */
/* copies out what setup built; any pointer may be NULL.  Qd [N x N], theta [N], GQ [N x M] (host buffers) */
int pqp_get_dual(pqp_handle *h, float *Qd, float *theta, float *GQ);
/* Fd [B x N] and Fp [B x M] of the last solve (host buffers, may be NULL) */
int pqp_get_linear_terms(pqp_handle *h, int B, float *Fd, float *Fp);
/* the handle's cudaStream_t (as void*), so callers can bracket calls with their own CUDA events */
void *pqp_get_stream(pqp_handle *h);
/* device time of the iteration kernel(s) of the last solve, in ms (CUDA events on the handle's stream) */
float pqp_last_solve_ms(pqp_handle *h);
/* device time, in ms, of the two GEMMs of convertToDual (GQ = Gp*Qp_inv, Qd = GQ*Gp'; PQP_CPU.c:492, :442) inside pqp_setup
 * (2*N*M*M + 2*N*N*M flop); 0 for a handle built by pqp_setup_dual */
float pqp_setup_gemm_ms(pqp_handle *h);
/* 1 when pqp_setup built Qd from the tiles of its upper triangle (Qp_inv symmetric element for element, FAST order, tensor-core
 * setup, exploit_symmetry): every element above the diagonal is one sum stored twice; the multiplied share of the N x N product
 * is the tiles of 128 x 192 that touch the upper triangle.  0 otherwise. */
int pqp_setup_mirrored(pqp_handle *h);
/* how many of this library's kernels the handle has launched so far */
long long pqp_launch_count(pqp_handle *h);
/* name of the iteration kernel the last solve used ("gemv_tma_stream", "gemv_strict", "batched_imma", "batched_simt", ...) */
const char *pqp_last_kernel(pqp_handle *h);
/* device copy of Qd (row stride = *ld floats), for zero-copy callers; valid until pqp_destroy */
const float *pqp_device_qd(pqp_handle *h, int *ld);

#ifdef __cplusplus
}
#endif
#endif /* PQP_B200_H */
