/*
 * pqp_internal.h -- launcher prototypes shared by the .cu files of libpqp_b200.so.
 * Nothing here is part of the public ABI (that is include/pqp.h).
 */
#ifndef PQP_INTERNAL_H
#define PQP_INTERNAL_H

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "pqp.h"

#define PQP_GEMV_THREADS 512
#define PQP_GEMV_WARPS (PQP_GEMV_THREADS / 32)

static inline int pqp_round_up(int x, int m) { return (x + m - 1) / m * m; }

/*
 * Experiment knobs (PQP_* environment variables) are read ONCE per handle, when it is created: pqp_setup /
 * pqp_setup_dual snapshot every PQP_* variable into the handle, each public entry point makes its handle's snapshot the calling
 * thread's current one, and every kernel-selection / geometry decision asks pqp_env() -- no getenv() on the solve path.
 * Outside a handle (pqp_matmul and the other stateless helpers) pqp_env() reads the process environment.
 */
const char *pqp_env(const char *name);

/* ---- setup kernels (pqp_setup_kernels.cu) ------------------------------------------------ */
/* C[a x c] (ldc) = A[a x b] (lda) * op(B); transB: B stored [c x b] (ldb) else [b x c] (ldb).
 * strict: one thread per element, k ascending, separately rounded mul/add (PQP_CPU.c:84-147). */
cudaError_t pqp_launch_matmul_strict(float *C, int ldc, const float *A, int lda, const float *B, int ldb, int transB,
				     int a, int b, int c, cudaStream_t s);
/* fp32 SIMT tiled GEMM (FAST mode without tensor cores; also the cross-check for the tcgen05 path) */
cudaError_t pqp_launch_matmul_simt(float *C, int ldc, const float *A, int lda, const float *B, int ldb, int transB,
				   int a, int b, int c, cudaStream_t s);
/* tcgen05 3xTF32 GEMM (pqp_gemm_umma.cu): C[a x c] = A[a x b] * Bt[c x b]', both operands K-major */
cudaError_t pqp_launch_gemm_umma(float *C, int ldc, const float *A, int lda, const float *Bt, int ldb, int a, int b, int c,
				 cudaStream_t s);
/* the same product as a warp-specialised pipeline (pqp_gemm_umma_ws.cu): loader/converter warps, one MMA warp, drain warps, 128 x 192 tiles */
int pqp_gemm_umma_ws_wanted(int a, int b, int c);
/* sym_bad != NULL and a == c: the product is symmetric in exact arithmetic provided *sym_bad (device memory: the count of unequal pairs
 * of the inner symmetric factor, pqp_launch_sym_check) is zero; then only the upper-triangle tiles are multiplied and mirrored */
cudaError_t pqp_launch_gemm_umma_ws(float *C, int ldc, const float *A, int lda, const float *Bt, int ldb, int a, int b, int c,
				    const unsigned *sym_bad, cudaStream_t s);
/* theta_i = max(sum_j max(0,-Q_ij), floor); strict: thread per row, j ascending */
cudaError_t pqp_launch_theta(float *theta, const float *Q, int ldq, int N, float floor_, int strict, cudaStream_t s);
/* out[c x r] (ldo) = in[r x c] (ldi) transposed */
cudaError_t pqp_launch_transpose(float *out, int ldo, const float *in, int ldi, int r, int c, cudaStream_t s);

/* ---- per-solve small kernels (pqp_setup_kernels.cu) --------------------------------------- */
/* Fp[b] = Fp1*D[b] + Fp2*X[b] - Fp3, reference order (computeFp, PQP_CPU.c:373-382).  D_stride 0: shared D.
 * nState == 0: Fp[b] = Fp_const. */
cudaError_t pqp_launch_fp(float *Fp, const float *Fp1, const float *Fp2, const float *Fp3, const float *Fp_const,
			  const float *D, int D_stride, const float *X, int B, int M, int nd, int nState,
			  cudaStream_t s);
/* Fd[b] = GQ*Fp[b] + Kp (computeFd, PQP_CPU.c:456-460); strict: k ascending per thread */
cudaError_t pqp_launch_fd(float *Fd, const float *GQ, const float *Fp, const float *Kp, int B, int N, int M,
			  int strict, cudaStream_t s);
/* Fd[b] += Kx*X[b] + Kd*D[b] (state-dependent constraint offsets; either matrix may be NULL); k ascending, separately rounded */
cudaError_t pqp_launch_fd_offsets(float *Fd, const float *Kx, const float *X, int nState, const float *Kd, const float *D, int D_stride,
				  int nd, int B, int N, cudaStream_t s);
/* Md[b] = Fp' Qp_inv Fp - Mp(x_b) (computeMd PQP_CPU.c:472-479, computeMp :395-428); Mp1==NULL: Mp = Mp0 */
cudaError_t pqp_launch_md(float *Md, float *tmp /* [B x M] scratch */, const float *Fp, const float *Qp_inv, const float *Mp1, const float *Mp2,
			  const float *Mp3, const float *Mp4, const float *Mp5, const float *Mp6, float Mp0,
			  const float *D, int D_stride, const float *X, int B, int M, int nd, int nState, cudaStream_t s);
/* U[b] = -(Qp_inv*(Gp'*Y[b] + Fp[b])) (computeUfromY, PQP_CPU.c:352-360); tmp [B x M] scratch */
cudaError_t pqp_launch_recover(float *U, float *tmp, const float *Y, int ldy, const float *Fp, const float *Gp,
			       const float *Qp_inv, int B, int N, int M, int strict, cudaStream_t s);
/* receding-horizon shift of the duals (out must not alias in) */
cudaError_t pqp_launch_shift_duals(float *out, const float *in, int B, int pH, int nI, float y_floor, cudaStream_t s);
/* Yn = Y .* (Qn*Y + Fdn) ./ (Qp*Y + Fdp), reference order (updateY2 + updY); all device pointers */
cudaError_t pqp_launch_update_y2_dense(float *Yn, const float *Y, const float *Qp, const float *Qn, const float *Fdp, const float *Fdn, int N,
				       cudaStream_t s);
/* mode 0: out[0] = 1/2 z'Az + F'z + m/2 (computeCost, PQP_CPU.c:648-666); mode 1: out[0] = z'Az - m (computeMd, :472-479);
 * reference order and promotions; tmp [n] scratch; all device pointers */
cudaError_t pqp_launch_quad_form(float *out, float *tmp, const float *z, const float *A, const float *F, const float *m, int n, int mode,
				 cudaStream_t s);
/* one acceleration / line-search step on B duals in place (computeph + computealphaY + updateY1, PQP_CPU.c:625-630, :545-588, with
 * computeph's self-addition read as += Fd): y += alpha*max(0, -(Qd y + Fd)); reference order; ws [3][B][N] scratch; 4 launches */
cudaError_t pqp_launch_accel_step(float *Y, int ldy, const float *Q, int ldq, const float *Fd, float *ws, int B, int N, cudaStream_t s);
/* fill n floats */
cudaError_t pqp_launch_fill(float *p, float v, size_t n, cudaStream_t s);

/* ---- single-problem iteration (pqp_gemv.cu) ------------------------------------------------- */
typedef struct pqp_gemv_args {
	const float *Q;     /* [N x ldq] signed Qd, padding columns zero */
	const float *QT;    /* [N x ldq] transpose (strict mode only) */
	int ldq, N;
	const float *theta; /* [N] */
	const float *Fd;    /* [N] */
	const float *Kp;    /* [N] or NULL */
	const float *Md;    /* [1] or NULL */
	float *ybuf0, *ybuf1; /* [ldq] each, padding zero; ybuf0 holds y_0 */
	int iters;          /* > 0: fixed count; <= 0: tolerance mode up to max_iters */
	int max_iters, check_every;
	float erc, eac, eaj, erj;
	unsigned *barrier;  /* zeroed before launch */
	float *partials;    /* [2][grid][8] */
	pqp_status *status; /* device, 1 entry */
	int *result_buf;    /* device int: which ybuf holds the answer */
	int grid;           /* CTAs (cooperative, <= SM count) */
	int resident_rows;  /* rows of each CTA's slab kept in shared memory across iterations */
} pqp_gemv_args;

cudaError_t pqp_gemv_smem_bytes(int N, int ldq, int grid, int resident_rows, size_t *bytes);
cudaError_t pqp_launch_gemv_persistent(const pqp_gemv_args *a, cudaStream_t s);
/* TMA-staged variant (pqp_gemv_tma.cu), fixed-count solves */
int pqp_gemv_tma_plan(int N, int ldq, int grid, size_t smem_budget, int *stages, int *resident, int *yc);
/* pk0/pk1: packet vectors [ldq x 8 B] for the flag-in-data y exchange, or NULL for the counter barrier; the
 * result is left in ybuf1 when packets are used, in ybuf[iters&1] otherwise */
cudaError_t pqp_launch_gemv_tma(const pqp_gemv_args *a, int stages, int resident, int pinned, int yc, void *pk0, void *pk1,
				cudaStream_t s);
/* one-CTA variant for N <= 128 (pqp_gemv_cta.cu): fixed count or run to tolerance; result left in ybuf1, status written */
int pqp_gemv_cta_supported(int N);
cudaError_t pqp_launch_gemv_cta(const pqp_gemv_args *a, cudaStream_t s);
/* B problems sharing Q, one block each: Fd / ybuf0 / ybuf1 (may be the same array) / Md / status of problem b at b * stride */
cudaError_t pqp_launch_gemv_cta_batch(const pqp_gemv_args *a, int B, int fd_stride, int y_stride, cudaStream_t s);
/* one-cluster variant for 128 < N <= 768 (pqp_gemv_cluster.cu): 16 (or 8) CTAs, y exchanged through distributed shared memory; fixed
 * count or run to tolerance; result left in ybuf1, status written */
int pqp_gemv_cluster_supported(int N);
cudaError_t pqp_launch_gemv_cluster(const pqp_gemv_args *a, cudaStream_t s);
/* B problems sharing Q, one cluster each: Fd / ybuf0 / ybuf1 (may be the same array) / Md / status of problem b at b * stride */
cudaError_t pqp_launch_gemv_cluster_batch(const pqp_gemv_args *a, int B, int fd_stride, int y_stride, cudaStream_t s);
/* register-resident variant for small N (pqp_gemv_small.cu); result left in ybuf1 */
int pqp_gemv_small_plan(int N, int ldq, int grid, int *wpr, int *cpt);
cudaError_t pqp_launch_gemv_small(const pqp_gemv_args *a, int wpr, int cpt, void *pk0, void *pk1, cudaStream_t s);
/* upper-triangle variant for a symmetric Qd (pqp_gemv_sym.cu), fixed-count solves; result left in ybuf1 */
typedef struct pqp_sym_plan {
	int nb, U, maxseg;           /* 128-wide blocks, 64x128 units of the upper triangle, most strips one CTA touches */
	int stages, resident, pinned;
	int tmem;                    /* units per CTA parked in tensor memory (0..8) */
	float *units;                /* device [U][64][128] */
	int *cta_u0, *cta_j0, *strip_c0, *strip_c1; /* device tables (pqp_gemv_sym_tables) */
	void *rowpart, *colpart;     /* device packet arrays */
	size_t rowpart_bytes, colpart_bytes;
} pqp_sym_plan;
void pqp_gemv_sym_counts(int N, int *nb, int *U);
size_t pqp_gemv_sym_units_bytes(int N);
int pqp_gemv_sym_plan(int N, int grid, size_t smem_budget, int *stages, int *resident);
int pqp_gemv_sym_tables(int N, int G, int *cta_u0, int *cta_j0, int *strip_c0, int *strip_c1);
cudaError_t pqp_launch_sym_check(const float *Q, int ldq, int N, unsigned *mismatch, cudaStream_t s);
/* pqp_setup, FAST order: when every pair (Q_ij, Q_ji) agrees to tol * sqrt(|Q_ii Q_jj|), both become their mean (else Q is left alone) */
cudaError_t pqp_launch_sym_mean(float *Q, int ldq, int N, float tol, unsigned *mismatch, cudaStream_t s);
cudaError_t pqp_launch_build_sym_units(float *units, const float *Q, int ldq, int N, cudaStream_t s);
cudaError_t pqp_launch_gemv_sym(const pqp_gemv_args *a, const pqp_sym_plan *pl, void *pk0, void *pk1, cudaStream_t s);
/* strict: one launch per iteration, thread i owns row i and walks k ascending over QT */
cudaError_t pqp_launch_gemv_strict_step(const pqp_gemv_args *a, const float *y_in, float *y_out, cudaStream_t s);
/* evaluation of the status quantities for one y (any mode); viol_out [B] (device, may be NULL) receives
 * max_i(-g_i - max(erc*Kp_i, eac)), the quantity compare() tests per row (PQP_CPU.c:334-343): feasible iff <= 0 */
cudaError_t pqp_launch_status(pqp_status *st, const float *Q, int ldq, int N, const float *Y, int ldy, const float *Fd,
			      const float *Md, const float *Kp, float erc, float eac, int B, int iters, float *viol_out, cudaStream_t s);

/* ---- batched iteration (pqp_batched.cu) ------------------------------------------------------ */
/* k-major pre-split operands: QpT/QnT [Kpad x Ipad], [k][i] = max(0,+-Qd[i][k]) + theta_i*(i==k) */
cudaError_t pqp_launch_build_split_t(float *QpT, float *QnT, int Kpad, int Ipad, const float *Q, int ldq,
				     const float *theta, int N, cudaStream_t s);
/* iters updates of B problems, Y [B x N] in place, Fd [B x N]; fp32 SIMT */
cudaError_t pqp_launch_batched_simt_split(const float *QpT, const float *QnT, int Kpad, int Ipad, int N, int B,
					  const float *Fd, float *Y, int iters, cudaStream_t s);
int pqp_batched_simt_supported(int N);
/* tcgen05 3xTF32 version (pqp_batched_umma.cu): pre-split, pre-tiled A operand + the loop */
int pqp_batched_umma_supported(int N);
size_t pqp_batched_umma_tiles_bytes(int N);
cudaError_t pqp_launch_build_umma_tiles(void *tiles, const float *Q, int ldq, const float *theta, int N, cudaStream_t s);
cudaError_t pqp_launch_batched_umma(const void *tiles, int N, int B, const float *Fd, float *Y, int iters, int cluster,
				    cudaStream_t s);
/* tcgen05 int8 digit-plane version (pqp_batched_imma.cu): error-free integer accumulation */
int pqp_batched_imma_supported(int N);
size_t pqp_batched_imma_tiles_bytes(int N);
size_t pqp_batched_imma_rowc_bytes(int N);
cudaError_t pqp_launch_build_imma_tiles(void *tiles, void *rowc, const float *Q, int ldq, const float *theta, int N, cudaStream_t s);
/* iters <= 0: run to the stop test of terminate() (tol != NULL), checked every check_every updates, per problem */
typedef struct pqp_imma_tol {
	int max_iters, check_every;
	float erc, eac, eaj, erj;
	const float *Kp;    /* [N] or NULL */
	const float *Md;    /* [B] or NULL */
	pqp_status *status; /* [B] device */
} pqp_imma_tol;
cudaError_t pqp_launch_batched_imma(const void *tiles, const void *rowc, int N, int B, const float *Fd, float *Y, int iters, int nb,
				    int cluster, size_t smem_optin, const pqp_imma_tol *tol, cudaStream_t s);
/* M tiles, padded K steps and K steps per ring stage of the digit-plane tile array */
void pqp_imma_geometry(int N, int *MT, int *NKS, int *ksc);
/* the same loop with the rows of Q split over a CTA pair sharing 64 problems (pqp_batched_imma_pair.cu); fixed count only */
int pqp_batched_imma_pair_supported(int N);
cudaError_t pqp_launch_batched_imma_pair(const void *tiles, const void *rowc, int N, int B, const float *Fd, float *Y, int iters,
					 size_t smem_optin, cudaStream_t s);
/* the loop for a Qd with the +/- row-pair structure of a box-constrained MPC dual (pqp_batched_imma_paired.cu): half the tensor
 * work, two interleaved groups of 32 problems per CTA pair; its own tile / row-constant arrays; fixed count only */
int pqp_batched_imma_paired_supported(int N);
size_t pqp_batched_imma_paired_tiles_bytes(int N);
size_t pqp_batched_imma_paired_rowc_bytes(int N);
cudaError_t pqp_launch_pair_struct_check(const float *Q, int ldq, int N, unsigned *bad, cudaStream_t s);
cudaError_t pqp_launch_build_imma_tiles_paired(void *tiles, void *rowc, const float *Q, int ldq, const float *theta, int N, cudaStream_t s);
/* the per-solve small steps inside that kernel (north star: "fused into the same kernel ... fused epilogue"):
 *   GQ != NULL  prologue: Fp_b = Fp1 D_b + Fp2 x_b - Fp3 (computeFp) and Fd_b = GQ Fp_b + Kp (computeFd), reference order, also
 *               written to Fp_out [B x M] / Fd_out [B x N]; the launcher's Fd argument is then not read
 *   U != NULL   epilogue: U_b = -Qp_inv (Gp' y_b + Fp_b) (computeUfromY), reference order; Fp [B x M] (= Fp_out when both are fused)
 *   y0_const    every dual starts at y_init; Y is output only */
typedef struct pqp_paired_fuse {
	const float *X, *D, *Fp1, *Fp2, *Fp3, *Fp_const, *GQ, *Kp;
	int nS, nd, D_stride, M;
	float *Fp_out, *Fd_out;
	const float *Gp, *Qp_inv, *Fp;
	float *U;
	int y0_const;
	float y_init;
	/* run to tolerance in chunks: resume state in / out, evaluation partials out (NULL: a plain fixed-count launch), compare()'s tolerances */
	const float *m_resume;
	float *m_out, *eval_part;
	const float *tol_Kp;
	float erc, eac;
} pqp_paired_fuse;
size_t pqp_paired_eval_part_floats(int B);
cudaError_t pqp_launch_paired_tol_decide(const float *part, const float *Md, pqp_status *st, unsigned *frozen, unsigned *newly, unsigned *remaining,
					 float *Yres, const float *Y, int B, int N, int count, int max_iters, float eaj, float erj, cudaStream_t s);
int pqp_batched_imma_paired_can_fuse(int N, int M, size_t smem_optin);
cudaError_t pqp_launch_batched_imma_paired(const void *tiles, const void *rowc, int N, int B, const float *Fd, float *Y, int iters,
					   size_t smem_optin, const pqp_paired_fuse *fz, cudaStream_t s);
#define PQP_BATCH_KPAD 16
#define PQP_BATCH_IPAD 128

#endif
