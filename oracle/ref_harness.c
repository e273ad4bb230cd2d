/*
 * oracle/ref_harness.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Builds the UNMODIFIED reference CPU program into a shared library by #including it
 * where it lies (REF_SRC = /root/reference/PQP_CPU.c, passed by oracle/Makefile) with
 * its main() renamed.  No reference source is copied into this repository; the output
 * goes to oracle/_ref/ (git-ignored, travels to the GPU box like any other built .so).
 *
 * The wrappers below only forward to the reference's own functions so that ctypes can
 * reach them with a stable prefix; the one loop written here (ref_iterate) is the
 * fixed-count loop of testing/CPU version/PQP_CPU_test.c:717-744 expressed with the
 * reference's own updateY2 + copyMatrix, because PQP_CPU.c's solveQuadraticDual only
 * has the run-to-stop loop.
 *
 * With -DREAL=double and REF_SRC pointing at a sed-made float->double twin (made in a
 * temp dir by the Makefile, never stored) this gives the "float64 twin" of SURVEY 8(c).
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <fcntl.h>

#ifndef REAL
#define REAL float
#endif
#ifndef SFX
#define SFX f32
#endif
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)

#define main pqp_cpu_ref_main
#include REF_SRC
#undef main

void FN(ref_dims)(int *out)
{
	out[0] = pHorizon; out[1] = nState; out[2] = nInput; out[3] = nOutput; out[4] = nDis;
}

/* input() opens "./example/..." relative to the cwd (PQP_CPU.c:764) */
int FN(ref_load_example)(const char *parent_dir, REAL *Qp_inv, REAL *Fp1, REAL *Fp2, REAL *Fp3,
			 REAL *Mp1, REAL *Mp2, REAL *Mp3, REAL *Mp4, REAL *Mp5, REAL *Mp6, REAL *Gp,
			 REAL *Kp, REAL *x, REAL *D, REAL *Theta, REAL *Z)
{
	char old[4096];
	if (!getcwd(old, sizeof old)) return -1;
	if (chdir(parent_dir)) return -2;
	input(Qp_inv, Fp1, Fp2, Fp3, Mp1, Mp2, Mp3, Mp4, Mp5, Mp6, Gp, Kp, x, D, Theta, Z);
	if (chdir(old)) return -3;
	return 0;
}

void FN(ref_matmul)(REAL *out, REAL *A, int tA, REAL *B, int tB, int a, int b, int c)
{
	matrixMultiply(out, A, tA, B, tB, a, b, c);
}
void FN(ref_compute_fp)(REAL *Fp, REAL *Fp1, REAL *Fp2, REAL *Fp3, REAL *D, REAL *x)
{
	computeFp(Fp, Fp1, Fp2, Fp3, D, x);
}
void FN(ref_compute_mp)(REAL *Mp, REAL *Mp1, REAL *Mp2, REAL *Mp3, REAL *Mp4, REAL *Mp5, REAL *Mp6,
			REAL *D, REAL *x)
{
	computeMp(Mp, Mp1, Mp2, Mp3, Mp4, Mp5, Mp6, D, x);
}
void FN(ref_convert_to_dual)(REAL *Qd, REAL *Fd, REAL *Md, REAL *Qp_inv, REAL *Gp, REAL *Kp, REAL *Fp,
			     REAL *Mp, int N, int M)
{
	convertToDual(Qd, Fd, Md, Qp_inv, Gp, Kp, Fp, Mp, N, M);
}
void FN(ref_gauss_jordan)(REAL *A, REAL *res, int n) { Gauss_Jordan(A, res, n); }
void FN(ref_recover_u)(REAL *U, REAL *Y, REAL *Fp, REAL *Gp, REAL *Qp_inv, int N, int M)
{
	computeUfromY(U, Y, Fp, Gp, Qp_inv, N, M);
}
REAL FN(ref_cost)(REAL *Z, REAL *Q, REAL *F, REAL *Mscalar, int n) { return computeCost(Z, Q, F, Mscalar, n); }
int FN(ref_terminate)(REAL *Y, REAL *Qd, REAL *Fd, REAL *Md, REAL *U, REAL *Qp, REAL *Qp_inv, REAL *Fp,
		      REAL *Mp, REAL *Gp, REAL *Kp, int N, int M)
{
	return terminate(Y, Qd, Fd, Md, U, Qp, Qp_inv, Fp, Mp, Gp, Kp, N, M);
}
void FN(ref_update_y2)(REAL *Yn, REAL *Y, REAL *Qdp_theta, REAL *Qdn_theta, REAL *Fd, REAL *Fdp, REAL *Fdn,
		       int N)
{
	updateY2(Yn, Y, Qdp_theta, Qdn_theta, Fd, Fdp, Fdn, N);
}

/* theta diag (as an N-vector) and the two split matrices, from the reference's own helpers */
void FN(ref_split)(REAL *Qdp_theta, REAL *Qdn_theta, REAL *theta_vec, REAL *Qd, int N)
{
	REAL *theta = newMatrix(N, N);
	computeTheta(theta, Qd, N);
	computeQdp_theta(Qdp_theta, Qd, theta, N);
	computeQdn_theta(Qdn_theta, Qd, theta, N);
	for (int i = 0; i < N; i++) theta_vec[i] = theta[(size_t)i * N + i];
	free(theta);
}

/* K updates from the Y passed in; init_y nonzero resets Y to 1000.0 first (PQP_CPU.c:710) */
void FN(ref_iterate)(REAL *Y, REAL *Qd, REAL *Fd, int N, long K, int init_y)
{
	REAL *theta = newMatrix(N, N);
	REAL *Qdp_theta = newMatrix(N, N);
	REAL *Qdn_theta = newMatrix(N, N);
	REAL *Y_next = newMatrix(N, 1);
	REAL *Fdn = newMatrix(N, 1);
	REAL *Fdp = newMatrix(N, 1);
	matrixPos(Fdp, Fd, N, 1);
	matrixNeg(Fdn, Fd, N, 1);
	computeTheta(theta, Qd, N);
	computeQdp_theta(Qdp_theta, Qd, theta, N);
	computeQdn_theta(Qdn_theta, Qd, theta, N);
	free(theta);
	if (init_y) initMat(Y, 1000.0, N);
	for (long h = 0; h < K; h++) {
		updateY2(Y_next, Y, Qdp_theta, Qdn_theta, Fd, Fdp, Fdn, N);
		copyMatrix(Y, Y_next, N, 1);
	}
	free(Qdp_theta); free(Qdn_theta); free(Y_next); free(Fdn); free(Fdp);
}

/* run fn with stdout captured into buf */
static int capture_begin(int *saved, char *tmpl)
{
	fflush(stdout);
	*saved = dup(1);
	int fd = mkstemp(tmpl);
	if (fd < 0) return -1;
	dup2(fd, 1);
	close(fd);
	return 0;
}
static void capture_end(int saved, const char *tmpl, char *buf, int cap)
{
	fflush(stdout);
	dup2(saved, 1);
	close(saved);
	FILE *f = fopen(tmpl, "r");
	size_t n = f ? fread(buf, 1, cap - 1, f) : 0;
	buf[n] = 0;
	if (f) fclose(f);
	unlink(tmpl);
}

/* solveQuadraticDual itself (PQP_CPU.c:694-750); returns the h it prints */
long FN(ref_solve_converge)(REAL *Y, REAL *Qd, REAL *Fd, REAL *Md, REAL *U, REAL *Qp, REAL *Qp_inv, REAL *Fp,
			    REAL *Mp, REAL *Gp, REAL *Kp, int N, int M)
{
	char tmpl[] = "/tmp/pqp_ref_XXXXXX", buf[512];
	int saved;
	if (capture_begin(&saved, tmpl)) return -1;
	solveQuadraticDual(Y, Qd, Fd, Md, U, Qp, Qp_inv, Fp, Mp, Gp, Kp, N, M);
	capture_end(saved, tmpl, buf, sizeof buf);
	const char *eq = strrchr(buf, '=');
	return eq ? atol(eq + 1) : -2;
}

/* the reference program end to end on <parent_dir>/example; its stdout goes to buf */
int FN(ref_main)(const char *parent_dir, char *buf, int cap)
{
	char old[4096], tmpl[] = "/tmp/pqp_ref_XXXXXX";
	int saved;
	if (!getcwd(old, sizeof old)) return -1;
	if (chdir(parent_dir)) return -2;
	if (capture_begin(&saved, tmpl)) return -3;
	pqp_cpu_ref_main();
	capture_end(saved, tmpl, buf, cap);
	if (chdir(old)) return -4;
	return 0;
}
