/*
 * tools/pqp_example.c -- the reference's program flow (PQP_CPU.c:935-1040: input -> Gauss_Jordan -> computeFp -> computeMp ->
 * convertToDual -> solveQuadraticDual -> computeUfromY -> print Jp, Jd, U*) on the B200 solver, through the reference's own
 * function names (libpqp_compat.so).  Prints exactly what PQP_CPU.c prints for example/ (SURVEY 8b.3).
 *
 *   pqp_example [example_dir] [--fast] [--iters K]
 *
 * Default: PQP_ORDER_STRICT (bit-identical arithmetic) and K = 312 updates, the count at which PQP_CPU.c itself stops on
 * example/ ("iterations = 313").  The reference's stop test passes on rounding noise of Jp + Jd, so its count is
 * a property of that build's float rounding, not of the problem; a fixed count is how its own testing/ harness runs too
 * (PQP_CPU_test.c:717).  Only the two scalar helpers the library leaves on the host (Qp for the printed Jp, Mp) are computed
 * here, in the reference's order: Gauss_Jordan PQP_CPU.c:251-326, computeMp :395-428.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pqp.h"

/* the reference's own entry points, provided by libpqp_compat.so */
void pqp_compat_set_dims(int pHorizon, int nState, int nInput, int nOutput, int nDis);
void pqp_compat_set_order(int order);
void pqp_compat_set_fixed_iters(long k);
void computeFp(float *Fp, float *Fp1, float *Fp2, float *Fp3, float *D, float *x);
void convertToDual(float *Qd, float *Fd, float *Md, float *Qp_inv, float *Gp, float *Kp, float *Fp, float *Mp, int N, int M);
void solveQuadraticDual(float *Y, float *Qd, float *Fd, float *Md, float *U, float *Qp, float *Qp_inv, float *Fp, float *Mp, float *Gp,
			float *Kp, int N, int M);
void computeUfromY(float *U, float *Y, float *Fp, float *Gp, float *Qp_inv, int N, int M);
float computeCost(float *Z, float *Q, float *F, float *M, int N);

/* inverse by Gauss-Jordan on [A | I], the elimination order of PQP_CPU.c:251-326 (row swap pass on column 0 included) */
static void invert(const float *A, float *res, int n)
{
	const int w = 2 * n;
	float *m = (float *)calloc((size_t)n * w, sizeof(float));
	for (int i = 0; i < n; i++) {
		for (int j = 0; j < n; j++) m[i * w + j] = A[i * n + j];
		m[i * w + n + i] = 1.0f;
	}
	for (int i = n - 1; i > 0; i--)
		if (m[(i - 1) * w] < m[i * w])
			for (int j = 0; j < w; j++) {
				const float t = m[i * w + j];
				m[i * w + j] = m[(i - 1) * w + j];
				m[(i - 1) * w + j] = t;
			}
	for (int i = 0; i < n; i++)
		for (int j = 0; j < n; j++)
			if (j != i) {
				const float f = m[j * w + i] / m[i * w + i];
				for (int k = 0; k < w; k++) m[j * w + k] -= m[i * w + k] * f;
			}
	for (int i = 0; i < n; i++) {
		const float piv = m[i * w + i];
		for (int j = 0; j < w; j++) m[i * w + j] = m[i * w + j] / piv;
	}
	for (int i = 0; i < n; i++)
		for (int j = 0; j < n; j++) res[i * n + j] = m[i * w + n + j];
	free(m);
}

/* row vector times matrix, then dot with v: (a' B) v, sums k ascending from zero like matrixMultiply (PQP_CPU.c:84-147) */
static float quad(const float *a, const float *B, const float *v, int na, int nv)
{
	float out = 0.0f;
	float *t = (float *)calloc((size_t)nv, sizeof(float));
	for (int j = 0; j < nv; j++) {
		float s = 0.0f;
		for (int k = 0; k < na; k++) s += a[k] * B[k * nv + j];
		t[j] = s;
	}
	for (int j = 0; j < nv; j++) out += t[j] * v[j];
	free(t);
	return out;
}
static float dot(const float *a, const float *b, int n)
{
	float s = 0.0f;
	for (int k = 0; k < n; k++) s += a[k] * b[k];
	return s;
}

int main(int argc, char **argv)
{
	const char *dir = "./example";
	int order = PQP_ORDER_STRICT;
	long iters = 312;
	for (int i = 1; i < argc; i++) {
		if (!strcmp(argv[i], "--fast")) order = PQP_ORDER_FAST;
		else if (!strcmp(argv[i], "--iters") && i + 1 < argc) iters = atol(argv[++i]);
		else dir = argv[i];
	}
	const int pHorizon = 1, nState = 29, nInput = 7, nOutput = 7, nDis = 1; /* PQP_CPU.c:13-17 */
	pqp_dims d;
	pqp_dims_mpc(&d, pHorizon, nState, nInput, nOutput, nDis);
	const int M = d.M, N = d.N, nd = d.nDisH;
	pqp_host_problem p;
	int rc = pqp_load_example(dir, &d, &p);
	if (rc) {
		fprintf(stderr, "pqp_example: cannot read %s: %s\n", dir, pqp_strerror(rc));
		return 1;
	}
	pqp_compat_set_dims(pHorizon, nState, nInput, nOutput, nDis);
	pqp_compat_set_order(order);
	pqp_compat_set_fixed_iters(iters);

	float *Qp = (float *)calloc((size_t)M * M, sizeof(float)), *Fp = (float *)calloc(M, sizeof(float));
	float *Qd = (float *)calloc((size_t)N * N, sizeof(float)), *Fd = (float *)calloc(N, sizeof(float));
	float *Y = (float *)calloc(N, sizeof(float)), *U = (float *)calloc(M, sizeof(float));
	float Mp[1] = { 0.0f }, Md[1] = { 0.0f };

	invert(p.Qp_inv, Qp, M);
	computeFp(Fp, (float *)p.Fp1, (float *)p.Fp2, (float *)p.Fp3, (float *)p.D, (float *)p.x);
	/* computeMp, PQP_CPU.c:395-428: every term halved as the code does */
	Mp[0] += quad(p.x, p.Mp1, p.x, nState, nState) / 2;
	Mp[0] += quad(p.D, p.Mp2, p.x, nd, nState) / 2;
	Mp[0] += dot(p.Mp4, p.x, nState) / 2;
	Mp[0] += quad(p.D, p.Mp3, p.D, nd, nd) / 2;
	Mp[0] += dot(p.Mp5, p.D, nd) / 2;
	Mp[0] += p.Mp6[0] / 2;

	convertToDual(Qd, Fd, Md, (float *)p.Qp_inv, (float *)p.Gp, (float *)p.Kp, Fp, Mp, N, M);
	solveQuadraticDual(Y, Qd, Fd, Md, U, Qp, (float *)p.Qp_inv, Fp, Mp, (float *)p.Gp, (float *)p.Kp, N, M);
	computeUfromY(U, Y, Fp, (float *)p.Gp, (float *)p.Qp_inv, N, M);

	const float Jp = computeCost(U, Qp, Fp, Mp, M);
	const float Jd = computeCost(Y, Qd, Fd, Md, N);
	printf("Jp = %f\n", Jp);
	printf("Jd = %f\n", Jd);
	printf("Printing U*\n");
	for (int i = 0; i < M; i++) printf("\t%f\n", U[i]);

	free(Qp); free(Fp); free(Qd); free(Fd); free(Y); free(U);
	pqp_free_problem(&p);
	return 0;
}
