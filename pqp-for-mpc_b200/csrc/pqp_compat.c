/*
 * pqp_compat.c -- libpqp_compat.so: the reference's own function signatures, backed by
 * libpqp_b200.so, so PQP_CPU.c's main() (or a test written like it) can link against the new
 * solver unchanged.  Each function cites the reference definition it stands in for.
 *
 * The reference's functions return void and print-and-continue; these do the same: on a
 * library error they print one line to stderr and leave the outputs untouched.  They are
 * synchronous and take host pointers, as PQP_CPU.c passes them.  Sizes that the reference
 * takes from #defines (computeFp, PQP_CPU.c:13-17) come from pqp_compat_set_dims().
 *
 * Order of arithmetic: pqp_compat_set_order(PQP_ORDER_STRICT) makes every result bit-identical
 * to PQP_CPU.c; the default PQP_ORDER_FAST is within 1e-5 (normwise) of it.
 *
 * Nothing is computed on the host here: every function forwards to a device entry point of libpqp_b200.so.
 * Stop test (g_fixed_iters == 0): the fused test on g = Qd y + Fd with the caller's Kp and Md -- terminate()'s
 * three conditions (PQP_CPU.c:673-687) without forming Jp; its first condition (Jp <= -Jd, no tolerance) is met by the
 * reference on rounding noise, so stop counts agree statistically, not exactly.
 */
#include "pqp.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static pqp_dims g_dims = { 7, 28, 29, 1, 1, 7, 7, 1 }; /* PQP_CPU.c:13-17, 940-941 */
static int g_order = PQP_ORDER_FAST;
static long g_fixed_iters = 0; /* 0: run to the stop test like PQP_CPU.c:718 */
static long g_last_h = 0;

void pqp_compat_set_dims(int pHorizon, int nState, int nInput, int nOutput, int nDis)
{
	pqp_dims_mpc(&g_dims, pHorizon, nState, nInput, nOutput, nDis);
}
void pqp_compat_set_order(int order) { g_order = order; }
/* > 0: solveQuadraticDual applies exactly this many updates (testing/CPU version/PQP_CPU_test.c:717) */
void pqp_compat_set_fixed_iters(long k) { g_fixed_iters = k; }
long pqp_compat_last_iterations(void) { return g_last_h; }

static void complain(const char *fn, int rc)
{
	fprintf(stderr, "pqp_compat: %s failed: %s (%s)\n", fn, pqp_strerror(rc), pqp_last_cuda_error());
}

static pqp_opts opts_now(void)
{
	pqp_opts o;
	pqp_default_opts(&o);
	o.order = g_order;
	o.check_every = 1;
	return o;
}

/* PQP_CPU.c:373 */
void computeFp(float *Fp, float *Fp1, float *Fp2, float *Fp3, float *D, float *x)
{
	/* sizes come from #defines in the reference (PQP_CPU.c:13-17): pqp_compat_set_dims().  On the device, reference order. */
	int rc = pqp_compute_fp(Fp, Fp1, Fp2, Fp3, D, x, g_dims.M, g_dims.nDisH, g_dims.nState, -1);
	if (rc) complain("computeFp/pqp_compute_fp", rc);
}

/* PQP_CPU.c:489.  Qd/Fd/Md out; Qp_inv [M x M], Gp [N x M], Kp [N], Fp [M], Mp [1] in. */
void convertToDual(float *Qd, float *Fd, float *Md, float *Qp_inv, float *Gp, float *Kp, float *Fp, float *Mp, int N, int M)
{
	pqp_dims d;
	memset(&d, 0, sizeof d);
	d.M = M;
	d.N = N;
	pqp_host_problem p;
	memset(&p, 0, sizeof p);
	p.Qp_inv = Qp_inv; p.Gp = Gp; p.Kp = Kp; p.Fp = Fp; p.Mp0 = Mp ? Mp[0] : 0.0f;
	pqp_opts o = opts_now();
	pqp_handle *h = NULL;
	int rc = pqp_setup(&h, &d, &p, &o);
	if (rc) { complain("convertToDual/pqp_setup", rc); return; }
	rc = pqp_get_dual(h, Qd, NULL, NULL);
	if (rc) complain("convertToDual/pqp_get_dual", rc);
	/* Fd = GQ*Fp + Kp is formed by the solve entry point: run one update on a scratch y, read Fd back */
	float *ytmp = (float *)malloc(sizeof(float) * (size_t)N);
	rc = pqp_solve_batch(h, NULL, NULL, 1, 1, NULL, ytmp, NULL);
	if (rc) complain("convertToDual/pqp_solve_batch", rc);
	else rc = pqp_get_linear_terms(h, 1, Fd, NULL);
	if (rc) complain("convertToDual/pqp_get_linear_terms", rc);
	free(ytmp);
	pqp_destroy(h);
	if (Md) { /* Md = Fp' Qp_inv Fp - Mp (computeMd, PQP_CPU.c:472-479): on the device, reference order */
		rc = pqp_compute_md(Md, Fp, Qp_inv, Mp, M, -1);
		if (rc) complain("convertToDual/pqp_compute_md", rc);
	}
}

/* PQP_CPU.c:694.  Y out [N], U out [M] (the last computeUfromY of terminate, PQP_CPU.c:675). */
void solveQuadraticDual(float *Y, float *Qd, float *Fd, float *Md, float *U, float *Qp, float *Qp_inv, float *Fp,
			float *Mp, float *Gp, float *Kp, int N, int M)
{
	(void)Qp; (void)Mp; /* Jp is never formed: Jp + Jd = y'(Qd y + Fd), so Gauss_Jordan's Qp and Mp are not needed */
	pqp_opts o = opts_now();
	pqp_handle *h = NULL;
	int rc = pqp_setup_dual(&h, N, Qd, M, Gp, Qp_inv, &o);
	if (rc) { complain("solveQuadraticDual/pqp_setup_dual", rc); return; }
	/* the caller's Kp and Md reach the stop test: compare()'s per-row tolerance max(erc*Kp_i, eac) (PQP_CPU.c:338) and the
	 * scale of the relative gap test |Jp + Jd| <= erj*|Jd| with Jd including Md/2 (PQP_CPU.c:684, :661) */
	if (Kp && (rc = pqp_set_constraint_bounds(h, Kp))) complain("solveQuadraticDual/pqp_set_constraint_bounds", rc);
	pqp_status st;
	memset(&st, 0, sizeof st);
	rc = pqp_solve_dual_full(h, Fd, Md, 1, (int)g_fixed_iters, NULL, Y, &st);
	if (rc) complain("solveQuadraticDual/pqp_solve_dual_full", rc);
	else if (U && Fp && Gp && Qp_inv) {
		rc = pqp_recover_primal(h, Y, Fp, 1, U);
		if (rc) complain("solveQuadraticDual/pqp_recover_primal", rc);
	}
	g_last_h = (long)st.iters + 1; /* the reference counts from 1 (PQP_CPU.c:714) */
	printf("Printing number of iterations = %ld\n", g_last_h); /* PQP_CPU.c:741 */
	pqp_destroy(h);
}

/* PQP_CPU.c:352 */
void computeUfromY(float *U, float *Y, float *Fp, float *Gp, float *Qp_inv, int N, int M)
{
	int rc = pqp_compute_u_from_y(U, Y, Fp, Gp, Qp_inv, N, M, 1, -1);
	if (rc) complain("computeUfromY/pqp_compute_u_from_y", rc);
}

/* PQP_CPU.c:648: J = 1/2 z'Qz + F'z + m/2, on the device in the reference's order and promotions */
float computeCost(float *Z, float *Q, float *F, float *M, int N)
{
	float J = 0;
	int rc = pqp_compute_cost(&J, Z, Q, F, M, N, -1);
	if (rc) complain("computeCost/pqp_compute_cost", rc);
	return J;
}

/*
 * PQP_CPU.c:603 (+ updY, :590-596, which the reference calls from inside it): one update from the two dense split matrices.
 * The solver proper never forms them (one signed Qd + a theta vector); this shim exists for single-step tests and is
 * bit-identical to the reference whatever pqp_compat_set_order says.  Fd is unused by the reference's body as well.
 */
void updateY2(float *Y_next, float *Y, float *Qdp_theta, float *Qdn_theta, float *Fd, float *Fdp, float *Fdn, int N)
{
	(void)Fd;
	int rc = pqp_update_y2(Y_next, Y, Qdp_theta, Qdn_theta, Fdp, Fdn, N, -1);
	if (rc) complain("updateY2/pqp_update_y2", rc);
}
