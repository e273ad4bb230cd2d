/*
 * pqp_io.c -- host-side data formats of the PQP path (plain C, no CUDA).
 *
 *   pqp_load_example          replaces input(), PQP_CPU.c:757-930
 *   pqp_load_testfile         replaces the reader in testing/CPU version/PQP_CPU_test.c:936-976
 *   pqp_generate_testproblem  the distribution of testing/test_generator.c:936-987, seeded
 *   pqp_write_testfile        the writer half of testing/test_generator.c:936-987
 */
#include "pqp.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

void pqp_dims_mpc(pqp_dims *d, int pHorizon, int nState, int nInput, int nOutput, int nDis)
{
	memset(d, 0, sizeof *d);
	d->pHorizon = pHorizon;
	d->nState = nState;
	d->nInput = nInput;
	d->nOutput = nOutput;
	d->nDis = nDis;
	d->M = pHorizon * nInput;      /* PQP_CPU.c:940 */
	d->N = 4 * pHorizon * nInput;  /* PQP_CPU.c:941 */
	d->nDisH = nDis * pHorizon;
}

const char *pqp_strerror(int code)
{
	switch (code) {
	case PQP_OK: return "ok";
	case PQP_ERR_INVALID: return "invalid argument";
	case PQP_ERR_NO_DEVICE: return "no usable sm_100 CUDA device (this library has no CPU fallback)";
	case PQP_ERR_CUDA: return "CUDA runtime error";
	case PQP_ERR_ALLOC: return "allocation failed";
	case PQP_ERR_IO: return "file missing or malformed";
	case PQP_ERR_UNSUPPORTED: return "shape not supported";
	default: return "unknown error";
	}
}

static float *falloc(size_t n)
{
	return (float *)calloc(n ? n : 1, sizeof(float));
}

/*
 * One example/ file: whitespace-separated decimal tokens, column-major (a MATLAB dump), ended by
 * a lone '#'.  Token (c, r) lands at dst[r*cols + c] -- the transposition every block of
 * input() performs (e.g. PQP_CPU.c:765-772).
 */
static int read_example_file(const char *dir, const char *name, int rows, int cols, float **out)
{
	char path[4096];
	if (snprintf(path, sizeof path, "%s/%s", dir, name) >= (int)sizeof path) return PQP_ERR_INVALID;
	FILE *f = fopen(path, "r");
	if (!f) return PQP_ERR_IO;
	float *dst = falloc((size_t)rows * cols);
	if (!dst) { fclose(f); return PQP_ERR_ALLOC; }
	for (int c = 0; c < cols; c++)
		for (int r = 0; r < rows; r++) {
			float v;
			if (fscanf(f, "%f", &v) != 1) {
				fclose(f);
				free(dst);
				return PQP_ERR_IO;
			}
			dst[(size_t)r * cols + c] = v;
		}
	fclose(f);
	*out = dst;
	return PQP_OK;
}

void pqp_free_problem(pqp_host_problem *p)
{
	if (!p) return;
	const float **fields[] = { &p->Qp_inv, &p->Gp, &p->Kp, &p->Fp1, &p->Fp2, &p->Fp3, &p->D, &p->Mp1, &p->Mp2,
				   &p->Mp3, &p->Mp4, &p->Mp5, &p->Mp6, &p->Fp, &p->x, &p->Z, &p->Theta };
	for (size_t i = 0; i < sizeof fields / sizeof fields[0]; i++) {
		free((void *)*fields[i]);
		*fields[i] = NULL;
	}
}

int pqp_load_example(const char *dir, const pqp_dims *d, pqp_host_problem *out)
{
	if (!dir || !d || !out || d->M <= 0 || d->N <= 0 || d->nState < 0 || d->nDisH < 0) return PQP_ERR_INVALID;
	memset(out, 0, sizeof *out);
	const int M = d->M, N = d->N, nS = d->nState, nd = d->nDisH, no = d->nOutput * d->pHorizon;
	struct { const char *name; int rows, cols; const float **dst; } files[] = {
		{ "Qp_inv.txt", M, M, &out->Qp_inv },  /* PQP_CPU.c:764-773 */
		{ "Fp1.txt", M, nd, &out->Fp1 },       /* :776-785 */
		{ "Fp2.txt", M, nS, &out->Fp2 },       /* :788-797 */
		{ "Fp3.txt", M, 1, &out->Fp3 },        /* :800-806 */
		{ "Mp1.txt", nS, nS, &out->Mp1 },      /* :809-818 */
		{ "Mp2.txt", nd, nS, &out->Mp2 },      /* :821-830 */
		{ "Mp3.txt", nd, nd, &out->Mp3 },      /* :833-842 */
		{ "Mp4.txt", nS, 1, &out->Mp4 },       /* :845-851 */
		{ "Mp5.txt", nd, 1, &out->Mp5 },       /* :854-860 */
		{ "Mp6.txt", 1, 1, &out->Mp6 },        /* :863-866 */
		{ "Gp.txt", N, M, &out->Gp },          /* :869-878 */
		{ "Kp.txt", N, 1, &out->Kp },          /* :881-887 */
		{ "Z.txt", no, nS, &out->Z },          /* :890-899 */
		{ "Theta.txt", no, nd, &out->Theta },  /* :902-911 */
		{ "D.txt", nd, 1, &out->D },           /* :914-920 */
		{ "x.txt", nS, 1, &out->x },           /* :923-929 */
	};
	for (size_t i = 0; i < sizeof files / sizeof files[0]; i++) {
		float *buf = NULL;
		int rc = read_example_file(dir, files[i].name, files[i].rows, files[i].cols, &buf);
		if (rc != PQP_OK) {
			pqp_free_problem(out);
			return rc;
		}
		*files[i].dst = buf;
	}
	return PQP_OK;
}

int pqp_load_testfile(const char *path, pqp_dims *d, pqp_host_problem *out)
{
	if (!path || !d || !out) return PQP_ERR_INVALID;
	memset(out, 0, sizeof *out);
	memset(d, 0, sizeof *d);
	FILE *f = fopen(path, "r");
	if (!f) return PQP_ERR_IO;
	int M = 0, N = 0, rc = PQP_ERR_IO;
	float *Qi = NULL, *Fp = NULL, *Kp = NULL, *Gp = NULL;
	if (fscanf(f, "%d%d", &M, &N) != 2 || M <= 0 || N <= 0) goto done; /* line 1: "M N" */
	Qi = falloc((size_t)M * M);
	Fp = falloc(M);
	Kp = falloc(N);
	Gp = falloc((size_t)N * M);
	if (!Qi || !Fp || !Kp || !Gp) { rc = PQP_ERR_ALLOC; goto done; }
	for (int i = 0; i < M; i++) /* diagonal of Qp_inv */
		if (fscanf(f, "%f", &Qi[(size_t)i * M + i]) != 1) goto done;
	for (int i = 0; i < M; i++)
		if (fscanf(f, "%f", &Fp[i]) != 1) goto done;
	if (fscanf(f, "%f", &out->Mp0) != 1) goto done;
	for (int i = 0; i < N; i++)
		if (fscanf(f, "%f", &Kp[i]) != 1) goto done;
	for (size_t e = 0; e < (size_t)N * M; e++) {
		int t;
		if (fscanf(f, "%d", &t) != 1) goto done;
		Gp[e] = (float)t; /* literal: -1 stays -1 (the reference reader turns it into +1) */
	}
	rc = PQP_OK;
done:
	fclose(f);
	if (rc != PQP_OK) {
		free(Qi); free(Fp); free(Kp); free(Gp);
		return rc;
	}
	d->M = M;
	d->N = N;
	out->Qp_inv = Qi;
	out->Fp = Fp;
	out->Kp = Kp;
	out->Gp = Gp;
	return PQP_OK;
}

static unsigned long long splitmix64(unsigned long long *s)
{
	unsigned long long z = (*s += 0x9E3779B97F4A7C15ULL);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

/* U[0,100) with six decimals, i.e. what fprintf("%f") of test_generator.c:944-945 leaves in the file */
static float gen_u100(unsigned long long *s)
{
	double u = (double)(splitmix64(s) >> 11) * (1.0 / 9007199254740992.0);
	return (float)(floor(u * 100.0 * 1e6 + 0.5) / 1e6);
}

int pqp_generate_testproblem(unsigned long long seed, int M, int N, pqp_dims *d, pqp_host_problem *out)
{
	if (M <= 0 || N <= 0 || !d || !out) return PQP_ERR_INVALID;
	memset(out, 0, sizeof *out);
	memset(d, 0, sizeof *d);
	float *Qi = falloc((size_t)M * M), *Fp = falloc(M), *Kp = falloc(N), *Gp = falloc((size_t)N * M);
	if (!Qi || !Fp || !Kp || !Gp) {
		free(Qi); free(Fp); free(Kp); free(Gp);
		return PQP_ERR_ALLOC;
	}
	unsigned long long s = seed;
	/* generator order: diag(Qp_inv), Fp, Mp, Kp, Gp (test_generator.c:942-984) */
	for (int i = 0; i < M; i++) Qi[(size_t)i * M + i] = gen_u100(&s);
	for (int i = 0; i < M; i++) Fp[i] = gen_u100(&s);
	out->Mp0 = gen_u100(&s);
	for (int i = 0; i < N; i++) Kp[i] = gen_u100(&s);
	for (size_t e = 0; e < (size_t)N * M; e++) {
		unsigned r = (unsigned)(splitmix64(&s) % 3u); /* 0 -> 0, 2 -> -1, else +1 (test_generator.c:970-981) */
		Gp[e] = r == 0 ? 0.0f : (r == 2 ? -1.0f : 1.0f);
	}
	d->M = M;
	d->N = N;
	out->Qp_inv = Qi;
	out->Fp = Fp;
	out->Kp = Kp;
	out->Gp = Gp;
	return PQP_OK;
}

int pqp_write_testfile(const char *path, const pqp_dims *d, const pqp_host_problem *p)
{
	if (!path || !d || !p || !p->Qp_inv || !p->Fp || !p->Kp || !p->Gp) return PQP_ERR_INVALID;
	FILE *f = fopen(path, "w");
	if (!f) return PQP_ERR_IO;
	const int M = d->M, N = d->N;
	fprintf(f, "%d %d\n", M, N);
	for (int i = 0; i < M; i++) fprintf(f, "%f ", p->Qp_inv[(size_t)i * M + i]);
	fprintf(f, "\n");
	for (int i = 0; i < M; i++) fprintf(f, "%f ", p->Fp[i]);
	fprintf(f, "\n%f\n", p->Mp0);
	for (int i = 0; i < N; i++) fprintf(f, "%f ", p->Kp[i]);
	fprintf(f, "\n");
	for (int i = 0; i < N; i++) {
		for (int j = 0; j < M; j++) fprintf(f, "%d ", (int)p->Gp[(size_t)i * M + j]);
		fprintf(f, "\n");
	}
	return fclose(f) == 0 ? PQP_OK : PQP_ERR_IO;
}
