/*
 * pqp_api.cu -- host side of the C ABI in include/pqp.h: handle, setup, solve, recovery.
 *
 * Mirrors the call sequence of PQP_CPU.c's main() (:988-999):
 *   input -> Gauss_Jordan -> computeFp -> computeMp -> convertToDual -> solveQuadraticDual -> computeUfromY
 * split into the x-independent part (pqp_setup: GQ, Qd, theta) and the per-state part
 * (pqp_solve_batch: Fp(x), Fd(x), the loop; pqp_recover_primal: U).  Gauss_Jordan is not needed:
 * Qp only feeds Jp inside terminate(), which SURVEY 3.3 replaces by reductions on g = Qd y + Fd.
 *
 * There is no CPU fallback anywhere in this file: without an sm_100 device every entry point
 * that computes returns PQP_ERR_NO_DEVICE.
 */
#include "pqp_internal.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static thread_local char g_cuda_err[512] = "";

/* ---- the handle's snapshot of the PQP_* environment (pqp_internal.h) ---------------------------------------------------------- */
extern char **environ;
struct pqp_env_snapshot {
	int n;
	char *kv[96]; /* "NAME=value" copies */
};
static thread_local const pqp_env_snapshot *g_env = NULL;

static void env_snapshot_take(pqp_env_snapshot *e)
{
	e->n = 0;
	for (char **p = environ; p && *p && e->n < (int)(sizeof e->kv / sizeof e->kv[0]); p++)
		if (!strncmp(*p, "PQP_", 4)) e->kv[e->n++] = strdup(*p);
}
static void env_snapshot_free(pqp_env_snapshot *e)
{
	for (int i = 0; i < e->n; i++) free(e->kv[i]);
	e->n = 0;
}
const char *pqp_env(const char *name)
{
	if (!g_env) return getenv(name);
	const size_t len = strlen(name);
	for (int i = 0; i < g_env->n; i++)
		if (g_env->kv[i] && !strncmp(g_env->kv[i], name, len) && g_env->kv[i][len] == '=') return g_env->kv[i] + len + 1;
	return NULL;
}

#define CK(call)                                                                                          \
	do {                                                                                              \
		cudaError_t e__ = (call);                                                                 \
		if (e__ != cudaSuccess) {                                                                 \
			snprintf(g_cuda_err, sizeof g_cuda_err, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, \
				 cudaGetErrorString(e__));                                                \
			return e__ == cudaErrorMemoryAllocation ? PQP_ERR_ALLOC : PQP_ERR_CUDA;           \
		}                                                                                         \
	} while (0)

struct pqp_handle {
	pqp_dims d;
	pqp_opts o;
	int device, num_sms;
	size_t smem_optin;
	cudaStream_t stream;
	cudaEvent_t ev0, ev1;
	int ev_valid;
	int ldq;
	int have_primal, have_fp_model;
	/* x-independent device data */
	float *Q, *QT, *theta, *GQ, *Gp, *Qp_inv, *Kp, *Fp1, *Fp2, *Fp3, *Fp_const, *D;
	float *Kx, *Kd; /* state / disturbance dependent part of Kp (NULL: constant Kp as in PQP_CPU.c) */
	float *Mp1, *Mp2, *Mp3, *Mp4, *Mp5, *Mp6;
	float Mp0;
	float *QpT, *QnT; /* batched operands, built on first batched solve */
	void *umma_tiles; /* pre-split, pre-tiled tf32 hi/lo operand of the tcgen05 3xTF32 batched kernel */
	void *imma_tiles, *imma_rowc; /* digit planes + row constants of the tcgen05 int8 (error-free) batched kernel */
	void *imma_ptiles, *imma_prowc; /* the same for the PAIRED scheme (Qd with the +/- row-pair structure of an MPC dual) */
	int paired_state; /* 0 not examined, 1 structure present and the arrays built, -1 unavailable */
	int Kpad, Ipad;
	/* per-batch workspace */
	int cap;
	float *X, *Db, *Fp, *Fd, *Md, *Y, *U, *Tmp;
	pqp_status *st;
	int fp_B; /* problems whose Fp is cached from the last solve */
	float *tol_ws;      /* run-to-tolerance on the paired kernel: evaluation partials | resume maxima [cap] | kept duals [cap x N] */
	unsigned *tol_flags; /* frozen [cap] | newly [cap] | remaining [1] */
	int tol_cap;
	unsigned *tol_rem_host; /* [2] pinned: "problems still running" after chunk k lands in slot k & 1 */
	cudaEvent_t tol_ev[2];  /* recorded behind that copy */
	float *acc_ws;      /* [3][cap][N] scratch of the acceleration step (opts.accelerate), allocated on first use */
	int acc_cap;
	int iters_base;     /* updates applied by earlier chunks of the solve in progress (reported in pqp_status.iters) */
	const float *cur_D; /* disturbance vectors of the solve being formed: h->D (stride 0) or h->Db (stride nDisH) */
	int cur_Dstride;
	/* single-problem loop state */
	float *ybuf0, *ybuf1, *partials;
	void *pk0, *pk1; /* {value, epoch} packet vectors of the flag-in-data y exchange */
	unsigned *barrier;
	int *result_buf;
	int gemv_grid, gemv_resident;
	int tma_ok, tma_stages, tma_resident, tma_yc, tma_pinned;
	int small_ok, small_wpr, small_cpt, small_grid;
	int cluster_state; /* one-cluster kernel for this handle: 0 not decided, 1 use it, -1 the multi-CTA kernel was faster here */
	int sym_state; /* 0 not examined, 1 Qd symmetric and the unit array built, -1 unavailable */
	pqp_sym_plan sym;
	int l2_window_set;
	long long launches;
	const char *last_kernel;
	pqp_env_snapshot env; /* PQP_* knobs as they were when the handle was created */
	int qd_mirrored;       /* pqp_setup built Qd from the tiles of its upper triangle (symmetric Qp_inv): both copies of a pair are one sum */
	float setup_gemm_ms;   /* device time of the two dual-construction GEMMs (GQ, Qd) in pqp_setup */
	size_t l2_limit_saved; /* cudaLimitPersistingL2CacheSize before this handle changed it */
	int l2_limit_changed;
};

/* every public entry point that takes a handle: its knobs are the calling thread's current ones until the next entry */
struct env_scope {
	const pqp_env_snapshot *prev;
	explicit env_scope(const pqp_handle *h) : prev(g_env) { g_env = h ? &h->env : NULL; }
	~env_scope() { g_env = prev; }
};

const char *pqp_last_cuda_error(void) { return g_cuda_err; }

void pqp_default_opts(pqp_opts *o)
{
	memset(o, 0, sizeof *o);
	o->theta_floor = 5.0f;  /* PQP_CPU.c:240 */
	o->y_init = 1000.0f;    /* PQP_CPU.c:710 */
	o->erc = o->eac = o->eaj = o->erj = 1e-6f; /* PQP_CPU.c:19-22 */
	o->order = PQP_ORDER_FAST;
	o->device = -1;
	o->max_iters = 100000;
	o->check_every = 8;
	o->batch_capacity = 1;
	o->use_tensor_cores = 1;
	o->l2_persist = 1;
	o->exploit_symmetry = 1;
	o->exploit_structure = 1;
	o->accelerate = 0;
}

int pqp_output_offsets(const pqp_dims *d, const float *Z, const float *Theta, float *Kx, float *Kd)
{
	if (!d) return PQP_ERR_INVALID;
	const int M = d->M, N = d->N, nS = d->nState, nd = d->nDisH, nOut = d->nOutput * d->pHorizon;
	if (N != 4 * M || nOut < 0 || nOut > M) return PQP_ERR_INVALID;
	if (Kx) {
		if (!Z && nS > 0 && nOut > 0) return PQP_ERR_INVALID;
		memset(Kx, 0, sizeof(float) * (size_t)N * nS);
		for (int r = 0; r < nOut; r++)
			for (int k = 0; k < nS; k++) {
				Kx[(size_t)(2 * M + r) * nS + k] = -Z[(size_t)r * nS + k];
				Kx[(size_t)(3 * M + r) * nS + k] = Z[(size_t)r * nS + k];
			}
	}
	if (Kd) {
		if (!Theta && nd > 0 && nOut > 0) return PQP_ERR_INVALID;
		memset(Kd, 0, sizeof(float) * (size_t)N * nd);
		for (int r = 0; r < nOut; r++)
			for (int k = 0; k < nd; k++) {
				Kd[(size_t)(2 * M + r) * nd + k] = -Theta[(size_t)r * nd + k];
				Kd[(size_t)(3 * M + r) * nd + k] = Theta[(size_t)r * nd + k];
			}
	}
	return PQP_OK;
}

int pqp_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	int ok = 0;
	for (int i = 0; i < n; i++) {
		int major = 0;
		if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) ok++;
	}
	return ok;
}

template <typename T> static int dalloc(T **p, size_t n)
{
	*p = NULL;
	if (n == 0) n = 1;
	cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
	if (e != cudaSuccess) {
		snprintf(g_cuda_err, sizeof g_cuda_err, "cudaMalloc(%zu bytes) -> %s", n * sizeof(T), cudaGetErrorString(e));
		cudaGetLastError();
		return PQP_ERR_ALLOC;
	}
	return PQP_OK;
}

/* allocate + copy from a host-or-device source (UVA sorts out the direction) */
static int upload(pqp_handle *h, float **dst, const float *src, size_t n)
{
	*dst = NULL;
	if (!src || n == 0) return PQP_OK;
	int rc = dalloc(dst, n);
	if (rc) return rc;
	CK(cudaMemcpyAsync(*dst, src, n * sizeof(float), cudaMemcpyDefault, h->stream));
	return PQP_OK;
}

/* the persisting-L2 carve-out is device-global state: remember what it was the first time this handle changes it, so that
 * pqp_destroy can put it back */
static cudaError_t set_l2_carveout(pqp_handle *h, size_t bytes)
{
	if (!h->l2_limit_changed) {
		size_t cur = 0;
		if (cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize) == cudaSuccess) {
			h->l2_limit_saved = cur;
			h->l2_limit_changed = 1;
		} else {
			cudaGetLastError();
		}
	}
	return cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, bytes);
}

static int open_device(pqp_handle *h, const pqp_opts *opts)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
		cudaGetLastError();
		return PQP_ERR_NO_DEVICE;
	}
	int dev = opts->device;
	if (dev < 0) CK(cudaGetDevice(&dev));
	if (dev >= n) return PQP_ERR_INVALID;
	int major = 0;
	CK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
	if (major != 10) return PQP_ERR_NO_DEVICE; /* the kernels are sm_100a-only */
	CK(cudaSetDevice(dev));
	h->device = dev;
	CK(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, dev));
	int optin = 0;
	CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
	h->smem_optin = (size_t)optin;
	if (pqp_env("PQP_L2_CARVEOUT_MB")) {
		/* experiment knob: size of the persisting-L2 carve-out (evict_last lines live there) */
		int maxp = 0;
		cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, dev);
		size_t want = (size_t)atoi(pqp_env("PQP_L2_CARVEOUT_MB")) << 20;
		if (want > (size_t)maxp) want = (size_t)maxp;
		CK(set_l2_carveout(h, want));
		if (pqp_env("PQP_VERBOSE")) fprintf(stderr, "pqp: persisting L2 max %d MB, set %zu MB\n", maxp >> 20, want >> 20);
	}
	CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
	CK(cudaEventCreate(&h->ev0));
	CK(cudaEventCreate(&h->ev1));
	return PQP_OK;
}

static int ensure_capacity(pqp_handle *h, int B)
{
	if (B <= h->cap) return PQP_OK;
	const int M = h->d.M, N = h->d.N;
	float **bufs[] = { &h->X, &h->Db, &h->Fp, &h->Fd, &h->Md, &h->Y, &h->U, &h->Tmp };
	for (size_t i = 0; i < sizeof bufs / sizeof bufs[0]; i++) {
		if (*bufs[i]) cudaFree(*bufs[i]);
		*bufs[i] = NULL;
	}
	if (h->st) cudaFree(h->st);
	h->st = NULL;
	h->cap = 0;
	h->fp_B = 0;
	int rc = 0;
	rc |= dalloc(&h->X, (size_t)B * (h->d.nState > 0 ? h->d.nState : 1));
	rc |= dalloc(&h->Db, (size_t)B * (h->d.nDisH > 0 ? h->d.nDisH : 1));
	rc |= dalloc(&h->Fp, (size_t)B * (M > 0 ? M : 1));
	rc |= dalloc(&h->Fd, (size_t)B * N);
	rc |= dalloc(&h->Md, (size_t)B);
	rc |= dalloc(&h->Y, (size_t)B * N);
	rc |= dalloc(&h->U, (size_t)B * (M > 0 ? M : 1));
	rc |= dalloc(&h->Tmp, (size_t)B * (M > 0 ? M : 1));
	rc |= dalloc(&h->st, (size_t)B);
	if (rc) return PQP_ERR_ALLOC;
	h->cap = B;
	return PQP_OK;
}

/* theta, transposes, loop state: everything that only needs Q on the device */
static int finish_setup(pqp_handle *h)
{
	const int N = h->d.N, ldq = h->ldq;
	const int strict = h->o.order == PQP_ORDER_STRICT;
	int rc;
	if ((rc = dalloc(&h->theta, N))) return rc;
	if (strict) {
		if ((rc = dalloc(&h->QT, (size_t)N * ldq))) return rc;
		CK(cudaMemsetAsync(h->QT, 0, (size_t)N * ldq * sizeof(float), h->stream));
		CK(pqp_launch_transpose(h->QT, ldq, h->Q, ldq, N, N, h->stream));
		CK(pqp_launch_theta(h->theta, h->QT, ldq, N, h->o.theta_floor, 1, h->stream));
		h->launches += 2;
	} else {
		CK(pqp_launch_theta(h->theta, h->Q, ldq, N, h->o.theta_floor, 0, h->stream));
		h->launches += 1;
	}
	if ((rc = dalloc(&h->ybuf0, ldq)) || (rc = dalloc(&h->ybuf1, ldq))) return rc;
	{
		double *a0 = NULL, *a1 = NULL;
		if ((rc = dalloc(&a0, ldq)) || (rc = dalloc(&a1, ldq))) return rc;
		h->pk0 = a0;
		h->pk1 = a1;
	}
	if ((rc = dalloc(&h->barrier, 4)) || (rc = dalloc(&h->result_buf, 4))) return rc;

	/* persistent-kernel geometry: one CTA per SM, but never fewer than 4 rows per CTA */
	int grid = h->num_sms;
	if (grid > (N + 3) / 4) grid = (N + 3) / 4;
	if (grid < 1) grid = 1;
	if (pqp_env("PQP_GEMV_GRID")) {
		int v = atoi(pqp_env("PQP_GEMV_GRID"));
		if (v >= 1 && v <= h->num_sms && v <= N) grid = v;
	}
	h->gemv_grid = grid;
	if ((rc = dalloc(&h->partials, (size_t)2 * h->num_sms * 16))) return rc; /* per-CTA evaluation slots / check packets of the loop kernels */
	/* rows of each slab that fit in shared memory next to y and the partial sums */
	size_t base;
	pqp_gemv_smem_bytes(N, ldq, grid, 0, &base);
	const size_t budget = h->smem_optin > 1024 ? h->smem_optin - 1024 : 0;
	int res = 0;
	if (base < budget) res = (int)((budget - base) / ((size_t)ldq * sizeof(float)));
	const int rows_max = (N + grid - 1) / grid + 1;
	if (res > rows_max) res = rows_max;
	if (base > budget) {
		h->gemv_grid = 0; /* y does not fit in shared memory: persistent kernel unavailable */
		res = 0;
	}
	const char *env = pqp_env("PQP_GEMV_RESIDENT");
	if (env) {
		int v = atoi(env);
		if (v >= 0 && v < res) res = v;
	}
	/* the LDG kernel only parks rows when the whole slab fits; partial residency belongs to the TMA kernel */
	if (res < rows_max && !env) res = 0;
	h->gemv_resident = res;

	/* register-resident kernel for small N */
	h->small_ok = 0;
	if (h->gemv_grid > 0 && !(pqp_env("PQP_GEMV_SMALL") && atoi(pqp_env("PQP_GEMV_SMALL")) == 0)) {
		/* as FEW CTAs as the 16-rows-per-CTA layout allows: the y exchange costs ~1 us with 32-64 participants and
		 * ~3.4 us with 148 (measured, tools/gemv_sweep.py N=1024), and the arithmetic is negligible either way */
		int sg = (N + 15) / 16;
		if (sg > h->num_sms) sg = h->num_sms;
		if (pqp_env("PQP_GEMV_GRID")) sg = h->gemv_grid;
		h->small_grid = sg;
		h->small_ok = pqp_gemv_small_plan(N, ldq, sg, &h->small_wpr, &h->small_cpt);
	}

	/* TMA-staged kernel: ring depth / residency / L2-pinned rows */
	h->tma_ok = 0;
	if (h->gemv_grid > 0) {
		h->tma_ok = pqp_gemv_tma_plan(N, ldq, h->gemv_grid, budget, &h->tma_stages, &h->tma_resident, &h->tma_yc);
		const char *e;
		if ((e = pqp_env("PQP_GEMV_TMA")) && atoi(e) == 0) h->tma_ok = 0;
		if (h->tma_ok) {
			const int total = h->tma_stages + h->tma_resident;
			if ((e = pqp_env("PQP_TMA_STAGES"))) {
				int v = atoi(e);
				if (v >= 2 && v <= total && h->tma_resident < rows_max) {
					h->tma_stages = v;
					h->tma_resident = total - v;
				}
			}
			if ((e = pqp_env("PQP_TMA_RESIDENT"))) {
				int v = atoi(e);
				if (v >= 0 && v < h->tma_resident) h->tma_resident = v;
			}
			/* rows per slab fetched evict_last: ~80% of L2 spread evenly over the CTAs (only matters when Q > L2) */
			int l2_bytes = 0;
			cudaDeviceGetAttribute(&l2_bytes, cudaDevAttrL2CacheSize, h->device);
			h->tma_pinned = 0;
			if (h->o.l2_persist && (size_t)N * ldq * sizeof(float) > (size_t)l2_bytes) {
				/* measured (tools/gemv_sweep.py, N=8192): 64 MB carve-out + ~60% of L2 pinned (16 rows/slab) is the best point */
				h->tma_pinned = (int)(0.6 * (double)l2_bytes / ((double)h->gemv_grid * ldq * sizeof(float)));
				int maxp = 0;
				cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, h->device);
				size_t want = (size_t)64 << 20;
				if (want > (size_t)maxp) want = (size_t)maxp;
				if (!pqp_env("PQP_L2_CARVEOUT_MB")) set_l2_carveout(h, want);
			}
			if ((e = pqp_env("PQP_L2_PIN_ROWS"))) h->tma_pinned = atoi(e) > 0 ? atoi(e) : 0;
			if (pqp_env("PQP_VERBOSE"))
				fprintf(stderr, "pqp: gemv_tma N=%d grid=%d stages=%d resident=%d pinned=%d yc=%d\n", N, h->gemv_grid,
					h->tma_stages, h->tma_resident, h->tma_pinned, h->tma_yc);
		}
	}

	if ((rc = ensure_capacity(h, h->o.batch_capacity > 0 ? h->o.batch_capacity : 1))) return rc;
	CK(cudaStreamSynchronize(h->stream));
	return PQP_OK;
}

static int l2_persist_window(pqp_handle *h, int enable)
{
	if (!h->o.l2_persist) return PQP_OK;
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, h->device));
	if (prop.persistingL2CacheMaxSize <= 0 || prop.accessPolicyMaxWindowSize <= 0) return PQP_OK;
	cudaStreamAttrValue v;
	memset(&v, 0, sizeof v);
	if (enable) {
		const size_t qbytes = (size_t)h->d.N * h->ldq * sizeof(float);
		size_t persist = (size_t)prop.persistingL2CacheMaxSize;
		const char *env = pqp_env("PQP_L2_PERSIST_MB");
		if (env) persist = (size_t)atoi(env) << 20;
		if (persist == 0) return PQP_OK;
		if (persist > (size_t)prop.persistingL2CacheMaxSize) persist = (size_t)prop.persistingL2CacheMaxSize;
		CK(set_l2_carveout(h, persist));
		size_t win = qbytes < (size_t)prop.accessPolicyMaxWindowSize ? qbytes : (size_t)prop.accessPolicyMaxWindowSize;
		v.accessPolicyWindow.base_ptr = (void *)h->Q;
		v.accessPolicyWindow.num_bytes = win;
		float ratio = (float)((double)persist / (double)win);
		v.accessPolicyWindow.hitRatio = ratio > 1.0f ? 1.0f : ratio;
		v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
		v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
	} else {
		v.accessPolicyWindow.num_bytes = 0;
	}
	CK(cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &v));
	h->l2_window_set = enable;
	return PQP_OK;
}

int pqp_setup(pqp_handle **out, const pqp_dims *dims, const pqp_host_problem *p, const pqp_opts *opts_in)
{
	if (!out || !dims || !p) return PQP_ERR_INVALID;
	*out = NULL;
	if (dims->M <= 0 || dims->N <= 0 || dims->nState < 0 || dims->nDisH < 0) return PQP_ERR_INVALID;
	if (!p->Qp_inv || !p->Gp || !p->Kp) return PQP_ERR_INVALID;
	if (dims->nState > 0 && (!p->Fp2 || !p->Fp3)) return PQP_ERR_INVALID;
	if (dims->nState > 0 && dims->nDisH > 0 && (!p->Fp1 || !p->D)) return PQP_ERR_INVALID;
	if (dims->nState == 0 && !p->Fp) return PQP_ERR_INVALID;
	pqp_opts o;
	if (opts_in) o = *opts_in; else pqp_default_opts(&o);
	if (o.order != PQP_ORDER_FAST && o.order != PQP_ORDER_STRICT) return PQP_ERR_INVALID;
	if (o.check_every < 1) o.check_every = 1;

	pqp_handle *h = (pqp_handle *)calloc(1, sizeof *h);
	if (!h) return PQP_ERR_ALLOC;
	h->d = *dims;
	h->o = o;
	h->last_kernel = "none";
	env_snapshot_take(&h->env);
	env_scope scope(h);
	int rc = open_device(h, &o);
	if (rc) { pqp_destroy(h); return rc; }

	const int M = dims->M, N = dims->N, nS = dims->nState, nd = dims->nDisH;
	h->ldq = pqp_round_up(N, 32);
	h->have_primal = 1;
	h->have_fp_model = 1;
	h->Mp0 = p->Mp0;
#define UP(field, n)                                                  \
	if ((rc = upload(h, &h->field, p->field, (size_t)(n)))) {     \
		pqp_destroy(h);                                       \
		return rc;                                            \
	}
	UP(Qp_inv, (size_t)M * M) UP(Gp, (size_t)N * M) UP(Kp, N)
	UP(Fp1, (size_t)M * nd) UP(Fp2, (size_t)M * nS) UP(Fp3, M) UP(D, nd)
	if (p->Mp1 && p->Mp2 && p->Mp3 && p->Mp4 && p->Mp5 && p->Mp6 && nS > 0) {
		UP(Mp1, (size_t)nS * nS) UP(Mp2, (size_t)nd * nS) UP(Mp3, (size_t)nd * nd) UP(Mp4, nS) UP(Mp5, nd) UP(Mp6, 1)
	}
#undef UP
	if (nS == 0 && (rc = upload(h, &h->Fp_const, p->Fp, M))) { pqp_destroy(h); return rc; }
	if (p->Kx && nS > 0 && (rc = upload(h, &h->Kx, p->Kx, (size_t)N * nS))) { pqp_destroy(h); return rc; }
	if (p->Kd && nd > 0 && (rc = upload(h, &h->Kd, p->Kd, (size_t)N * nd))) { pqp_destroy(h); return rc; }

	const int strict = o.order == PQP_ORDER_STRICT;
	if ((rc = dalloc(&h->GQ, (size_t)N * M)) || (rc = dalloc(&h->Q, (size_t)N * h->ldq))) { pqp_destroy(h); return rc; }
	cudaError_t e = cudaMemsetAsync(h->Q, 0, (size_t)N * h->ldq * sizeof(float), h->stream);
	/* GQ = Gp*Qp_inv (PQP_CPU.c:492), Qd = GQ*Gp' (PQP_CPU.c:442) */
	const int tensor_setup = !strict && o.use_tensor_cores && N >= 64 && M >= 32;
	float *QiT = NULL; /* tcgen05 3xTF32: both operands K-major, so Qp_inv goes in transposed (allocated before the timed span) */
	if (tensor_setup && (rc = dalloc(&QiT, (size_t)M * M))) { pqp_destroy(h); return rc; }
	/* Qd = Gp Qp_inv Gp' is symmetric in exact arithmetic whenever Qp_inv is: then only the tiles of the upper triangle are multiplied
	 * and every element above the diagonal is stored twice (pqp_gemm_umma_ws.cu) -- half the tensor work, and a Qd that is symmetric
	 * bit for bit.  Whether Qp_inv is symmetric is counted on the device and read by the GEMM there: no host round trip. */
	const char *esym = pqp_env("PQP_GEMM_SYM");
	const int sym_build = tensor_setup && o.exploit_symmetry && pqp_gemm_umma_ws_wanted(N, M, N) && !(esym && atoi(esym) == 0);
	unsigned *qsym_bad = NULL, qsym_bad_host = 1;
	if (sym_build && (rc = dalloc(&qsym_bad, 1))) { cudaFree(QiT); pqp_destroy(h); return rc; }
	if (e == cudaSuccess) e = cudaEventRecord(h->ev0, h->stream);
	if (e == cudaSuccess) {
		if (strict) {
			e = pqp_launch_matmul_strict(h->GQ, M, h->Gp, M, h->Qp_inv, M, 0, N, M, M, h->stream);
			if (e == cudaSuccess) e = pqp_launch_matmul_strict(h->Q, h->ldq, h->GQ, M, h->Gp, M, 1, N, M, N, h->stream);
		} else if (tensor_setup) {
			e = pqp_launch_transpose(QiT, M, h->Qp_inv, M, M, M, h->stream);
			if (e == cudaSuccess) e = pqp_launch_gemm_umma(h->GQ, M, h->Gp, M, QiT, M, N, M, M, h->stream);
			if (sym_build) {
				if (e == cudaSuccess) e = pqp_launch_sym_check(h->Qp_inv, M, M, qsym_bad, h->stream);
				if (e == cudaSuccess) e = pqp_launch_gemm_umma_ws(h->Q, h->ldq, h->GQ, M, h->Gp, M, N, M, N, qsym_bad, h->stream);
				if (e == cudaSuccess) e = cudaMemcpyAsync(&qsym_bad_host, qsym_bad, sizeof qsym_bad_host, cudaMemcpyDeviceToHost, h->stream);
				h->launches += 1;
			} else if (e == cudaSuccess) {
				e = pqp_launch_gemm_umma(h->Q, h->ldq, h->GQ, M, h->Gp, M, N, M, N, h->stream);
			}
			h->launches += 1;
		} else {
			e = pqp_launch_matmul_simt(h->GQ, M, h->Gp, M, h->Qp_inv, M, 0, N, M, M, h->stream);
			if (e == cudaSuccess) e = pqp_launch_matmul_simt(h->Q, h->ldq, h->GQ, M, h->Gp, M, 1, N, M, N, h->stream);
		}
		h->launches += 2;
		if (e == cudaSuccess) e = cudaEventRecord(h->ev1, h->stream);
		if (e == cudaSuccess) e = cudaEventSynchronize(h->ev1);
		if (e == cudaSuccess) e = cudaEventElapsedTime(&h->setup_gemm_ms, h->ev0, h->ev1);
	}
	if (qsym_bad) cudaFree(qsym_bad);
	h->qd_mirrored = sym_build && e == cudaSuccess && qsym_bad_host == 0;
	if (QiT) cudaFree(QiT);
	if (e != cudaSuccess) {
		snprintf(g_cuda_err, sizeof g_cuda_err, "setup GEMMs -> %s", cudaGetErrorString(e));
		pqp_destroy(h);
		return PQP_ERR_CUDA;
	}
	if (!strict && o.exploit_symmetry && !h->qd_mirrored) {
		/* (a Qd built from its upper triangle is symmetric already.)  Gp Qp_inv Gp' is symmetric in exact arithmetic; its two fp32 copies of an element are different sums.  One value per
		 * pair (their mean) when all pairs agree to rounding: then every loop sees a symmetric Qd and the single-problem loop can
		 * run from the upper triangle.  A Qd whose copies differ by more than rounding (an unsymmetric Qp_inv) is left as computed. */
		unsigned *bad_dev = NULL, bad = 0;
		if ((rc = dalloc(&bad_dev, 1))) { pqp_destroy(h); return rc; }
		e = pqp_launch_sym_mean(h->Q, h->ldq, N, 1e-5f, bad_dev, h->stream);
		if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, bad_dev, sizeof bad, cudaMemcpyDeviceToHost, h->stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
		cudaFree(bad_dev);
		h->launches += 2;
		if (e != cudaSuccess) {
			snprintf(g_cuda_err, sizeof g_cuda_err, "symmetry pass -> %s", cudaGetErrorString(e));
			pqp_destroy(h);
			return PQP_ERR_CUDA;
		}
		if (pqp_env("PQP_VERBOSE")) fprintf(stderr, "pqp: setup: %u pairs of Qd differ by more than rounding%s\n", bad, bad ? " (left as computed)" : " (both copies set to their mean)");
	}
	if ((rc = finish_setup(h))) { pqp_destroy(h); return rc; }
	*out = h;
	return PQP_OK;
}

int pqp_setup_dual(pqp_handle **out, int N, const float *Qd, int M, const float *Gp, const float *Qp_inv,
		   const pqp_opts *opts_in)
{
	if (!out || N <= 0 || !Qd) return PQP_ERR_INVALID;
	*out = NULL;
	pqp_opts o;
	if (opts_in) o = *opts_in; else pqp_default_opts(&o);
	if (o.order != PQP_ORDER_FAST && o.order != PQP_ORDER_STRICT) return PQP_ERR_INVALID;
	if (o.check_every < 1) o.check_every = 1;
	pqp_handle *h = (pqp_handle *)calloc(1, sizeof *h);
	if (!h) return PQP_ERR_ALLOC;
	h->d.N = N;
	h->d.M = (Gp && Qp_inv && M > 0) ? M : 0;
	h->o = o;
	h->last_kernel = "none";
	env_snapshot_take(&h->env);
	env_scope scope(h);
	int rc = open_device(h, &o);
	if (rc) { pqp_destroy(h); return rc; }
	h->ldq = pqp_round_up(N, 32);
	h->have_primal = h->d.M > 0;
	if (h->have_primal) {
		if ((rc = upload(h, &h->Gp, Gp, (size_t)N * M)) || (rc = upload(h, &h->Qp_inv, Qp_inv, (size_t)M * M))) {
			pqp_destroy(h);
			return rc;
		}
	}
	if ((rc = dalloc(&h->Q, (size_t)N * h->ldq))) { pqp_destroy(h); return rc; }
	cudaError_t e = cudaMemsetAsync(h->Q, 0, (size_t)N * h->ldq * sizeof(float), h->stream);
	if (e == cudaSuccess)
		e = cudaMemcpy2DAsync(h->Q, (size_t)h->ldq * sizeof(float), Qd, (size_t)N * sizeof(float), (size_t)N * sizeof(float), N,
				      cudaMemcpyDefault, h->stream);
	if (e != cudaSuccess) {
		snprintf(g_cuda_err, sizeof g_cuda_err, "upload Qd -> %s", cudaGetErrorString(e));
		pqp_destroy(h);
		return PQP_ERR_CUDA;
	}
	if ((rc = finish_setup(h))) { pqp_destroy(h); return rc; }
	*out = h;
	return PQP_OK;
}

void pqp_destroy(pqp_handle *h)
{
	if (!h) return;
	if (h->stream) {
		cudaSetDevice(h->device);
		cudaStreamSynchronize(h->stream);
		if (h->l2_window_set) l2_persist_window(h, 0);
		if (h->l2_limit_changed) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, h->l2_limit_saved);
	}
	void *ptrs[] = { h->Q, h->QT, h->theta, h->GQ, h->Gp, h->Qp_inv, h->Kp, h->Fp1, h->Fp2, h->Fp3, h->Fp_const, h->D,
			 h->Kx, h->Kd, h->Mp1, h->Mp2, h->Mp3, h->Mp4, h->Mp5, h->Mp6, h->QpT, h->QnT, h->umma_tiles, h->imma_tiles, h->imma_rowc, h->imma_ptiles, h->imma_prowc, h->acc_ws, h->tol_ws, h->tol_flags, h->X, h->Db, h->Fp, h->Fd, h->Md, h->Y,
			 h->U, h->Tmp, h->st, h->ybuf0, h->ybuf1, h->partials, h->barrier, h->result_buf, h->pk0, h->pk1,
			 h->sym.units, h->sym.cta_u0, h->sym.cta_j0, h->sym.strip_c0, h->sym.strip_c1, h->sym.rowpart, h->sym.colpart };
	for (size_t i = 0; i < sizeof ptrs / sizeof ptrs[0]; i++)
		if (ptrs[i]) cudaFree(ptrs[i]);
	if (h->tol_rem_host) {
		cudaFreeHost(h->tol_rem_host);
		cudaEventDestroy(h->tol_ev[0]);
		cudaEventDestroy(h->tol_ev[1]);
	}
	if (h->ev0) cudaEventDestroy(h->ev0);
	if (h->ev1) cudaEventDestroy(h->ev1);
	if (h->stream) cudaStreamDestroy(h->stream);
	cudaGetLastError();
	env_snapshot_free(&h->env);
	free(h);
}

/* ---- the loop -------------------------------------------------------------------------------- */

/*
 * The upper-triangle loop (pqp_gemv_sym.cu) needs Qd symmetric element for element; examined once per handle, on the
 * first single-problem solve that could use it.  Returns PQP_OK with h->sym_state set to 1 or -1.
 */
static int ensure_sym(pqp_handle *h)
{
	if (h->sym_state) return PQP_OK;
	h->sym_state = -1;
	const int N = h->d.N, G = h->gemv_grid;
	const char *e;
	if ((e = pqp_env("PQP_GEMV_SYM")) && atoi(e) == 0) return PQP_OK;
	if (!h->o.exploit_symmetry || h->o.order == PQP_ORDER_STRICT || G <= 0 || h->small_ok) return PQP_OK;
	const size_t budget = h->smem_optin > 1024 ? h->smem_optin - 1024 : 0;
	pqp_sym_plan pl;
	memset(&pl, 0, sizeof pl);
	if (!pqp_gemv_sym_plan(N, G, budget, &pl.stages, &pl.resident)) return PQP_OK;
	unsigned *bad_dev = NULL, bad = 1;
	int rc;
	if ((rc = dalloc(&bad_dev, 1))) return rc;
	cudaError_t ce0 = pqp_launch_sym_check(h->Q, h->ldq, N, bad_dev, h->stream);
	h->launches++;
	if (ce0 == cudaSuccess) ce0 = cudaMemcpyAsync(&bad, bad_dev, sizeof bad, cudaMemcpyDeviceToHost, h->stream);
	if (ce0 == cudaSuccess) ce0 = cudaStreamSynchronize(h->stream);
	cudaFree(bad_dev);
	CK(ce0);
	if (pqp_env("PQP_VERBOSE")) fprintf(stderr, "pqp: symmetry test of Qd: %u unequal pairs\n", bad);
	if (bad) return PQP_OK;

	pqp_gemv_sym_counts(N, &pl.nb, &pl.U);
	int *tab = (int *)malloc(sizeof(int) * ((size_t)2 * G + 1 + 2 * (size_t)pl.nb));
	if (!tab) return PQP_ERR_ALLOC;
	int *cta_u0 = tab, *cta_j0 = tab + G + 1, *strip_c0 = cta_j0 + G, *strip_c1 = strip_c0 + pl.nb;
	pl.maxseg = pqp_gemv_sym_tables(N, G, cta_u0, cta_j0, strip_c0, strip_c1);
	const int umax = (pl.U + G - 1) / G;
	if ((e = pqp_env("PQP_SYM_RESIDENT"))) {
		int v = atoi(e);
		if (v >= 0 && v < pl.resident) pl.resident = v;
	}
	/* streamed units per CTA fetched evict_last: everything when the streamed part of the triangle fits in ~80% of L2, else a
	 * fraction of L2 spread evenly over the CTAs */
	int l2_bytes = 0;
	cudaDeviceGetAttribute(&l2_bytes, cudaDevAttrL2CacheSize, h->device);
	/* eight more units per CTA can live in the SM's tensor memory (256 KB, otherwise idle in this loop).  Measured (B200): a unit read
	 * back from tensor memory costs about what a streamed unit costs once the copy pipelines are in steady state, so parking pays
	 * where it takes a large share of the traffic away -- N=4096 (everything on chip) 7.6 -> 6.9 us/update, N=6144 (8 of 15
	 * parked) 14.9 -> 13.0 -- and not at N=8192 (8 of 28: 19.3 -> 19.4-19.7): used while at least half of the units outside shared
	 * memory fit */
	pl.tmem = (umax - pl.resident <= 16) ? 8 : 0;
	if ((e = pqp_env("PQP_SYM_TMEM"))) {
		int v = atoi(e);
		if (v >= 0 && v <= 8) pl.tmem = v;
	}
	if (pl.tmem > umax - pl.resident) pl.tmem = umax - pl.resident > 0 ? umax - pl.resident : 0;
	const int streamed = umax - pl.resident - pl.tmem > 0 ? umax - pl.resident - pl.tmem : 0;
	const double unit_bytes = 64.0 * 128.0 * 4.0;
	double frac = 0.5; /* measured at N=8192 (29 units per CTA): 10-16 units per CTA evict_last is a plateau (19.6 us/update); 0: 25.2, 8: 20.5, 18: 20.1, 28: 25.9 */
	if ((e = pqp_env("PQP_SYM_PIN_FRAC"))) frac = atof(e);
	pl.pinned = streamed;
	if ((double)streamed * G * unit_bytes > 0.8 * (double)l2_bytes) pl.pinned = (int)(frac * (double)l2_bytes / ((double)G * unit_bytes));
	if (!h->o.l2_persist) pl.pinned = 0;
	if ((e = pqp_env("PQP_SYM_PIN"))) pl.pinned = atoi(e) > 0 ? atoi(e) : 0;

	const size_t nT = (size_t)pl.nb * (pl.nb + 1) / 2;
	pl.rowpart_bytes = 2 * nT * 128 * 16;
	pl.colpart_bytes = 2 * (size_t)G * pl.maxseg * 128 * 16;
	unsigned char *rp = NULL, *cp = NULL;
	if ((rc = dalloc(&pl.units, pqp_gemv_sym_units_bytes(N) / sizeof(float))) || (rc = dalloc(&pl.cta_u0, (size_t)G + 1)) ||
	    (rc = dalloc(&pl.cta_j0, (size_t)G)) || (rc = dalloc(&pl.strip_c0, (size_t)pl.nb)) || (rc = dalloc(&pl.strip_c1, (size_t)pl.nb)) ||
	    (rc = dalloc(&rp, pl.rowpart_bytes)) || (rc = dalloc(&cp, pl.colpart_bytes))) {
		free(tab);
		void *ptrs[] = { pl.units, pl.cta_u0, pl.cta_j0, pl.strip_c0, pl.strip_c1, rp, cp };
		for (size_t i = 0; i < sizeof ptrs / sizeof ptrs[0]; i++)
			if (ptrs[i]) cudaFree(ptrs[i]);
		cudaGetLastError();
		return PQP_OK; /* no room for the second copy of the triangle: the full-matrix kernels still run */
	}
	pl.rowpart = rp;
	pl.colpart = cp;
	cudaMemcpyAsync(pl.cta_u0, cta_u0, sizeof(int) * ((size_t)G + 1), cudaMemcpyHostToDevice, h->stream);
	cudaMemcpyAsync(pl.cta_j0, cta_j0, sizeof(int) * (size_t)G, cudaMemcpyHostToDevice, h->stream);
	cudaMemcpyAsync(pl.strip_c0, strip_c0, sizeof(int) * (size_t)pl.nb, cudaMemcpyHostToDevice, h->stream);
	cudaMemcpyAsync(pl.strip_c1, strip_c1, sizeof(int) * (size_t)pl.nb, cudaMemcpyHostToDevice, h->stream);
	cudaError_t ce = pqp_launch_build_sym_units(pl.units, h->Q, h->ldq, N, h->stream);
	if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
	free(tab);
	h->sym = pl; /* owned by the handle from here on (freed in pqp_destroy) */
	CK(ce);
	h->launches++;
	if (pqp_env("PQP_VERBOSE"))
		fprintf(stderr, "pqp: gemv_sym N=%d grid=%d units=%d (%d per CTA) stages=%d resident=%d tmem=%d pinned=%d maxseg=%d\n", N, G, pl.U, umax,
			pl.stages, pl.resident, pl.tmem, pl.pinned, pl.maxseg);
	h->sym_state = 1;
	return PQP_OK;
}

/* one problem: Fd (device, N), y0 (device N, or NULL -> y_init); result left in *y_res (device) */
static int run_single(pqp_handle *h, const float *Fd, const float *Md, const float *y0, int iters, const float **y_res,
		      pqp_status *st_dev, int want_status)
{
	const int N = h->d.N, ldq = h->ldq;
	CK(cudaMemsetAsync(h->ybuf0, 0, (size_t)ldq * sizeof(float), h->stream));
	CK(cudaMemsetAsync(h->ybuf1, 0, (size_t)ldq * sizeof(float), h->stream));
	if (y0) {
		CK(cudaMemcpyAsync(h->ybuf0, y0, (size_t)N * sizeof(float), cudaMemcpyDefault, h->stream));
	} else {
		CK(pqp_launch_fill(h->ybuf0, h->o.y_init, N, h->stream));
		h->launches++;
	}

	pqp_gemv_args a;
	memset(&a, 0, sizeof a);
	a.Q = h->Q; a.QT = h->QT; a.ldq = ldq; a.N = N; a.theta = h->theta; a.Fd = Fd; a.Kp = h->Kp; a.Md = Md;
	a.ybuf0 = h->ybuf0; a.ybuf1 = h->ybuf1; a.iters = iters; a.max_iters = h->o.max_iters;
	a.check_every = h->o.check_every; a.erc = h->o.erc; a.eac = h->o.eac; a.eaj = h->o.eaj; a.erj = h->o.erj;
	a.barrier = h->barrier; a.partials = h->partials; a.status = st_dev; a.result_buf = h->result_buf;
	a.grid = h->gemv_grid; a.resident_rows = h->gemv_resident;

	if (h->o.order == PQP_ORDER_STRICT) {
		h->last_kernel = "gemv_strict";
		float *bufs[2] = { h->ybuf0, h->ybuf1 };
		int done = 0;
		if (iters > 0) {
			for (int k = 0; k < iters; k++) CK(pqp_launch_gemv_strict_step(&a, bufs[k & 1], bufs[(k + 1) & 1], h->stream));
			h->launches += iters;
			done = iters;
			if (want_status) {
				CK(pqp_launch_status(st_dev, h->Q, ldq, N, bufs[done & 1], ldq, Fd, Md, h->Kp, h->o.erc, h->o.eac, 1, done + h->iters_base, NULL, h->stream));
				h->launches++;
			}
		} else {
			/* tolerance mode in reference order: test on the host every check_every updates */
			pqp_status hs;
			memset(&hs, 0, sizeof hs);
			for (;;) {
				float viol = INFINITY; /* max_i(-g_i - max(erc*Kp_i, eac)): the per-row tolerance of compare(), PQP_CPU.c:338 */
				CK(pqp_launch_status(st_dev, h->Q, ldq, N, bufs[done & 1], ldq, Fd, Md, h->Kp, h->o.erc, h->o.eac, 1, done, h->partials, h->stream));
				h->launches++;
				CK(cudaMemcpyAsync(&hs, st_dev, sizeof hs, cudaMemcpyDeviceToHost, h->stream));
				CK(cudaMemcpyAsync(&viol, h->partials, sizeof viol, cudaMemcpyDeviceToHost, h->stream));
				CK(cudaStreamSynchronize(h->stream));
				/* the same three conditions the fused kernels test (gemv_cta/small/sym, batched_imma) */
				const int conv = viol <= 0.0f && fabsf(hs.gap) <= h->o.eaj && fabsf(hs.gap) <= h->o.erj * fabsf(hs.Jd);
				if (conv || done >= h->o.max_iters) {
					hs.converged = conv;
					CK(cudaMemcpyAsync(st_dev, &hs, sizeof hs, cudaMemcpyHostToDevice, h->stream));
					break;
				}
				for (int k = 0; k < h->o.check_every && done < h->o.max_iters; k++, done++)
					CK(pqp_launch_gemv_strict_step(&a, bufs[done & 1], bufs[(done + 1) & 1], h->stream));
				h->launches += h->o.check_every;
			}
		}
		*y_res = bufs[done & 1];
		return PQP_OK;
	}

	if (pqp_gemv_cluster_supported(N) && h->cluster_state == 0) {
		/* Decided once per handle by a measurement, not by a table: each kernel at 32 and at 160 updates (CUDA events; y is set up again
		 * before every launch, so the trial leaves nothing behind).  On the B200s measured the cluster wins below N = 512 by 1.1-1.6x, but its
		 * speed rests on SM-to-SM latency inside one GPC, which the floorsweeping of an individual part may change; a box where the
		 * multi-CTA kernel is faster keeps it.  PQP_GEMV_CLUSTER=1 skips the trial, =0 never uses the cluster. */
		const char *e = pqp_env("PQP_GEMV_CLUSTER");
		h->cluster_state = 1;
		const int cta_alt = pqp_gemv_cta_supported(N) && !(pqp_env("PQP_GEMV_CTA") && atoi(pqp_env("PQP_GEMV_CTA")) == 0);
		if (!(e && atoi(e) == 1) && (cta_alt || (h->small_ok && h->gemv_grid > 0))) {
			pqp_gemv_args ta = a;
			ta.status = h->st; /* scratch: the real solve writes it again */
			/* the cost of an UPDATE, not of a launch: the kernels differ in how they bring the matrix on chip, so each is timed at two
			 * counts and the difference counts (rep 0 warms both up) */
			float t_c[2] = { 0.0f, 0.0f }, t_s[2] = { 0.0f, 0.0f };
			const int counts[2] = { 32, 160 };
			for (int rep = 0; rep < 3; rep++) {
				const int k = rep == 0 ? 0 : rep - 1;
				pqp_gemv_args tb = ta;
				tb.iters = counts[k];
				CK(cudaEventRecord(h->ev0, h->stream));
				CK(pqp_launch_gemv_cluster(&tb, h->stream));
				CK(cudaEventRecord(h->ev1, h->stream));
				CK(cudaEventSynchronize(h->ev1));
				CK(cudaEventElapsedTime(&t_c[k], h->ev0, h->ev1));
				tb.grid = h->small_grid;
				CK(cudaEventRecord(h->ev0, h->stream));
				if (cta_alt) CK(pqp_launch_gemv_cta(&tb, h->stream));
				else CK(pqp_launch_gemv_small(&tb, h->small_wpr, h->small_cpt, h->pk0, h->pk1, h->stream));
				CK(cudaEventRecord(h->ev1, h->stream));
				CK(cudaEventSynchronize(h->ev1));
				CK(cudaEventElapsedTime(&t_s[k], h->ev0, h->ev1));
				h->launches += 2;
			}
			const float ms_cluster = fmaxf(t_c[1] - t_c[0], 1e-6f), ms_small = fmaxf(t_s[1] - t_s[0], 1e-6f); /* per 128 updates */
			/* Hysteresis, so that the choice does not flip from handle to handle on noise (and with it the last bits of the results): up to
			 * N = 320, where the cluster wins by 1.6x on most parts, it stays unless it is 1.25x slower here; above, where it wins by 10 %
			 * at best and was measured 1.2-1.5x slower on the GPUs of one 8-GPU box, it has to be 5 % faster to be taken. */
			if (N <= 320 ? (1.25f * ms_small < ms_cluster) : (1.05f * ms_cluster > ms_small)) h->cluster_state = -1;
			if (pqp_env("PQP_VERBOSE"))
				fprintf(stderr, "pqp: one-cluster kernel %.1f us, the alternative %.1f us per 128 updates at N=%d: using the %s\n", 1e3 * ms_cluster,
					1e3 * ms_small, N, h->cluster_state == 1 ? "cluster" : "alternative");
			/* the trials read the caller's y_0 and wrote ybuf1; the launchers reset their own exchange state */
			CK(cudaMemsetAsync(h->ybuf1, 0, (size_t)ldq * sizeof(float), h->stream));
		}
	}
	if (pqp_gemv_cluster_supported(N) && h->cluster_state == 1) {
		/* a mid-size problem (one condensed-MPC QP): one thread-block cluster, y exchanged through distributed shared memory */
		h->last_kernel = iters > 0 ? "gemv_cluster" : "gemv_cluster_tol";
		CK(pqp_launch_gemv_cluster(&a, h->stream));
		h->launches++;
		*y_res = h->ybuf1;
		return PQP_OK;
	}
	if (pqp_gemv_cta_supported(N) && !(pqp_env("PQP_GEMV_CTA") && atoi(pqp_env("PQP_GEMV_CTA")) == 0)) {
		/* a problem that fits one thread block: no exchange through L2 at all */
		h->last_kernel = iters > 0 ? "gemv_cta" : "gemv_cta_tol";
		CK(pqp_launch_gemv_cta(&a, h->stream));
		h->launches++;
		*y_res = h->ybuf1;
		return PQP_OK;
	}
	if (h->gemv_grid <= 0) return PQP_ERR_UNSUPPORTED;
	if (h->small_ok && !(iters <= 0 && pqp_env("PQP_GEMV_SMALL_TOL") && atoi(pqp_env("PQP_GEMV_SMALL_TOL")) == 0)) {
		/* fixed count, or run to tolerance with the stop test evaluated in the kernel every check_every updates */
		h->last_kernel = iters > 0 ? "gemv_small_registers" : "gemv_small_registers_tol";
		a.grid = h->small_grid;
		CK(pqp_launch_gemv_small(&a, h->small_wpr, h->small_cpt, h->pk0, h->pk1, h->stream));
		h->launches++;
		*y_res = h->ybuf1;
		return PQP_OK;
	}
	if (!(iters <= 0 && pqp_env("PQP_GEMV_SYM_TOL") && atoi(pqp_env("PQP_GEMV_SYM_TOL")) == 0)) {
		/* fixed count, or run to tolerance with the stop test evaluated by the owners every check_every updates */
		int rc = ensure_sym(h);
		if (rc) return rc;
		if (h->sym_state == 1) {
			const int all_res = h->sym.resident + h->sym.tmem >= (h->sym.U + h->gemv_grid - 1) / h->gemv_grid;
			h->last_kernel = iters > 0 ? (all_res ? "gemv_sym_resident" : "gemv_sym_stream") : (all_res ? "gemv_sym_resident_tol" : "gemv_sym_stream_tol");
			CK(pqp_launch_gemv_sym(&a, &h->sym, h->pk0, h->pk1, h->stream));
			h->launches++;
			*y_res = h->ybuf1;
			return PQP_OK;
		}
	}
	if (iters > 0 && h->tma_ok) {
		h->last_kernel = h->tma_resident >= (N + h->gemv_grid - 1) / h->gemv_grid + 1 ? "gemv_tma_resident" : "gemv_tma_stream";
		const int ll = !(pqp_env("PQP_GEMV_LL") && atoi(pqp_env("PQP_GEMV_LL")) == 0);
		CK(pqp_launch_gemv_tma(&a, h->tma_stages, h->tma_resident, h->tma_pinned, h->tma_yc, ll ? h->pk0 : NULL, ll ? h->pk1 : NULL,
				       h->stream));
		h->launches++;
		*y_res = ll ? h->ybuf1 : ((iters & 1) ? h->ybuf1 : h->ybuf0);
		return PQP_OK;
	} else {
		h->last_kernel = h->gemv_resident > 0 ? "gemv_persistent_resident" : "gemv_persistent_stream";
		CK(pqp_launch_gemv_persistent(&a, h->stream));
	}
	h->launches++;
	/* which ping-pong buffer holds the result: known on the host for fixed counts */
	if (iters > 0) {
		*y_res = (iters & 1) ? h->ybuf1 : h->ybuf0;
	} else {
		int which = 0;
		CK(cudaMemcpyAsync(&which, h->result_buf, sizeof which, cudaMemcpyDeviceToHost, h->stream));
		CK(cudaStreamSynchronize(h->stream));
		*y_res = which ? h->ybuf1 : h->ybuf0;
	}
	return PQP_OK;
}

/* which batched engine: 0 fp32 SIMT, 1 tcgen05 int8 digit planes (error-free accumulation; the default), 2 tcgen05 3xTF32 */
enum { BATCH_SIMT = 0, BATCH_IMMA = 1, BATCH_UMMA = 2 };
static int batched_engine(const pqp_handle *h)
{
	/*
	 * use_tensor_cores: 0 -> SIMT.  1 (default) -> the int8 kernel: exact integer accumulation in TMEM, accuracy of
	 * PQP_CPU.c's own fp32 arithmetic.  2 -> the 3xTF32 kernel: its fp32 accumulator truncates on every step, which leaves the loop
	 * 5-100x above the fp32 noise floor after 1000 updates, so it is opt-in only.  PQP_BATCHED=simt|imma|umma overrides.
	 */
	const char *e = pqp_env("PQP_BATCHED");
	int want = h->o.use_tensor_cores <= 0 ? BATCH_SIMT : (h->o.use_tensor_cores >= 2 ? BATCH_UMMA : BATCH_IMMA);
	if (e) want = !strcmp(e, "simt") ? BATCH_SIMT : (!strcmp(e, "umma") ? BATCH_UMMA : (!strcmp(e, "imma") ? BATCH_IMMA : want));
	if (pqp_env("PQP_BATCHED_UMMA")) want = atoi(pqp_env("PQP_BATCHED_UMMA")) ? BATCH_UMMA : BATCH_SIMT; /* older knob */
	if (want == BATCH_IMMA && !pqp_batched_imma_supported(h->d.N)) want = BATCH_SIMT;
	if (want == BATCH_UMMA && !pqp_batched_umma_supported(h->d.N)) want = BATCH_SIMT;
	return want;
}

static int ensure_imma_tiles(pqp_handle *h)
{
	if (h->imma_tiles) return PQP_OK;
	unsigned char *t = NULL, *r = NULL;
	int rc;
	if ((rc = dalloc(&t, pqp_batched_imma_tiles_bytes(h->d.N))) || (rc = dalloc(&r, pqp_batched_imma_rowc_bytes(h->d.N)))) return rc;
	h->imma_tiles = t;
	h->imma_rowc = r;
	CK(pqp_launch_build_imma_tiles(h->imma_tiles, h->imma_rowc, h->Q, h->ldq, h->theta, h->d.N, h->stream));
	h->launches++;
	return PQP_OK;
}

/*
 * The PAIRED batched scheme needs Qd[sigma(i)][j] == -Qd[i][j] == Qd[i][sigma(j)] element for element (pqp_batched_imma_pair.cu);
 * examined once per handle, on the first batched solve that could use it.  Leaves h->paired_state at 1 or -1.
 */
static int ensure_imma_paired(pqp_handle *h)
{
	if (h->paired_state) return PQP_OK;
	h->paired_state = -1;
	const int N = h->d.N;
	const char *e;
	if ((e = pqp_env("PQP_IMMA_PAIRED")) && atoi(e) == 0) return PQP_OK;
	if (!h->o.exploit_structure || h->o.order == PQP_ORDER_STRICT || !pqp_batched_imma_paired_supported(N)) return PQP_OK;
	unsigned *bad_dev = NULL, bad = 1;
	int rc;
	if ((rc = dalloc(&bad_dev, 1))) return rc;
	cudaError_t ce = pqp_launch_pair_struct_check(h->Q, h->ldq, N, bad_dev, h->stream);
	h->launches++;
	if (ce == cudaSuccess) ce = cudaMemcpyAsync(&bad, bad_dev, sizeof bad, cudaMemcpyDeviceToHost, h->stream);
	if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
	cudaFree(bad_dev);
	CK(ce);
	if (pqp_env("PQP_VERBOSE")) fprintf(stderr, "pqp: +/- row-pair structure of Qd: %u violations\n", bad);
	if (bad) return PQP_OK;
	unsigned char *t = NULL, *r = NULL;
	if ((rc = dalloc(&t, pqp_batched_imma_paired_tiles_bytes(N))) || (rc = dalloc(&r, pqp_batched_imma_paired_rowc_bytes(N)))) {
		if (t) cudaFree(t);
		return rc;
	}
	h->imma_ptiles = t;
	h->imma_prowc = r;
	CK(pqp_launch_build_imma_tiles_paired(h->imma_ptiles, h->imma_prowc, h->Q, h->ldq, h->theta, N, h->stream));
	h->launches++;
	h->paired_state = 1;
	return PQP_OK;
}

static int ensure_umma_tiles(pqp_handle *h)
{
	if (h->umma_tiles) return PQP_OK;
	unsigned char *t = NULL;
	int rc = dalloc(&t, pqp_batched_umma_tiles_bytes(h->d.N));
	if (rc) return rc;
	h->umma_tiles = t;
	CK(pqp_launch_build_umma_tiles(h->umma_tiles, h->Q, h->ldq, h->theta, h->d.N, h->stream));
	h->launches++;
	return PQP_OK;
}

static int ensure_batched_operands(pqp_handle *h)
{
	if (h->QpT) return PQP_OK;
	const int N = h->d.N;
	h->Kpad = pqp_round_up(N, PQP_BATCH_KPAD);
	h->Ipad = pqp_round_up(N, PQP_BATCH_IPAD);
	int rc;
	if ((rc = dalloc(&h->QpT, (size_t)h->Kpad * h->Ipad)) || (rc = dalloc(&h->QnT, (size_t)h->Kpad * h->Ipad))) return rc;
	CK(pqp_launch_build_split_t(h->QpT, h->QnT, h->Kpad, h->Ipad, h->Q, h->ldq, h->theta, N, h->stream));
	h->launches++;
	return PQP_OK;
}

enum { FUSE_REFRESH = 1, FUSE_RECOVER = 2 };

/*
 * Can this solve run as ONE launch of the paired-rows tensor-core kernel with the h(x) refresh (and the recovery) inside it?
 * Needs: a batch, fixed count, FAST order, the int8 engine, a Qd with the +/- row-pair structure, the handle's own Fp model,
 * no state-dependent constraint offsets (Kx / Kd: the separate kernels keep those), room for the scratch in the operand ring.
 */
/* Batches up to this size go to one thread-block cluster per problem when the size allows (pqp_gemv_cluster.cu; four 16-SM clusters run
 * at a time on a B200, measured at N = 480 per 1000 updates: B = 4 1.0 ms, 8 2.4-2.6, 16 4.2-5.2, 18 5.2-6.4); above, the tensor-core
 * kernels' 8.7 ms for any batch up to 4096 is the shorter or as short. */
#define PQP_CLUSTER_BATCH_MAX 16
static int fused_path(pqp_handle *h, int B, int iters)
{
	const char *e;
	if (B <= PQP_CLUSTER_BATCH_MAX || h->o.order == PQP_ORDER_STRICT || !h->have_fp_model || h->Kx || h->Kd || (iters > 0 && h->o.accelerate > 0)) return 0;
	if (iters <= 0 && pqp_env("PQP_IMMA_PAIRED_TOL") && atoi(pqp_env("PQP_IMMA_PAIRED_TOL")) == 0) return 0;
	if ((e = pqp_env("PQP_IMMA_FUSE")) && atoi(e) == 0) return 0;
	if ((e = pqp_env("PQP_IMMA_PAIR")) && atoi(e) == 0) return 0;
	if (batched_engine(h) != BATCH_IMMA || !pqp_batched_imma_pair_supported(h->d.N)) return 0;
	if (ensure_imma_paired(h) || h->paired_state != 1) return 0;
	return pqp_batched_imma_paired_can_fuse(h->d.N, h->d.M, h->smem_optin);
}

/*
 * Run to tolerance (iters <= 0) on the paired-rows kernel: chunks of check_every updates followed by ONE evaluation pass inside the
 * same launch (g = Qd y + Fd from the sums the loop forms anyway), a tiny decide kernel that applies terminate()'s test per problem
 * (PQP_CPU.c:673-687) and keeps the duals of every problem that passes exactly as they stand, one 4-byte read-back of "how many are
 * still running" per chunk.  The next chunk resumes from the device-resident duals with the scale state the uninterrupted loop would
 * have (m_resume), so a problem's result equals the fixed-count solve at its own count bit for bit (tested).  Problems that have
 * passed keep being multiplied along with their batch mates (their kept result is not touched).
 */
static int tol_chunked(pqp_handle *h, int B, const float *Y0, float *Y, pqp_status *st, int fuse_refresh, const float *Md)
{
	const int N = h->d.N, M = h->d.M;
	if (h->tol_cap < h->cap) {
		if (h->tol_ws) cudaFree(h->tol_ws);
		if (h->tol_flags) cudaFree(h->tol_flags);
		h->tol_ws = NULL;
		h->tol_flags = NULL;
		h->tol_cap = 0;
		int rc;
		if ((rc = dalloc(&h->tol_ws, pqp_paired_eval_part_floats(h->cap) + (size_t)h->cap + (size_t)h->cap * N)) ||
		    (rc = dalloc(&h->tol_flags, (size_t)2 * h->cap + 1)))
			return rc;
		h->tol_cap = h->cap;
	}
	if (!h->tol_rem_host) {
		CK(cudaMallocHost((void **)&h->tol_rem_host, 2 * sizeof(unsigned)));
		CK(cudaEventCreateWithFlags(&h->tol_ev[0], cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&h->tol_ev[1], cudaEventDisableTiming));
	}
	float *part = h->tol_ws, *mres = part + pqp_paired_eval_part_floats(h->cap), *Yres = mres + h->cap;
	unsigned *frozen = h->tol_flags, *newly = frozen + h->cap, *remaining = newly + h->cap;
	CK(cudaMemsetAsync(frozen, 0, (size_t)B * sizeof(unsigned), h->stream));
	if (Y0) CK(cudaMemcpyAsync(h->Y, Y0, (size_t)B * N * sizeof(float), cudaMemcpyDefault, h->stream));
	const int every = h->o.check_every > 0 ? h->o.check_every : 1, cap_it = h->o.max_iters > 0 ? h->o.max_iters : 1;
	/* Chunk k+1 is queued BEFORE the count of chunk k is read: the device runs chunk after chunk without waiting for the host, and the
	 * host learns one chunk late that everybody had passed.  The chunk too many is harmless -- a problem that has passed is frozen: its
	 * kept duals and status are never touched again -- and costs check_every updates once per solve instead of a stream round trip
	 * per chunk. */
	int count = 0, first = 1, k = 0;
	CK(cudaEventRecord(h->ev0, h->stream));
	for (;;) {
		const int n = cap_it - count < every ? cap_it - count : every;
		pqp_paired_fuse fz;
		memset(&fz, 0, sizeof fz);
		fz.M = M;
		if (first && fuse_refresh) {
			fz.X = h->X; fz.D = h->cur_D; fz.D_stride = h->cur_Dstride; fz.nS = h->d.nState; fz.nd = h->d.nDisH;
			fz.Fp1 = h->Fp1; fz.Fp2 = h->Fp2; fz.Fp3 = h->Fp3; fz.Fp_const = h->Fp_const; fz.GQ = h->GQ; fz.Kp = h->Kp;
			fz.Fp_out = h->Fp; fz.Fd_out = h->Fd;
		}
		fz.y0_const = first && Y0 == NULL;
		fz.y_init = h->o.y_init;
		fz.m_resume = first ? NULL : mres;
		fz.m_out = mres;
		fz.eval_part = part;
		fz.tol_Kp = h->Kp; fz.erc = h->o.erc; fz.eac = h->o.eac;
		CK(pqp_launch_batched_imma_paired(h->imma_ptiles, h->imma_prowc, N, B, h->Fd, h->Y, n, h->smem_optin, &fz, h->stream));
		h->launches++;
		count += n;
		if (first && fuse_refresh) {
			h->fp_B = B;
			if (Md) { /* Md needs the Fp the kernel formed */
				CK(pqp_launch_md(h->Md, h->Tmp, h->Fp, h->Qp_inv, h->Mp1, h->Mp2, h->Mp3, h->Mp4, h->Mp5, h->Mp6, h->Mp0, h->cur_D, h->cur_Dstride, h->X, B, M,
						 h->d.nDisH, h->d.nState, h->stream));
				h->launches++;
			}
		}
		CK(pqp_launch_paired_tol_decide(part, Md, h->st, frozen, newly, remaining, Yres, h->Y, B, N, count, cap_it, h->o.eaj, h->o.erj, h->stream));
		h->launches += 2;
		CK(cudaMemcpyAsync(&h->tol_rem_host[k & 1], remaining, sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
		CK(cudaEventRecord(h->tol_ev[k & 1], h->stream));
		first = 0;
		if (k > 0) {
			CK(cudaEventSynchronize(h->tol_ev[(k - 1) & 1]));
			if (h->tol_rem_host[(k - 1) & 1] == 0) break; /* everybody had passed a chunk ago */
		}
		if (count >= cap_it) break; /* the cap: the decide kernel has frozen whoever was left */
		k++;
	}
	CK(cudaEventRecord(h->ev1, h->stream));
	h->ev_valid = 1;
	h->last_kernel = "batched_imma_paired_tol";
	CK(cudaMemcpyAsync(h->Y, Yres, (size_t)B * N * sizeof(float), cudaMemcpyDeviceToDevice, h->stream)); /* the kept duals are the result */
	if (Y) CK(cudaMemcpyAsync(Y, h->Y, (size_t)B * N * sizeof(float), cudaMemcpyDefault, h->stream));
	if (st) CK(cudaMemcpyAsync(st, h->st, (size_t)B * sizeof(pqp_status), cudaMemcpyDefault, h->stream));
	return PQP_OK;
}

/* Fd already in h->Fd [B x N]; Md in h->Md when want_status.  fuse (FUSE_*): the paired kernel forms Fp / Fd itself from h->X
 * (form_linear_terms was told to skip them) and, with FUSE_RECOVER, leaves U in h->U. */
static int run_loop(pqp_handle *h, int B, int iters, const float *Y0, float *Y, pqp_status *st, int fuse)
{
	const int N = h->d.N;
	const int strict = h->o.order == PQP_ORDER_STRICT;
	const int want_status = st != NULL;
	const float *Md = ((want_status || iters <= 0) && h->have_fp_model) ? h->Md : NULL; /* tolerance mode needs Jd, hence Md */
	h->ev_valid = 0;

	/* one block per problem (N <= 64), measured against the tensor-core kernel per 300 updates: N = 28 0.06 / 0.40 / 3.0 ms at B = 16 / 4096 /
	 * 32768 against 1.3 / 1.0 / 7.2; N = 64 0.10 / 0.14 ms at B = 16 / 256 against 1.4 / 1.1, but 1.8 against 1.1 at B = 4096 */
	const int cta_batch_max = pqp_env("PQP_CTA_BATCH_MAX") ? atoi(pqp_env("PQP_CTA_BATCH_MAX")) : (N <= 32 ? (1 << 30) : 2048);
	const int engine = batched_engine(h);
	/* fixed count: any batched engine; run-to-tolerance (iters <= 0): the int8 engine evaluates the stop test per problem itself */
	const int batched = B > 1 && !strict &&
			    (iters > 0 ? (engine != BATCH_SIMT || pqp_batched_simt_supported(N)) : engine == BATCH_IMMA);
	if (iters <= 0 && B > PQP_CLUSTER_BATCH_MAX && !strict && engine == BATCH_IMMA && pqp_batched_imma_pair_supported(N) &&
	    !(pqp_env("PQP_IMMA_PAIR") && atoi(pqp_env("PQP_IMMA_PAIR")) == 0) && !(pqp_env("PQP_IMMA_PAIRED_TOL") && atoi(pqp_env("PQP_IMMA_PAIRED_TOL")) == 0)) {
		int rc = ensure_imma_paired(h);
		if (rc) return rc;
		if (h->paired_state == 1) return tol_chunked(h, B, Y0, Y, st, fuse & FUSE_REFRESH, Md);
	}
	if (fuse) {
		/* one launch: refresh -> loop -> recovery (pqp_batched_imma_paired.cu) */
		const int M = h->d.M;
		if (Y0) CK(cudaMemcpyAsync(h->Y, Y0, (size_t)B * N * sizeof(float), cudaMemcpyDefault, h->stream));
		pqp_paired_fuse fz;
		memset(&fz, 0, sizeof fz);
		fz.X = h->X; fz.D = h->cur_D; fz.D_stride = h->cur_Dstride; fz.nS = h->d.nState; fz.nd = h->d.nDisH; fz.M = M;
		fz.Fp1 = h->Fp1; fz.Fp2 = h->Fp2; fz.Fp3 = h->Fp3; fz.Fp_const = h->Fp_const; fz.GQ = h->GQ; fz.Kp = h->Kp;
		fz.Fp_out = h->Fp; fz.Fd_out = h->Fd;
		if (fuse & FUSE_RECOVER) {
			fz.Gp = h->Gp; fz.Qp_inv = h->Qp_inv; fz.Fp = h->Fp; fz.U = h->U;
		}
		fz.y0_const = Y0 == NULL;
		fz.y_init = h->o.y_init;
		CK(cudaEventRecord(h->ev0, h->stream));
		CK(pqp_launch_batched_imma_paired(h->imma_ptiles, h->imma_prowc, N, B, h->Fd, h->Y, iters, h->smem_optin, &fz, h->stream));
		CK(cudaEventRecord(h->ev1, h->stream));
		h->ev_valid = 1;
		h->launches++;
		h->last_kernel = "batched_imma_paired";
		h->fp_B = B;
		if (want_status) {
			/* Md needs the Fp the kernel formed: after it.  Jd = 1/2 y'Qd y + Fd'y + Md/2 (computeCost, PQP_CPU.c:648-666) */
			CK(pqp_launch_md(h->Md, h->Tmp, h->Fp, h->Qp_inv, h->Mp1, h->Mp2, h->Mp3, h->Mp4, h->Mp5, h->Mp6, h->Mp0, h->cur_D, h->cur_Dstride, h->X, B, M,
					 h->d.nDisH, h->d.nState, h->stream));
			CK(pqp_launch_status(h->st, h->Q, h->ldq, N, h->Y, N, h->Fd, h->Md, h->Kp, h->o.erc, h->o.eac, B, iters + h->iters_base, NULL, h->stream));
			h->launches += 2;
		}
	} else if (B > 1 && B <= cta_batch_max && N <= 64 && !strict && h->o.accelerate <= 0 && !pqp_env("PQP_BATCHED") && pqp_gemv_cta_supported(N) &&
		   !(pqp_env("PQP_GEMV_CTA") && atoi(pqp_env("PQP_GEMV_CTA")) == 0)) {
		/* Problems that fit one thread block (pqp_gemv_cta.cu), any number of them: one block each, many blocks per SM */
		if (Y0) {
			if (Y0 != h->Y) CK(cudaMemcpyAsync(h->Y, Y0, (size_t)B * N * sizeof(float), cudaMemcpyDefault, h->stream));
		} else {
			CK(pqp_launch_fill(h->Y, h->o.y_init, (size_t)B * N, h->stream));
			h->launches++;
		}
		pqp_gemv_args a;
		memset(&a, 0, sizeof a);
		a.Q = h->Q; a.ldq = h->ldq; a.N = N; a.theta = h->theta; a.Fd = h->Fd; a.Kp = h->Kp; a.Md = Md;
		a.ybuf0 = h->Y; a.ybuf1 = h->Y; /* in place: a block reads all of its y_0 before it writes */
		a.iters = iters; a.max_iters = h->o.max_iters; a.check_every = h->o.check_every;
		a.erc = h->o.erc; a.eac = h->o.eac; a.eaj = h->o.eaj; a.erj = h->o.erj;
		a.status = h->st; a.result_buf = h->result_buf;
		CK(cudaEventRecord(h->ev0, h->stream));
		CK(pqp_launch_gemv_cta_batch(&a, B, N, N, h->stream));
		CK(cudaEventRecord(h->ev1, h->stream));
		h->ev_valid = 1;
		h->launches++;
		h->last_kernel = iters > 0 ? "gemv_cta_batch" : "gemv_cta_batch_tol";
	} else if (B > 1 && B <= PQP_CLUSTER_BATCH_MAX && !strict && h->o.accelerate <= 0 && !pqp_env("PQP_BATCHED") && pqp_gemv_cluster_supported(N)) {
		/* A handful of problems of a single controller's size: one thread-block cluster each (pqp_gemv_cluster.cu), as many at a time as
		 * the device has GPCs.  The tensor-core kernels pay the latency of a 64-problem tile whatever B is (8.7 ms per 1000 updates at
		 * N = 480; the single-CTA kernel that used to serve B <= 32: 21 ms); four clusters at a time finish four problems per 1.0 ms. */
		if (Y0) {
			if (Y0 != h->Y) CK(cudaMemcpyAsync(h->Y, Y0, (size_t)B * N * sizeof(float), cudaMemcpyDefault, h->stream));
		} else {
			CK(pqp_launch_fill(h->Y, h->o.y_init, (size_t)B * N, h->stream));
			h->launches++;
		}
		pqp_gemv_args a;
		memset(&a, 0, sizeof a);
		a.Q = h->Q; a.ldq = h->ldq; a.N = N; a.theta = h->theta; a.Fd = h->Fd; a.Kp = h->Kp; a.Md = Md;
		a.ybuf0 = h->Y; a.ybuf1 = h->Y; /* in place: every CTA reads y_0 before the cluster's first barrier and writes its rows after the loop */
		a.iters = iters; a.max_iters = h->o.max_iters; a.check_every = h->o.check_every;
		a.erc = h->o.erc; a.eac = h->o.eac; a.eaj = h->o.eaj; a.erj = h->o.erj;
		a.status = h->st; a.result_buf = h->result_buf;
		CK(cudaEventRecord(h->ev0, h->stream));
		CK(pqp_launch_gemv_cluster_batch(&a, B, N, N, h->stream));
		CK(cudaEventRecord(h->ev1, h->stream));
		h->ev_valid = 1;
		h->launches++;
		h->last_kernel = iters > 0 ? "gemv_cluster_batch" : "gemv_cluster_batch_tol";
	} else if (batched) {
		int rc = engine == BATCH_IMMA ? ensure_imma_tiles(h) : (engine == BATCH_UMMA ? ensure_umma_tiles(h) : ensure_batched_operands(h));
		if (rc) return rc;
		if (Y0) {
			if (Y0 != h->Y) CK(cudaMemcpyAsync(h->Y, Y0, (size_t)B * N * sizeof(float), cudaMemcpyDefault, h->stream));
		} else {
			CK(pqp_launch_fill(h->Y, h->o.y_init, (size_t)B * N, h->stream));
			h->launches++;
		}
		CK(cudaEventRecord(h->ev0, h->stream));
		if (engine == BATCH_IMMA) {
			/* problems per CTA: 32 (measured faster than 64 at every batch size: 195k vs 103k solves/s at B=4096, 225k vs 165k
			 * at B=32768; the 64-problem tile is kept as an experiment knob) */
			int nb = 32;
			if (pqp_env("PQP_IMMA_NB")) nb = atoi(pqp_env("PQP_IMMA_NB")) == 64 ? 64 : 32;
			int cluster = 1;
			if (pqp_env("PQP_IMMA_CLUSTER")) cluster = atoi(pqp_env("PQP_IMMA_CLUSTER"));
			if (cluster != 1 && cluster != 2 && cluster != 4 && cluster != 8 && cluster != 16) cluster = 1;
			/* fixed count, more than one 32-problem tile, at least two M tiles: the CTA-pair kernel (64 problems per pair, rows of Q
			 * split over the two SMs) moves half the operand bytes per problem through shared memory */
			const int pair = iters > 0 && B > PQP_CLUSTER_BATCH_MAX && pqp_batched_imma_pair_supported(N) &&
					 (pqp_env("PQP_IMMA_PAIR") ? atoi(pqp_env("PQP_IMMA_PAIR")) != 0 : 1);
			if (pair) {
				/* a Qd with the +/- row-pair structure of a box-constrained MPC dual: half the tensor work (PAIRED instantiation) */
				int rc2 = ensure_imma_paired(h);
				if (rc2) return rc2;
				if (h->paired_state == 1) {
					CK(pqp_launch_batched_imma_paired(h->imma_ptiles, h->imma_prowc, N, B, h->Fd, h->Y, iters, h->smem_optin, NULL, h->stream));
					h->last_kernel = "batched_imma_paired";
				} else {
					CK(pqp_launch_batched_imma_pair(h->imma_tiles, h->imma_rowc, N, B, h->Fd, h->Y, iters, h->smem_optin, h->stream));
					h->last_kernel = "batched_imma_pair";
				}
			} else {
			pqp_imma_tol t;
			t.max_iters = h->o.max_iters; t.check_every = h->o.check_every;
			t.erc = h->o.erc; t.eac = h->o.eac; t.eaj = h->o.eaj; t.erj = h->o.erj;
			t.Kp = h->Kp; t.Md = Md; t.status = h->st;
			CK(pqp_launch_batched_imma(h->imma_tiles, h->imma_rowc, N, B, h->Fd, h->Y, iters, nb, cluster, h->smem_optin, &t, h->stream));
			h->last_kernel = "batched_imma";
			}
		} else if (engine == BATCH_UMMA) {
			int cluster = 4;
			if (pqp_env("PQP_UMMA_CLUSTER")) cluster = atoi(pqp_env("PQP_UMMA_CLUSTER"));
			if (cluster != 1 && cluster != 2 && cluster != 4 && cluster != 8) cluster = 1;
			CK(pqp_launch_batched_umma(h->umma_tiles, N, B, h->Fd, h->Y, iters, cluster, h->stream));
			h->last_kernel = "batched_umma";
		} else {
			CK(pqp_launch_batched_simt_split(h->QpT, h->QnT, h->Kpad, h->Ipad, N, B, h->Fd, h->Y, iters, h->stream));
			h->last_kernel = "batched_simt";
		}
		CK(cudaEventRecord(h->ev1, h->stream));
		h->ev_valid = 1;
		h->launches++;
		if (want_status && iters > 0) { /* in tolerance mode the kernel wrote the status of every problem itself */
			CK(pqp_launch_status(h->st, h->Q, h->ldq, N, h->Y, N, h->Fd, Md, h->Kp, h->o.erc, h->o.eac, B, iters + h->iters_base, NULL, h->stream));
			h->launches++;
		}
	} else {
		/* single-problem kernel, one problem after the other */
		if (h->o.l2_persist && pqp_env("PQP_L2_WINDOW") && !strict && !h->l2_window_set) {
			int rc = l2_persist_window(h, 1);
			if (rc) return rc;
		}
		CK(cudaEventRecord(h->ev0, h->stream));
		for (int b = 0; b < B; b++) {
			const float *y_res = NULL;
			int rc = run_single(h, h->Fd + (size_t)b * N, Md ? Md + b : NULL, Y0 ? Y0 + (size_t)b * N : NULL, iters, &y_res,
					    h->st + b, want_status);
			if (rc) return rc;
			CK(cudaMemcpyAsync(h->Y + (size_t)b * N, y_res, (size_t)N * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
		}
		CK(cudaEventRecord(h->ev1, h->stream));
		h->ev_valid = 1;
	}
	if (Y) CK(cudaMemcpyAsync(Y, h->Y, (size_t)B * N * sizeof(float), cudaMemcpyDefault, h->stream));
	if (st) CK(cudaMemcpyAsync(st, h->st, (size_t)B * sizeof(pqp_status), cudaMemcpyDefault, h->stream));
	return PQP_OK;
}

/*
 * opts.accelerate = P > 0, fixed count: the K multiplicative updates run in chunks of P (each chunk on the loop kernel the plain
 * solve would use) with one acceleration / line-search step between chunks (pqp_launch_accel_step; PQP_CPU.c:545-588, :625-630 --
 * the branch the reference leaves dead behind `if(1)`, :721-735), not counted in K and not applied after the last update.
 */
static int run_loop_any(pqp_handle *h, int B, int iters, const float *Y0, float *Y, pqp_status *st, int fuse)
{
	const int P = h->o.accelerate;
	h->iters_base = 0;
	if (P <= 0 || iters <= 0 || iters <= P) return run_loop(h, B, iters, Y0, Y, st, fuse);
	const int N = h->d.N;
	if (h->acc_cap < h->cap) {
		if (h->acc_ws) cudaFree(h->acc_ws);
		h->acc_ws = NULL;
		h->acc_cap = 0;
		int rc = dalloc(&h->acc_ws, (size_t)3 * h->cap * N);
		if (rc) return rc;
		h->acc_cap = h->cap;
	}
	int done = 0;
	const float *y0 = Y0;
	while (done < iters) {
		const int n = iters - done < P ? iters - done : P, last = done + n == iters;
		h->iters_base = done;
		int rc = run_loop(h, B, n, y0, last ? Y : NULL, last ? st : NULL, 0);
		if (rc) return rc;
		done += n;
		y0 = h->Y;
		if (!last) {
			CK(pqp_launch_accel_step(h->Y, N, h->Q, h->ldq, h->Fd, h->acc_ws, B, N, h->stream));
			h->launches += 4;
		}
	}
	h->iters_base = 0;
	return PQP_OK;
}

/* skip_fp_fd: the loop kernel forms Fp, Fd (and Md is launched behind it) -- only X and D are brought to the device here */
static int form_linear_terms(pqp_handle *h, const float *X, const float *D, int B, int want_status, int skip_fp_fd)
{
	const int M = h->d.M, N = h->d.N, nS = h->d.nState, nd = h->d.nDisH;
	const int strict = h->o.order == PQP_ORDER_STRICT;
	if (!h->have_fp_model) return PQP_ERR_INVALID;
	if (nS > 0 && !X) return PQP_ERR_INVALID;
	if (nS > 0) CK(cudaMemcpyAsync(h->X, X, (size_t)B * nS * sizeof(float), cudaMemcpyDefault, h->stream));
	const float *Dd = h->D;
	int Dstride = 0;
	if (D && nd > 0) {
		CK(cudaMemcpyAsync(h->Db, D, (size_t)B * nd * sizeof(float), cudaMemcpyDefault, h->stream));
		Dd = h->Db;
		Dstride = nd;
	}
	h->cur_D = Dd;
	h->cur_Dstride = Dstride;
	if (skip_fp_fd) return PQP_OK;
	CK(pqp_launch_fp(h->Fp, h->Fp1, h->Fp2, h->Fp3, h->Fp_const, Dd, Dstride, h->X, B, M, nd, nS, h->stream));
	CK(pqp_launch_fd(h->Fd, h->GQ, h->Fp, h->Kp, B, N, M, strict, h->stream));
	h->launches += 2;
	if (h->Kx || h->Kd) { /* Kp(x, D) = Kp + Kx*x + Kd*D */
		CK(pqp_launch_fd_offsets(h->Fd, h->Kx, h->X, nS, h->Kd, Dd, Dstride, nd, B, N, h->stream));
		h->launches++;
	}
	h->fp_B = B;
	if (want_status) {
		CK(pqp_launch_md(h->Md, h->Tmp, h->Fp, h->Qp_inv, h->Mp1, h->Mp2, h->Mp3, h->Mp4, h->Mp5, h->Mp6, h->Mp0, Dd, Dstride, h->X, B, M,
				 nd, nS, h->stream));
		h->launches++;
	}
	return PQP_OK;
}

int pqp_solve_batch(pqp_handle *h, const float *X, const float *D, int B, int iters, const float *Y0, float *Y,
		    pqp_status *st)
{
	if (!h || B <= 0) return PQP_ERR_INVALID;
	env_scope scope(h);
	CK(cudaSetDevice(h->device));
	int rc = ensure_capacity(h, B);
	if (rc) return rc;
	const int fuse = fused_path(h, B, iters) ? FUSE_REFRESH : 0;
	if ((rc = form_linear_terms(h, X, D, B, st != NULL || iters <= 0, fuse))) return rc;
	if ((rc = run_loop_any(h, B, iters, Y0, Y, st, fuse))) return rc;
	CK(cudaStreamSynchronize(h->stream));
	return PQP_OK;
}

int pqp_solve_dual_full(pqp_handle *h, const float *Fd, const float *Md, int B, int iters, const float *Y0, float *Y, pqp_status *st)
{
	if (!h || B <= 0 || !Fd) return PQP_ERR_INVALID;
	env_scope scope(h);
	CK(cudaSetDevice(h->device));
	int rc = ensure_capacity(h, B);
	if (rc) return rc;
	CK(cudaMemcpyAsync(h->Fd, Fd, (size_t)B * h->d.N * sizeof(float), cudaMemcpyDefault, h->stream));
	if (Md) CK(cudaMemcpyAsync(h->Md, Md, (size_t)B * sizeof(float), cudaMemcpyDefault, h->stream));
	h->fp_B = 0;
	const int saved = h->have_fp_model;
	h->have_fp_model = Md != NULL; /* without Md, Jd is reported (and tested) without the constant Md/2 */
	rc = run_loop_any(h, B, iters, Y0, Y, st, 0);
	h->have_fp_model = saved;
	if (rc) return rc;
	CK(cudaStreamSynchronize(h->stream));
	return PQP_OK;
}

int pqp_solve_dual(pqp_handle *h, const float *Fd, int B, int iters, const float *Y0, float *Y, pqp_status *st)
{
	return pqp_solve_dual_full(h, Fd, NULL, B, iters, Y0, Y, st);
}

int pqp_set_constraint_bounds(pqp_handle *h, const float *Kp)
{
	if (!h) return PQP_ERR_INVALID;
	CK(cudaSetDevice(h->device));
	if (!Kp) {
		if (h->Kp) cudaFree(h->Kp);
		h->Kp = NULL;
		return PQP_OK;
	}
	if (!h->Kp) {
		int rc = dalloc(&h->Kp, (size_t)h->d.N);
		if (rc) return rc;
	}
	CK(cudaMemcpyAsync(h->Kp, Kp, (size_t)h->d.N * sizeof(float), cudaMemcpyDefault, h->stream));
	CK(cudaStreamSynchronize(h->stream));
	return PQP_OK;
}

static int recover_on_device(pqp_handle *h, const float *Ydev, int ldy, const float *Fp, int B, float *U)
{
	const int M = h->d.M, N = h->d.N;
	if (!h->have_primal || M <= 0) return PQP_ERR_INVALID;
	if (Fp) {
		CK(cudaMemcpyAsync(h->Fp, Fp, (size_t)B * M * sizeof(float), cudaMemcpyDefault, h->stream));
		h->fp_B = B;
	} else if (h->fp_B < B) {
		return PQP_ERR_INVALID; /* no cached Fp for these problems */
	}
	CK(pqp_launch_recover(h->U, h->Tmp, Ydev, ldy, h->Fp, h->Gp, h->Qp_inv, B, N, M, h->o.order == PQP_ORDER_STRICT, h->stream));
	h->launches += 2;
	if (U) CK(cudaMemcpyAsync(U, h->U, (size_t)B * M * sizeof(float), cudaMemcpyDefault, h->stream));
	return PQP_OK;
}

int pqp_recover_primal(pqp_handle *h, const float *Y, const float *Fp, int B, float *U)
{
	if (!h || !Y || !U || B <= 0) return PQP_ERR_INVALID;
	env_scope scope(h);
	CK(cudaSetDevice(h->device));
	if (B > h->cap) {
		/* growing the workspace would drop the cached Fp; only legal when the caller supplies Fp */
		if (!Fp) return PQP_ERR_INVALID;
		int rc = ensure_capacity(h, B);
		if (rc) return rc;
	}
	CK(cudaMemcpyAsync(h->Y, Y, (size_t)B * h->d.N * sizeof(float), cudaMemcpyDefault, h->stream));
	int rc = recover_on_device(h, h->Y, h->d.N, Fp, B, U);
	if (rc) return rc;
	CK(cudaStreamSynchronize(h->stream));
	return PQP_OK;
}

int pqp_solve_batch_primal(pqp_handle *h, const float *X, const float *D, int B, int iters, const float *Y0, float *Y,
			   float *U, pqp_status *st)
{
	if (!h || B <= 0 || !U) return PQP_ERR_INVALID;
	env_scope scope(h);
	CK(cudaSetDevice(h->device));
	int rc = ensure_capacity(h, B);
	if (rc) return rc;
	const int fuse = (h->have_primal && h->d.M > 0 && fused_path(h, B, iters)) ? (iters > 0 ? (FUSE_REFRESH | FUSE_RECOVER) : FUSE_REFRESH) : 0;
	if ((rc = form_linear_terms(h, X, D, B, st != NULL || iters <= 0, fuse))) return rc;
	if ((rc = run_loop_any(h, B, iters, Y0, Y, st, fuse))) return rc;
	if (fuse & FUSE_RECOVER) CK(cudaMemcpyAsync(U, h->U, (size_t)B * h->d.M * sizeof(float), cudaMemcpyDefault, h->stream)); /* the kernel's epilogue left U in h->U */
	else if ((rc = recover_on_device(h, h->Y, h->d.N, NULL, B, U))) return rc;
	CK(cudaStreamSynchronize(h->stream));
	return PQP_OK;
}

/* ---- receding-horizon warm start -------------------------------------------------------------- */
int pqp_shift_duals(pqp_handle *h, const float *Y, int B, float y_floor, float *Ynext)
{
	if (!h || !Y || !Ynext || B <= 0) return PQP_ERR_INVALID;
	env_scope scope(h);
	const int N = h->d.N, pH = h->d.pHorizon, nI = h->d.nInput;
	if (pH <= 0 || nI <= 0 || 4 * pH * nI != N) return PQP_ERR_INVALID; /* not an MPC-structured dual */
	CK(cudaSetDevice(h->device));
	if (B > h->cap) {
		int rc = ensure_capacity(h, B);
		if (rc) return rc;
	}
	/* staged through the workspace: h->Fd is rewritten by every solve, so it is free between solves */
	CK(cudaMemcpyAsync(h->Fd, Y, (size_t)B * N * sizeof(float), cudaMemcpyDefault, h->stream));
	CK(pqp_launch_shift_duals(h->Y, h->Fd, B, pH, nI, y_floor, h->stream));
	h->launches++;
	CK(cudaMemcpyAsync(Ynext, h->Y, (size_t)B * N * sizeof(float), cudaMemcpyDefault, h->stream));
	CK(cudaStreamSynchronize(h->stream));
	return PQP_OK;
}

/* ---- updateY2 + updY (PQP_CPU.c:603-618, 590-596) on the reference's dense operands ---- */
int pqp_update_y2(float *Y_next, const float *Y, const float *Qdp_theta, const float *Qdn_theta, const float *Fdp, const float *Fdn, int N,
		  int device)
{
	if (!Y_next || !Y || !Qdp_theta || !Qdn_theta || !Fdp || !Fdn || N <= 0) return PQP_ERR_INVALID;
	if (pqp_device_count() == 0) return PQP_ERR_NO_DEVICE;
	if (device >= 0) CK(cudaSetDevice(device));
	float *dQp = NULL, *dQn = NULL, *dv = NULL; /* dv: Y | Fdp | Fdn | Yn */
	int rc = PQP_OK;
	if ((rc = dalloc(&dQp, (size_t)N * N)) || (rc = dalloc(&dQn, (size_t)N * N)) || (rc = dalloc(&dv, (size_t)4 * N))) goto done;
	{
		const size_t nn = (size_t)N * N * sizeof(float), nv = (size_t)N * sizeof(float);
		cudaError_t e = cudaMemcpy(dQp, Qdp_theta, nn, cudaMemcpyDefault);
		if (e == cudaSuccess) e = cudaMemcpy(dQn, Qdn_theta, nn, cudaMemcpyDefault);
		if (e == cudaSuccess) e = cudaMemcpy(dv, Y, nv, cudaMemcpyDefault);
		if (e == cudaSuccess) e = cudaMemcpy(dv + N, Fdp, nv, cudaMemcpyDefault);
		if (e == cudaSuccess) e = cudaMemcpy(dv + 2 * (size_t)N, Fdn, nv, cudaMemcpyDefault);
		if (e == cudaSuccess) e = pqp_launch_update_y2_dense(dv + 3 * (size_t)N, dv, dQp, dQn, dv + N, dv + 2 * (size_t)N, N, 0);
		if (e == cudaSuccess) e = cudaMemcpy(Y_next, dv + 3 * (size_t)N, nv, cudaMemcpyDefault);
		if (e != cudaSuccess) {
			snprintf(g_cuda_err, sizeof g_cuda_err, "pqp_update_y2 -> %s", cudaGetErrorString(e));
			cudaGetLastError();
			rc = PQP_ERR_CUDA;
		}
	}
done:
	if (dQp) cudaFree(dQp);
	if (dQn) cudaFree(dQn);
	if (dv) cudaFree(dv);
	return rc;
}

/* ---- stateless reference helpers on the device (what the compat library's computeFp / computeCost / convertToDual(Md) /
 * computeUfromY call: no host arithmetic anywhere in the product) ------------------------------------------------------------- */
struct dev_scratch {
	float *p[8];
	int n;
	dev_scratch() : n(0) { memset(p, 0, sizeof p); }
	~dev_scratch()
	{
		for (int i = 0; i < n; i++)
			if (p[i]) cudaFree(p[i]);
	}
	/* device buffer of `count` floats, filled from `src` (host or device) when src != NULL */
	float *get(size_t count, const float *src, int *rc)
	{
		float *d = NULL;
		if (*rc) return NULL;
		if ((*rc = dalloc(&d, count))) return NULL;
		p[n++] = d;
		if (src && cudaMemcpy(d, src, count * sizeof(float), cudaMemcpyDefault) != cudaSuccess) {
			snprintf(g_cuda_err, sizeof g_cuda_err, "upload of %zu floats failed", count);
			cudaGetLastError();
			*rc = PQP_ERR_CUDA;
		}
		return d;
	}
};

static int finish_stateless(const char *what, cudaError_t e, float *dst, const float *src_dev, size_t count)
{
	if (e == cudaSuccess) e = cudaStreamSynchronize(0);
	if (e == cudaSuccess) e = cudaMemcpy(dst, src_dev, count * sizeof(float), cudaMemcpyDefault);
	if (e != cudaSuccess) {
		snprintf(g_cuda_err, sizeof g_cuda_err, "%s -> %s", what, cudaGetErrorString(e));
		cudaGetLastError();
		return PQP_ERR_CUDA;
	}
	return PQP_OK;
}

int pqp_compute_fp(float *Fp, const float *Fp1, const float *Fp2, const float *Fp3, const float *D, const float *x, int M, int nDisH, int nState,
		   int device)
{
	if (!Fp || !Fp3 || M <= 0 || nDisH < 0 || nState < 0 || (nDisH > 0 && (!Fp1 || !D)) || (nState > 0 && (!Fp2 || !x))) return PQP_ERR_INVALID;
	if (pqp_device_count() == 0) return PQP_ERR_NO_DEVICE;
	if (device >= 0) CK(cudaSetDevice(device));
	dev_scratch sc;
	int rc = PQP_OK;
	float *dFp = sc.get((size_t)M, NULL, &rc), *d1 = sc.get((size_t)M * (nDisH > 0 ? nDisH : 1), nDisH > 0 ? Fp1 : NULL, &rc),
	      *d2 = sc.get((size_t)M * (nState > 0 ? nState : 1), nState > 0 ? Fp2 : NULL, &rc), *d3 = sc.get((size_t)M, Fp3, &rc),
	      *dD = sc.get((size_t)(nDisH > 0 ? nDisH : 1), nDisH > 0 ? D : NULL, &rc), *dx = sc.get((size_t)(nState > 0 ? nState : 1), nState > 0 ? x : NULL, &rc);
	if (rc) return rc;
	if (nState == 0) { /* the kernel reads "nState == 0" as "Fp is a constant": keep computeFp's formula with an empty x term */
		CK(cudaMemset(dx, 0, sizeof(float)));
		CK(cudaMemset(d2, 0, (size_t)M * sizeof(float)));
		nState = 1;
	}
	return finish_stateless("pqp_compute_fp", pqp_launch_fp(dFp, d1, d2, d3, NULL, dD, 0, dx, 1, M, nDisH, nState, 0), Fp, dFp, (size_t)M);
}

int pqp_compute_cost(float *J, const float *Z, const float *Q, const float *F, const float *Mc, int N, int device)
{
	if (!J || !Z || !Q || !F || N <= 0) return PQP_ERR_INVALID;
	if (pqp_device_count() == 0) return PQP_ERR_NO_DEVICE;
	if (device >= 0) CK(cudaSetDevice(device));
	dev_scratch sc;
	int rc = PQP_OK;
	float *dz = sc.get((size_t)N, Z, &rc), *dQ = sc.get((size_t)N * N, Q, &rc), *dF = sc.get((size_t)N, F, &rc), *dm = sc.get(1, Mc, &rc),
	      *dt = sc.get((size_t)N, NULL, &rc), *dJ = sc.get(1, NULL, &rc);
	if (rc) return rc;
	return finish_stateless("pqp_compute_cost", pqp_launch_quad_form(dJ, dt, dz, dQ, dF, Mc ? dm : NULL, N, 0, 0), J, dJ, 1);
}

int pqp_compute_md(float *Md, const float *Fp, const float *Qp_inv, const float *Mp, int M, int device)
{
	if (!Md || !Fp || !Qp_inv || M <= 0) return PQP_ERR_INVALID;
	if (pqp_device_count() == 0) return PQP_ERR_NO_DEVICE;
	if (device >= 0) CK(cudaSetDevice(device));
	dev_scratch sc;
	int rc = PQP_OK;
	float *dz = sc.get((size_t)M, Fp, &rc), *dQ = sc.get((size_t)M * M, Qp_inv, &rc), *dm = sc.get(1, Mp, &rc), *dt = sc.get((size_t)M, NULL, &rc),
	      *dJ = sc.get(1, NULL, &rc);
	if (rc) return rc;
	return finish_stateless("pqp_compute_md", pqp_launch_quad_form(dJ, dt, dz, dQ, NULL, Mp ? dm : NULL, M, 1, 0), Md, dJ, 1);
}

int pqp_compute_u_from_y(float *U, const float *Y, const float *Fp, const float *Gp, const float *Qp_inv, int N, int M, int B, int device)
{
	if (!U || !Y || !Fp || !Gp || !Qp_inv || N <= 0 || M <= 0 || B <= 0) return PQP_ERR_INVALID;
	if (pqp_device_count() == 0) return PQP_ERR_NO_DEVICE;
	if (device >= 0) CK(cudaSetDevice(device));
	dev_scratch sc;
	int rc = PQP_OK;
	float *dY = sc.get((size_t)B * N, Y, &rc), *dFp = sc.get((size_t)B * M, Fp, &rc), *dG = sc.get((size_t)N * M, Gp, &rc),
	      *dQi = sc.get((size_t)M * M, Qp_inv, &rc), *dT = sc.get((size_t)B * M, NULL, &rc), *dU = sc.get((size_t)B * M, NULL, &rc);
	if (rc) return rc;
	return finish_stateless("pqp_compute_u_from_y", pqp_launch_recover(dU, dT, dY, N, dFp, dG, dQi, B, N, M, 1, 0), U, dU, (size_t)B * M);
}

/* ---- matrixMultiply (PQP_CPU.c:84-147) on the device ---------------------------------------- */
int pqp_matmul(float *out, const float *A, int tA, const float *B, int tB, int a, int b, int c, int engine, int device)
{
	if (!out || !A || !B || a <= 0 || b <= 0 || c <= 0) return PQP_ERR_INVALID;
	if (engine != PQP_MM_STRICT && engine != PQP_MM_SIMT && engine != PQP_MM_TENSOR) return PQP_ERR_INVALID;
	if (pqp_device_count() == 0) return PQP_ERR_NO_DEVICE;
	if (device >= 0) CK(cudaSetDevice(device));
	float *dA = NULL, *dB = NULL, *dC = NULL, *dT = NULL;
	int rc = PQP_OK;
	cudaStream_t s = 0;
	if ((rc = dalloc(&dA, (size_t)a * b)) || (rc = dalloc(&dB, (size_t)b * c)) || (rc = dalloc(&dC, (size_t)a * c)) ||
	    (rc = dalloc(&dT, (size_t)(a > c ? a : c) * b)))
		goto done;
	{
		cudaError_t e = cudaMemcpy(dA, A, (size_t)a * b * sizeof(float), cudaMemcpyDefault);
		if (e == cudaSuccess) e = cudaMemcpy(dB, B, (size_t)b * c * sizeof(float), cudaMemcpyDefault);
		/* normalise to A [a x b] row-major (K-major) */
		const float *Ak = dA;
		if (e == cudaSuccess && tA) { /* stored [b x a] */
			e = pqp_launch_transpose(dT, b, dA, a, b, a, s);
			if (e == cudaSuccess) e = cudaMemcpyAsync(dA, dT, (size_t)a * b * sizeof(float), cudaMemcpyDeviceToDevice, s);
		}
		/* and B as Bt [c x b] (K-major) */
		if (e == cudaSuccess && !tB) { /* stored [b x c] */
			e = pqp_launch_transpose(dT, b, dB, c, b, c, s);
			if (e == cudaSuccess) e = cudaMemcpyAsync(dB, dT, (size_t)b * c * sizeof(float), cudaMemcpyDeviceToDevice, s);
		}
		if (e == cudaSuccess) {
			if (engine == PQP_MM_STRICT) e = pqp_launch_matmul_strict(dC, c, Ak, b, dB, b, 1, a, b, c, s);
			else if (engine == PQP_MM_SIMT) e = pqp_launch_matmul_simt(dC, c, Ak, b, dB, b, 1, a, b, c, s);
			else e = pqp_launch_gemm_umma(dC, c, Ak, b, dB, b, a, b, c, s);
		}
		if (e == cudaSuccess) e = cudaStreamSynchronize(s);
		if (e == cudaSuccess) e = cudaMemcpy(out, dC, (size_t)a * c * sizeof(float), cudaMemcpyDefault);
		if (e != cudaSuccess) {
			snprintf(g_cuda_err, sizeof g_cuda_err, "pqp_matmul -> %s", cudaGetErrorString(e));
			cudaGetLastError();
			rc = PQP_ERR_CUDA;
		}
	}
done:
	cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dT);
	return rc;
}

/* ---- introspection ---------------------------------------------------------------------------- */
int pqp_get_dual(pqp_handle *h, float *Qd, float *theta, float *GQ)
{
	if (!h) return PQP_ERR_INVALID;
	CK(cudaSetDevice(h->device));
	const int N = h->d.N, M = h->d.M;
	if (Qd)
		CK(cudaMemcpy2DAsync(Qd, (size_t)N * sizeof(float), h->Q, (size_t)h->ldq * sizeof(float), (size_t)N * sizeof(float), N,
				     cudaMemcpyDefault, h->stream));
	if (theta) CK(cudaMemcpyAsync(theta, h->theta, (size_t)N * sizeof(float), cudaMemcpyDefault, h->stream));
	if (GQ) {
		if (!h->GQ) return PQP_ERR_INVALID;
		CK(cudaMemcpyAsync(GQ, h->GQ, (size_t)N * M * sizeof(float), cudaMemcpyDefault, h->stream));
	}
	CK(cudaStreamSynchronize(h->stream));
	return PQP_OK;
}

int pqp_get_linear_terms(pqp_handle *h, int B, float *Fd, float *Fp)
{
	if (!h || B <= 0 || B > h->cap) return PQP_ERR_INVALID;
	CK(cudaSetDevice(h->device));
	if (Fd) CK(cudaMemcpyAsync(Fd, h->Fd, (size_t)B * h->d.N * sizeof(float), cudaMemcpyDefault, h->stream));
	if (Fp) {
		if (h->fp_B < B) return PQP_ERR_INVALID;
		CK(cudaMemcpyAsync(Fp, h->Fp, (size_t)B * h->d.M * sizeof(float), cudaMemcpyDefault, h->stream));
	}
	CK(cudaStreamSynchronize(h->stream));
	return PQP_OK;
}

void *pqp_get_stream(pqp_handle *h) { return h ? (void *)h->stream : NULL; }

float pqp_last_solve_ms(pqp_handle *h)
{
	if (!h || !h->ev_valid) return -1.0f;
	float ms = -1.0f;
	if (cudaEventSynchronize(h->ev1) != cudaSuccess) return -1.0f;
	if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) return -1.0f;
	return ms;
}

float pqp_setup_gemm_ms(pqp_handle *h) { return h ? h->setup_gemm_ms : -1.0f; }
int pqp_setup_mirrored(pqp_handle *h) { return h ? h->qd_mirrored : 0; }
long long pqp_launch_count(pqp_handle *h) { return h ? h->launches : 0; }
const char *pqp_last_kernel(pqp_handle *h) { return h ? h->last_kernel : "none"; }
const float *pqp_device_qd(pqp_handle *h, int *ld)
{
	if (!h) return NULL;
	if (ld) *ld = h->ldq;
	return h->Q;
}
