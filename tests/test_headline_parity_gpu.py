"""GPU (-m gpu): parity AT THE HEADLINE SIZES, against the oracle itself (round-1 review, item 1).

What the measurements say (B200, tools/headline_parity_probe.py, gpurun_out/r2_parity_probe.log):

  * N = 8192: PQP_CPU.c's own float arithmetic (the oracle) is 2.0e-5 from its float64 twin after 100 updates even on a
    well-conditioned dual (M = N = 8192, Qd > 0) -- every row sum is 8192 sequentially rounded fp32 additions.  The CUDA
    loops (upper-triangle and full-matrix) land 3e-6 .. 5e-6 from the float64 twin: CLOSER to exact arithmetic than the
    reference is, and 1.9e-5 from the float oracle, which is the oracle's own noise.  The north star's "<= 1e-5" is therefore
    asserted where it can hold -- against the float64 twin -- and the distance to the float oracle is held to the rule of
    DESIGN.md 4 (within twice the oracle's own noise).  Active sets must be identical.
  * C4 shape (N = 480, 1000 updates, the bench's states): the oracle's noise is 1e-5 .. 1e-4 per state (mid-transient,
    rank-deficient dual); the int8 tensor-core loop is closer to the float64 twin than the oracle in the median.
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from conftest import active_set, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _oracles(oracle32, oracle64, Qd, Fd, K):
    """float oracle and float64 twin side by side (ctypes releases the GIL)"""
    with ThreadPoolExecutor(2) as ex:
        a, b = ex.submit(oracle32.solve_fixed, Qd, Fd, K), ex.submit(oracle64.solve_fixed, Qd, Fd, K)
        return a.result()[0], b.result()[0]


def _strict_dual(pqp, d, prob):
    """Qd, Fd in the reference's order (bit-identical to PQP_CPU.c wherever the oracle was run against it: every golden case)"""
    with pqp.Solver(d, prob, order=pqp.ORDER_STRICT) as s:
        s.solve(iters=1, status=False)
        Qd, _, _ = s.dual(want_gq=False)
        Fd, _ = s.linear_terms(1)
    return Qd, Fd[0]


def _check_single(pqp, oracle32, oracle64, seed, M, N, K, kernels, need_f64_tol):
    prob, d = pqp.generate_testproblem(seed, M, N)
    Qd, Fd = _strict_dual(pqp, d, prob)
    y32, y64 = _oracles(oracle32, oracle64, Qd, Fd, K)
    e_fd = relerr(y32, y64)
    out = {}
    for sym, kernel in kernels:
        with pqp.Solver(d, prob, exploit_symmetry=sym) as s:      # end to end: the handle's own tensor-core dual
            Y, U, st = s.solve(iters=K, primal=True)
            assert s.last_kernel == kernel, s.last_kernel
        e_gf, e_gd = relerr(Y[0], y32), relerr(Y[0], y64)
        print(f"N={N} M={M} K={K} {kernel}: err(gpu,f32)={e_gf:.2e} err(gpu,f64)={e_gd:.2e} err(f32,f64)={e_fd:.2e}")
        assert e_gf <= TOL or (e_gf <= 2 * e_fd and e_gd <= max(e_fd, TOL)), (kernel, e_gf, e_gd, e_fd)
        if need_f64_tol:
            assert e_gd <= TOL, (kernel, e_gd)       # the north star's 1e-5, against exact arithmetic
        assert e_gd <= e_fd, (kernel, e_gd, e_fd)     # and never further from it than the reference's own float arithmetic
        assert np.array_equal(active_set(Y[0]), active_set(y32)), kernel
        assert st["iters"][0] == K and np.all(np.isfinite(U))
        out[kernel] = (e_gf, e_gd, e_fd)
    return out


def test_n8192_well_conditioned_against_the_oracle(pqp, oracle32, oracle64):
    """N = 8192 with M = N (the generator's Qd = Gp diag(q) Gp' is then positive definite), 100 updates, both single-problem
    loops end to end (tensor-core setup + loop) against PQP_CPU's arithmetic run on the reference-order dual."""
    _check_single(pqp, oracle32, oracle64, 4242, 8192, 8192, 100, ((1, "gemv_sym_stream"), (0, "gemv_tma_stream")), need_f64_tol=True)


def test_c3_bench_instance_against_the_oracle(pqp, oracle32, oracle64):
    """The instance bench.py times (N = 8192, M = 2048, seed 12346; rank-deficient, N = 4M as in PQP_CPU.c:940-941), 60 updates,
    against the oracle itself (round 1 compared it with the STRICT kernel and a bare 5e-5)."""
    _check_single(pqp, oracle32, oracle64, 12346, 2048, 8192, 60, ((1, "gemv_sym_stream"), (0, "gemv_tma_stream")), need_f64_tol=True)


def test_c2_bench_instance_against_the_oracle(pqp, oracle32, oracle64):
    """Config C2 as benched: N = 1024, M = 512 (the report's 2:1 shape; rank-deficient), 1000 updates."""
    _check_single(pqp, oracle32, oracle64, 12345, 512, 1024, 1000, ((1, "gemv_small_registers"),), need_f64_tol=False)


def test_c4_bench_states_against_the_oracle(pqp, oracle32, oracle64):
    """Config C4 as benched (N = 480, 1000 updates, bench_problems.BENCH_X_SCALE states): a seeded sample of 64 of 256 states,
    each against the oracle and its float64 twin.  Per state: within 3x the oracle's own distance to float64 (floor 1e-5),
    identical active set, U bit-identical to computeUfromY on the GPU's y; over the sample: the tensor-core loop is no
    further from float64 than the oracle in the median; no state is degenerate (NaN or all zero)."""
    from bench_problems import BENCH_MIN_VIOLATED, BENCH_X_SCALE, condensed_mpc
    B, K = 256, 1000
    prob, d, X = condensed_mpc(2024, 30, 12, 4, n_states=B, x_scale=BENCH_X_SCALE, min_violated=BENCH_MIN_VIOLATED)
    with pqp.Solver(d, prob, batch_capacity=B) as s:
        Y, U, st = s.solve(X, iters=K, primal=True)
        assert s.last_kernel.startswith("batched_imma"), s.last_kernel
        Qd, _, _ = s.dual()
        Fd, Fp = s.linear_terms(B)
    assert np.isfinite(Y).all() and not (Y == 0).all(axis=1).any()
    idx = np.random.default_rng(64).choice(B, 64, replace=False)
    with ThreadPoolExecutor(16) as ex:
        r32 = list(ex.map(lambda b: oracle32.solve_fixed(Qd, Fd[b], K)[0], idx))
        r64 = list(ex.map(lambda b: oracle64.solve_fixed(Qd, Fd[b], K)[0], idx))
    e = []
    for b, y32, y64 in zip(idx, r32, r64):
        e_gf, e_gd, e_fd = relerr(Y[b], y32), relerr(Y[b], y64), relerr(y32, y64)
        assert e_gf <= TOL or e_gd <= 3 * max(e_fd, TOL), (int(b), e_gf, e_gd, e_fd)
        assert np.array_equal(active_set(Y[b], 1e-5), active_set(y32, 1e-5)), int(b)
        assert np.array_equal(U[b], oracle32.recover_u(Y[b], Fp[b], prob["Gp"], prob["Qp_inv"])), int(b)
        e.append((e_gf, e_gd, e_fd))
    e = np.array(e)
    act = np.array([active_set(y, 1e-5).sum() for y in r32])
    print("C4 sample of 64: worst err(gpu,f32) %.2e err(gpu,f64) %.2e err(f32,f64) %.2e; medians %.2e %.2e %.2e; active constraints median %d of %d"
          % (*e.max(0), *np.median(e, 0), int(np.median(act)), d.N))
    assert np.median(e[:, 1]) <= 1.5 * np.median(e[:, 2])


def test_c4_bench_batch_has_no_degenerate_state(pqp):
    """All 4096 states of the benched batch after 1000 updates: finite, none trivially zero (round 1's x_scale = 60 batch had
    13 NaN and 142 all-zero states, 3.8 % of the timed work), a non-trivial share of the 480 constraints active."""
    from bench_problems import BENCH_MIN_VIOLATED, BENCH_X_SCALE, condensed_mpc
    prob, d, X = condensed_mpc(2024, 30, 12, 4, n_states=4096, x_scale=BENCH_X_SCALE, min_violated=BENCH_MIN_VIOLATED)
    with pqp.Solver(d, prob, batch_capacity=4096) as s:
        Y, U, _ = s.solve(X, iters=1000, primal=True, status=False)
    assert np.isfinite(Y).all() and np.isfinite(U).all() and np.all(Y >= 0)
    assert not (Y == 0).all(axis=1).any()
    act = (Y > 1e-5 * Y.max(axis=1, keepdims=True)).sum(axis=1)
    print("C4 bench batch: active constraints per state: min %d median %d max %d of %d" % (act.min(), np.median(act), act.max(), d.N))
    assert act.min() >= 1 and np.median(act) >= 0.04 * d.N


def test_reference_order_run_to_tolerance_uses_the_per_row_feasibility_tolerance(pqp):
    """PQP_ORDER_STRICT with iters <= 0: the feasibility part of the stop test is compare()'s per-row max(erc*Kp_i, eac)
    (PQP_CPU.c:338), as in the fused kernels.  Instance: a generator problem with every Kp_i >= 50 and erc = 0.02, eac = 1e-6,
    so each row tolerates a violation of >= 1 while eac alone tolerates none: the reference-order loop stops where the fast
    loop stops (same check, or the neighbouring one), at a point that round 1's `min_slack >= -eac` would have rejected --
    it then ran to max_iters, because the sequentially rounded sums of PQP_CPU.c stall at a KKT residual of ~0.1-2 on these
    instances (measured: gpurun_out/r2 probes), far above 1e-6."""
    prob, d = pqp.generate_testproblem(5, 400, 150)
    prob = dict(prob, Kp=(50.0 + 0.5 * prob["Kp"]).astype(np.float32))
    opts = dict(erc=0.02, eac=1e-6, eaj=1e30, erj=1e-5, check_every=8, max_iters=20000)
    with pqp.Solver(d, prob, **opts) as s:
        Yf, _, stf = s.solve(iters=0)
        assert stf["converged"][0] == 1 and stf["iters"][0] < 20000
    with pqp.Solver(d, prob, order=pqp.ORDER_STRICT, **opts) as s:
        Ys, _, sts = s.solve(iters=0)
        assert s.last_kernel == "gemv_strict" and sts["converged"][0] == 1
        print("stop counts: fast", int(stf["iters"][0]), "reference order", int(sts["iters"][0]), "min_slack", float(sts["min_slack"][0]))
        assert sts["min_slack"][0] < -1e-6          # feasible only within the per-row tolerance
        assert abs(int(sts["iters"][0]) - int(stf["iters"][0])) <= 8, (sts["iters"][0], stf["iters"][0])
        Yk, _, _ = s.solve(iters=int(sts["iters"][0]))
        assert np.array_equal(Ys, Yk)


def test_stateless_reference_helpers_run_on_the_device_bit_for_bit(pqp, oracle32):
    """pqp_compute_fp / pqp_compute_cost / pqp_compute_md / pqp_compute_u_from_y (what libpqp_compat.so's computeFp, computeCost,
    convertToDual and computeUfromY call; nothing of them is computed on the host any more) against the oracle's restatement of
    PQP_CPU.c:373-382, :648-666, :472-479, :352-360: bit-identical."""
    import ctypes as C
    from conftest import EXAMPLE_DIR
    L = pqp.lib()
    prob, d = pqp.load_example(EXAMPLE_DIR)
    p = pqp._as_ptr
    Fp = np.zeros(d.M, np.float32)
    assert L.pqp_compute_fp(p(Fp), p(prob["Fp1"]), p(prob["Fp2"]), p(prob["Fp3"]), p(prob["D"]), p(prob["x"]), d.M, d.nDisH, d.nState, -1) == 0
    assert np.array_equal(Fp, oracle32.compute_fp(prob["Fp1"], prob["Fp2"], prob["Fp3"], prob["D"], prob["x"]))
    rng = np.random.default_rng(3)
    for n in (1, 7, 28, 300):
        A = rng.standard_normal((n, n)).astype(np.float32)
        Q = (A @ A.T).astype(np.float32)
        z, F = rng.uniform(-30, 900, n).astype(np.float32), rng.uniform(-50, 50, n).astype(np.float32)
        m = np.array([rng.uniform(-1e5, 1e5)], np.float32)
        J = np.zeros(1, np.float32)
        assert L.pqp_compute_cost(p(J), p(z), p(Q), p(F), p(m), n, -1) == 0
        assert J[0] == np.float32(oracle32.cost(z, Q, F, float(m[0]))), n
        Md = np.zeros(1, np.float32)
        assert L.pqp_compute_md(p(Md), p(z), p(Q), p(m), n, -1) == 0
        G = rng.integers(-1, 2, (2 * n, n)).astype(np.float32)
        _, _, md_ref, _ = oracle32.convert_to_dual(Q, G, np.zeros(2 * n, np.float32), z, float(m[0]), want_qd=False)
        assert Md[0] == np.float32(md_ref), n
        y = rng.uniform(0, 1000, (3, 2 * n)).astype(np.float32)
        fp3 = rng.uniform(-50, 50, (3, n)).astype(np.float32)
        U = np.zeros((3, n), np.float32)
        assert L.pqp_compute_u_from_y(p(U), p(y), p(fp3), p(G), p(Q), 2 * n, n, 3, -1) == 0
        for b in range(3):
            assert np.array_equal(U[b], oracle32.recover_u(y[b], fp3[b], G, Q)), (n, b)
