"""GPU (-m gpu): Brand's acceleration / line-search step as an opt-in (pqp_opts.accelerate; SURVEY 8(f)4, first half).

The reference carries it as dead code behind `if(1)` (PQP_CPU.c:721-735: computeph :625-630, computealphaY :545-575, updateY1
:579-588).  The oracle restates it with the one evident fix (computeph adds ph to itself where the gradient needs Fd) and the
library runs the same arithmetic in the same order, so the reference-order solve must equal the oracle bit for bit; the fast
solve is held to the tolerance rule of DESIGN.md 4; and the step must actually accelerate.
"""
import numpy as np
import pytest

from conftest import EXAMPLE_DIR, RANDOM_CASES, active_set, golden_problem, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-5


def test_oracle_step_is_a_descent_step_that_keeps_the_duals_non_negative(oracle32, oracle64, gold_random):
    g = gold_random
    Qd, Fd = g["s104_Qd"], g["s104_Fd"]
    y, _ = oracle32.solve_fixed(Qd, Fd, 120)
    yn, alpha = oracle32.accel_step(y, Qd, Fd)
    J = lambda v: 0.5 * v.astype(np.float64) @ Qd.astype(np.float64) @ v + Fd.astype(np.float64) @ v
    assert alpha >= 0 and np.all(yn >= y) and J(yn) <= J(y) + 1e-9 * abs(J(y))


@pytest.mark.parametrize("seed,M,N,K", RANDOM_CASES[:4])
def test_reference_order_solve_with_acceleration_is_bit_identical(pqp, oracle32, gold_random, seed, M, N, K):
    g, t = gold_random, f"s{seed}"
    prob = golden_problem(g, seed)
    for every in (1, 7, K // 2):
        want = oracle32.solve_accel(g[f"{t}_Qd"], g[f"{t}_Fd"], K, every)
        with pqp.Solver(pqp.dims_plain(M, N), prob, order=pqp.ORDER_STRICT, accelerate=every) as s:
            Y, U, st = s.solve(iters=K, primal=True)
            assert np.array_equal(Y[0], want), (seed, every)
            assert st["iters"][0] == K
        with pqp.Solver(Qd=g[f"{t}_Qd"], order=pqp.ORDER_STRICT, accelerate=every) as s:
            Y, _, _ = s.solve(Fd=g[f"{t}_Fd"], iters=K)
            assert np.array_equal(Y[0], want), (seed, every)


def test_example_with_acceleration(pqp, oracle32, gold_example):
    g = gold_example
    prob, d = pqp.load_example(EXAMPLE_DIR)
    want = oracle32.solve_accel(g["Qd"], g["Fd"], 100, 10)
    with pqp.Solver(d, prob, order=pqp.ORDER_STRICT, accelerate=10) as s:
        Y, _, _ = s.solve(prob["x"][None], iters=100)
        assert np.array_equal(Y[0], want)
    with pqp.Solver(d, prob, accelerate=10) as s:
        Y, _, _ = s.solve(prob["x"][None], iters=100)
        assert relerr(Y[0], want) <= TOL


def test_fast_and_batched_solves_with_acceleration(pqp, oracle32, oracle64):
    from bench_problems import condensed_mpc
    prob, d, X = condensed_mpc(3, 6, 5, 2, n_states=37)
    K, every = 90, 15
    with pqp.Solver(d, prob, accelerate=every) as s:
        Y, U, st = s.solve(X, iters=K, primal=True)
        assert s.last_kernel.startswith("batched")
        Qd, _, _ = s.dual()
        Fd, Fp = s.linear_terms(37)
        assert np.all(st["iters"] == K)
        Y1, _, _ = s.solve(X[5][None], iters=K)            # the single-problem loop, same option
    for b in (0, 5, 36):
        y32, y64 = oracle32.solve_accel(Qd, Fd[b], K, every), oracle64.solve_accel(Qd, Fd[b], K, every)
        e_gf, e_gd, e_fd = relerr(Y[b], y32), relerr(Y[b], y64), relerr(y32, y64)
        assert e_gf <= TOL or (e_gf <= 2 * e_fd and e_gd <= max(2 * e_fd, TOL)), (b, e_gf, e_gd, e_fd)
        assert np.array_equal(active_set(Y[b], 1e-5), active_set(y32, 1e-5))
        assert np.array_equal(U[b], oracle32.recover_u(Y[b], Fp[b], prob["Gp"], prob["Qp_inv"]))
    y32 = oracle32.solve_accel(Qd, Fd[5], K, every)
    assert relerr(Y1[0], y32) <= max(TOL, 2 * relerr(y32, oracle64.solve_accel(Qd, Fd[5], K, every)))


def test_acceleration_reaches_a_smaller_kkt_residual_in_the_same_number_of_updates(pqp):
    """16 small condensed-MPC states (the instance family of the stop-count comparison): KKT residual ||min(y, Qd y + Fd)||inf
    after the same 400 multiplicative updates, with a line-search step every 10 updates and without."""
    from bench_problems import condensed_mpc
    mp, md, X = condensed_mpc(11, pH=6, nS=4, nI=2, n_states=16, x_scale=25.0)
    K = 400
    with pqp.Solver(md, mp, batch_capacity=16) as s:
        _, _, st0 = s.solve(X, iters=K)
    with pqp.Solver(md, mp, batch_capacity=16, accelerate=10) as s:
        _, _, st1 = s.solve(X, iters=K)
    k0, k1 = st0["kkt"].astype(np.float64), st1["kkt"].astype(np.float64)
    ratio = k1 / np.maximum(k0, 1e-30)
    print("KKT residual after %d updates: plain median %.3e, accelerated median %.3e; ratio median %.3f, worst %.3f"
          % (K, np.median(k0), np.median(k1), np.median(ratio), ratio.max()))
    assert np.all(np.isfinite(k1)) and np.median(ratio) <= 1.0
