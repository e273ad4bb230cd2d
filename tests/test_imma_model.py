"""CPU (-m "not gpu"): the arithmetic scheme of the int8 tensor-core batched kernel, restated in numpy
(tests/imma_model.py), against the oracle and its float64 twin.  This is the accuracy argument for running the
batched PQP loop on integer tensor cores; the GPU tests then only need to show the kernel is bit-identical to the model.

Tolerance: the rule of DESIGN.md 4 (normwise 1e-5, or within the oracle's own fp32 noise on the instance).
"""
import numpy as np
import pytest

from conftest import active_set, relerr
from imma_model import run, slice_rows_u8, slice_y_s8

TOL = 1e-5


def _mpc_duals(oracle32, seed, pH, nS, nI, B):
    from bench_problems import condensed_mpc
    prob, d, X = condensed_mpc(seed, pH, nS, nI, n_states=B)
    Qd, Fds = None, []
    for b in range(B):
        Fp = oracle32.compute_fp(prob["Fp1"], prob["Fp2"], prob["Fp3"], prob["D"], X[b])
        Qd, Fd, _, _ = oracle32.convert_to_dual(prob["Qp_inv"], prob["Gp"], prob["Kp"], Fp, 0.0)
        Fds.append(Fd)
    return Qd, oracle32.theta(Qd), np.stack(Fds)


def test_digit_planes_reconstruct_their_operands():
    rng = np.random.default_rng(1)
    A = rng.uniform(0, 3, (40, 40)).astype(np.float32)
    A[rng.uniform(size=A.shape) < 0.5] = 0
    np.fill_diagonal(A, 0)
    A0, A1, A2, rs = slice_rows_u8(A)
    for P in (A0, A1, A2):
        assert P.min() >= 0 and P.max() <= 255 and np.array_equal(P, np.rint(P))
    back = (A0 * 65536 + A1 * 256 + A2) * (rs[:, None].astype(np.float64) / 65536.0)
    assert np.abs(back - A).max() <= 2.0 ** -24 * A.max(axis=1).max() * 1.0001  # 24-bit fixed point per row
    Y = rng.uniform(0, 1000, (40, 7)).astype(np.float32)
    Y[3] = 1e-30
    Y0, Y1, Y2, isc = slice_y_s8(Y)
    for P in (Y0, Y1, Y2):
        assert P.min() >= -128 and P.max() <= 127 and np.array_equal(P, np.rint(P))
    back = (Y0 * 65536 + Y1 * 256 + Y2) * isc[None, :].astype(np.float64)
    assert np.abs(back - Y).max() <= 2.0 ** -22 * Y.max()  # 22-bit fixed point per problem, round to nearest


@pytest.mark.parametrize("pH,nS,nI,K", [(6, 5, 2, 60), (5, 4, 3, 200), (9, 4, 3, 120)])
def test_model_matches_oracle_small(oracle32, oracle64, pH, nS, nI, K):
    B = 6
    Qd, th, Fd = _mpc_duals(oracle32, 3, pH, nS, nI, B)
    Y = run(Qd, th, Fd, K)
    for b in range(B):
        y32, _ = oracle32.solve_fixed(Qd, Fd[b], K)
        y64, _ = oracle64.solve_fixed(Qd, Fd[b], K)
        e_gf, e_gd, e_fd = relerr(Y[b], y32), relerr(Y[b], y64), relerr(y32, y64)
        assert e_gf <= TOL or (e_gf <= 2 * e_fd and e_gd <= max(e_fd, TOL)), (b, e_gf, e_gd, e_fd)
        assert np.array_equal(active_set(Y[b], 1e-5), active_set(y32, 1e-5))


def test_model_c4_shape_is_at_the_oracles_noise_floor(oracle32, oracle64):
    """Config C4 shape (N=480), 3 states, 150 updates: the int8 scheme is no further from the float64 twin than the fp32
    oracle itself (measured over 24 states and 1000 updates in tools/ozaki_emulate.py: ratio 0.3 - 2.2, median < 1)."""
    Qd, th, Fd = _mpc_duals(oracle32, 2024, 30, 12, 4, 3)
    K = 150
    Y = run(Qd, th, Fd, K)
    for b in range(3):
        y32, _ = oracle32.solve_fixed(Qd, Fd[b], K)
        y64, _ = oracle64.solve_fixed(Qd, Fd[b], K)
        assert relerr(Y[b], y64) <= max(2 * relerr(y32, y64), TOL)
        assert np.array_equal(active_set(Y[b], 1e-5), active_set(y64, 1e-5))


def test_model_degenerate_problems_follow_the_reference():
    """No active constraint: y underflows to exactly 0 and stays finite.  A dual that hits 0 while its constraint is violated
    gives 0/0 in the reference (PQP_CPU.c:594) and NaN in every component one update later; the model does the same."""
    Qd = np.array([[2, -1], [-1, 2]], np.float32)
    th = np.array([5, 5], np.float32)
    Y = run(Qd, th, np.array([[1.0, 2.0]], np.float32), 3000)
    assert np.all(Y == 0)
    Y = run(Qd, th, np.array([[-1.0, 2.0]], np.float32), 5, Y0=np.zeros((1, 2), np.float32))
    assert np.isnan(Y).all()


def test_oracle_duals_of_the_mpc_layout_have_the_pair_structure(oracle32):
    """Gp = [I; -I; C Gam; -C Gam] (the reference's N = 4*pHorizon*nInput layout; example/Gp.txt is [I; -I; 0; 0]): PQP_CPU.c's own
    float Qd satisfies Qd[sigma(i)][j] == -Qd[i][j] == Qd[i][sigma(j)] element for element -- negating an operand row negates
    every product and every partial sum exactly -- which is what the PAIRED tensor-core scheme relies on."""
    from imma_model import has_pair_structure
    Qd, th, Fd = _mpc_duals(oracle32, 3, 6, 5, 2, 1)
    assert has_pair_structure(Qd)
    g = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "golden_example.npz"))
    assert has_pair_structure(g["Qd"])
    rng = np.random.default_rng(0)
    A = rng.standard_normal((8, 8)).astype(np.float32)
    assert not has_pair_structure((A @ A.T).astype(np.float32))


@pytest.mark.parametrize("pH,nS,nI,K", [(6, 5, 2, 60), (9, 4, 3, 120)])
def test_paired_model_matches_oracle_and_the_plain_model(oracle32, oracle64, pH, nS, nI, K):
    """The PAIRED arithmetic (half the products: only the representative rows are multiplied) against the oracle and its float64
    twin with the tolerance rule of DESIGN.md 4, and next to the plain scheme: the two differ only in fp32 roundings of the
    a_ii terms and in per-row scales, i.e. at the 1e-6 level."""
    from imma_model import run_paired
    B = 5
    Qd, th, Fd = _mpc_duals(oracle32, 3, pH, nS, nI, B)
    Yp = run_paired(Qd, th, Fd, K)
    Y = run(Qd, th, Fd, K)
    for b in range(B):
        y32, _ = oracle32.solve_fixed(Qd, Fd[b], K)
        y64, _ = oracle64.solve_fixed(Qd, Fd[b], K)
        e_gf, e_gd, e_fd = relerr(Yp[b], y32), relerr(Yp[b], y64), relerr(y32, y64)
        assert e_gf <= TOL or (e_gf <= 2 * e_fd and e_gd <= max(e_fd, TOL)), (b, e_gf, e_gd, e_fd)
        assert np.array_equal(active_set(Yp[b], 1e-5), active_set(y32, 1e-5))
        assert relerr(Yp[b], Y[b]) <= max(TOL, 2 * e_fd)


def test_paired_model_c4_shape(oracle32, oracle64):
    from imma_model import run_paired
    Qd, th, Fd = _mpc_duals(oracle32, 2024, 30, 12, 4, 3)
    K = 150
    Y = run_paired(Qd, th, Fd, K)
    for b in range(3):
        y32, _ = oracle32.solve_fixed(Qd, Fd[b], K)
        y64, _ = oracle64.solve_fixed(Qd, Fd[b], K)
        assert relerr(Y[b], y64) <= max(2 * relerr(y32, y64), TOL)
        assert np.array_equal(active_set(Y[b], 1e-5), active_set(y64, 1e-5))
