"""CPU tests: the oracle restatement (oracle/pqp_oracle.c) pinned against
  (a) the reference's own known answers for example/ (U* in PQP_GPU_optimized_coarsened.cu:1209-1215, Jp = -Jd,
      313 printed iterations),
  (b) the golden vectors generated from the unmodified reference (tests/golden/make_golden.py),
  (c) the unmodified reference itself (oracle/_ref) when that library is present."""
import numpy as np
import pytest

from conftest import EXAMPLE_DIR, RANDOM_CASES, golden_problem, relerr
from oracle.oracle import Oracle, Reference

# hand-pasted by the reference's authors, PQP_GPU_optimized_coarsened.cu:1209-1215 (same block PQP_Fusion.cu:1372-1378)
REF_U_STAR = np.array([-6.398985, -10.646729, -4.792132, -7.027614, -4.792255, -10.643004, -6.398996])
# SURVEY 8(c) intermediates dumped from PQP_CPU.c
REF_FP = np.array([6.41107655, 10.709012, 4.79723597, 7.05250216, 4.79735994, 10.705266, 6.41108751])
REF_FD_HEAD = np.array([26.399, 30.6467, 24.7921, 27.0276, 24.7923, 30.643, 26.399, 13.601, 9.35327, 15.2079, 12.9724,
                        15.2077, 9.357, 13.601])


def _example_chain(o):
    p = o.load_example(EXAMPLE_DIR)
    Fp = o.compute_fp(p["Fp1"], p["Fp2"], p["Fp3"], p["D"], p["x"])
    Mp = o.compute_mp(*[p[k] for k in ("Mp1", "Mp2", "Mp3", "Mp4", "Mp5", "Mp6", "D", "x")])
    Qd, Fd, Md, GQ = o.convert_to_dual(p["Qp_inv"], p["Gp"], p["Kp"], Fp, Mp)
    return p, Fp, Mp, Qd, Fd, Md


def test_example_known_answers(oracle32):
    o = oracle32
    p, Fp, Mp, Qd, Fd, Md = _example_chain(o)
    np.testing.assert_allclose(Fp, REF_FP, rtol=1e-6)
    assert Mp == 312097.0
    np.testing.assert_allclose(Md, -311691.469, rtol=1e-7)
    np.testing.assert_allclose(Fd[:14], REF_FD_HEAD, rtol=2e-5)
    assert np.all(Fd[14:] == 0)
    assert np.all(Qd[14:] == 0) and np.all(Qd[:, 14:] == 0)
    assert np.all(o.theta(Qd) == 5.0)
    Qp = o.gauss_jordan(p["Qp_inv"])
    Y, U, h = o.solve_converge(Qd, Fd, Md, Qp, p["Qp_inv"], Fp, Mp, p["Gp"], p["Kp"])
    assert h == 313  # "Printing number of iterations = 313" -> 312 updates
    U = o.recover_u(Y, Fp, p["Gp"], p["Qp_inv"])
    np.testing.assert_allclose(U, REF_U_STAR, atol=1.5e-6)
    Jp, Jd = o.cost(U, Qp, Fp, Mp), o.cost(Y, Qd, Fd, Md)
    assert Jp == 155845.734375 and Jd == -155845.734375  # strong duality as the reference prints it
    assert np.all(Y[14:] == 1000.0)  # zero rows of Qd: those duals never move


def test_example_matches_golden_bit_for_bit(oracle32, gold_example):
    o, g = oracle32, gold_example
    p, Fp, Mp, Qd, Fd, Md = _example_chain(o)
    assert np.array_equal(Fp, g["Fp"]) and Mp == float(g["Mp"]) and Md == float(g["Md"])
    assert np.array_equal(Qd, g["Qd"]) and np.array_equal(Fd, g["Fd"])
    assert np.array_equal(o.theta(Qd), g["theta"])
    assert np.array_equal(o.gauss_jordan(p["Qp_inv"]), g["Qp"])
    for K in (1, 2, 10, 100, 312):
        Y, _ = o.solve_fixed(Qd, Fd, K)
        assert np.array_equal(Y, g[f"Y_K{K}"]), K
    Y, U, h = o.solve_converge(Qd, Fd, Md, g["Qp"], p["Qp_inv"], Fp, Mp, p["Gp"], p["Kp"])
    assert h == int(g["h"]) and np.array_equal(Y, g["Y_conv"])
    assert np.array_equal(o.recover_u(Y, Fp, p["Gp"], p["Qp_inv"]), g["U_conv"])
    # the reference's printed output carries the same numbers
    out = str(g["stdout"])
    assert "iterations = 313" in out and "Jp = 155845.734375" in out and "Jd = -155845.734375" in out
    printed = np.array([float(t) for t in out.split("Printing U*")[1].split()])
    np.testing.assert_allclose(printed, REF_U_STAR, atol=1.5e-6)


def test_float64_twin_matches_golden(oracle64, gold_example):
    g = gold_example
    for K in (100, 312):
        Y, _ = oracle64.solve_fixed(g["Qd"], g["Fd"], K)
        assert np.array_equal(Y, g[f"Y64_K{K}"])


@pytest.mark.parametrize("seed,M,N,K", RANDOM_CASES)
def test_random_matches_golden_bit_for_bit(oracle32, oracle64, gold_random, seed, M, N, K):
    g, t = gold_random, f"s{seed}"
    prob = golden_problem(g, seed)
    Qd, Fd, Md, GQ = oracle32.convert_to_dual(prob["Qp_inv"], prob["Gp"], prob["Kp"], prob["Fp"], prob["Mp0"])
    assert np.array_equal(Qd, g[f"{t}_Qd"]) and np.array_equal(Fd, g[f"{t}_Fd"]) and Md == float(g[f"{t}_Md"])
    Y, th = oracle32.solve_fixed(Qd, Fd, K)
    assert np.array_equal(th, g[f"{t}_theta"]) and np.array_equal(Y, g[f"{t}_Y"])
    assert np.array_equal(oracle32.recover_u(Y, prob["Fp"], prob["Gp"], prob["Qp_inv"]), g[f"{t}_U"])
    Y64, _ = oracle64.solve_fixed(Qd, Fd, K)
    assert np.array_equal(Y64, g[f"{t}_Y64"])


@pytest.mark.parametrize("name", ["test1", "test3"])
def test_reference_testfiles_match_golden_bit_for_bit(oracle32, oracle64, gold_testfiles, name):
    """testing/sample test/test1.txt (M=500, N=1500) and test3.txt (M=800, N=1200) of the reference: the restatement against
    what the unmodified reference makes of them (100 updates)."""
    from conftest import problem_from_testfile
    g = gold_testfiles
    prob = problem_from_testfile(g, name)
    Qd, Fd, Md, GQ = oracle32.convert_to_dual(prob["Qp_inv"], prob["Gp"], prob["Kp"], prob["Fp"], prob["Mp0"])
    assert np.array_equal(Fd, g[f"{name}_Fd"]) and np.array_equal(Qd.sum(1), g[f"{name}_Qd_rowsum"])
    assert np.array_equal(np.diag(Qd), g[f"{name}_Qd_diag"])
    K = int(g[f"{name}_K"])
    Y, th = oracle32.solve_fixed(Qd, Fd, K)
    assert np.array_equal(th, g[f"{name}_theta"]) and np.array_equal(Y, g[f"{name}_Y"])
    assert np.array_equal(oracle32.recover_u(Y, prob["Fp"], prob["Gp"], prob["Qp_inv"]), g[f"{name}_U"])
    Y64, _ = oracle64.solve_fixed(Qd, Fd, K)
    assert np.array_equal(Y64, g[f"{name}_Y64"])


def test_split_matrices_and_single_update(oracle32, gold_random):
    g, t = gold_random, "s101"
    Qd, Fd = g[f"{t}_Qd"], g[f"{t}_Fd"]
    th = oracle32.theta(Qd)
    P, Nn = oracle32.split(Qd, th)
    assert np.array_equal(P - np.diag(th), np.maximum(Qd, 0)) or np.allclose(P - np.diag(th), np.maximum(Qd, 0))
    assert np.all(P >= 0) and np.all(Nn >= 0)
    off = ~np.eye(Qd.shape[0], dtype=bool)
    assert np.all((P * Nn)[off] == 0)  # disjoint supports off the diagonal
    Y0 = np.full(Qd.shape[0], 1000.0, np.float32)
    Y1 = oracle32.update_y2(Y0, P, Nn, Fd)
    Yk, _ = oracle32.solve_fixed(Qd, Fd, 1)
    assert np.array_equal(Y1, Yk)
    # iterate() continues a solve exactly
    Y5, _ = oracle32.solve_fixed(Qd, Fd, 5)
    assert np.array_equal(oracle32.iterate(Y1, Qd, Fd, 4), Y5)


def test_matmul_transpose_variants(oracle32):
    rng = np.random.default_rng(0)
    A, B = rng.standard_normal((5, 7)).astype(np.float32), rng.standard_normal((7, 3)).astype(np.float32)
    want = A.astype(np.float64) @ B.astype(np.float64)
    for tA in (0, 1):
        for tB in (0, 1):
            got = oracle32.matmul(A.T.copy() if tA else A, tA, B.T.copy() if tB else B, tB, 5, 7, 3)
            np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)


@pytest.mark.skipif(not Reference.available(), reason="oracle/_ref not built here (no /root/reference)")
def test_restatement_equals_unmodified_reference():
    """Every function of the restatement, bit for bit, against PQP_CPU.c compiled where it lies."""
    o, r = Oracle(np.float32), Reference(np.float32)
    rng = np.random.default_rng(7)
    for (M, N) in [(16, 40), (50, 30), (33, 65)]:
        Qi = np.diag(rng.uniform(0, 100, M)).astype(np.float32)
        Qi += (rng.uniform(-1, 1, (M, M)) * 0.01).astype(np.float32)  # dense Qp_inv too
        Fp = rng.uniform(0, 100, M).astype(np.float32)
        Kp = rng.uniform(0, 100, N).astype(np.float32)
        Gp = np.where((g := rng.integers(0, 3, (N, M))) == 2, -1, g).astype(np.float32)
        a, b = r.convert_to_dual(Qi, Gp, Kp, Fp, 3.0), o.convert_to_dual(Qi, Gp, Kp, Fp, 3.0)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]
        P, Nn, th = r.split(a[0])
        th2 = o.theta(b[0])
        P2, N2 = o.split(b[0], th2)
        assert np.array_equal(th, th2) and np.array_equal(P, P2) and np.array_equal(Nn, N2)
        y = rng.uniform(0, 10, N).astype(np.float32)
        assert np.array_equal(r.update_y2(y, P, Nn, a[1]), o.update_y2(y, P2, N2, b[1]))
        assert np.array_equal(r.solve_fixed(a[0], a[1], 40), o.solve_fixed(b[0], b[1], 40)[0])
        Y = r.solve_fixed(a[0], a[1], 40)
        assert np.array_equal(r.recover_u(Y, Fp, Gp, Qi), o.recover_u(Y, Fp, Gp, Qi))
        assert np.array_equal(r.gauss_jordan(Qi), o.gauss_jordan(Qi))
        assert r.cost(Y, a[0], a[1], a[2]) == o.cost(Y, b[0], b[1], b[2])
    # float64 twin
    o64, r64 = Oracle(np.float64), Reference(np.float64)
    assert np.array_equal(r64.solve_fixed(a[0], a[1], 40), o64.solve_fixed(a[0], a[1], 40)[0])


@pytest.mark.skipif(not Reference.available(), reason="oracle/_ref not built here (no /root/reference)")
def test_reference_program_output_on_golden_copy_of_example():
    """The unmodified program run on tests/golden/example prints the known answers (the copy is faithful)."""
    import os
    r = Reference(np.float32)
    out = r.main_stdout(os.path.dirname(EXAMPLE_DIR))
    assert "iterations = 313" in out and "Jp = 155845.734375" in out
