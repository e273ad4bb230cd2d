"""GPU tests (-m gpu) of the upper-triangle loop (pqp_gemv_sym.cu): a single problem whose fp32 Qd is
symmetric element for element is iterated from the strictly upper triangle alone (2N^2 bytes per update).  Same sums as
updateY2 (PQP_CPU.c:603-618) in another order, so the bar is the FAST one of test_parity_gpu.py: the oracle on the same
inputs, the tolerance rule of DESIGN.md 4, identical active sets; plus reproducibility, warm starts, ragged sizes, and the
guarantee that a Qd which is NOT exactly symmetric never takes this path.
"""
import numpy as np
import pytest

from conftest import active_set, relerr
from test_parity_gpu import TOL, check_fast

pytestmark = pytest.mark.gpu


def spd_dual(rng, N, rank):
    """A symmetric positive semidefinite Qd = A A' with both signs off the diagonal, formed symmetric bit for bit."""
    A = rng.standard_normal((N, rank)).astype(np.float32)
    Q = (A.astype(np.float64) @ A.astype(np.float64).T).astype(np.float32)
    Q = np.triu(Q) + np.triu(Q, 1).T
    return np.ascontiguousarray(Q)


@pytest.mark.parametrize("N,rank,K", [(2560, 3000, 40), (3001, 3500, 30), (2563, 700, 30)])
def test_sym_against_the_oracle(pqp, oracle32, oracle64, N, rank, K):
    """Well-conditioned (rank > N), ragged (N = 3001: not a multiple of 4, 24 tiles of which the last is partial) and
    rank-deficient (rank < N) duals: STRICT is the oracle bit for bit, the upper-triangle loop meets the FAST tolerance."""
    rng = np.random.default_rng(N)
    Qd = spd_dual(rng, N, rank)
    Fd = rng.uniform(-50, 50, N).astype(np.float32)
    y32, th32 = oracle32.solve_fixed(Qd, Fd, K)
    y64, _ = oracle64.solve_fixed(Qd, Fd, K)
    with pqp.Solver(Qd=Qd, order=pqp.ORDER_STRICT) as s:
        Ys, _, _ = s.solve(Fd=Fd, iters=K)
        assert np.array_equal(Ys[0], y32)
    with pqp.Solver(Qd=Qd) as s:
        Y, _, st = s.solve(Fd=Fd, iters=K)
        assert s.last_kernel.startswith("gemv_sym"), s.last_kernel
        e = check_fast(Y[0], y32, y64, f"sym N={N}")
        print(f"sym N={N}: errs (gpu-f32, gpu-f64, f32-f64) = {e}")
        assert np.array_equal(active_set(Y[0]), active_set(y32))
        assert st["iters"][0] == K and np.all(np.isfinite(Y)) and np.all(Y >= 0)
        # reproducible bit for bit
        Y2, _, _ = s.solve(Fd=Fd, iters=K)
        assert np.array_equal(Y, Y2)
    with pqp.Solver(Qd=Qd, exploit_symmetry=0) as s:
        Yf, _, stf = s.solve(Fd=Fd, iters=K)
        assert not s.last_kernel.startswith("gemv_sym"), s.last_kernel
        check_fast(Yf[0], y32, y64, f"full N={N}")
        assert relerr(Y[0], Yf[0]) <= max(TOL, 3 * relerr(y32, y64))
        # the status block (evaluated by each kernel on its own g = Qd y + Fd) agrees between the two loops
        for k in ("gap", "Jd", "kkt"):
            np.testing.assert_allclose(st[k][0], stf[k][0], rtol=2e-3, err_msg=k)
        assert abs(st["min_slack"][0] - stf["min_slack"][0]) <= 2e-3 * max(1.0, abs(stf["min_slack"][0]))


def test_sym_generator_instance_warm_start_and_chunking(pqp):
    """The reference generator's shape (testing/test_generator.c: Qp_inv diagonal, Gp in {0,+-1}, N = 4M): its Qd is symmetric
    as computed.  K1 updates then K2 more from Y0 equals K1+K2 in one go, bit for bit; the primal comes back finite."""
    prob, d = pqp.generate_testproblem(777, 640, 2560)
    with pqp.Solver(d, prob) as s:
        Ya, _, _ = s.solve(iters=70)
        assert s.last_kernel.startswith("gemv_sym"), s.last_kernel
        Yb, Ub, _ = s.solve(iters=50, Y0=Ya, primal=True)
        Yc, Uc, stc = s.solve(iters=120, primal=True)
        assert np.array_equal(Yb, Yc) and np.array_equal(Ub, Uc)
        assert np.all(np.isfinite(Uc)) and stc["iters"][0] == 120
    with pqp.Solver(d, prob, exploit_symmetry=0) as s:
        Yf, _, _ = s.solve(iters=120)
        assert not s.last_kernel.startswith("gemv_sym")
        assert relerr(Yc[0], Yf[0]) <= 5e-5   # rank-deficient shape: the oracle's own fp32 noise is ~2e-5 here
        assert np.array_equal(active_set(Yc[0], 1e-4), active_set(Yf[0], 1e-4))


def test_sym_never_used_for_an_unsymmetric_dual(pqp):
    """One element of the lower triangle off by one ulp: the handle must keep to the full-matrix loop, and give exactly what
    it gives with the option switched off."""
    rng = np.random.default_rng(9)
    N = 2560
    Qd = spd_dual(rng, N, 3000)
    Qd[1700, 3] = np.nextafter(Qd[1700, 3], np.float32(np.inf))
    Fd = rng.uniform(-50, 50, N).astype(np.float32)
    with pqp.Solver(Qd=Qd) as s:
        Y, _, _ = s.solve(Fd=Fd, iters=20)
        k1 = s.last_kernel
    with pqp.Solver(Qd=Qd, exploit_symmetry=0) as s:
        Y0, _, _ = s.solve(Fd=Fd, iters=20)
        k0 = s.last_kernel
    assert not k1.startswith("gemv_sym") and k1 == k0
    assert np.array_equal(Y, Y0)


def test_sym_signed_zero_and_zero_rows(pqp, oracle32):
    """-0.0 against +0.0 across the diagonal still counts as symmetric (their contributions are equal), and rows of Qd that are
    entirely zero with Fd = 0 keep their duals at the initial value (example rows 14-27 behave like that)."""
    rng = np.random.default_rng(10)
    N = 2560
    Qd = spd_dual(rng, N, 3000)
    dead = [5, 1290, 2559]
    for i in dead:
        Qd[i, :] = 0.0
        Qd[:, i] = 0.0
    Qd[7, 900] = -0.0
    Qd[900, 7] = 0.0
    Fd = rng.uniform(-50, 50, N).astype(np.float32)
    Fd[dead] = 0.0
    y32, _ = oracle32.solve_fixed(Qd, Fd, 25)
    with pqp.Solver(Qd=Qd) as s:
        Y, _, _ = s.solve(Fd=Fd, iters=25)
        assert s.last_kernel.startswith("gemv_sym"), s.last_kernel
        assert np.all(Y[0][dead] == np.float32(1000.0))
        assert relerr(Y[0], y32) <= TOL


def test_sym_run_to_tolerance(pqp):
    """iters <= 0 on the upper-triangle loop: the owners form the stop test of terminate() (PQP_CPU.c:673-687, on g = Qd y + Fd)
    every check_every updates; the run stops at a check, bit-identical to the fixed-count solve at the reported count, at the
    same place (to within a few checks: the sums are ordered differently) as the full-matrix tolerance kernel; a cap that is
    too small reports exactly max_iters updates, unconverged."""
    rng = np.random.default_rng(21)
    N = 2560
    Qd = spd_dual(rng, N, 3000) / np.float32(3000.0)
    Qd[np.arange(N), np.arange(N)] += np.float32(0.5)     # well conditioned: converges in a few hundred updates
    Fd = rng.uniform(-50, 50, N).astype(np.float32)
    opts = dict(eaj=1e30, erj=1e-5, eac=1e-3, erc=1e-3, check_every=8, max_iters=20000)
    with pqp.Solver(Qd=Qd, **opts) as s:
        Y, _, st = s.solve(Fd=Fd, iters=0)
        assert s.last_kernel.startswith("gemv_sym") and s.last_kernel.endswith("_tol"), s.last_kernel
        it = int(st["iters"][0])
        print("sym run-to-tolerance:", it, "updates, gap", st["gap"][0], "min_slack", st["min_slack"][0])
        assert st["converged"][0] == 1 and 0 < it < 20000 and it % 8 == 0
        assert st["min_slack"][0] >= -1e-3 and abs(st["gap"][0]) <= 1e-5 * abs(st["Jd"][0])
        Yf, _, stf = s.solve(Fd=Fd, iters=it)
        assert s.last_kernel.startswith("gemv_sym") and not s.last_kernel.endswith("_tol")
        assert np.array_equal(Y, Yf)
        np.testing.assert_allclose(st["gap"][0], stf["gap"][0], rtol=1e-6)
        # the previous check had not passed yet
        if it >= 8:
            _, _, stp = s.solve(Fd=Fd, iters=it - 8)
            assert not (stp["min_slack"][0] >= -1e-3 and abs(stp["gap"][0]) <= 1e-5 * abs(stp["Jd"][0]))
    with pqp.Solver(Qd=Qd, exploit_symmetry=0, **opts) as s:
        Y0, _, st0 = s.solve(Fd=Fd, iters=0)
        assert not s.last_kernel.startswith("gemv_sym")
        # the criterion is crossed slowly at the tail, and the two loops round differently: the counts agree to a few checks
        assert st0["converged"][0] == 1 and abs(int(st0["iters"][0]) - it) <= max(8, 0.02 * it)
        assert relerr(Y[0], Y0[0]) <= 1e-4
    with pqp.Solver(Qd=Qd, **dict(opts, max_iters=40)) as s:
        Yc, _, stc = s.solve(Fd=Fd, iters=0)
        assert stc["converged"][0] == 0 and stc["iters"][0] == 40
        Yd, _, _ = s.solve(Fd=Fd, iters=40)
        assert np.array_equal(Yc, Yd)


def test_sym_other_grids_and_sizes(pqp):
    """The smallest size that leaves the on-chip kernel (N = 2369: 19 blocks, the last one 65 rows wide), a size whose last
    float4 of y is partial, and a grid of 64 CTAs (PQP_GEMV_GRID): the upper-triangle loop against the full-matrix loop, plus
    the same bits from run to run.  Nothing here depends on the SM count except the cut of the unit ranges."""
    import os
    rng = np.random.default_rng(33)
    for N, grid in ((2369, None), (2817, None), (2560, 64), (3330, 100)):
        Qd = spd_dual(rng, N, N + 300)
        Fd = rng.uniform(-50, 50, N).astype(np.float32)
        if grid:
            os.environ["PQP_GEMV_GRID"] = str(grid)
        try:
            with pqp.Solver(Qd=Qd) as s:
                Y, _, st = s.solve(Fd=Fd, iters=25)
                assert s.last_kernel.startswith("gemv_sym"), (N, grid, s.last_kernel)
                Y2, _, _ = s.solve(Fd=Fd, iters=25)
                assert np.array_equal(Y, Y2)
            with pqp.Solver(Qd=Qd, exploit_symmetry=0) as s:
                Yf, _, stf = s.solve(Fd=Fd, iters=25)
        finally:
            os.environ.pop("PQP_GEMV_GRID", None)
        assert relerr(Y[0], Yf[0]) <= TOL, (N, grid, relerr(Y[0], Yf[0]))
        np.testing.assert_allclose(st["Jd"][0], stf["Jd"][0], rtol=1e-4)


def test_sym_several_right_hand_sides_in_one_call(pqp):
    """pqp_solve_dual with B = 3 on a dual too large for the batched tensor-core kernel: three runs of the upper-triangle
    loop on the same unit array, each equal to its own single solve (the per-row constants are rebuilt per launch)."""
    rng = np.random.default_rng(44)
    N = 2560
    Qd = spd_dual(rng, N, 2900)
    Fd = rng.uniform(-50, 50, (3, N)).astype(np.float32)
    with pqp.Solver(Qd=Qd, batch_capacity=3) as s:
        Y, _, st = s.solve(Fd=Fd, iters=20)
        assert s.last_kernel.startswith("gemv_sym"), s.last_kernel
        assert Y.shape == (3, N) and np.all(st["iters"] == 20)
        for b in range(3):
            Yb, _, stb = s.solve(Fd=Fd[b], iters=20)
            assert np.array_equal(Y[b], Yb[0]), b
            np.testing.assert_allclose(st["Jd"][b], stb["Jd"][0], rtol=1e-6)


def test_sym_random_sizes_against_the_full_matrix_loop(pqp):
    """Indexing stress: random generator instances (N in 2369..9000, ragged M, random update counts) through the
    upper-triangle loop and through the full-matrix loop: same y to rounding, same dual cost, same bits from run to run
    (tools/sym_stress.py runs a longer list)."""
    rng = np.random.default_rng(2)
    for t in range(8):
        N = int(rng.integers(2369, 9000))
        M = int(rng.integers(max(8, N // 6), N // 2))
        K = int(rng.integers(3, 40))
        prob, d = pqp.generate_testproblem(500 + t, M, N)
        with pqp.Solver(d, prob) as s:
            Y, _, st = s.solve(iters=K)
            assert s.last_kernel.startswith("gemv_sym"), (N, M, s.last_kernel)
            Y2, _, _ = s.solve(iters=K)
            assert np.array_equal(Y, Y2), (N, M, K)
        with pqp.Solver(d, prob, exploit_symmetry=0) as s:
            Yf, _, stf = s.solve(iters=K)
        assert np.all(np.isfinite(Y)) and relerr(Y[0], Yf[0]) <= 2e-5, (N, M, K, relerr(Y[0], Yf[0]))
        np.testing.assert_allclose(st["Jd"][0], stf["Jd"][0], rtol=1e-5)


def test_sym_units_in_tensor_memory_change_nothing(pqp):
    """Up to eight units per SM are parked in tensor memory (PQP_SYM_TMEM; by default below N ~ 6800) instead of being streamed: where a unit is
    kept does not enter the arithmetic, so 0, 3 and 8 parked units must give the same bits -- fixed count and run to tolerance,
    at a size that is fully on chip with them (N = 3001), one that is half streamed (6100) and the C3 size."""
    import os
    for N, M, K in ((3001, 700, 30), (6100, 1500, 12), (8192, 2048, 6)):
        prob, d = pqp.generate_testproblem(77, M, N)
        got = {}
        for tm in ("0", "3", "8"):
            os.environ["PQP_SYM_TMEM"] = tm
            try:
                with pqp.Solver(d, prob, check_every=4, max_iters=K + 3) as s:  # the stop test is evaluated, never met: ends at the cap
                    Y, _, st = s.solve(iters=K)
                    assert s.last_kernel.startswith("gemv_sym"), (N, tm, s.last_kernel)
                    Yt, _, stt = s.solve(iters=0)
                    got[tm] = (Y.copy(), st["Jd"][0], Yt.copy(), int(stt["iters"][0]))
            finally:
                os.environ.pop("PQP_SYM_TMEM", None)
        for tm in ("3", "8"):
            assert np.array_equal(got["0"][0], got[tm][0]), (N, tm)
            assert got["0"][1] == got[tm][1], (N, tm)
            assert np.array_equal(got["0"][2], got[tm][2]) and got["0"][3] == got[tm][3], (N, tm)


def test_setup_gives_a_dense_symmetric_hessian_one_value_per_pair(pqp, oracle32, oracle64):
    """pqp_setup, FAST order: Qd = Gp Qp_inv Gp' with a DENSE symmetric Qp_inv (the generator's is diagonal) comes out of the GEMM
    with its two copies of an element differing in the last bits; the setup replaces each pair by its mean when all pairs agree
    to rounding, so this problem class reaches the upper-triangle loop too.  The result stays within the FAST tolerance of the
    oracle on the same inputs.  exploit_symmetry = 0 keeps the matrix as computed; an unsymmetric Qp_inv is never touched."""
    rng = np.random.default_rng(91)
    M, N, K = 300, 2400, 30
    A = rng.standard_normal((M, M)).astype(np.float32)
    Qi = (A.astype(np.float64) @ A.astype(np.float64).T / M + np.eye(M)).astype(np.float32)
    Qi = np.ascontiguousarray(np.triu(Qi) + np.triu(Qi, 1).T)
    prob = dict(Qp_inv=Qi, Gp=rng.integers(-1, 2, (N, M)).astype(np.float32), Kp=rng.uniform(0, 100, N).astype(np.float32),
                Fp=rng.uniform(0, 100, M).astype(np.float32), Mp0=0.0)
    d = pqp.dims_plain(M, N)
    with pqp.Solver(d, prob, exploit_symmetry=0) as s:
        Q0, _, _ = s.dual()
        Y0, _, _ = s.solve(iters=K)
        assert not s.last_kernel.startswith("gemv_sym")
    with pqp.Solver(d, prob) as s:
        Q1, _, _ = s.dual()
        Y1, U1, st = s.solve(iters=K, primal=True)
        assert s.last_kernel.startswith("gemv_sym"), s.last_kernel
    assert np.array_equal(Q1, Q1.T)
    assert not np.array_equal(Q0, Q0.T), "the GEMM output was symmetric already: the test instance does not exercise the mean"
    assert relerr(Q1, Q0) <= 1e-6 and np.array_equal(np.diag(Q1), np.diag(Q0))
    Qd, Fd, Md, _ = oracle32.convert_to_dual(prob["Qp_inv"], prob["Gp"], prob["Kp"], prob["Fp"], 0.0)
    y32, _ = oracle32.solve_fixed(Qd, Fd, K)
    y64, _ = oracle64.solve_fixed(Qd.astype(np.float64), Fd.astype(np.float64), K)
    check_fast(Y1[0], y32, y64, "dense Qp_inv, upper-triangle loop")
    check_fast(Y0[0], y32, y64, "dense Qp_inv, full-matrix loop")
    assert np.array_equal(active_set(Y1[0]), active_set(y32))
    assert np.array_equal(U1[0], oracle32.recover_u(Y1[0], prob["Fp"], prob["Gp"], prob["Qp_inv"]))
    # an unsymmetric Qp_inv: Qd is not symmetric even in exact arithmetic -> left exactly as computed
    bad = dict(prob, Qp_inv=np.ascontiguousarray(Qi + np.triu(np.full((M, M), 0.01, np.float32), 1)))
    with pqp.Solver(d, bad, exploit_symmetry=0) as s:
        Qa, _, _ = s.dual()
    with pqp.Solver(d, bad) as s:
        Qb, _, _ = s.dual()
        s.solve(iters=3)
        assert not s.last_kernel.startswith("gemv_sym")
    assert np.array_equal(Qa, Qb)


def test_full_matrix_loop_rows_in_tensor_memory_change_nothing(pqp):
    """The full-matrix TMA loop parks eight rows per slab in tensor memory (PQP_TMA_TMEM, long-row instantiation: 4096 < N <= 8192);
    where a row is kept does not enter the arithmetic: 0, 5 and 8 parked rows give the same bits, also on a ragged size."""
    import os
    for N, M, K in ((6100, 1500, 10), (8192, 2048, 5), (4800, 1200, 12)):
        prob, d = pqp.generate_testproblem(78, M, N)
        got = {}
        for tm in ("0", "5", "8"):
            os.environ["PQP_TMA_TMEM"] = tm
            try:
                with pqp.Solver(d, prob, exploit_symmetry=0) as s:
                    Y, _, st = s.solve(iters=K)
                    assert s.last_kernel.startswith("gemv_tma"), (N, tm, s.last_kernel)
                    got[tm] = (Y.copy(), st["Jd"][0], st["gap"][0])
            finally:
                os.environ.pop("PQP_TMA_TMEM", None)
        for tm in ("5", "8"):
            assert np.array_equal(got["0"][0], got[tm][0]) and got["0"][1:] == got[tm][1:], (N, tm)
