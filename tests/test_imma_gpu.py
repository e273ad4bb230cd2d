"""GPU (-m gpu): the batched PQP loop on the int8 tensor cores (pqp_batched_imma.cu) through the C ABI.

 1. bit-identical to its numpy model (tests/imma_model.py) -- integer accumulation is exact and the fp32 epilogue is
    deterministic, so anything but equality is a kernel bug (descriptors, barriers, digit layout);
 2. against the oracle and its float64 twin with the tolerance rule of DESIGN.md 4, at the config-C4 shape;
 3. full-size C4 (B=4096, 1000 updates): size-independent properties, and the degenerate problems behave like the reference.
"""
import os

import numpy as np
import pytest

from conftest import active_set, relerr
from imma_model import has_pair_structure, run as model_run, run_paired as model_run_paired

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(autouse=True)
def _tensor_core_kernels_for_tiny_problems(monkeypatch):
    """Batches of problems that fit one thread block (N <= 64) are served by one block per problem; this file is about the
    tensor-core kernels, also on the tiny shapes that make good edge cases for them."""
    monkeypatch.setenv("PQP_CTA_BATCH_MAX", "0")


def _mpc(seed, pH, nS, nI, B):
    from bench_problems import condensed_mpc
    return condensed_mpc(seed, pH, nS, nI, n_states=B)


@pytest.mark.parametrize("pH,nS,nI,B,K", [(6, 5, 2, 37, 60), (9, 4, 3, 70, 40), (31, 4, 4, 33, 25), (30, 12, 4, 96, 120),
                                          (4, 4, 1, 5, 30)])
def test_imma_is_bit_identical_to_its_model(pqp, pH, nS, nI, B, K):
    prob, d, X = _mpc(11, pH, nS, nI, B)
    with pqp.Solver(d, prob, batch_capacity=B) as s:
        Y, _, _ = s.solve(X, iters=K, status=False)
        assert s.last_kernel in ("batched_imma", "batched_imma_pair", "batched_imma_paired")
        kernel = s.last_kernel
        Qd, th, _ = s.dual()
        Fd, _ = s.linear_terms(B, want_fp=False)
    n = min(B, 40)
    model = model_run_paired if kernel == "batched_imma_paired" else model_run
    assert np.array_equal(Y[:n], model(Qd, th, Fd[:n], K), equal_nan=True)


@pytest.mark.parametrize("pH,nS,nI,B,K", [(30, 12, 4, 130, 150), (17, 4, 4, 97, 40), (13, 3, 5, 64, 30), (32, 4, 4, 33, 20), (16, 4, 4, 70, 25)])
def test_paired_rows_kernel_is_bit_identical_to_its_model(pqp, pH, nS, nI, B, K):
    """Duals of the reference's MPC layout (constraint rows in +/- pairs) run on the PAIRED instantiation of the CTA-pair kernel:
    only the N/2 representative rows go through the tensor cores, each epilogue thread updates a row and its partner.  The
    handle's own tensor-core Qd has the structure element for element (checked on the device, and here with numpy); the kernel
    equals tests/imma_model.py::run_paired bit for bit for N = 480, 272, 260 (two representatives in the second M tile), 512
    (full tiles), 256 < N only, ragged batches; exploit_structure = 0 falls back to all N rows."""
    prob, d, X = _mpc(11, pH, nS, nI, B)
    with pqp.Solver(d, prob, batch_capacity=B) as s:
        Y, _, _ = s.solve(X, iters=K, status=False)
        Qd, th, _ = s.dual()
        Fd, _ = s.linear_terms(B, want_fp=False)
        assert has_pair_structure(Qd)
        if d.N > 256:
            assert s.last_kernel == "batched_imma_paired", s.last_kernel
            n = min(B, 70)
            assert np.array_equal(Y[:n], model_run_paired(Qd, th, Fd[:n], K), equal_nan=True)
        else:
            assert s.last_kernel == "batched_imma_pair", s.last_kernel
    with pqp.Solver(d, prob, batch_capacity=B, exploit_structure=0) as s:
        Y0, _, _ = s.solve(X, iters=K, status=False)
        assert s.last_kernel == "batched_imma_pair", s.last_kernel
        assert np.array_equal(Y0[:40], model_run(Qd, th, Fd[:40], K), equal_nan=True)
    fin = np.isfinite(Y0).all(axis=1) & np.isfinite(Y).all(axis=1)
    assert relerr(Y[fin], Y0[fin]) <= 1e-4          # the two schemes agree to rounding (each is held to the oracle elsewhere)


@pytest.mark.parametrize("pH,nS,nI,B,K", [(30, 12, 4, 130, 150), (17, 4, 4, 97, 40), (9, 4, 4, 64, 30), (32, 4, 4, 33, 20)])
def test_single_cta_and_pair_kernels_are_bit_identical(pqp, monkeypatch, pH, nS, nI, B, K):
    """The CTA-pair kernel (two SMs share 64 problems, each multiplies half the M tiles, digit planes exchanged through DSMEM)
    and the single-CTA kernel run the same arithmetic: same bits, for 2, 3 and 4 M tiles and ragged batches."""
    prob, d, X = _mpc(7, pH, nS, nI, B)
    out = {}
    monkeypatch.setenv("PQP_IMMA_PAIRED", "0")     # all N rows in both kernels (the paired-rows scheme has its own test)
    for pair in ("0", "1"):
        monkeypatch.setenv("PQP_IMMA_PAIR", pair)
        with pqp.Solver(d, prob, batch_capacity=B) as s:
            out[pair], _, _ = s.solve(X, iters=K, status=False)
            assert s.last_kernel == ("batched_imma_pair" if pair == "1" else "batched_imma")
    assert np.array_equal(out["0"], out["1"], equal_nan=True)


def test_imma_warm_start_and_chunking_are_exact(pqp):
    prob, d, X = _mpc(5, 8, 4, 3, 50)
    with pqp.Solver(d, prob, batch_capacity=50) as s:
        Ya, _, _ = s.solve(X, iters=70, status=False)
        Yb, _, _ = s.solve(X, iters=30, Y0=Ya, status=False)
        Yc, _, _ = s.solve(X, iters=100, status=False)
        assert np.array_equal(Yb, Yc)
        Yd, _, _ = s.solve(X[:37], iters=100, status=False)   # ragged tail, different CTA population
        assert np.array_equal(Yd, Yc[:37])
        # up to 16 problems of this size run on one thread-block cluster each (plain fp32, the single-problem arithmetic): the same
        # answers to rounding, and again independent of who else is in the batch
        Ye, _, _ = s.solve(X[:13], iters=100, status=False)
        assert s.last_kernel == "gemv_cluster_batch", s.last_kernel
        assert np.abs(Ye - Yc[:13]).max() <= 2e-5 * np.abs(Yc[:13]).max()
        Yf, _, _ = s.solve(X[:5], iters=100, status=False)
        assert np.array_equal(Yf, Ye[:5])
        Yg, _, _ = s.solve(X[:1], iters=100, status=False)
        assert np.array_equal(Yg, Ye[:1])


def test_c4_shape_against_oracle(pqp, oracle32, oracle64):
    """Config C4 shape: N=480, 1000 fixed updates; 6 of 200 states checked against the oracle (0.5 s each).

    Tolerance: 1e-5 normwise where the instance allows it.  Several of these states are ill-conditioned enough that the
    oracle's own fp32 rounding noise err(f32,f64) is 2e-5 - 5e-5 after 1000 updates (state 12: 4.3e-5); there any fp32-level
    implementation lands at a comparable but different distance from the float64 twin (the SIMT kernel included), so the bound
    is 3x the oracle's own noise, with identical active sets."""
    prob, d, X = _mpc(2024, 30, 12, 4, 200)
    K = 1000
    with pqp.Solver(d, prob, batch_capacity=200) as s:
        Y, U, st = s.solve(X, iters=K, primal=True)
        assert s.last_kernel == "batched_imma_paired"
        Qd, th, _ = s.dual()
        Fd, Fp = s.linear_terms(200)
    worst = 0.0
    for b in (0, 7, 12, 23, 101, 199):
        y32, _ = oracle32.solve_fixed(Qd, Fd[b], K)
        y64, _ = oracle64.solve_fixed(Qd, Fd[b], K)
        if not np.isfinite(y32).all():
            assert not np.isfinite(Y[b]).all()
            continue
        e_gf, e_gd, e_fd = relerr(Y[b], y32), relerr(Y[b], y64), relerr(y32, y64)
        assert e_gf <= TOL or e_gd <= 3 * max(e_fd, TOL / 3), (b, e_gf, e_gd, e_fd)
        assert np.array_equal(active_set(Y[b], 1e-5), active_set(y32, 1e-5)), b
        assert np.array_equal(U[b], oracle32.recover_u(Y[b], Fp[b], prob["Gp"], prob["Qp_inv"]))
        worst = max(worst, e_gd / max(e_fd, 1e-12))
    print("worst err(gpu,f64)/err(f32,f64) over the sample:", worst)


def test_c4_full_size_properties(pqp, oracle32):
    """B=4096 states, 1000 updates (the bench configuration): y >= 0; the KKT residual shrinks with more updates; problems
    with no active constraint end at exactly 0; the 13 states whose only violated constraint underflows in fp32 end as NaN in
    every component, exactly as PQP_CPU.c does on them (0/0 at :594, then 0*NaN in every row)."""
    prob, d, X = _mpc(2024, 30, 12, 4, 4096)
    with pqp.Solver(d, prob, batch_capacity=4096) as s:
        Y, U, st = s.solve(X, iters=1000, primal=True)
        assert s.last_kernel == "batched_imma_paired"
        Fd, _ = s.linear_terms(4096, want_fp=False)
        Qd, _, _ = s.dual()
        Y2, _, st2 = s.solve(X, iters=1500)
    fin = np.isfinite(Y).all(axis=1)
    assert (~fin).sum() < 0.01 * len(Y)
    assert np.isnan(Y[~fin]).all()
    assert np.all(Y[fin] >= 0)
    no_active = (Fd > 0).all(axis=1)
    assert np.all(Y[no_active & fin] == 0)
    ok = fin & np.isfinite(Y2).all(axis=1)
    assert np.median(st2["kkt"][ok]) <= np.median(st["kkt"][ok])
    b = int(np.where(~fin)[0][0])
    y32, _ = oracle32.solve_fixed(Qd, Fd[b], 1000)
    assert np.isnan(y32).all()


def test_engines_agree_within_tolerance(pqp, oracle32, oracle64):
    """The three batched engines on the same states: int8 tensor (default), fp32 SIMT (use_tensor_cores=0), 3xTF32 tensor
    (opt-in, use_tensor_cores=2: documented as above the parity tolerance after many updates, so it is only required to
    stay within 1e-3 here)."""
    prob, d, X = _mpc(2, 10, 6, 3, 64)
    K = 80
    out = {}
    for name, tc in (("imma", 1), ("simt", 0), ("umma", 2)):
        with pqp.Solver(d, prob, batch_capacity=64, use_tensor_cores=tc) as s:
            out[name], _, _ = s.solve(X, iters=K, status=False)
            assert s.last_kernel.startswith("batched_" + name)
            Qd, th, _ = s.dual()
            Fd, _ = s.linear_terms(64, want_fp=False)
    for b in (0, 31, 63):
        y32, _ = oracle32.solve_fixed(Qd, Fd[b], K)
        y64, _ = oracle64.solve_fixed(Qd, Fd[b], K)
        e_fd = relerr(y32, y64)
        for name in ("imma", "simt"):
            e_gf, e_gd = relerr(out[name][b], y32), relerr(out[name][b], y64)
            assert e_gf <= TOL or (e_gf <= 2 * e_fd and e_gd <= max(2 * e_fd, TOL)), (name, b, e_gf, e_gd, e_fd)
        assert relerr(out["umma"][b], y32) <= 1e-3


def test_batched_run_to_tolerance(pqp):
    """iters <= 0 with B > 1: the int8 kernel evaluates the stop test of terminate() (PQP_CPU.c:673-687, on g = Qd y + Fd,
    SURVEY 3.3) per problem every check_every updates and freezes a problem at exactly the y that passed.  Properties:
    a problem's result is bit-identical to the fixed-count run stopped at its own iteration count; the single-problem kernel
    stops at (or within one check of) the same count; the reported status is the one of that y."""
    prob, d, X = _mpc(2024, 6, 5, 2, 70)
    with pqp.Solver(d, prob, batch_capacity=70, eaj=1e-3, erj=1e-6, check_every=8, max_iters=20000) as s:
        Y, U, st = s.solve(X, iters=0, primal=True)
        assert s.last_kernel == "batched_imma"
        assert st["converged"].all() and np.all(st["iters"] % 8 == 0) and st["iters"].min() >= 8 and st["iters"].max() < 20000
        assert np.all(np.abs(st["gap"]) <= 1e-3) and np.all(st["min_slack"] >= -1e-3)
        for b in (0, 35, 69):
            Yf, _, stf = s.solve(X, iters=int(st["iters"][b]))
            assert np.array_equal(Yf[b], Y[b])
            assert abs(stf["Jd"][b] - st["Jd"][b]) <= 1e-5 * abs(st["Jd"][b]) + 1e-3
            assert abs(stf["gap"][b] - st["gap"][b]) <= 2e-2 and abs(stf["kkt"][b] - st["kkt"][b]) <= 1e-3
        for b in (0, 69):
            Y1, _, st1 = s.solve(X[b][None], iters=0)
            assert s.last_kernel.startswith("gemv_")
            assert st1["converged"][0] == 1 and abs(int(st1["iters"][0]) - int(st["iters"][b])) <= 8
            assert relerr(Y1[0], Y[b]) <= 1e-4


def test_batched_run_to_tolerance_retires_degenerate_states(pqp):
    """C4 shape: states on which the reference itself ends in 0/0 are retired unconverged; the others converge and stop early."""
    prob, d, X = _mpc(2024, 30, 12, 4, 1100)
    with pqp.Solver(d, prob, batch_capacity=1100, eaj=1e-2, erj=1e-6, check_every=16, max_iters=6000) as s:
        Y, _, st = s.solve(X, iters=0)
        nan_states = ~np.isfinite(Y).all(axis=1)
        assert nan_states.sum() >= 1 and not st["converged"][nan_states].any()   # state 1068 is one of them (test_c4_full_size_properties)
        assert st["iters"][nan_states].max() < 6000
        good = ~nan_states
        assert st["converged"][good].mean() > 0.9
        assert np.median(st["iters"][good]) < 3000


def test_receding_horizon_warm_start(pqp, oracle32):
    """the MPC loop the report describes.  pqp_shift_duals moves each of the four constraint blocks one horizon
    step forward; warm-started periods reach the same controls as cold-started ones in fewer updates; the first period equals
    the plain solve; and the shift itself is checked against numpy."""
    sys_path_tools = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools")
    import sys
    sys.path.insert(0, sys_path_tools)
    from mpc_closed_loop import closed_loop
    from bench_problems import condensed_mpc
    pH, nS, nI, B, T = 8, 4, 3, 64, 5
    prob, d, X, plant = condensed_mpc(5, pH, nS, nI, n_states=B, x_scale=20.0, return_plant=True)
    opts = dict(batch_capacity=B, eaj=1e-3, erj=1e-6, check_every=8, max_iters=20000)
    with pqp.Solver(d, prob, **opts) as s:
        Y, _, _ = s.solve(X, iters=50)
        Ys = s.shift_duals(Y, 0.25)
        ref = Y.reshape(B, 4, pH, nI).copy()
        ref[:, :, :-1] = ref[:, :, 1:]
        assert np.array_equal(Ys, np.maximum(ref.reshape(B, -1), np.float32(0.25)))
        Xc, Uc, stc = closed_loop(s, plant, d, prob, X, T, warm=False)
        Xw, Uw, stw = closed_loop(s, plant, d, prob, X, T, warm=True)
    ok = np.logical_and.reduce([st["converged"] == 1 for st in stc + stw])
    assert ok.mean() >= 0.9, ok.mean()
    assert np.array_equal(Uc[0], Uw[0])                                # the first period has nothing to warm-start from
    scale = np.abs(Uc[:, ok]).max()
    assert np.abs(Uc[:, ok] - Uw[:, ok]).max() <= 5e-3 * scale         # same controls up to the stop tolerance
    its_c = sum(int(st["iters"][ok].sum()) for st in stc[1:])
    its_w = sum(int(st["iters"][ok].sum()) for st in stw[1:])
    assert its_w < 0.8 * its_c, (its_w, its_c)
    assert np.abs(Xw[-1][ok]).mean() < np.abs(Xw[0][ok]).mean()        # the loop regulates


def test_stopping_points_against_the_reference_stop_test(pqp, oracle32):
    """SURVEY 8(f)3, statistically: terminate() (PQP_CPU.c:673-687, oracle restatement, float32, the reference's 1e-6) against the
    fused stop test of the CUDA loop on g = Qd y + Fd, checked after every update, on 16 small condensed-MPC states.  The
    reference's first condition carries no tolerance (it waits for Jp <= -Jd, which float32 reaches on rounding noise), so
    the counts are compared as a distribution: every state stops, within a few percent of the reference's count in the
    median and never far from it, at the same primal solution (tools/convergence_stats.py prints the table)."""
    from bench_problems import condensed_mpc
    n, cap, tol = 16, 30000, 1e-6
    mp, md, X = condensed_mpc(11, pH=6, nS=4, nI=2, n_states=n, x_scale=25.0)
    o = oracle32
    Qp = o.gauss_jordan(mp["Qp_inv"])
    ref = []
    for b in range(n):
        Fp = o.compute_fp(mp["Fp1"], mp["Fp2"], mp["Fp3"], mp["D"], X[b])
        Mp = o.compute_mp(*[mp[k] for k in ("Mp1", "Mp2", "Mp3", "Mp4", "Mp5", "Mp6")], mp["D"], X[b])
        Qd, Fd, Md, _ = o.convert_to_dual(mp["Qp_inv"], mp["Gp"], mp["Kp"], Fp, Mp)
        Y, U, h = o.solve_converge(Qd, Fd, Md, Qp, mp["Qp_inv"], Fp, Mp, mp["Gp"], mp["Kp"], max_h=cap)
        assert h < cap
        ref.append((h - 1, U))
    with pqp.Solver(md, mp, eaj=tol, erj=tol, erc=tol, eac=tol, check_every=1, max_iters=cap, batch_capacity=n) as s:
        Y, U, st = s.solve(X, iters=0, primal=True)
    assert np.all(st["converged"] == 1)
    ratio = np.array([st["iters"][b] / max(ref[b][0], 1) for b in range(n)])
    print("ours / reference update counts: median %.3f min %.3f max %.3f" % (np.median(ratio), ratio.min(), ratio.max()))
    assert 0.9 <= np.median(ratio) <= 1.1 and ratio.min() >= 0.7 and ratio.max() <= 1.5
    for b in range(n):
        assert np.abs(U[b] - ref[b][1]).max() <= 2e-4 * max(np.abs(ref[b][1]).max(), 1.0), b


@pytest.mark.parametrize("pH,nS,nI,B", [(30, 12, 4, 200), (17, 4, 4, 97), (32, 4, 4, 65)])
def test_refresh_and_recovery_fused_into_the_loop_kernel(pqp, oracle32, monkeypatch, pH, nS, nI, B):
    """pqp_solve_batch_primal on the paired-rows kernel is ONE launch: computeFp + computeFd (PQP_CPU.c:373-382, :456-460) in its
    prologue, computeUfromY (:352-360) in its epilogue, all in the reference's order -- bit-identical to the separate
    fp / fd / recover kernels (PQP_IMMA_FUSE=0), Fd and Fp bit-identical to the oracle's, U bit-identical to the oracle's
    computeUfromY on the GPU's y; warm starts (Y0) and per-problem disturbances (D) included."""
    prob, d, X = _mpc(5, pH, nS, nI, B)
    rng = np.random.default_rng(1)
    Dm = (prob["D"][None, :] + 0.05 * rng.standard_normal((B, d.nDisH))).astype(np.float32)
    K = 40
    out = {}
    for fuse in ("1", "0"):
        monkeypatch.setenv("PQP_IMMA_FUSE", fuse)
        with pqp.Solver(d, prob, batch_capacity=B) as s:
            s.solve(X, iters=1, primal=True, status=False)           # the one-time launches (structure test, digit planes) happen here
            l0 = s.launch_count
            Y, U, _ = s.solve(X, iters=K, primal=True, status=False)
            n1 = s.launch_count - l0
            assert s.last_kernel == "batched_imma_paired"
            Fd, Fp = s.linear_terms(B)
            Y2, U2, st2 = s.solve(X, iters=K, primal=True, D=Dm, Y0=Y)        # warm start, per-problem D, status
            Fd2, Fp2 = s.linear_terms(B)
            out[fuse] = (Y, U, Fd, Fp, Y2, U2, Fd2, Fp2, st2, n1)
    f, u = out["1"], out["0"]
    assert f[9] == 1 and u[9] >= 6, (f[9], u[9])          # one launch against fp, fd, fill, loop, recover x 2
    for a, b in zip(f[:8], u[:8]):
        assert np.array_equal(a, b, equal_nan=True)
    for k in ("iters", "min_slack", "gap", "Jd", "kkt"):
        assert np.array_equal(f[8][k], u[8][k], equal_nan=True), k
    for b in (0, B // 2, B - 1):
        Fpb = oracle32.compute_fp(prob["Fp1"], prob["Fp2"], prob["Fp3"], prob["D"], X[b])
        _, Fdb, _, _ = oracle32.convert_to_dual(prob["Qp_inv"], prob["Gp"], prob["Kp"], Fpb, 0.0, want_qd=False)
        assert np.array_equal(f[3][b], Fpb)
        # Fd = GQ Fp + Kp (a cancelling sum) with the handle's tensor-core GQ: against the oracle's reference-order GQ only to rounding
        assert relerr(f[2][b], Fdb) <= 2e-5
        if np.isfinite(f[0][b]).all():
            assert np.array_equal(f[1][b], oracle32.recover_u(f[0][b], Fpb, prob["Gp"], prob["Qp_inv"]))
        Fpd = oracle32.compute_fp(prob["Fp1"], prob["Fp2"], prob["Fp3"], Dm[b], X[b])
        assert np.array_equal(f[7][b], Fpd)


def test_run_to_tolerance_on_the_paired_rows_kernel(pqp, monkeypatch):
    """iters <= 0 on the paired-rows kernel (N > 256, MPC layout): chunks of check_every updates + an evaluation pass in the same
    launch, terminate()'s test per problem in a decide kernel, the duals of a problem that passes kept exactly as they stand.
    Properties: every problem's result is bit-identical to the fixed-count solve at its own count (the chunks resume with the
    scale state of the uninterrupted loop); the status block is the one of that y; the single-CTA tolerance kernel (all N rows,
    exact-maximum scales) stops within one check of the same counts at the same solution; U comes from the kept duals."""
    prob, d, X = _mpc(2024, 30, 12, 4, 100)
    X = (X * 0.5).astype(np.float32)                      # x_scale 30: every state converges within a few thousand updates
    opts = dict(batch_capacity=100, eaj=1e-2, erj=1e-6, check_every=16, max_iters=8000)
    with pqp.Solver(d, prob, **opts) as s:
        Y, U, st = s.solve(X, iters=0, primal=True)
        assert s.last_kernel == "batched_imma_paired_tol", s.last_kernel
        fin = np.isfinite(Y).all(axis=1)
        assert st["converged"][fin].mean() > 0.9 and np.all(st["iters"] % 16 == 0) and st["iters"].min() >= 16
        conv = st["converged"] == 1
        assert np.all(np.abs(st["gap"][conv]) <= 1e-2)
        for b in (0, 41, 99):
            k = int(st["iters"][b])
            Yf, _, stf = s.solve(X, iters=k)
            assert s.last_kernel == "batched_imma_paired"
            assert np.array_equal(Yf[b], Y[b], equal_nan=True), (b, k)
            if conv[b]:
                assert abs(stf["gap"][b] - st["gap"][b]) <= 2e-2 and abs(stf["kkt"][b] - st["kkt"][b]) <= 1e-3
        Fd, Fp = s.linear_terms(100)
    monkeypatch.setenv("PQP_IMMA_PAIRED_TOL", "0")
    with pqp.Solver(d, prob, **opts) as s:
        Y1, U1, st1 = s.solve(X, iters=0, primal=True)
        assert s.last_kernel == "batched_imma"
    both = conv & (st1["converged"] == 1)
    assert both.mean() > 0.9
    assert np.median(np.abs(st["iters"][both].astype(int) - st1["iters"][both].astype(int))) <= 16
    assert relerr(U[both], U1[both]) <= 1e-3
