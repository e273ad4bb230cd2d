"""CPU tests of the host side: file formats, the seeded generator, and that libpqp_b200.so loads
and exports every symbol include/pqp.h declares.  No compute call is made without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import EXAMPLE_DIR, GOLDEN, ROOT


def test_library_exports_every_declared_symbol(pqp):
    L = pqp.lib()
    header = open(os.path.join(ROOT, "include", "pqp.h")).read()
    declared = set(re.findall(r"\b(pqp_[a-z0-9_]+)\s*\(", header)) - {"pqp_error"}  # the enum name appears in prose
    assert declared == set(pqp.ABI_SYMBOLS), declared ^ set(pqp.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name
    compat = C.CDLL(pqp.COMPAT_PATH)
    for name in ("convertToDual", "solveQuadraticDual", "computeUfromY", "computeFp", "computeCost", "updateY2",
                 "pqp_compat_set_dims", "pqp_compat_set_order", "pqp_compat_set_fixed_iters"):
        assert hasattr(compat, name), name


def test_struct_layouts_match_header(pqp):
    # sizes the C side reports via behaviour: defaults land in the right fields
    o = pqp.default_opts()
    assert (o.theta_floor, o.y_init) == (5.0, 1000.0)           # PQP_CPU.c:240, :710
    assert abs(o.erc - 1e-6) < 1e-12 and abs(o.erj - 1e-6) < 1e-12  # PQP_CPU.c:19-22
    assert o.order == pqp.ORDER_FAST and o.device == -1 and o.check_every >= 1
    d = pqp.dims_mpc(**pqp.EXAMPLE_DIMS)
    assert (d.M, d.N, d.nDisH) == (7, 28, 1)                     # PQP_CPU.c:940-941
    d = pqp.dims_mpc(30, 12, 4, 4, 1)
    assert (d.M, d.N) == (120, 480)                              # config C4
    assert C.sizeof(pqp.Status) == pqp.STATUS_DTYPE.itemsize == 24


def test_output_offsets_helper(pqp):
    """pqp_output_offsets (host only): Kx = [0; 0; -Z; +Z], Kd = [0; 0; -Theta; +Theta] on the example's row blocks, and the
    shapes it must refuse."""
    prob, d = pqp.load_example(EXAMPLE_DIR)
    M, N, nO = d.M, d.N, d.nOutput * d.pHorizon
    Kx, Kd = pqp.output_offsets(d, prob["Z"], prob["Theta"])
    assert Kx.shape == (N, d.nState) and Kd.shape == (N, d.nDisH)
    Z, Th = prob["Z"].reshape(nO, d.nState), prob["Theta"].reshape(nO, d.nDisH)
    assert np.array_equal(Kx[2 * M:2 * M + nO], -Z) and np.array_equal(Kx[3 * M:3 * M + nO], Z)
    assert np.array_equal(Kd[2 * M:2 * M + nO], -Th) and np.array_equal(Kd[3 * M:3 * M + nO], Th)
    assert not Kx[:2 * M].any() and not Kd[:2 * M].any()
    bad = pqp.dims_plain(d.M, d.N + 1)
    with pytest.raises(pqp.PQPError):
        pqp.output_offsets(bad, prob["Z"], prob["Theta"])


def test_example_loader_matches_oracle_loader(pqp, oracle32):
    got, d = pqp.load_example(EXAMPLE_DIR)
    want = oracle32.load_example(EXAMPLE_DIR)
    for k, v in want.items():
        assert np.array_equal(got[k], v), k
    # spot KATs of the column-major -> row-major transposition
    assert got["Gp"].shape == (28, 7) and got["Gp"][0, 0] == 1.0 and got["Gp"][7, 0] == -1.0
    assert np.all(got["Gp"][14:] == 0) and np.all(got["Kp"][:14] == 20.0) and np.all(got["Kp"][14:] == 0.0)
    assert abs(got["Qp_inv"][0, 0] - 0.998114) < 1e-6 and abs(got["D"][0] - 312.15) < 1e-4


def test_example_loader_errors(pqp, tmp_path):
    with pytest.raises(pqp.PQPError) as e:
        pqp.load_example(str(tmp_path))
    assert e.value.code == -5
    # truncated file -> IO error, not garbage
    for f in os.listdir(EXAMPLE_DIR):
        (tmp_path / f).write_text(open(os.path.join(EXAMPLE_DIR, f)).read())
    (tmp_path / "Gp.txt").write_text("1.0 0.0 #")
    with pytest.raises(pqp.PQPError):
        pqp.load_example(str(tmp_path))


def test_testfile_reader_is_literal(pqp):
    prob, d = pqp.load_testfile(os.path.join(GOLDEN, "test2.txt"))
    assert (d.M, d.N) == (100, 400)
    tok = open(os.path.join(GOLDEN, "test2.txt")).read().split()
    assert abs(prob["Qp_inv"][0, 0] - float(tok[2])) < 1e-6
    assert np.count_nonzero(prob["Qp_inv"] - np.diag(np.diag(prob["Qp_inv"]))) == 0
    assert set(np.unique(prob["Gp"])) == {-1.0, 0.0, 1.0}     # -1 stays -1 (reference reader maps it to +1)
    kp_file = np.array(tok[2 + 100 + 100 + 1: 2 + 100 + 100 + 1 + 400], np.float32)
    assert np.array_equal(prob["Kp"], kp_file)                 # Kp from the file (reference overwrites with rand())
    gp_file = np.array(tok[2 + 100 + 100 + 1 + 400:], np.float32).reshape(400, 100)
    assert np.array_equal(prob["Gp"], gp_file)


def test_generator_is_seeded_and_roundtrips(pqp, tmp_path, gold_random):
    a, d = pqp.generate_testproblem(101, 32, 64)
    b, _ = pqp.generate_testproblem(101, 32, 64)
    c, _ = pqp.generate_testproblem(102, 32, 64)
    for k in ("Qp_inv", "Fp", "Kp", "Gp"):
        assert np.array_equal(a[k], b[k]) and not np.array_equal(a[k], c[k])
    # stable against the committed golden inputs
    assert np.array_equal(np.diag(a["Qp_inv"]), gold_random["s101_Qp_inv_diag"])
    assert np.array_equal(a["Gp"], gold_random["s101_Gp"].astype(np.float32))
    # distribution of testing/test_generator.c:936-987
    big, _ = pqp.generate_testproblem(7, 200, 600)
    assert 0 <= big["Kp"].min() and big["Kp"].max() < 100 and 40 < big["Kp"].mean() < 60
    frac = [(big["Gp"] == v).mean() for v in (-1, 0, 1)]
    assert all(abs(f - 1 / 3) < 0.01 for f in frac)
    path = str(tmp_path / "t.txt")
    pqp.write_testfile(path, a, d)
    r, d2 = pqp.load_testfile(path)
    assert (d2.M, d2.N) == (32, 64)
    for k in ("Qp_inv", "Fp", "Kp", "Gp"):
        assert np.array_equal(a[k], r[k]), k
    assert abs(a["Mp0"] - r["Mp0"]) < 1e-6


def test_no_cpu_fallback(pqp):
    """Without a B200 the solver refuses to compute (it must never silently run on the CPU)."""
    if pqp.device_count() > 0:
        pytest.skip("a GPU is present")
    prob, d = pqp.load_example(EXAMPLE_DIR)
    with pytest.raises(pqp.PQPError) as e:
        pqp.Solver(d, prob)
    assert e.value.code == -2
    with pytest.raises(pqp.PQPError):
        pqp.Solver(Qd=np.eye(4, dtype=np.float32))
    # the stand-alone device entry points refuse as well
    I4 = np.eye(4, dtype=np.float32)
    v4 = np.ones(4, np.float32)
    with pytest.raises(pqp.PQPError) as e:
        pqp.update_y2(v4, I4, I4, v4, v4)
    assert e.value.code == -2
    with pytest.raises(pqp.PQPError) as e:
        pqp.matmul(I4, I4)
    assert e.value.code == -2


def test_product_never_references_the_oracle():
    """The shipped sources must not include, link or import anything under oracle/."""
    pkg = os.path.join(ROOT, "pqp-for-mpc_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".c", ".cu", ".h", ".cuh", ".py")) or f == "Makefile":
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.lower(), (dirpath, f)


def test_numpy_port_of_the_generator_is_bit_identical():
    """bench_problems.generator_instance (what `bench.py --impl reference` uses, so that the CPU arm never loads the product
    library) against pqp_generate_testproblem, element for element."""
    import pqp_for_mpc_b200 as pqp
    from bench_problems import generator_instance
    for seed, M, N in ((12345, 512, 1024), (1, 1, 1), (7, 3, 5), (2**63 + 11, 40, 17)):
        a, da = generator_instance(seed, M, N)
        b, db = pqp.generate_testproblem(seed, M, N)
        assert (da.M, da.N) == (db.M, db.N) == (M, N)
        for k in ("Qp_inv", "Fp", "Kp", "Gp"):
            assert np.array_equal(a[k], b[k]), (seed, k)
        assert a["Mp0"] == b["Mp0"]


def test_bench_states_are_seeded_and_clearly_constrained():
    """bench_problems.condensed_mpc with the bench's state recipe: deterministic, shard-independent, and every state has at
    least BENCH_MIN_VIOLATED rows of Fd(x) below -1 (no trivially-zero problem); the recipe with min_violated = 0 is the
    round-1 stream unchanged."""
    from bench_problems import BENCH_MIN_VIOLATED, BENCH_X_SCALE, condensed_mpc
    p1, d1, X1 = condensed_mpc(2024, 6, 5, 2, n_states=3000, x_scale=BENCH_X_SCALE, min_violated=BENCH_MIN_VIOLATED)
    p2, d2, X2 = condensed_mpc(2024, 6, 5, 2, n_states=3000, x_scale=BENCH_X_SCALE, min_violated=BENCH_MIN_VIOLATED)
    assert np.array_equal(X1, X2) and (d1.M, d1.N, d1.nDisH) == (12, 48, 6)
    GQ = p1["Gp"].astype(np.float64) @ p1["Qp_inv"].astype(np.float64)
    Fp = (p1["Fp1"].astype(np.float64) @ p1["D"])[None] + X1.astype(np.float64) @ p1["Fp2"].T.astype(np.float64) - p1["Fp3"][None]
    Fd = Fp @ GQ.T + p1["Kp"][None]
    assert ((Fd < -0.99).sum(axis=1) >= BENCH_MIN_VIOLATED).all()
    _, _, X3 = condensed_mpc(2024, 6, 5, 2, n_states=3000)
    _, _, X4 = condensed_mpc(2024, 6, 5, 2, n_states=3000, x_scale=60.0, min_violated=0)
    assert np.array_equal(X3, X4)
