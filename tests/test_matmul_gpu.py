"""GPU tests of pqp_matmul = the reference's matrixMultiply (PQP_CPU.c:84-147) on the device, three engines:
STRICT must equal the oracle bit for bit (all four transpose variants, ragged shapes); SIMT and the tcgen05 3xTF32
TENSOR engine are held to fp32-level accuracy against a float64 product."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu

SHAPES = [(1, 1, 1), (7, 28, 7), (33, 17, 65), (128, 32, 128), (129, 40, 127), (300, 257, 200), (512, 512, 512)]


@pytest.mark.parametrize("a,b,c", SHAPES)
def test_strict_engine_is_bit_identical_to_matrixMultiply(pqp, oracle32, a, b, c):
    rng = np.random.default_rng(a * 1000 + b * 10 + c)
    A = rng.standard_normal((a, b)).astype(np.float32) * 10
    B = rng.standard_normal((b, c)).astype(np.float32)
    for tA in (0, 1):
        for tB in (0, 1):
            As = np.ascontiguousarray(A.T) if tA else A
            Bs = np.ascontiguousarray(B.T) if tB else B
            want = oracle32.matmul(As, tA, Bs, tB, a, b, c)
            got = pqp.matmul(As, Bs, tA=bool(tA), tB=bool(tB), engine=pqp.MM_STRICT)
            assert np.array_equal(got, want), (tA, tB)


@pytest.mark.parametrize("a,b,c", SHAPES + [(1024, 2048, 1024), (8192, 2048, 256), (257, 65, 193), (1000, 100, 1000), (640, 96, 384)])
@pytest.mark.parametrize("engine", ["simt", "tensor"])
def test_fast_engines_fp32_accuracy(pqp, a, b, c, engine):
    rng = np.random.default_rng(a + b + c)
    A = (rng.standard_normal((a, b)) * rng.uniform(0.1, 100, (a, 1))).astype(np.float32)
    Bt = rng.standard_normal((c, b)).astype(np.float32)
    want = A.astype(np.float64) @ Bt.astype(np.float64).T
    got = pqp.matmul(A, Bt, tB=True, engine=pqp.MM_TENSOR if engine == "tensor" else pqp.MM_SIMT)
    # entrywise error relative to the magnitude of the terms that were summed (what fp32 accumulation can promise)
    scale = np.abs(A).astype(np.float64) @ np.abs(Bt).astype(np.float64).T
    err = np.abs(got - want) / np.maximum(scale, 1e-30)
    assert err.max() <= 2e-6, f"{engine} max scaled error {err.max():.3e}"
    assert relerr(got, want) <= 4e-6  # normwise, with sign cancellation in the random operands


def test_tensor_engine_nonnegative_operands_match_simt(pqp):
    """The regime of the PQP loop: all operands >= 0, no cancellation -> 3xTF32 is within a few fp32 ulps."""
    rng = np.random.default_rng(3)
    A = rng.uniform(0, 10, (480, 480)).astype(np.float32)
    Bt = rng.uniform(0, 1000, (64, 480)).astype(np.float32)
    want = A.astype(np.float64) @ Bt.astype(np.float64).T
    t = pqp.matmul(A, Bt, tB=True, engine=pqp.MM_TENSOR)
    s = pqp.matmul(A, Bt, tB=True, engine=pqp.MM_SIMT)
    et, es = np.abs(t - want).max() / np.abs(want).max(), np.abs(s - want).max() / np.abs(want).max()
    print(f"relative error: tensor {et:.3e}  simt {es:.3e}")
    assert et <= 2e-6 and es <= 2e-6


def test_setup_uses_tensor_cores_and_matches_oracle_dual(pqp, gold_random):
    from conftest import golden_problem
    g, t = gold_random, "s104"
    prob = golden_problem(g, 104)
    with pqp.Solver(pqp.dims_plain(300, 200), prob, use_tensor_cores=1) as s:
        Qd, th, GQ = s.dual()
    with pqp.Solver(pqp.dims_plain(300, 200), prob, use_tensor_cores=0) as s:
        Qd0, th0, GQ0 = s.dual()
    assert relerr(Qd, g[f"{t}_Qd"]) <= 2e-6 and relerr(Qd0, g[f"{t}_Qd"]) <= 2e-6
    assert relerr(th, g[f"{t}_theta"]) <= 2e-6


@pytest.mark.parametrize("N,M", [(640, 96), (1000, 130), (2048, 256)])
def test_setup_builds_a_symmetric_dual_from_its_upper_triangle(pqp, N, M):
    """Qd = Gp Qp_inv Gp' (PQP_CPU.c:492, :442) with a symmetric, dense Qp_inv: pqp_setup multiplies only the tiles of the upper
    triangle and stores every element above the diagonal twice.  The upper triangle must equal the full product's bit for bit, the
    lower must be its mirror image, and the whole must be the float64 product to fp32 accuracy.  An unsymmetric Qp_inv (not a QP,
    but the reference would multiply it) takes the full product."""
    rng = np.random.default_rng(N + M)
    R = rng.standard_normal((M, M))
    Qs = (R @ R.T / M + np.eye(M)).astype(np.float32)
    Qs = ((Qs + Qs.T) * np.float32(0.5)).astype(np.float32)
    assert np.array_equal(Qs, Qs.T)
    Gp = rng.standard_normal((N, M)).astype(np.float32)
    prob = dict(Qp_inv=Qs, Gp=Gp, Kp=np.ones(N, np.float32), Fp=rng.standard_normal(M).astype(np.float32), Mp0=0.0)
    with pqp.Solver(pqp.dims_plain(M, N), prob) as s:
        Qd, _, _ = s.dual()
    with pqp.Solver(pqp.dims_plain(M, N), prob, exploit_symmetry=0) as s:
        Qfull, _, _ = s.dual()
    assert np.array_equal(Qd, Qd.T)
    iu = np.triu_indices(N)
    assert np.array_equal(Qd[iu], Qfull[iu])
    want = Gp.astype(np.float64) @ Qs.astype(np.float64) @ Gp.astype(np.float64).T
    assert relerr(Qd, want) <= 4e-6
    Qu = Qs.copy()
    Qu[0, 1] += np.float32(0.25)
    prob_u = dict(prob, Qp_inv=Qu)
    with pqp.Solver(pqp.dims_plain(M, N), prob_u) as s:
        Qd_u, _, _ = s.dual()
    want_u = Gp.astype(np.float64) @ Qu.astype(np.float64) @ Gp.astype(np.float64).T
    assert relerr(Qd_u, want_u) <= 4e-6 and not np.array_equal(Qd_u, Qd_u.T)
