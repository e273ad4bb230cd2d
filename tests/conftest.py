import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
EXAMPLE_DIR = os.path.join(GOLDEN, "example")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are never silently green on a CPU box: they are skipped with an explicit reason, and on a GPU
    box a missing libpqp_b200.so is an error (the product has no fallback path to hide behind)."""
    try:
        import pqp_for_mpc_b200 as pqp
        ndev = pqp.device_count()
    except Exception:
        ndev = -1
    if ndev == 0:
        skip = pytest.mark.skip(reason="no sm_100 CUDA device in this container (run under gpurun)")
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(skip)


@pytest.fixture(scope="session")
def pqp():
    import pqp_for_mpc_b200 as m
    if not os.path.exists(m.LIB_PATH):
        m.build()
    return m


@pytest.fixture(scope="session")
def oracle32():
    from oracle.oracle import Oracle
    return Oracle(np.float32)


@pytest.fixture(scope="session")
def oracle64():
    from oracle.oracle import Oracle
    return Oracle(np.float64)


@pytest.fixture(scope="session")
def gold_example():
    return np.load(os.path.join(GOLDEN, "golden_example.npz"))


@pytest.fixture(scope="session")
def gold_random():
    return np.load(os.path.join(GOLDEN, "golden_random.npz"))


@pytest.fixture(scope="session")
def gold_testfiles():
    """The reference's larger generated instances test1.txt / test3.txt (tests/golden/make_golden_testfiles.py)."""
    return np.load(os.path.join(GOLDEN, "golden_testfiles.npz"))


def problem_from_testfile(g, name):
    return dict(Qp_inv=np.diag(g[f"{name}_Qp_inv_diag"]).astype(np.float32), Fp=g[f"{name}_Fp"], Kp=g[f"{name}_Kp"],
                Gp=g[f"{name}_Gp"].astype(np.float32), Mp0=float(g[f"{name}_Mp0"]))


def relerr(a, b):
    """normwise relative error ||a-b||inf / ||b||inf (the tolerance metric of DESIGN.md)"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))


def active_set(y, rel=1e-6):
    y = np.asarray(y, np.float64)
    return y > rel * np.abs(y).max()


RANDOM_CASES = [(101, 32, 64, 50), (102, 100, 40, 50), (103, 64, 256, 200), (104, 300, 200, 200), (105, 200, 257, 100),
                (106, 7, 5, 30)]


def golden_problem(gold_random, seed):
    t = f"s{seed}"
    return dict(Qp_inv=np.diag(gold_random[f"{t}_Qp_inv_diag"]).astype(np.float32), Fp=gold_random[f"{t}_Fp"],
                Kp=gold_random[f"{t}_Kp"], Gp=gold_random[f"{t}_Gp"].astype(np.float32), Mp0=float(gold_random[f"{t}_Mp0"]))
