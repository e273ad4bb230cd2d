"""Golden vectors for the reference's two larger generated instances, `testing/sample test/test1.txt` (M=500, N=1500) and
`test3.txt` (M=800, N=1200): the parsed inputs in compact form (Qp_inv is diagonal in the generator's files, Gp is {0,+-1})
and what the UNMODIFIED reference (oracle/_ref = /root/reference/PQP_CPU.c compiled where it lies) and its float64 twin make
of them after 100 updates.  The text files themselves (4 MB) stay in /root/reference; the text format is covered by test2.txt.
Run in the authoring container: python tests/golden/make_golden_testfiles.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import Reference, build  # noqa: E402
import pqp_for_mpc_b200 as pqp  # noqa: E402  (host-side loader only; no GPU needed)

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/testing/sample test"
K = 100


def main():
    build()
    assert Reference.available(), "reference library not built (needs /root/reference)"
    r32, r64 = Reference(np.float32), Reference(np.float64)
    out = {}
    for name in ("test1", "test3"):
        prob, d = pqp.load_testfile(os.path.join(SRC, name + ".txt"))
        Qi = prob["Qp_inv"].reshape(d.M, d.M)
        assert np.count_nonzero(Qi - np.diag(np.diag(Qi))) == 0, "Qp_inv not diagonal: store it whole"
        Gp = prob["Gp"].reshape(d.N, d.M)
        assert np.array_equal(Gp, Gp.astype(np.int8).astype(np.float32))
        Qd, Fd, Md = r32.convert_to_dual(prob["Qp_inv"], prob["Gp"], prob["Kp"], prob["Fp"], prob["Mp0"])
        _, _, theta = r32.split(Qd)
        Y = r32.solve_fixed(Qd, Fd, K)
        Y64 = r64.solve_fixed(Qd, Fd, K)
        U = r32.recover_u(Y, prob["Fp"], prob["Gp"], prob["Qp_inv"])
        out.update({f"{name}_M": np.int64(d.M), f"{name}_N": np.int64(d.N), f"{name}_K": np.int64(K),
                    f"{name}_Qp_inv_diag": np.diag(Qi).copy(), f"{name}_Gp": Gp.astype(np.int8), f"{name}_Fp": prob["Fp"],
                    f"{name}_Kp": prob["Kp"], f"{name}_Mp0": np.float32(prob["Mp0"]), f"{name}_Fd": Fd, f"{name}_theta": theta,
                    f"{name}_Qd_rowsum": Qd.sum(1), f"{name}_Qd_diag": np.diag(Qd).copy(), f"{name}_Y": Y, f"{name}_Y64": Y64, f"{name}_U": U})
        act = int((Y > 1e-6 * np.abs(Y).max()).sum())
        print(f"{name}: M={d.M} N={d.N} K={K} |Y|max={np.abs(Y).max():.4g} active={act}/{d.N} "
              f"err(f32,f64)={np.abs(Y - Y64).max() / np.abs(Y64).max():.2e} symmetric={bool(np.array_equal(Qd, Qd.T))}")
    path = os.path.join(HERE, "golden_testfiles.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
