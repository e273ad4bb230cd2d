"""Generates the golden vectors from the UNMODIFIED reference (oracle/_ref, i.e. /root/reference/PQP_CPU.c
compiled where it lies).  Run in the authoring container: python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle, Reference, build  # noqa: E402
import pqp_for_mpc_b200 as pqp  # noqa: E402  (host-side generator/loaders only; no GPU needed)

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

# (seed, M, N, K): generator-distribution instances (testing/test_generator.c:936-987)
RANDOM_CASES = [(101, 32, 64, 50), (102, 100, 40, 50), (103, 64, 256, 200), (104, 300, 200, 200), (105, 200, 257, 100),
                (106, 7, 5, 30)]


def main():
    build()
    assert Reference.available(), "reference library not built (needs /root/reference)"
    r32, r64 = Reference(np.float32), Reference(np.float64)

    # ---- the shipped example ------------------------------------------------------------------
    p = r32.load_example(REF)
    Fp = r32.compute_fp(p["Fp1"], p["Fp2"], p["Fp3"], p["D"], p["x"])
    Mp = r32.compute_mp(*[p[k] for k in ("Mp1", "Mp2", "Mp3", "Mp4", "Mp5", "Mp6", "D", "x")])
    Qd, Fd, Md = r32.convert_to_dual(p["Qp_inv"], p["Gp"], p["Kp"], Fp, Mp)
    Qp = r32.gauss_jordan(p["Qp_inv"])
    P, Nn, theta = r32.split(Qd)
    Yc, Uc, h = r32.solve_converge(Qd, Fd, Md, Qp, p["Qp_inv"], Fp, Mp, p["Gp"], p["Kp"])
    U = r32.recover_u(Yc, Fp, p["Gp"], p["Qp_inv"])
    Jp, Jd = r32.cost(U, Qp, Fp, Mp), r32.cost(Yc, Qd, Fd, Md)
    fixed = {f"Y_K{K}": r32.solve_fixed(Qd, Fd, K) for K in (1, 2, 10, 100, 312)}
    fixed64 = {f"Y64_K{K}": r64.solve_fixed(Qd, Fd, K) for K in (100, 312)}
    np.savez_compressed(os.path.join(HERE, "golden_example.npz"), Fp=Fp, Mp=np.float32(Mp), Qd=Qd, Fd=Fd,
                        Md=np.float32(Md), Qp=Qp, theta=theta, Y_conv=Yc, U_conv=U, h=np.int64(h), Jp=np.float32(Jp),
                        Jd=np.float32(Jd), stdout=np.array(r32.main_stdout(REF)), **fixed, **fixed64)
    print("example: h =", h, "Jp =", Jp, "Jd =", Jd, "U =", U)

    # ---- generator-distribution instances ---------------------------------------------------------
    out = {}
    for seed, M, N, K in RANDOM_CASES:
        prob, d = pqp.generate_testproblem(seed, M, N)
        Qd, Fd, Md = r32.convert_to_dual(prob["Qp_inv"], prob["Gp"], prob["Kp"], prob["Fp"], prob["Mp0"])
        _, _, theta = r32.split(Qd)
        Y = r32.solve_fixed(Qd, Fd, K)
        Y64 = r64.solve_fixed(Qd, Fd, K)
        U = r32.recover_u(Y, prob["Fp"], prob["Gp"], prob["Qp_inv"])
        tag = f"s{seed}"
        out.update({f"{tag}_Qp_inv_diag": np.diag(prob["Qp_inv"]).copy(), f"{tag}_Fp": prob["Fp"], f"{tag}_Kp": prob["Kp"],
                    f"{tag}_Gp": prob["Gp"].astype(np.int8), f"{tag}_Mp0": np.float32(prob["Mp0"]), f"{tag}_Qd": Qd,
                    f"{tag}_Fd": Fd, f"{tag}_Md": np.float32(Md), f"{tag}_theta": theta, f"{tag}_Y": Y, f"{tag}_Y64": Y64,
                    f"{tag}_U": U, f"{tag}_K": np.int64(K)})
        act = int((Y > 1e-6 * np.abs(Y).max()).sum())
        print(f"seed {seed} M={M} N={N} K={K}: |Y|max={np.abs(Y).max():.4g} active={act}/{N} "
              f"err(f32,f64)={np.abs(Y - Y64).max() / np.abs(Y64).max():.2e}")
    # the reference's own generated file, read literally (-1 stays -1, Kp from the file)
    prob, d = pqp.load_testfile(os.path.join(HERE, "test2.txt"))
    Qd, Fd, Md = r32.convert_to_dual(prob["Qp_inv"], prob["Gp"], prob["Kp"], prob["Fp"], prob["Mp0"])
    Y = r32.solve_fixed(Qd, Fd, 100)
    out.update(test2_Fd=Fd, test2_Y=Y, test2_Y64=r64.solve_fixed(Qd, Fd, 100), test2_K=np.int64(100),
               test2_U=r32.recover_u(Y, prob["Fp"], prob["Gp"], prob["Qp_inv"]), test2_Qd_rowsum=Qd.sum(1))
    np.savez_compressed(os.path.join(HERE, "golden_random.npz"), **out)
    for f in ("golden_example.npz", "golden_random.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
