"""CPU: accuracy of the int8 digit-plane scheme (tests/imma_model.py) against the oracle and its float64 twin on C4 problems.
usage: python tools/ozaki_emulate.py [K=1000] [B=24] [ybits=22]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.oracle import Oracle  # noqa: E402
from imma_model import run, slice_rows_u8, slice_y_s8  # noqa: E402,F401


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    ybits = int(sys.argv[3]) if len(sys.argv) > 3 else 22
    from bench_problems import condensed_mpc
    prob, d, X = condensed_mpc(2024, 30, 12, 4, n_states=B)
    o32, o64 = Oracle(np.float32), Oracle(np.float64)
    Fds = []
    Qd = None
    for b in range(B):
        Fp = o32.compute_fp(prob["Fp1"], prob["Fp2"], prob["Fp3"], prob["D"], X[b])
        Qd, Fd, _, _ = o32.convert_to_dual(prob["Qp_inv"], prob["Gp"], prob["Kp"], Fp, 0.0)
        Fds.append(Fd)
    Fds = np.stack(Fds)
    theta = o32.theta(Qd)
    Yz = run(Qd, theta, Fds, K, ybits)
    print(f"K={K} ybits={ybits}:  problem  ymax   e(i8,f64)  e(f32,f64)  e(i8,f32)  active-set(i8==f64, f32==f64)")
    worst = 0.0
    for b in range(B):
        y64, _ = o64.solve_fixed(Qd, Fds[b], K)
        y32, _ = o32.solve_fixed(Qd, Fds[b], K)
        n = np.abs(y64).max()
        if n == 0:
            print(f"{b:4d} ymax=0  i8 max {np.abs(Yz[b]).max()}")
            continue
        ez, ef, ezf = np.abs(Yz[b] - y64).max() / n, np.abs(y32 - y64).max() / n, np.abs(Yz[b] - y32).max() / n
        a64, a32, az = y64 > 1e-6 * n, y32 > 1e-6 * n, Yz[b] > 1e-6 * n
        worst = max(worst, ez / max(ef, 1e-12))
        print(f"{b:4d} {n:9.4g}  {ez:.2e}   {ef:.2e}   {ezf:.2e}   {np.array_equal(az, a64)} {np.array_equal(a32, a64)}")
    print("worst e(i8,f64)/e(f32,f64) =", worst)


if __name__ == "__main__":
    main()
