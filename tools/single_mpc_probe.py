"""Experiment driver (GPU box): latency of ONE condensed-MPC QP (config C4's shape as a single problem, N = 4*pH*nI) through
pqp_solve_batch (B = 1): device time of the loop kernel and host-to-host time.  usage: python tools/single_mpc_probe.py [pH] [iters]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pqp_for_mpc_b200 as pqp
from bench_problems import BENCH_MIN_VIOLATED, BENCH_X_SCALE, condensed_mpc

iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
for pH in ([int(sys.argv[1])] if len(sys.argv) > 1 else [9, 16, 30, 40, 48]):
    prob, d, X = condensed_mpc(2024, pH, 12, 4, n_states=4, x_scale=BENCH_X_SCALE, min_violated=BENCH_MIN_VIOLATED)
    res = {}
    for cl in ("1", "0"):
        os.environ["PQP_GEMV_CLUSTER"] = cl
        with pqp.Solver(d, prob, batch_capacity=1) as s:
            x = X[:1]
            for _ in range(3):
                Y, U, st = s.solve(x, iters=iters, primal=True)
            ms, wall = [], []
            for _ in range(10):
                t0 = time.perf_counter()
                Y, U, st = s.solve(x, iters=iters, primal=True)
                wall.append(time.perf_counter() - t0)
                ms.append(s.last_solve_ms)
            Yt, Ut, stt = s.solve(x, iters=0, primal=True)
            res[cl] = (Y.copy(), U.copy(), st.copy(), Yt.copy(), stt.copy())
            print(f"N={d.N:4d} M={d.M:3d} PQP_GEMV_CLUSTER={cl}: kernel {s.last_kernel:28s} {1e3 * min(ms) / iters:.3f} us/update "
                  f"({min(ms):.3f} ms per {iters}-update solve), host to host {1e3 * min(wall):.3f} ms; to tolerance: {int(stt['iters'][0])} updates, "
                  f"converged {int(stt['converged'][0])}", flush=True)
    a, b = res["1"], res["0"]
    sc = max(np.abs(b[0]).max(), 1e-30)
    print(f"        cluster vs multi-CTA kernel: max|dY|/max|Y| = {np.abs(a[0] - b[0]).max() / sc:.2e}, max|dU|/max|U| = "
          f"{np.abs(a[1] - b[1]).max() / max(np.abs(b[1]).max(), 1e-30):.2e}; status gap {float(a[2]['gap'][0]):.4e} / {float(b[2]['gap'][0]):.4e}, "
          f"Jd {float(a[2]['Jd'][0]):.6e} / {float(b[2]['Jd'][0]):.6e}, kkt {float(a[2]['kkt'][0]):.3e} / {float(b[2]['kkt'][0]):.3e}; tolerance runs: "
          f"max|dY|/max|Y| = {np.abs(a[3] - b[3]).max() / max(np.abs(b[3]).max(), 1e-30):.2e}", flush=True)
