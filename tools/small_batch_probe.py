"""Experiment driver (GPU box): latency of SMALL batches of condensed-MPC problems (config C4's shape, B = 1 .. 128) through
pqp_solve_batch_primal: which kernel serves them and how long 1000 updates take.  usage: python tools/small_batch_probe.py [iters]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pqp_for_mpc_b200 as pqp
from bench_problems import BENCH_MIN_VIOLATED, BENCH_X_SCALE, condensed_mpc

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
prob, d, X = condensed_mpc(2024, 30, 12, 4, n_states=128, x_scale=BENCH_X_SCALE, min_violated=BENCH_MIN_VIOLATED)
ref = None
for B in (1, 2, 4, 8, 9, 16, 18, 32, 36, 64, 128):
    with pqp.Solver(d, prob, batch_capacity=B) as s:
        for _ in range(2):
            Y, U, st = s.solve(X[:B], iters=iters, primal=True)
        ms, wall = [], []
        for _ in range(5):
            t0 = time.perf_counter()
            Y, U, st = s.solve(X[:B], iters=iters, primal=True)
            wall.append(time.perf_counter() - t0)
            ms.append(s.last_solve_ms)
        if ref is None:
            ref = Y[0].copy()
        print(f"B={B:4d}: kernel {s.last_kernel:24s} {min(ms):8.3f} ms per {iters}-update batch ({B / (min(ms) * 1e-3):9.0f} solves/s), host to host "
              f"{1e3 * min(wall):.3f} ms; problem 0 vs B=1: {np.abs(Y[0] - ref).max() / np.abs(ref).max():.1e}", flush=True)
