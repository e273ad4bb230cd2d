"""Experiment driver (GPU box): per-iteration time of the single-problem loop under env toggles.
usage: python tools/gemv_sweep.py N M iters [name=value ...]   (each name=value run is one Solver)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pqp_for_mpc_b200 as pqp

N, M, iters = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
variants = sys.argv[4:] or [""]
rng = np.random.default_rng(0)
# a dual straight from numpy: the sweep measures the loop, not the setup
G = rng.integers(0, 3, (N, M)).astype(np.float32)
G[G == 2] = -1
q = rng.uniform(0, 100, M).astype(np.float32)
Qd = (G * q) @ G.T
Fd = (rng.uniform(0, 100, N) + (G * q) @ rng.uniform(0, 100, M)).astype(np.float32)
ldq = (N + 31) // 32 * 32
bytes_iter = 4.0 * N * ldq + 16.0 * N
ref = None
for v in variants:
    env = dict(kv.split("=") for kv in v.split(",") if kv)
    for k, val in env.items():
        os.environ[k] = val
    opts = {}
    if "l2" in env:
        opts["l2_persist"] = int(env["l2"])
    with pqp.Solver(Qd=Qd, **opts) as s:
        ms = []
        for rep in range(4):
            Y, _, st = s.solve(Fd=Fd, iters=iters, status=False)
            ms.append(s.last_solve_ms)
        best = min(ms[1:])
        us = 1e3 * best / iters
        if ref is None:
            ref = Y.copy()
        same = bool(np.array_equal(ref, Y))
        print(f"{v or 'default':40s} kernel={s.last_kernel:28s} {us:8.2f} us/iter  {bytes_iter / us / 1e3:8.1f} GB/s  "
              f"same_result={same}", flush=True)
    for k in env:
        os.environ.pop(k, None)
