"""Experiment driver (GPU box): batched loop variants.  usage: python tools/batched_sweep.py pH nS nI B iters [name=value,...]..."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pqp_for_mpc_b200 as pqp
from bench_problems import condensed_mpc

pH, nS, nI, B, iters = map(int, sys.argv[1:6])
variants = sys.argv[6:] or [""]
prob, d, X = condensed_mpc(2024, pH, nS, nI, n_states=B)
ref = None
for v in variants:
    env = dict(kv.split("=") for kv in v.split(",") if kv)
    os.environ.update(env)
    with pqp.Solver(d, prob, batch_capacity=B) as s:
        ms = []
        for rep in range(3):
            Y, _, _ = s.solve(X, iters=iters, status=False)
            ms.append(s.last_solve_ms)
        best = min(ms[1:])
        if ref is None:
            ref = Y.copy()
        err = np.abs(Y - ref).max(axis=1) / np.maximum(np.abs(ref).max(axis=1), 1e-30)
        flops = 4.0 * d.N * d.N * B * iters
        print(f"{v or 'default':44s} kernel={s.last_kernel:14s} {1e3 * best / iters:8.2f} us/iter  {B / (best * 1e-3):10.0f} solves/s  "
              f"{flops / (best * 1e-3) / 1e12:7.1f} TFLOP/s(fp32-eq)  max relerr vs first = {err.max():.2e}  finite={np.isfinite(Y).all()}", flush=True)
    for k in env:
        os.environ.pop(k, None)
