"""Stress driver (GPU box) for the one-cluster kernel (pqp_gemv_cluster.cu): hundreds of short solves, fixed count and run to
tolerance with checks every 8 / every update, each mode in its own process (a protocol bug traps and takes the context with it).
usage: python tools/cluster_stress.py"""
import os, sys, numpy as np, subprocess
sys.path.insert(0, "/root/repo")
if len(sys.argv) < 2:
    for mode in ("fixed", "tol8", "tol1", "fixed", "tol8"):
        r = subprocess.run([sys.executable, __file__, mode], capture_output=True, text=True)
        print(mode, (r.stdout.strip().splitlines() or ["-"])[-1], "|", r.stderr.strip()[-120:].replace("\n", " "), flush=True)
    sys.exit(0)
import pqp_for_mpc_b200 as pqp
from bench_problems import condensed_mpc, BENCH_X_SCALE, BENCH_MIN_VIOLATED
mode = sys.argv[1]
prob, d, X = condensed_mpc(2024, 9, 12, 4, n_states=4, x_scale=BENCH_X_SCALE, min_violated=BENCH_MIN_VIOLATED)
n = 0
for rep in range(8):
    for cnt in range(3, 43):
        if mode == "fixed":
            with pqp.Solver(d, prob, batch_capacity=1) as s:
                Y, U, st = s.solve(X[:1], iters=cnt, primal=True)
        else:
            with pqp.Solver(d, prob, batch_capacity=1, max_iters=cnt, check_every=int(mode[3:])) as s:
                Y, U, st = s.solve(X[:1], iters=0, primal=True)
        n += 1
        print(f"ok {n} solves, last count {cnt}", flush=True)
