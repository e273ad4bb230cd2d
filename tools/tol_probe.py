"""GPU box: run-to-tolerance in the batched int8 kernel.  usage: python tools/tol_probe.py pH nS nI B eaj erj [check_every] [max_iters]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pqp_for_mpc_b200 as pqp
from bench_problems import condensed_mpc
pH, nS, nI, B = map(int, sys.argv[1:5])
eaj, erj = float(sys.argv[5]), float(sys.argv[6])
ce = int(sys.argv[7]) if len(sys.argv) > 7 else 8
mi = int(sys.argv[8]) if len(sys.argv) > 8 else 20000
prob, d, X = condensed_mpc(2024, pH, nS, nI, n_states=B)
with pqp.Solver(d, prob, batch_capacity=B, eaj=eaj, erj=erj, check_every=ce, max_iters=mi) as s:
    t0 = time.time()
    Y, U, st = s.solve(X, iters=0, primal=True)
    dt = time.time() - t0
    print("kernel", s.last_kernel, "solve ms", s.last_solve_ms, "wall", dt)
    it = st["iters"]
    print("converged", int(st["converged"].sum()), "of", B, " iters: min", it.min(), "median", int(np.median(it)), "max", it.max())
    print("gap |max| among converged", np.abs(st["gap"][st["converged"] == 1]).max() if st["converged"].any() else None,
          " min_slack min", st["min_slack"][st["converged"] == 1].min() if st["converged"].any() else None)
    # tolerance mode == fixed mode stopped at iters_b, bit for bit
    for b in (0, B // 2, B - 1):
        Yf, _, stf = s.solve(X, iters=int(it[b]))
        print(" problem", b, "iters", it[b], "conv", st["converged"][b], "bit-identical to fixed-count:", np.array_equal(Yf[b], Y[b], equal_nan=True),
              "| gap", st["gap"][b], "vs status kernel", stf["gap"][b], "| Jd", st["Jd"][b], stf["Jd"][b], "| min_slack", st["min_slack"][b], stf["min_slack"][b],
              "| kkt", st["kkt"][b], stf["kkt"][b])
    # the same problems one at a time through the single-problem kernel
    for b in (0, B - 1):
        Y1, _, st1 = s.solve(X[b][None], iters=0)
        print(" single-problem kernel:", s.last_kernel, "iters", st1["iters"][0], "conv", st1["converged"][0], "relerr y", np.abs(Y1[0] - Y[b]).max() / max(np.abs(Y[b]).max(), 1e-30))
