"""Stress driver (GPU box) for the one-cluster kernel (pqp_gemv_cluster.cu): hundreds of short solves over its shape classes, single
problems and small batches, fixed count and run to tolerance with checks every 8 / every update, each mode in its own process (a protocol
bug traps and takes the context with it).  Every solve is compared with the multi-CTA / tensor-core path on the same input.
usage: python tools/cluster_stress.py"""
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) < 2:
    for mode in ("fixed", "tol8", "tol1"):
        r = subprocess.run([sys.executable, __file__, mode], capture_output=True, text=True)
        print(mode, (r.stdout.strip().splitlines() or ["-"])[-1], "|", r.stderr.strip()[-160:].replace("\n", " "), flush=True)
    sys.exit(0)
import pqp_for_mpc_b200 as pqp  # noqa: E402
from bench_problems import condensed_mpc  # noqa: E402

mode = sys.argv[1]
n = 0
worst = 0.0
for pH, nI in ((17, 1), (25, 1), (9, 4), (19, 4), (30, 4), (32, 4)):   # N = 68, 100, 144, 304, 480, 512
    prob, d, X = condensed_mpc(2024 + pH, pH, 6, nI, n_states=16, x_scale=30.0)
    for B in (1, 3, 16):
        for cnt in list(range(3, 23)) + [40, 97]:
            out = {}
            for cl in ("1", "0"):
                os.environ["PQP_GEMV_CLUSTER"] = cl
                kw = {} if mode == "fixed" else dict(max_iters=cnt, check_every=int(mode[3:]), eaj=1e-30, erj=1e-30)
                with pqp.Solver(d, prob, batch_capacity=B, **kw) as s:
                    Y, U, st = s.solve(X[:B], iters=cnt if mode == "fixed" else 0, primal=True)
                    out[cl] = (Y, s.last_kernel)
                    assert int(st["iters"][0]) == cnt
            assert out["1"][1].startswith("gemv_cluster"), out["1"][1]
            fin = np.isfinite(out["0"][0]).all(1)
            err = np.abs(out["1"][0][fin] - out["0"][0][fin]).max() / max(np.abs(out["0"][0][fin]).max(), 1e-30)
            worst = max(worst, err)
            assert err <= 1e-4, (d.N, B, cnt, err)
            n += 1
            print(f"ok {n} solves (N={d.N} B={B} count {cnt}), worst difference to the other path {worst:.1e}", flush=True)
