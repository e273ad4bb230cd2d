"""GPU box: which problems of the C4 batch end non-finite, per engine."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pqp_for_mpc_b200 as pqp
from bench_problems import condensed_mpc
prob, d, X = condensed_mpc(2024, 30, 12, 4, n_states=4096)
for eng in ("simt", "imma"):
    os.environ["PQP_BATCHED"] = eng
    with pqp.Solver(d, prob, batch_capacity=4096) as s:
        Y, _, _ = s.solve(X, iters=1000, status=False)
        bad = np.where(~np.isfinite(Y).all(axis=1))[0]
        print(eng, s.last_kernel, "non-finite problems:", bad.tolist()[:40], "count", len(bad))
        for b in bad[:3]:
            print("  problem", b, "nan count", np.isnan(Y[b]).sum(), "inf", np.isinf(Y[b]).sum(), "finite max", np.nanmax(np.where(np.isfinite(Y[b]), Y[b], 0)))
        allzero = np.where((Y == 0).all(axis=1))[0]
        print("  all-zero problems:", len(allzero))
