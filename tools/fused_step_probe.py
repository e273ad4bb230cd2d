"""Experiment driver (GPU box): a whole pqp_solve_batch_primal step (device-resident X, Y, U) with the refresh / recovery inside the
loop kernel against the separate kernels.  usage: python tools/fused_step_probe.py [B] [iters]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pqp_for_mpc_b200 as pqp
from bench_problems import BENCH_MIN_VIOLATED, BENCH_X_SCALE, condensed_mpc

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
prob, d, X = condensed_mpc(2024, 30, 12, 4, n_states=B, x_scale=BENCH_X_SCALE, min_violated=BENCH_MIN_VIOLATED)
Xd = torch.from_numpy(X).cuda()
Yd = torch.empty((B, d.N), dtype=torch.float32, device="cuda")
Ud = torch.empty((B, d.M), dtype=torch.float32, device="cuda")
res = {}
for fuse in ("1", "0"):
    os.environ["PQP_IMMA_FUSE"] = fuse
    with pqp.Solver(d, prob, batch_capacity=B) as s:
        def step():
            rc = pqp.lib().pqp_solve_batch_primal(s.handle, pqp._as_ptr(Xd.data_ptr()), None, B, iters, None, pqp._as_ptr(Yd.data_ptr()),
                                                  pqp._as_ptr(Ud.data_ptr()), None)
            assert rc == 0
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        l0 = s.launch_count
        t0 = time.perf_counter()
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        res[fuse] = (Yd.cpu().numpy().copy(), Ud.cpu().numpy().copy())
        print(f"PQP_IMMA_FUSE={fuse}: {1e3 * dt:.3f} ms per step ({B / dt:.0f} solves/s), loop kernel {s.last_solve_ms:.3f} ms, "
              f"{(s.launch_count - l0) // 5} launches per step, kernel {s.last_kernel}", flush=True)
print("Y identical:", np.array_equal(res["1"][0], res["0"][0], equal_nan=True), " U identical:", np.array_equal(res["1"][1], res["0"][1], equal_nan=True))
