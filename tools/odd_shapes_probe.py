import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
import pqp_for_mpc_b200 as pqp
from oracle.oracle import Oracle
o32, o64 = Oracle(np.float32), Oracle(np.float64)
rng = np.random.default_rng(9)
for N, B, K in ((8, 5, 30), (16, 33, 30), (129, 70, 40), (512, 65, 20), (520, 3, 20), (600, 40, 15), (1030, 2, 10)):
    A = rng.standard_normal((N, max(1, N // 2))).astype(np.float32)
    Qd = (A @ A.T).astype(np.float32)
    Fd = rng.uniform(-50, 50, (B, N)).astype(np.float32)
    with pqp.Solver(Qd=Qd, batch_capacity=B) as s:
        Y, _, st = s.solve(Fd=Fd, iters=K)
        k = s.last_kernel
        worst = 0
        for b in (0, B - 1):
            y32, _ = o32.solve_fixed(Qd, Fd[b], K); y64, _ = o64.solve_fixed(Qd, Fd[b], K)
            e = np.abs(Y[b] - y64).max() / np.abs(y64).max(); ef = np.abs(y32 - y64).max() / np.abs(y64).max()
            worst = max(worst, e / max(ef, 1e-7))
        Yt, _, stt = s.solve(Fd=Fd[:2], iters=0)
        print(f"N={N:5d} B={B:3d} kernel={k:20s} worst err ratio vs oracle noise {worst:.2f}  tol-mode kernel={s.last_kernel} iters={stt['iters'].tolist()} conv={stt['converged'].tolist()}", flush=True)
