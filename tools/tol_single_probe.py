import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import pqp_for_mpc_b200 as pqp
for N, M in ((1024, 2048), (2048, 4096), (480, 0)):
    if M:
        prob, d = pqp.generate_testproblem(12345, M, N)
        with pqp.Solver(d, prob, eaj=1e-2, erj=1e-6, check_every=8, max_iters=200000) as s:
            Y, U, st = s.solve(iters=0, primal=True)
            ms_tol, k_tol, it = s.last_solve_ms, s.last_kernel, int(st["iters"][0])
            Yf, _, _ = s.solve(iters=max(it, 1))
            print(f"N={N}: tolerance mode {k_tol}: {it} updates in {ms_tol:.3f} ms = {1e3*ms_tol/max(it,1):.2f} us/update (converged {st['converged'][0]}); fixed-count {s.last_kernel}: {1e3*s.last_solve_ms/max(it,1):.2f} us/update; same y: {np.abs(Y-Yf).max()/np.abs(Yf).max():.2e}")
