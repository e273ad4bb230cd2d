import sys, numpy as np
sys.path.insert(0, "/root/repo")
import pqp_for_mpc_b200 as pqp
prob, d = pqp.generate_testproblem(12345, 2048, 1024)
for ce in (1, 8, 64, 1000):
    with pqp.Solver(d, prob, eaj=1e-30, erj=1e-30, check_every=ce, max_iters=20000) as s:
        Y, _, st = s.solve(iters=0)
        print(f"check_every={ce:5d}: {s.last_kernel} {1e3*s.last_solve_ms/20000:.2f} us/update iters {st['iters'][0]}")
