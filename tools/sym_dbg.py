"""Timing experiments on the upper-triangle loop (results invalid under PQP_SYM_DBG != 0): python tools/sym_dbg.py N [dbg ...]
The switches exist only in a library built with -DPQP_SYM_DEBUG (make EXTRA=-DPQP_SYM_DEBUG); PQP_SYM_PROF=1 works in any build."""
import os, sys
sys.path.insert(0, "/root/repo")
import pqp_for_mpc_b200 as pqp
N = int(sys.argv[1]); K = 200
dbgs = [int(a) for a in sys.argv[2:]] or [0, 2, 8, 10, 11, 15]
prob, d = pqp.generate_testproblem(12346, max(N // 4, 8), N)
os.environ["PQP_SYM_PROF"] = "1"
with pqp.Solver(d, prob) as s:
    s.solve(iters=2)
    for dbg in dbgs:
        os.environ["PQP_SYM_DBG"] = str(dbg)
        print(f"dbg={dbg}", flush=True)
        s.solve(iters=K)
        print(f"   {s.last_kernel} {1e3 * s.last_solve_ms / K:.2f} us/update", flush=True)
