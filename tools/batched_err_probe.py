import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pqp_for_mpc_b200 as pqp
from bench_problems import condensed_mpc
from oracle.oracle import Oracle
K = int(sys.argv[1]) if len(sys.argv) > 1 else 100
B = 64
prob, d, X = condensed_mpc(2024, 30, 12, 4, n_states=B)
o32, o64 = Oracle(np.float32), Oracle(np.float64)
res = {}
for name, env in (("simt", {"PQP_BATCHED_UMMA": "0"}), ("umma", {"PQP_UMMA_CLUSTER": "1"})):
    os.environ.update(env)
    with pqp.Solver(d, prob, batch_capacity=B) as s:
        res[name], _, _ = s.solve(X, iters=K, status=False)
        Qd, th, _ = s.dual()
        Fd, _ = s.linear_terms(B, want_fp=False)
    for k in env: os.environ.pop(k)
print("problem  ymax      e(simt,f64) e(umma,f64) e(f32,f64)  worst-component info (umma)")
for b in range(0, B, 4):
    y64, _ = o64.solve_fixed(Qd, Fd[b], K)
    y32, _ = o32.solve_fixed(Qd, Fd[b], K)
    n = np.abs(y64).max()
    es, eu, ef = (np.abs(res["simt"][b] - y64).max() / n, np.abs(res["umma"][b] - y64).max() / n, np.abs(y32 - y64).max() / n)
    j = int(np.argmax(np.abs(res["umma"][b] - y64)))
    print(f"{b:4d} {n:10.4g}  {es:.2e}   {eu:.2e}   {ef:.2e}   comp {j}: y64={y64[j]:.6g} umma={res['umma'][b][j]:.6g} signed rel {(res['umma'][b][j]-y64[j])/y64[j]:+.2e}")
