"""GPU box: the int8 digit-plane batched kernel against its numpy model (bit for bit) and the float64 twin.
usage: python tools/imma_check.py pH nS nI B iters [name=value,...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pqp_for_mpc_b200 as pqp  # noqa: E402
from bench_problems import condensed_mpc  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
import ozaki_emulate as oz  # noqa: E402

pH, nS, nI, B, iters = map(int, sys.argv[1:6])
env = dict(kv.split("=") for kv in (sys.argv[6].split(",") if len(sys.argv) > 6 else []) if kv)
os.environ.update(env)
os.environ.setdefault("PQP_BATCHED", "imma")
prob, d, X = condensed_mpc(2024, pH, nS, nI, n_states=B)
with pqp.Solver(d, prob, batch_capacity=B) as s:
    Y, _, _ = s.solve(X, iters=iters, status=False)
    print("kernel", s.last_kernel, "ms", s.last_solve_ms, "N", d.N, "B", B, "iters", iters, flush=True)
    Qd, th, _ = s.dual()
    Fd, _ = s.linear_terms(B, want_fp=False)
nchk = min(B, 48)
Ye = oz.run(Qd, th, Fd[:nchk], iters)
same = np.array_equal(Y[:nchk], Ye, equal_nan=True)
diff = np.abs(Y[:nchk] - Ye)
print("bit-identical to the numpy model:", same, " max abs diff", np.nanmax(diff), " first mismatch", np.argwhere(~(diff == 0))[:3].tolist())
o32, o64 = Oracle(np.float32), Oracle(np.float64)
worst = 0.0
for b in range(0, nchk, max(1, nchk // 12)):
    y64, _ = o64.solve_fixed(Qd, Fd[b], iters)
    y32, _ = o32.solve_fixed(Qd, Fd[b], iters)
    n = np.abs(y64).max()
    if n == 0:
        continue
    eg, ef = np.abs(Y[b] - y64).max() / n, np.abs(y32 - y64).max() / n
    worst = max(worst, eg / max(ef, 1e-12))
    print(f"  problem {b:3d} ymax {n:9.4g}  e(gpu,f64) {eg:.2e}  e(f32,f64) {ef:.2e}  e(gpu,f32) {np.abs(Y[b]-y32).max()/n:.2e}  "
          f"active set same: {np.array_equal(Y[b] > 1e-6*n, y64 > 1e-6*n)}")
print("worst e(gpu,f64)/e(f32,f64):", worst)
