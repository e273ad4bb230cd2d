"""Stopping points: the reference's terminate() (PQP_CPU.c:673-687, through the oracle restatement, float32 and its float64
twin) against the fused stop test of the CUDA loops (run to tolerance, checked after every update), on small condensed-MPC
states.  SURVEY.md 8(f)3: the comparison is statistical -- the reference's first condition (Jp > -Jd: keep going) has no
tolerance, so in float32 it fires on rounding noise in Jp + Jd; the fused test is two-sided on y'g = Jp + Jd.
usage: python tools/convergence_stats.py [n_states] [tol]"""
import sys
import numpy as np
sys.path.insert(0, "/root/repo")
import pqp_for_mpc_b200 as pqp
from bench_problems import condensed_mpc
from oracle.oracle import Oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
tol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-6
mp, md, X = condensed_mpc(11, pH=6, nS=4, nI=2, n_states=n, x_scale=25.0)
cap = 30000
rows = []
for dtype in (np.float32, np.float64):
    o = Oracle(dtype)
    Qp = o.gauss_jordan(mp["Qp_inv"])
    for b in range(n):
        Fp = o.compute_fp(mp["Fp1"], mp["Fp2"], mp["Fp3"], mp["D"], X[b])
        Mp = o.compute_mp(*[mp[k] for k in ("Mp1", "Mp2", "Mp3", "Mp4", "Mp5", "Mp6")], mp["D"], X[b])
        Qd, Fd, Md, _ = o.convert_to_dual(mp["Qp_inv"], mp["Gp"], mp["Kp"], Fp, Mp)
        Y, U, h = o.solve_converge(Qd, Fd, Md, Qp, mp["Qp_inv"], Fp, Mp, mp["Gp"], mp["Kp"], max_h=cap, tol=tol)
        rows.append((dtype.__name__, b, h - 1, U.astype(np.float64)))
ref32 = {b: (k, U) for t, b, k, U in rows if t == "float32"}
ref64 = {b: (k, U) for t, b, k, U in rows if t == "float64"}
with pqp.Solver(md, mp, eaj=tol, erj=tol, erc=tol, eac=tol, check_every=1, max_iters=cap, batch_capacity=n) as s:
    Y, U, st = s.solve(X, iters=0, primal=True)
    kern = s.last_kernel
print(f"{n} states, N={md.N}, tol={tol:g}, cap {cap}; kernel {kern}")
print(" state   ref f32   ref f64      ours  conv   |U-Uref32|inf/|U|inf")
ratios = []
for b in range(n):
    k32, U32 = ref32[b]
    k64, _ = ref64[b]
    ko, cv = int(st["iters"][b]), int(st["converged"][b])
    du = np.abs(U[b] - U32).max() / max(np.abs(U32).max(), 1e-30)
    print(f"{b:6d} {k32:9d} {k64:9d} {ko:9d} {cv:5d}   {du:.2e}")
    if cv and k32 < cap - 1:
        ratios.append(ko / max(k32, 1))
if ratios:
    r = np.array(ratios)
    print(f"ours / reference(f32) update counts over {r.size} states that both stopped: median {np.median(r):.2f}, min {r.min():.2f}, max {r.max():.2f}")
