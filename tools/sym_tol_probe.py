"""Run-to-tolerance in the upper-triangle loop: agreement with the fixed-count solve at the reported count, time per update.
usage: python tools/sym_tol_probe.py [N ...]"""
import os, sys, numpy as np
sys.path.insert(0, "/root/repo")
import pqp_for_mpc_b200 as pqp

for N in [int(a) for a in sys.argv[1:]] or [2560, 4096, 8192]:
    prob, d = pqp.generate_testproblem(12346, 2 * N, N)     # M = 2N: well conditioned, converges
    for ce in (8, 64):
        with pqp.Solver(d, prob, eaj=1e30, erj=1e-4, check_every=ce, max_iters=3000) as s:
            s.solve(iters=2)                                  # one-time symmetry test / unit array off the clock
            Y, _, st = s.solve(iters=0)
            k, ms, it = s.last_kernel, s.last_solve_ms, int(st["iters"][0])
            Yf, _, stf = s.solve(iters=max(it, 1))
            same = bool(np.array_equal(Y, Yf))
            print(f"N={N} check_every={ce}: {k} converged={int(st['converged'][0])} after {it} updates, {1e3 * ms / max(it, 1):.2f} us/update; "
                  f"fixed-count at that count: {s.last_kernel} {1e3 * s.last_solve_ms / max(it, 1):.2f} us/update, bit-identical {same}; "
                  f"gap {st['gap'][0]:.4g} Jd {st['Jd'][0]:.6g} (fixed: {stf['gap'][0]:.4g} {stf['Jd'][0]:.6g})", flush=True)
    os.environ["PQP_GEMV_SYM_TOL"] = "0"
    with pqp.Solver(d, prob, eaj=1e30, erj=1e-4, check_every=8, max_iters=3000) as s:
        s.solve(iters=2)
        Y0, _, st0 = s.solve(iters=0)
        print(f"N={N} full-matrix tolerance kernel: {s.last_kernel} {int(st0['iters'][0])} updates, {1e3 * s.last_solve_ms / max(int(st0['iters'][0]), 1):.2f} us/update", flush=True)
    del os.environ["PQP_GEMV_SYM_TOL"]
