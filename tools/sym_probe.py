"""Upper-triangle loop (pqp_gemv_sym.cu) against the full-matrix TMA loop: agreement, reproducibility, time per update.
usage: python tools/sym_probe.py [N ...]   (env PQP_SYM_ITERS, default 200)"""
import os, sys, numpy as np
sys.path.insert(0, "/root/repo")
import pqp_for_mpc_b200 as pqp

Ns = [int(a) for a in sys.argv[1:]] or [2560, 3001, 4096, 8192]
K = int(os.environ.get("PQP_SYM_ITERS", "200"))
for N in Ns:
    M = max(N // 4, 8)
    prob, d = pqp.generate_testproblem(12346, M, N)
    res = {}
    for mode in ("1", "0"):
        os.environ["PQP_GEMV_SYM"] = mode
        with pqp.Solver(d, prob) as s:
            Y, _, st = s.solve(iters=K)
            Y2, _, _ = s.solve(iters=K)
            ms = s.last_solve_ms
            res[mode] = (Y.copy(), s.last_kernel, 1e3 * ms / K, bool(np.array_equal(Y, Y2)), st)
    a, b = res["1"], res["0"]
    err = np.linalg.norm(a[0] - b[0]) / max(np.linalg.norm(b[0]), 1e-30)
    print(f"RESULT N={N}: {a[1]} {a[2]:.2f} us/update (repro {a[3]}) | {b[1]} {b[2]:.2f} us/update | normwise diff {err:.2e} "
          f"gap {a[4]['gap'][0]:.4g} vs {b[4]['gap'][0]:.4g} nan {int(np.isnan(a[0]).sum())}", flush=True)
