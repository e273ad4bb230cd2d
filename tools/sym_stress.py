"""Random sizes through the upper-triangle loop against the full-matrix loop (indexing stress): python tools/sym_stress.py [count] [seed]"""
import sys, numpy as np
sys.path.insert(0, "/root/repo")
import pqp_for_mpc_b200 as pqp

cnt = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
worst = 0.0
for t in range(cnt):
    N = int(rng.integers(2369, 9000))
    M = int(rng.integers(max(8, N // 6), N // 2))
    K = int(rng.integers(3, 40))
    prob, d = pqp.generate_testproblem(1000 + t, M, N)
    with pqp.Solver(d, prob) as s:
        Y, _, st = s.solve(iters=K)
        k1 = s.last_kernel
        Y2, _, _ = s.solve(iters=K)
    with pqp.Solver(d, prob, exploit_symmetry=0) as s:
        Yf, _, stf = s.solve(iters=K)
    e = np.abs(Y - Yf).max() / np.abs(Yf).max()
    worst = max(worst, e)
    ok = k1.startswith("gemv_sym") and np.array_equal(Y, Y2) and e <= 2e-5 and np.isfinite(Y).all()
    print(f"N={N} M={M} K={K}: {k1} err {e:.2e} Jd {st['Jd'][0]:.6g} vs {stf['Jd'][0]:.6g} {'ok' if ok else 'FAIL'}", flush=True)
    assert ok
print("worst", worst)
