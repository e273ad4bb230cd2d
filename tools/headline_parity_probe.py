"""Experiment driver (GPU box): parity numbers at the headline sizes, printed so that tests/test_headline_parity_gpu.py can state
its bounds from measurements.  usage: python tools/headline_parity_probe.py [wc|c3|c4|c2]..."""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pqp_for_mpc_b200 as pqp
from bench_problems import BENCH_MIN_VIOLATED, BENCH_X_SCALE, condensed_mpc
from oracle.oracle import Oracle

o32, o64 = Oracle(np.float32), Oracle(np.float64)


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def act(y, rel=1e-6):
    y = np.asarray(y, np.float64)
    return y > rel * np.abs(y).max()


def both(Qd, Fd, K):
    with ThreadPoolExecutor(2) as ex:
        f32, f64 = ex.submit(o32.solve_fixed, Qd, Fd, K), ex.submit(o64.solve_fixed, Qd, Fd, K)
        return f32.result()[0], f64.result()[0]


def single(tag, seed, M, N, K):
    t0 = time.time()
    prob, d = pqp.generate_testproblem(seed, M, N)
    with pqp.Solver(d, prob, order=pqp.ORDER_STRICT) as s:
        s.solve(iters=1, status=False)
        Qd, th, _ = s.dual(want_gq=False)
        Fd, _ = s.linear_terms(1)
    print(tag, "strict setup", time.time() - t0, flush=True)
    t0 = time.time()
    y32, y64 = both(Qd, Fd[0], K)
    print(tag, "oracles", time.time() - t0, "err(f32,f64)", relerr(y32, y64), "active", int(act(y32).sum()), "of", N, flush=True)
    for sym in (1, 0):
        with pqp.Solver(d, prob, exploit_symmetry=sym) as s:
            Y, U, st = s.solve(iters=K, primal=True)
            Qf, thf, _ = s.dual(want_gq=False)
            print(tag, s.last_kernel, "Qd err", relerr(Qf, Qd), "err(gpu,f32)", relerr(Y[0], y32), "err(gpu,f64)", relerr(Y[0], y64),
                  "active same(1e-6)", bool(np.array_equal(act(Y[0]), act(y32))), "(1e-5)", bool(np.array_equal(act(Y[0], 1e-5), act(y32, 1e-5))),
                  "(1e-4)", bool(np.array_equal(act(Y[0], 1e-4), act(y32, 1e-4))), flush=True)
        # the loop alone, fed the reference-order dual
        with pqp.Solver(Qd=Qd, exploit_symmetry=sym) as s:
            Y, _, _ = s.solve(Fd=Fd[0], iters=K)
            print(tag, "fed strict Qd:", s.last_kernel, "err(gpu,f32)", relerr(Y[0], y32), "err(gpu,f64)", relerr(Y[0], y64), flush=True)


def c4():
    B, K = 256, 1000
    prob, d, X = condensed_mpc(2024, 30, 12, 4, n_states=B, x_scale=BENCH_X_SCALE, min_violated=BENCH_MIN_VIOLATED)
    with pqp.Solver(d, prob, batch_capacity=B) as s:
        Y, U, st = s.solve(X, iters=K, primal=True)
        Qd, th, _ = s.dual()
        Fd, Fp = s.linear_terms(B)
        print("c4 kernel", s.last_kernel)
    idx = np.random.default_rng(64).choice(B, 64, replace=False)
    with ThreadPoolExecutor(16) as ex:
        r32 = list(ex.map(lambda b: o32.solve_fixed(Qd, Fd[b], K)[0], idx))
        r64 = list(ex.map(lambda b: o64.solve_fixed(Qd, Fd[b], K)[0], idx))
    rows = []
    for b, y32, y64 in zip(idx, r32, r64):
        rows.append((int(b), relerr(Y[b], y32), relerr(Y[b], y64), relerr(y32, y64), bool(np.array_equal(act(Y[b], 1e-5), act(y32, 1e-5))), int(act(y32, 1e-5).sum())))
    rows.sort(key=lambda r: -r[2])
    for r in rows[:12]:
        print("c4 state %4d err(gpu,f32) %.2e err(gpu,f64) %.2e err(f32,f64) %.2e active same %s active %d" % r)
    a = np.array([[r[1], r[2], r[3]] for r in rows])
    print("c4 worst:", a.max(0), "median:", np.median(a, 0), "ratio gd/fd worst", (a[:, 1] / a[:, 2]).max(), "all active sets same", all(r[4] for r in rows),
          "nan", int((~np.isfinite(Y).all(1)).sum()), "zero", int((Y == 0).all(1).sum()))


for w in sys.argv[1:] or ["c2", "c4", "c3", "wc"]:
    if w == "wc":
        single("wc N=8192 M=8192", 4242, 8192, 8192, 100)
    elif w == "c3":
        single("c3 N=8192 M=2048", 12346, 2048, 8192, 60)
    elif w == "c2":
        single("c2 N=1024 M=512", 12345, 512, 1024, 1000)
    elif w == "c4":
        c4()
