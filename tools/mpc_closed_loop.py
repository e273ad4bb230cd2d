"""Receding-horizon driver: B independent closed loops x+ = A x + B u*(x) + E d on the GPU solver.

Every control period solves the B condensed QPs to tolerance with the batched tensor-core kernel, warm-started from the
previous period's duals shifted by one horizon step (pqp_shift_duals); the first nInput entries of U are applied to the plant.
usage (GPU box): python tools/mpc_closed_loop.py [B=4096] [periods=20] [pH=30 nS=12 nI=4]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pqp_for_mpc_b200 as pqp  # noqa: E402
from bench_problems import condensed_mpc  # noqa: E402


def closed_loop(solver, plant, d, prob, X0, periods, warm=True, iters=0, y_floor=1e-3):
    """Returns (X trajectory [periods+1 x B x nS], U applied [periods x B x nI], status per period).
    A loop whose QP the multiplicative update cannot solve (the 0/0 states of PQP_CPU.c:594, NaN duals) gets u = 0 for that
    period and restarts from y = y_init."""
    A, Bm, E = plant
    nI = d.nInput
    X = [X0.astype(np.float32)]
    Us, sts, Y = [], [], None
    dist = float(prob["D"][0])
    for _ in range(periods):
        Y0 = solver.shift_duals(Y, y_floor) if (warm and Y is not None) else None
        Y, U, st = solver.solve(X[-1], iters=iters, Y0=Y0, primal=True)
        bad = ~np.isfinite(U).all(axis=1)
        U[bad] = 0.0
        Y[bad] = 1000.0
        u = U[:, :nI]
        Us.append(u.copy())
        sts.append(st.copy())
        X.append((X[-1] @ A.T + u @ Bm.T + dist * E[:, 0][None, :]).astype(np.float32))
    return np.stack(X), np.stack(Us), sts


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    periods = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    pH, nS, nI = (int(v) for v in sys.argv[3:6]) if len(sys.argv) > 5 else (30, 12, 4)
    prob, d, X, plant = condensed_mpc(2024, pH, nS, nI, n_states=B, x_scale=20.0, return_plant=True)
    for warm in (False, True):
        with pqp.Solver(d, prob, batch_capacity=B, eaj=1e-2, erj=1e-6, check_every=16, max_iters=6000) as s:
            t0 = time.perf_counter()
            Xs, Us, sts = closed_loop(s, plant, d, prob, X, periods, warm=warm)
            dt = time.perf_counter() - t0
        its = np.array([st["iters"].mean() for st in sts])
        conv = np.mean([st["converged"].mean() for st in sts])
        print(f"{'warm' if warm else 'cold'} start: {periods} periods x {B} loops in {dt:.2f} s = {B * periods / dt:,.0f} closed-loop QP solves/s; "
              f"mean updates per solve by period: {np.round(its[:6]).astype(int).tolist()} ... overall {its.mean():.0f}; converged {100 * conv:.1f} %; "
              f"|x| median: {np.median(np.abs(Xs[0])):.2f} -> {np.median(np.abs(Xs[-1])):.2f}")


if __name__ == "__main__":
    main()
