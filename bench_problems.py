"""Synthetic condensed linear-MPC instances for configs C4/C5 (bench and test INPUT synthesis only).

The reference ships exactly one instance (example/, pHorizon=1) and no MPC generator; SURVEY 7 "Hard parts"
asks the bench to synthesise the condensed matrices and document how.  This is that recipe, in the
reference's own conventions (PQP_CPU.c:5-6, 373-382, 940-941):

    x_{k+1} = A x_k + B u_k + E d_k,   y_k = C x_k           (nState, nInput, nOutput = nInput, nDis = 1)
    stacked over the horizon:  Xs = Phi x + Gam U + Gd D
    cost   1/2 U'Qp U + Fp(x)'U + 1/2 Mp,   Qp = Gam'Qb Gam + Rb,  Fp(x) = Fp1 D + Fp2 x - Fp3
           Fp2 = Gam'Qb Phi,  Fp1 = Gam'Qb Gd,  Fp3 = Gam'Qb r (tracking offset)
    constraints  Gp U <= Kp,  Gp = [I; -I; C Gam; -C Gam]   (N = 4*pHorizon*nInput rows, M = pHorizon*nInput)

Bounds umax = 20 (the example's Kp, example/Kp.txt), ymax = 30 and states x ~ N(0, 60^2) put the instance in the
regime of the shipped example: the dual starts at 1000 (PQP_CPU.c:710), theta sits at its floor 5, and 1000
updates bring the KKT residual to ~1e-3 with a few percent of the 480 constraints active (11-60 depending on x).

Everything is built in float64 with a seeded numpy Generator and rounded once to float32.
"""
from __future__ import annotations

import numpy as np


# The bench instance of configs C4 / C5 (bench.py): states wide enough that EVERY problem has active constraints.  With the
# x_scale = 60 the round-1 bench used, 142 of 4096 states had no violated constraint at all (the solve is trivially y -> 0) and 13
# ended in 0/0 -- the reference's own fp32 behaviour when the only violated constraint's dual underflows, PQP_CPU.c:594 -- i.e.
# 3.8 % of the timed work was degenerate.  x_scale = 150 with at least 4 rows of Fd(x) below -1 per state: measured on the 4096
# states of seed 2024 after 1000 updates (float32 numpy model): 0 NaN, 0 all-zero, 38 of 480 constraints active in the median
# (8 %), every state still feasible (the float64 iteration reaches a KKT residual of 1e-12).  The degenerate states stay covered
# by tests/test_imma_gpu.py, which keeps x_scale = 60.
BENCH_X_SCALE = 150.0
BENCH_MIN_VIOLATED = 4


def condensed_mpc(seed: int, pH: int = 30, nS: int = 12, nI: int = 4, n_states: int = 4096, x_scale: float = 60.0,
                  umax: float = 20.0, ymax: float = 30.0, return_plant: bool = False, min_violated: int = 0):
    """Returns (problem dict for pqp.Solver, pqp Dims, X [n_states x nS] float32) and, with return_plant, the plant
    (A, B, E) the condensed matrices were built from (for closed-loop drivers: x+ = A x + B u + E d).
    min_violated > 0: a state is redrawn (same seeded stream) until at least that many rows of Fd(x) = GQ Fp(x) + Kp are
    below -1, i.e. until the unconstrained optimum violates that many constraints clearly -- no trivially-zero problems."""
    import pqp_for_mpc_b200 as pqp

    rng = np.random.default_rng(seed)
    A = rng.standard_normal((nS, nS))
    A *= 0.95 / np.abs(np.linalg.eigvals(A)).max()  # stable, spectral radius 0.95
    Bm = rng.standard_normal((nS, nI))
    Cm = rng.standard_normal((nI, nS)) / np.sqrt(nS)
    E = rng.standard_normal((nS, 1)) * 0.1
    Qb1 = np.eye(nS) + 0.1 * (lambda W: W @ W.T)(rng.standard_normal((nS, nS))) / nS
    Rb1 = 0.1 * np.eye(nI)

    M, N = pH * nI, 4 * pH * nI
    Phi = np.zeros((pH * nS, nS))
    Gam = np.zeros((pH * nS, M))
    Gd = np.zeros((pH * nS, pH))
    Ak = np.eye(nS)
    pows = [np.eye(nS)]
    for k in range(pH):
        pows.append(pows[-1] @ A)
    for k in range(pH):
        Phi[k * nS:(k + 1) * nS] = pows[k + 1]
        for j in range(k + 1):
            Gam[k * nS:(k + 1) * nS, j * nI:(j + 1) * nI] = pows[k - j] @ Bm
            Gd[k * nS:(k + 1) * nS, j:j + 1] = pows[k - j] @ E
    Qb = np.kron(np.eye(pH), Qb1)
    Rb = np.kron(np.eye(pH), Rb1)
    Cs = np.kron(np.eye(pH), Cm)

    Qp = Gam.T @ Qb @ Gam + Rb
    Qp = 0.5 * (Qp + Qp.T)
    Qp_inv = np.linalg.inv(Qp)
    r = rng.standard_normal(pH * nS) * 0.5
    Fp2 = Gam.T @ Qb @ Phi
    Fp1 = Gam.T @ Qb @ Gd
    Fp3 = Gam.T @ Qb @ r
    D = rng.standard_normal(pH) * 0.1
    CG = Cs @ Gam
    Gp = np.vstack([np.eye(M), -np.eye(M), CG, -CG])
    Kp = np.concatenate([np.full(2 * M, umax), np.full(2 * M, ymax)])

    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    prob = dict(Qp_inv=f32(Qp_inv), Gp=f32(Gp), Kp=f32(Kp), Fp1=f32(Fp1), Fp2=f32(Fp2), Fp3=f32(Fp3), D=f32(D),
                Mp1=f32(Phi.T @ Qb @ Phi), Mp2=f32(Gd.T @ Qb @ Phi), Mp3=f32(Gd.T @ Qb @ Gd),
                Mp4=f32(-2 * Phi.T @ Qb @ r), Mp5=f32(-2 * Gd.T @ Qb @ r), Mp6=f32([r @ Qb @ r]))
    d = pqp.Dims()   # pqp_dims_mpc (PQP_CPU.c:940-941) spelled out, so that input synthesis never loads the product library
    d.pHorizon, d.nState, d.nInput, d.nOutput, d.nDis = pH, nS, nI, nI, 1
    d.M, d.N, d.nDisH = pH * nI, 4 * pH * nI, pH
    X = rng.standard_normal((n_states, nS)) * x_scale
    if min_violated > 0:
        GQ = Gp @ Qp_inv
        Sx, s0 = GQ @ Fp2, GQ @ (Fp1 @ D - Fp3) + Kp          # Fd(x) = Sx x + s0
        CH = 65536                                              # fixed chunking: the redraw order is part of the recipe
        for c0 in range(0, n_states, CH):
            Xc = X[c0:c0 + CH]
            for _ in range(64):
                bad = np.nonzero(((Xc @ Sx.T + s0) < -1.0).sum(axis=1) < min_violated)[0]
                if bad.size == 0:
                    break
                Xc[bad] = rng.standard_normal((bad.size, nS)) * x_scale
            else:
                raise ValueError("condensed_mpc: could not draw states with enough violated constraints")
    X = f32(X)
    prob["x"] = X[0].copy()
    if return_plant:
        return prob, d, X, (f32(A), f32(Bm), f32(E))
    return prob, d, X


def shard_range(total: int, world: int, rank: int):
    """Contiguous shard [lo, hi) of `total` independent problems for `rank` of `world`."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def generator_instance(seed: int, M: int, N: int):
    """numpy port of pqp_generate_testproblem (csrc/pqp_io.c; the distribution of testing/test_generator.c:936-987 from a
    seeded splitmix64 stream), bit for bit (tests/test_io_abi.py).  It exists so that `bench.py --impl reference` can build
    the instance the product arm times WITHOUT loading libpqp_b200.so.  Returns (problem dict, Dims)."""
    import pqp_for_mpc_b200 as pqp

    n = 2 * M + 1 + N + N * M
    with np.errstate(over="ignore"):
        st = np.uint64(seed) + np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
        z = (st ^ (st >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)

    def u100(v):   # gen_u100: U[0,100) with six decimals, as fprintf("%f") leaves it (test_generator.c:944-945)
        u = (v >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
        return (np.floor(u * 100.0 * 1e6 + 0.5) / 1e6).astype(np.float32)

    q, fp, mp0, kp = u100(z[:M]), u100(z[M:2 * M]), u100(z[2 * M:2 * M + 1]), u100(z[2 * M + 1:2 * M + 1 + N])
    r = (z[2 * M + 1 + N:] % np.uint64(3)).astype(np.int8)
    Gp = np.where(r == 0, 0.0, np.where(r == 2, -1.0, 1.0)).astype(np.float32).reshape(N, M)
    d = pqp.Dims()
    d.M, d.N = M, N
    return dict(Qp_inv=np.diag(q).astype(np.float32), Fp=fp, Kp=kp, Gp=Gp, Mp0=float(mp0[0])), d
