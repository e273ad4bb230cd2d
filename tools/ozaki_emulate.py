"""CPU emulation (numpy, exact integer arithmetic) of the int8-sliced batched PQP update that
pqp_batched_imma.cu runs on the tensor cores, to check its accuracy against the float64 twin before / beside the
GPU.  usage: python tools/ozaki_emulate.py [K=1000] [B=24] [ybits=22]

Scheme (error-free accumulation, Ozaki-style):
  off-diagonal Q+/Q- rows  -> 24-bit unsigned fixed point relative to a per-row scale, three u8 slices A0,A1,A2
  y of one problem         -> `ybits`-bit fixed point relative to a power-of-two per-problem scale >= max y,
                              three signed-digit s8 slices Y0,Y1,Y2 (round to nearest: dropped terms are zero mean)
  w0 = A0*Y0, w1 = A0*Y1 + A1*Y0, w2 = A0*Y2 + A1*Y1 + A2*Y0   exact in int32
  S = (w0*65536 + w1*256 + w2) * rowscale * yscale; num = S- + (Q-_ii+theta_i) y_i + F-, den likewise  (fp32)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.oracle import Oracle  # noqa: E402


def _f32_bits(x):
    return np.asarray(x, np.float32).view(np.uint32)


def slice_rows_u8(Aoff):
    """Aoff [N x N] >= 0 float32, zero diagonal -> (A0, A1, A2 float64 integer planes, row scale 2^(e-8) float32 [N]).
    Mirrors build_imma_tiles_kernel: rmax < 2^e, e = biased exponent - 126; entries scaled by 2^(24-e)."""
    rmax = Aoff.max(axis=1).astype(np.float32)
    ex = (_f32_bits(rmax) >> 23).astype(np.int64)
    ok = (rmax > 0) & (ex >= 32) & (ex <= 254)
    scale = np.where(ok, np.exp2(150.0 - ex), 0.0)
    rs = np.where(ok, np.exp2(ex - 134.0), 0.0).astype(np.float32)
    a = np.rint(Aoff.astype(np.float64) * scale[:, None])
    assert a.max() < 2 ** 24
    A0 = np.floor(a / 65536)
    A1 = np.floor((a - A0 * 65536) / 256)
    A2 = a - A0 * 65536 - A1 * 256
    return A0, A1, A2, rs


def slice_y_s8(Y, ybits=22):
    """Y [N x B] >= 0 float32 -> signed-digit planes and per-problem power-of-two inverse scale (problem_scales())."""
    ymax = Y.max(axis=0).astype(np.float32)
    ex = np.clip((_f32_bits(ymax) >> 23).astype(np.int64), 22, 254)
    sc = np.exp2(148.0 - ex + (ybits - 22))
    isc = np.exp2(ex - 148.0 - (ybits - 22)).astype(np.float32)
    with np.errstate(invalid="ignore", over="ignore"):
        b = np.rint(Y.astype(np.float64) * sc[None, :])
    b = np.nan_to_num(b, nan=0.0, posinf=2.0 ** 31 - 1)
    Y2 = ((b + 128) % 256) - 128
    b1 = (b - Y2) / 256
    Y1 = ((b1 + 128) % 256) - 128
    Y0 = (b1 - Y1) / 256
    return Y0, Y1, Y2, isc


def run(Qd, theta, Fd, K, ybits=22, y_init=1000.0, Y0=None):
    """K updates of B problems; Fd [B x N]; returns Y [B x N].  Operation for operation what batched_imma_kernel does."""
    N, B = Qd.shape[0], Fd.shape[0]
    f32, f64 = np.float32, np.float64
    Qp = np.maximum(Qd, 0).astype(f32)
    Qn = np.maximum(-Qd, 0).astype(f32)
    dp = (np.diag(Qp) + theta.astype(f32)).astype(f32)
    dn = (np.diag(Qn) + theta.astype(f32)).astype(f32)
    np.fill_diagonal(Qp, 0)
    np.fill_diagonal(Qn, 0)
    P = slice_rows_u8(Qp)
    Nn = slice_rows_u8(Qn)
    Fp_ = np.maximum(Fd, 0).astype(f32).T
    Fn_ = np.maximum(-Fd, 0).astype(f32).T
    Y = np.full((N, B), y_init, f32) if Y0 is None else np.ascontiguousarray(Y0.T, f32)
    for _ in range(K):
        Yd0, Yd1, Yd2, isc = slice_y_s8(Y, ybits)
        out = []
        for (A0, A1, A2, rs) in (Nn, P):
            w0 = A0 @ Yd0
            w1 = A0 @ Yd1 + A1 @ Yd0
            w2 = A0 @ Yd2 + A1 @ Yd1 + A2 @ Yd0
            assert max(np.abs(w0).max(), np.abs(w1).max(), np.abs(w2).max()) < 2 ** 31
            f0, f1, f2 = w0.astype(f32), w1.astype(f32), w2.astype(f32)       # I2F, round to nearest
            inner = (f1.astype(f64) * 256.0 + f2.astype(f64)).astype(f32)     # fmaf: exact in f64, one rounding
            t = (f0.astype(f64) * 65536.0 + inner.astype(f64)).astype(f32)    # fmaf
            S = ((t * rs[:, None]).astype(f32) * isc[None, :]).astype(f32)    # exact power-of-two scalings
            out.append(S)
        num = ((out[0] + (dn[:, None] * Y).astype(f32)).astype(f32) + Fn_).astype(f32)
        den = ((out[1] + (dp[:, None] * Y).astype(f32)).astype(f32) + Fp_).astype(f32)
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            Y = ((num / den).astype(f32) * Y).astype(f32)
        # a problem whose largest dual is not finite turns NaN as a whole (what the reference's dense sums do one update later)
        Y[:, ~np.isfinite(Y).all(axis=0)] = np.nan
    return np.ascontiguousarray(Y.T)


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    ybits = int(sys.argv[3]) if len(sys.argv) > 3 else 22
    from bench_problems import condensed_mpc
    prob, d, X = condensed_mpc(2024, 30, 12, 4, n_states=B)
    o32, o64 = Oracle(np.float32), Oracle(np.float64)
    Fds = []
    Qd = None
    for b in range(B):
        Fp = o32.compute_fp(prob["Fp1"], prob["Fp2"], prob["Fp3"], prob["D"], X[b])
        Qd, Fd, _, _ = o32.convert_to_dual(prob["Qp_inv"], prob["Gp"], prob["Kp"], Fp, 0.0)
        Fds.append(Fd)
    Fds = np.stack(Fds)
    theta = o32.theta(Qd)
    Yz = run(Qd, theta, Fds, K, ybits)
    print(f"K={K} ybits={ybits}:  problem  ymax   e(i8,f64)  e(f32,f64)  e(i8,f32)  active-set(i8==f64, f32==f64)")
    worst = 0.0
    for b in range(B):
        y64, _ = o64.solve_fixed(Qd, Fds[b], K)
        y32, _ = o32.solve_fixed(Qd, Fds[b], K)
        n = np.abs(y64).max()
        if n == 0:
            print(f"{b:4d} ymax=0  i8 max {np.abs(Yz[b]).max()}")
            continue
        ez, ef, ezf = np.abs(Yz[b] - y64).max() / n, np.abs(y32 - y64).max() / n, np.abs(Yz[b] - y32).max() / n
        a64, a32, az = y64 > 1e-6 * n, y32 > 1e-6 * n, Yz[b] > 1e-6 * n
        worst = max(worst, ez / max(ef, 1e-12))
        print(f"{b:4d} {n:9.4g}  {ez:.2e}   {ef:.2e}   {ezf:.2e}   {np.array_equal(az, a64)} {np.array_equal(a32, a64)}")
    print("worst e(i8,f64)/e(f32,f64) =", worst)


if __name__ == "__main__":
    main()
