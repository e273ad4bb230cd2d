"""Numpy model of pqp-for-mpc_b200/csrc/pqp_batched_imma.cu (the int8 digit-plane batched PQP loop), operation for operation.

NOT the oracle (that is oracle/): this restates what OUR kernel computes -- integer digit products exactly, the fp32 epilogue
with the same operation order -- so that the GPU result can be required to be bit-identical to it (tests/test_imma_gpu.py),
and so that the accuracy of the scheme against the oracle and its float64 twin can be checked without a GPU
(tests/test_imma_model.py).

Scheme (error-free accumulation, Ozaki-style splitting):
  off-diagonal Q+/Q- rows  -> 24-bit unsigned fixed point relative to a per-row power-of-two scale, three u8 digit planes A0,A1,A2
  y of one problem         -> 22-bit fixed point relative to a per-problem power-of-two scale > max y, three signed s8 digit planes
                              Y0,Y1,Y2 (round to nearest: the dropped cross terms are zero mean), re-derived every iteration
  w0 = A0*Y0, w1 = A0*Y1 + A1*Y0, w2 = A0*Y2 + A1*Y1 + A2*Y0   exact in int32 (tensor cores, kind::i8)
  S = (w0*65536 + w1*256 + w2) * rowscale * yscale; num = S- + (Q-_ii+theta_i) y_i + F-, den likewise, y <- (num/den)*y  (fp32)
"""
import numpy as np


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


