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



def has_pair_structure(Qd):
    """The test pair_struct_check_kernel applies: Qd[sigma(i)][j] == -Qd[i][j] == Qd[i][sigma(j)] for the representatives
    R = [0, N/4) u [N/2, 3N/4), sigma(i) = i + N/4, and a non-negative diagonal there."""
    N = Qd.shape[0]
    if N % 4:
        return False
    Mq = N // 4
    R = np.r_[0:Mq, 2 * Mq:3 * Mq]
    S = R + Mq
    sig = np.empty(N, np.int64)
    sig[R], sig[S] = S, R
    A = Qd[R]
    return bool(np.array_equal(Qd[S], -A) and np.array_equal(A[:, sig], -A) and (np.diag(Qd)[R] >= 0).all())


def _bitmax(Y):
    """per-problem maximum of |y| as the kernel forms it: on the bit patterns, so NaN > inf > every finite value"""
    return (Y.view(np.uint32) & np.uint32(0x7fffffff)).max(axis=0).view(np.float32)


def _scale_from(bound_bits_f32, ybits=22):
    """problem_scales(): quantise with 2^(22-f), 2^f > bound; returns (sc float64, isc float32)"""
    ex = np.clip((_f32_bits(bound_bits_f32) >> 23).astype(np.int64), 22, 254)
    return np.exp2(148.0 - ex + (ybits - 22)), np.exp2(ex - 148.0 - (ybits - 22)).astype(np.float32)


def _digits(Y, sc):
    with np.errstate(invalid="ignore", over="ignore"):
        b = np.rint(Y.astype(np.float64) * sc[None, :])
    b = np.nan_to_num(b, nan=0.0, posinf=2.0 ** 31 - 1)
    Y2 = ((b + 128) % 256) - 128
    b1 = (b - Y2) / 256
    Y1 = ((b1 + 128) % 256) - 128
    Y0 = (b1 - Y1) / 256
    return Y0, Y1, Y2


def run_paired(Qd, theta, Fd, K, ybits=22, y_init=1000.0, Y0=None):
    """K updates of B problems whose Qd has the +/- row-pair structure, operation for operation what
    batched_imma_pair_kernel<PAIRED = true> does (pqp_batched_imma_pair.cu, paired_epilogue):
      * only the N/2 representative rows are multiplied (matrix 0 = their Q- rows, matrix 1 = their Q+ rows, columns in the order
        [R | sigma(R)], the two a_ii elements left out); a row and its partner are updated together from the sums S2, S1;
      * the scale of the new digits comes from the bound (2 M_t + c)(1 + 2^-16) >= max y_{t+1}, M_t = max y_t, c = max_i F-_i/theta_i
        (the first scale from the exact maximum of y_0); a problem turns NaN as a whole one update after a non-finite dual appears."""
    N, B = Qd.shape[0], Fd.shape[0]
    f32, f64 = np.float32, np.float64
    Mq, nh = N // 4, N // 2
    R = np.r_[0:Mq, 2 * Mq:3 * Mq]
    S_ = R + Mq
    perm = np.r_[R, S_]                          # K position k <-> column perm[k]
    Qr = Qd[R][:, perm].astype(f32)              # representative rows, permuted columns
    M0 = np.maximum(-Qr, 0).astype(f32)          # Q- rows; the (i, sigma(i)) element (K position nh + i) left out
    M1 = np.maximum(Qr, 0).astype(f32)           # Q+ rows; the diagonal (K position i) left out
    ii = np.arange(nh)
    M0[ii, nh + ii] = 0
    M1[ii, ii] = 0
    P0, P1 = slice_rows_u8(M0), slice_rows_u8(M1)
    a = np.diag(Qd)[R].astype(f32)
    th_r, th_s = theta[R].astype(f32), theta[S_].astype(f32)
    dp_r, dp_s = (a + th_r).astype(f32), (a + th_s).astype(f32)
    FdT = Fd.T.astype(f32)
    Fr, Fs = FdT[R], FdT[S_]
    with np.errstate(divide="ignore", invalid="ignore"):
        c = np.maximum((np.maximum(-Fr, 0) / th_r[:, None]).astype(f32), (np.maximum(-Fs, 0) / th_s[:, None]).astype(f32)).max(axis=0)
    Y = np.full((N, B), y_init, f32) if Y0 is None else np.ascontiguousarray(Y0.T, f32)
    m = lambda u, v: (u * v).astype(f32)
    p = lambda u, v: (u + v).astype(f32)
    M = _bitmax(Y)
    sc, isc = _scale_from(M, ybits)
    Y[:, ~np.isfinite(M)] = np.nan
    for _ in range(K):
        Yd0, Yd1, Yd2 = _digits(Y, sc)
        Yd0, Yd1, Yd2 = Yd0[perm], Yd1[perm], Yd2[perm]
        out = []
        for (A0, A1, A2, rs) in (P0, P1):
            w0 = A0 @ Yd0
            w1 = A0 @ Yd1 + A1 @ Yd0
            w2 = A0 @ Yd2 + A1 @ Yd1 + A2 @ Yd0
            assert max(np.abs(w0).max(), np.abs(w1).max(), np.abs(w2).max()) < 2 ** 31
            f0, f1, f2 = w0.astype(f32), w1.astype(f32), w2.astype(f32)
            inner = (f1.astype(f64) * 256.0 + f2.astype(f64)).astype(f32)
            t = (f0.astype(f64) * 65536.0 + inner.astype(f64)).astype(f32)
            out.append(((t * rs[:, None]).astype(f32) * isc[None, :]).astype(f32))
        S2, S1 = out
        yi, ys = Y[R], Y[S_]
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            num_i = p(S2, p(p(m(th_r[:, None], yi), m(a[:, None], ys)), np.maximum(-Fr, 0)))
            den_i = p(S1, p(m(dp_r[:, None], yi), np.maximum(Fr, 0)))
            num_s = p(S1, p(p(m(th_s[:, None], ys), m(a[:, None], yi)), np.maximum(-Fs, 0)))
            den_s = p(S2, p(m(dp_s[:, None], ys), np.maximum(Fs, 0)))
            Yn = np.empty_like(Y)
            Yn[R] = ((num_i / den_i).astype(f32) * yi).astype(f32)
            Yn[S_] = ((num_s / den_s).astype(f32) * ys).astype(f32)
            # scale of the new digits from M_t = max y_t (the iterate just consumed) and c; non-finite M_t: the problem turns NaN
            bound = (((2.0 * M.astype(f64) + c.astype(f64)).astype(f32)).astype(f64) * f64(f32(1.0000152587890625))).astype(f32)
        Yn[:, ~np.isfinite(M)] = np.nan
        sc, isc = _scale_from(bound, ybits)
        Y = Yn
        M = _bitmax(Y)
    return np.ascontiguousarray(Y.T)
