"""CPU tests of the host-side partition of the upper-triangle loop (pqp_gemv_sym.cu: which units, strips and packets every
CTA owns).  The kernel's packet addressing relies on exactly these invariants; no GPU is touched."""
import ctypes as C

import numpy as np
import pytest


def tables(pqp, N, G):
    L = pqp.lib()
    nbmax = (N + 127) // 128
    u0 = np.zeros(G + 1, np.int32)
    j0 = np.zeros(G, np.int32)
    c0 = np.zeros(nbmax, np.int32)
    c1 = np.zeros(nbmax, np.int32)
    nb, U = C.c_int(0), C.c_int(0)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    maxseg = L.pqp_internal_sym_tables(N, G, ip(u0), ip(j0), ip(c0), ip(c1), C.byref(nb), C.byref(U))
    return maxseg, nb.value, U.value, u0, j0, c0, c1


@pytest.mark.parametrize("N", [2369, 2560, 3001, 4096, 6000, 8192, 12288, 16384])
@pytest.mark.parametrize("G", [148, 132, 64, 37])
def test_unit_ranges_and_strip_tables(pqp, N, G):
    maxseg, nb, U, u0, j0, c0, c1 = tables(pqp, N, G)
    assert nb == (N + 127) // 128 and U == nb * (nb + 1)       # two 64-row units per 128x128 tile of the upper triangle
    assert maxseg >= 1
    assert u0[0] == 0 and u0[G] == U and np.all(np.diff(u0) >= 1)   # contiguous, non-empty, covering
    first = lambda J: J * (J + 1)                                # strip J holds units [J(J+1), (J+1)(J+2))
    strip_of = lambda u: int((np.sqrt(4.0 * u + 1.0) - 1.0) / 2.0 + 1e-9)
    segs = 0
    for c in range(G):
        J = j0[c]
        assert first(J) <= u0[c] < first(J + 1)                  # cta_j0 = strip of the CTA's first unit
        Jl = strip_of(u0[c + 1] - 1)
        assert first(Jl) <= u0[c + 1] - 1 < first(Jl + 1)
        segs = max(segs, Jl - J + 1)
    assert segs == maxseg                                        # room for every (CTA, strip) column packet
    for J in range(nb):
        a, b = first(J), first(J + 1) - 1
        assert u0[c0[J]] <= a < u0[c0[J] + 1] and u0[c1[J]] <= b < u0[c1[J] + 1]   # first / last CTA touching strip J
        assert c0[J] <= c1[J]
    # equal cost (a unit 1, a strip end 1.7): no CTA carries much more than the mean
    ends = np.array([first(J + 1) - 1 for J in range(nb)])
    cost = np.array([(u0[c + 1] - u0[c]) + 1.7 * np.count_nonzero((ends >= u0[c]) & (ends < u0[c + 1])) for c in range(G)])
    assert cost.max() <= cost.mean() + 2.8


def test_too_few_units_is_refused(pqp):
    maxseg, nb, U, *_ = tables(pqp, 512, 148)                    # 4 blocks -> 20 units for 148 CTAs
    assert maxseg == 0 and U == 20
