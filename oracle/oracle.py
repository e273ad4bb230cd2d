"""ctypes front-end of the CPU checker.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, bench.py's cpu_baseline / ``--impl reference`` legs and
``__graft_entry__.smoke()`` may import this module (it is the checker, never the thing
measured as the product or shipped).  Two back-ends with the same method names:

* ``Oracle(np.float32 | np.float64)`` -- our restatement, ``oracle/libpqp_oracle.so``
  (``pqp_oracle.c``; every function cites the PQP_CPU.c lines it follows);
* ``Reference(np.float32 | np.float64)`` -- the unmodified reference compiled from
  ``/root/reference/PQP_CPU.c`` into ``oracle/_ref/`` (``ref_harness.c``); ``available()`` is
  False when that library was not built / did not travel.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EXAMPLE_DIMS = dict(pHorizon=1, nState=29, nInput=7, nOutput=7, nDis=1)  # PQP_CPU.c:13-17


def build(quiet: bool = True) -> None:
    """Compile the restatement, and the reference harness when /root/reference exists."""
    subprocess.run(["make", "-C", HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class _Base:
    def __init__(self, path: str, dtype, prefix: str):
        self.dtype = np.dtype(dtype)
        self.sfx = "f32" if self.dtype == np.float32 else "f64"
        self.real = C.c_float if self.dtype == np.float32 else C.c_double
        self.lib = C.CDLL(path, mode=os.RTLD_LOCAL)
        self.prefix = prefix

    def _f(self, name, restype=None):
        fn = getattr(self.lib, f"{self.prefix}_{name}_{self.sfx}")
        fn.restype = restype
        return fn

    def arr(self, x):
        return np.ascontiguousarray(np.asarray(x, dtype=self.dtype))

    def zeros(self, *shape):
        return np.zeros(shape, dtype=self.dtype)

    # ---- shared surface -------------------------------------------------------------
    def matmul(self, A, tA, B, tB, a, b, c):
        A, B = self.arr(A), self.arr(B)
        out = self.zeros(a, c)
        self._f("matmul")(_ptr(out), _ptr(A), C.c_int(tA), _ptr(B), C.c_int(tB),
                          C.c_int(a), C.c_int(b), C.c_int(c))
        return out

    def gauss_jordan(self, A):
        A = self.arr(A)
        n = A.shape[0]
        res = self.zeros(n, n)
        self._f("gauss_jordan")(_ptr(A), _ptr(res), C.c_int(n))
        return res

    def recover_u(self, Y, Fp, Gp, Qp_inv):
        Y, Fp, Gp, Qp_inv = map(self.arr, (Y, Fp, Gp, Qp_inv))
        N, M = Gp.shape
        U = self.zeros(M)
        self._f("recover_u")(_ptr(U), _ptr(Y), _ptr(Fp), _ptr(Gp), _ptr(Qp_inv), C.c_int(N), C.c_int(M))
        return U

    def cost(self, Z, Q, F, m):
        Z, Q, F = map(self.arr, (Z, Q, F))
        ms = self.arr([m])
        return float(self._f("cost", self.real)(_ptr(Z), _ptr(Q), _ptr(F), _ptr(ms), C.c_int(Z.size)))


class Oracle(_Base):
    """Our restatement (pqp_oracle.c)."""

    def __init__(self, dtype=np.float32):
        path = os.path.join(HERE, "libpqp_oracle.so")
        if not os.path.exists(path):
            build()
        super().__init__(path, dtype, "orc")

    def load_example(self, directory, dims=EXAMPLE_DIMS):
        pH, nS, nI, nO, nD = (dims[k] for k in ("pHorizon", "nState", "nInput", "nOutput", "nDis"))
        M, N, nd, no = pH * nI, 4 * pH * nI, nD * pH, nO * pH
        p = dict(Qp_inv=self.zeros(M, M), Fp1=self.zeros(M, nd), Fp2=self.zeros(M, nS), Fp3=self.zeros(M),
                 Mp1=self.zeros(nS, nS), Mp2=self.zeros(nd, nS), Mp3=self.zeros(nd, nd), Mp4=self.zeros(nS),
                 Mp5=self.zeros(nd), Mp6=self.zeros(1), Gp=self.zeros(N, M), Kp=self.zeros(N),
                 x=self.zeros(nS), D=self.zeros(nd), Theta=self.zeros(no, nd), Z=self.zeros(no, nS))
        order = ["Qp_inv", "Fp1", "Fp2", "Fp3", "Mp1", "Mp2", "Mp3", "Mp4", "Mp5", "Mp6", "Gp", "Kp", "x", "D",
                 "Theta", "Z"]
        fn = self._f("load_example", C.c_int)
        rc = fn(directory.encode(), *(C.c_int(v) for v in (pH, nS, nI, nO, nD)), *(_ptr(p[k]) for k in order))
        if rc != 0:
            raise OSError(f"orc_load_example({directory}) failed rc={rc}")
        return p

    def compute_fp(self, Fp1, Fp2, Fp3, D, x):
        Fp1, Fp2, Fp3, D, x = map(self.arr, (Fp1, Fp2, Fp3, D, x))
        M = Fp3.size
        Fp = self.zeros(M)
        self._f("compute_fp")(_ptr(Fp), _ptr(Fp1), _ptr(Fp2), _ptr(Fp3), _ptr(D), _ptr(x),
                              C.c_int(M), C.c_int(D.size), C.c_int(x.size))
        return Fp

    def compute_mp(self, Mp1, Mp2, Mp3, Mp4, Mp5, Mp6, D, x):
        a = list(map(self.arr, (Mp1, Mp2, Mp3, Mp4, Mp5, Mp6, D, x)))
        return float(self._f("compute_mp", self.real)(*(_ptr(v) for v in a), C.c_int(a[6].size), C.c_int(a[7].size)))

    def convert_to_dual(self, Qp_inv, Gp, Kp, Fp, Mp, want_qd=True):
        Qp_inv, Gp, Kp, Fp = map(self.arr, (Qp_inv, Gp, Kp, Fp))
        N, M = Gp.shape
        Qd = self.zeros(N, N) if want_qd else None
        Fd, Md, GQ, mp = self.zeros(N), self.zeros(1), self.zeros(N, M), self.arr([Mp])
        self._f("convert_to_dual")(_ptr(Qd) if want_qd else None, _ptr(Fd), _ptr(Md), _ptr(GQ), _ptr(Qp_inv),
                                   _ptr(Gp), _ptr(Kp), _ptr(Fp), _ptr(mp), C.c_int(N), C.c_int(M))
        return Qd, Fd, float(Md[0]), GQ

    def theta(self, Qd, floor=5.0):
        Qd = self.arr(Qd)
        N = Qd.shape[0]
        th = self.zeros(N)
        self._f("theta")(_ptr(th), _ptr(Qd), C.c_int(N), self.real(floor))
        return th

    def split(self, Qd, theta):
        Qd, theta = self.arr(Qd), self.arr(theta)
        N = Qd.shape[0]
        P, Nn = self.zeros(N, N), self.zeros(N, N)
        self._f("split")(_ptr(P), _ptr(Nn), _ptr(Qd), _ptr(theta), C.c_int(N))
        return P, Nn

    def update_y2(self, Y, Qdp_theta, Qdn_theta, Fd):
        Y, P, Nn, Fd = map(self.arr, (Y, Qdp_theta, Qdn_theta, Fd))
        Fdp, Fdn = np.maximum(Fd, 0).astype(self.dtype), np.maximum(-Fd, 0).astype(self.dtype)
        Yn = self.zeros(Y.size)
        self._f("update_y2")(_ptr(Yn), _ptr(Y), _ptr(P), _ptr(Nn), _ptr(Fdp), _ptr(Fdn), C.c_int(Y.size))
        return Yn

    def solve_fixed(self, Qd, Fd, K, y_init=1000.0, theta_floor=5.0):
        Qd, Fd = self.arr(Qd), self.arr(Fd)
        N = Fd.size
        Y, th = self.zeros(N), self.zeros(N)
        self._f("solve_fixed")(_ptr(Y), _ptr(Qd), _ptr(Fd), C.c_int(N), C.c_long(K), self.real(y_init),
                               self.real(theta_floor), _ptr(th))
        return Y, th

    def iterate(self, Y, Qd, Fd, K, theta_floor=5.0):
        Y, Qd, Fd = self.arr(Y).copy(), self.arr(Qd), self.arr(Fd)
        self._f("iterate")(_ptr(Y), _ptr(Qd), _ptr(Fd), C.c_int(Fd.size), C.c_long(K), self.real(theta_floor))
        return Y

    def accel_step(self, Y, Qd, Fd):
        """computeph + computealphaY + updateY1 (PQP_CPU.c:625-630, :545-588) with computeph's `+= ph` read as `+= Fd`;
        returns (Y + alpha*ph, alpha)."""
        Y, Qd, Fd = self.arr(Y), self.arr(Qd), self.arr(Fd)
        Yn, alpha = self.zeros(Fd.size), self.zeros(1)
        self._f("accel_step")(_ptr(Yn), _ptr(alpha), _ptr(Y), _ptr(Qd), _ptr(Fd), C.c_int(Fd.size))
        return Yn, float(alpha[0])

    def solve_accel(self, Qd, Fd, K, every, y_init=1000.0, theta_floor=5.0):
        """K multiplicative updates with one acceleration step after every `every`-th of them (not after the last)."""
        Qd, Fd = self.arr(Qd), self.arr(Fd)
        Y = self.zeros(Fd.size)
        self._f("solve_accel")(_ptr(Y), _ptr(Qd), _ptr(Fd), C.c_int(Fd.size), C.c_long(K), C.c_long(every), self.real(y_init),
                               self.real(theta_floor))
        return Y

    def solve_converge(self, Qd, Fd, Md, Qp, Qp_inv, Fp, Mp, Gp, Kp, y_init=1000.0, theta_floor=5.0,
                       max_h=10 ** 7, tol=None):
        """tol=None: the reference's 1e-6 (PQP_CPU.c:19-22); a number: the same test with erc = eac = eaj = erj = tol."""
        Qd, Fd, Qp, Qp_inv, Fp, Gp, Kp = map(self.arr, (Qd, Fd, Qp, Qp_inv, Fp, Gp, Kp))
        N, M = Gp.shape
        Y, U, md, mp = self.zeros(N), self.zeros(M), self.arr([Md]), self.arr([Mp])
        args = (_ptr(Y), _ptr(U), _ptr(Qd), _ptr(Fd), _ptr(md), _ptr(Qp), _ptr(Qp_inv), _ptr(Fp), _ptr(mp),
                _ptr(Gp), _ptr(Kp), C.c_int(N), C.c_int(M), self.real(y_init), self.real(theta_floor), C.c_long(max_h))
        if tol is None:
            h = self._f("solve_converge", C.c_long)(*args)
        else:
            h = self._f("solve_converge_tol", C.c_long)(*args, C.c_double(tol))
        return Y, U, int(h)


class Reference(_Base):
    """The unmodified reference (PQP_CPU.c) behind oracle/_ref/libpqp_ref*.so."""

    @staticmethod
    def path(dtype=np.float32):
        name = "libpqp_ref.so" if np.dtype(dtype) == np.float32 else "libpqp_ref_f64.so"
        return os.path.join(HERE, "_ref", name)

    @classmethod
    def available(cls, dtype=np.float32):
        return os.path.exists(cls.path(dtype))

    def __init__(self, dtype=np.float32):
        super().__init__(self.path(dtype), dtype, "ref")

    def dims(self):
        out = (C.c_int * 5)()
        self._f("dims")(out)
        return dict(zip(("pHorizon", "nState", "nInput", "nOutput", "nDis"), out))

    def load_example(self, parent_dir):
        d = self.dims()
        pH, nS, nI, nO, nD = (d[k] for k in ("pHorizon", "nState", "nInput", "nOutput", "nDis"))
        M, N, nd, no = pH * nI, 4 * pH * nI, nD * pH, nO * pH
        p = dict(Qp_inv=self.zeros(M, M), Fp1=self.zeros(M, nd), Fp2=self.zeros(M, nS), Fp3=self.zeros(M),
                 Mp1=self.zeros(nS, nS), Mp2=self.zeros(nd, nS), Mp3=self.zeros(nd, nd), Mp4=self.zeros(nS),
                 Mp5=self.zeros(nd), Mp6=self.zeros(1), Gp=self.zeros(N, M), Kp=self.zeros(N),
                 x=self.zeros(nS), D=self.zeros(nd), Theta=self.zeros(no, nd), Z=self.zeros(no, nS))
        order = ["Qp_inv", "Fp1", "Fp2", "Fp3", "Mp1", "Mp2", "Mp3", "Mp4", "Mp5", "Mp6", "Gp", "Kp", "x", "D",
                 "Theta", "Z"]
        rc = self._f("load_example", C.c_int)(parent_dir.encode(), *(_ptr(p[k]) for k in order))
        if rc != 0:
            raise OSError(f"ref_load_example({parent_dir}) failed rc={rc}")
        return p

    def compute_fp(self, Fp1, Fp2, Fp3, D, x):
        Fp1, Fp2, Fp3, D, x = map(self.arr, (Fp1, Fp2, Fp3, D, x))
        Fp = self.zeros(Fp3.size)
        self._f("compute_fp")(_ptr(Fp), _ptr(Fp1), _ptr(Fp2), _ptr(Fp3), _ptr(D), _ptr(x))
        return Fp

    def compute_mp(self, Mp1, Mp2, Mp3, Mp4, Mp5, Mp6, D, x):
        a = list(map(self.arr, (Mp1, Mp2, Mp3, Mp4, Mp5, Mp6, D, x)))
        mp = self.zeros(1)
        self._f("compute_mp")(_ptr(mp), *(_ptr(v) for v in a))
        return float(mp[0])

    def convert_to_dual(self, Qp_inv, Gp, Kp, Fp, Mp):
        Qp_inv, Gp, Kp, Fp = map(self.arr, (Qp_inv, Gp, Kp, Fp))
        N, M = Gp.shape
        Qd, Fd, Md, mp = self.zeros(N, N), self.zeros(N), self.zeros(1), self.arr([Mp])
        self._f("convert_to_dual")(_ptr(Qd), _ptr(Fd), _ptr(Md), _ptr(Qp_inv), _ptr(Gp), _ptr(Kp), _ptr(Fp),
                                   _ptr(mp), C.c_int(N), C.c_int(M))
        return Qd, Fd, float(Md[0])

    def split(self, Qd):
        Qd = self.arr(Qd)
        N = Qd.shape[0]
        P, Nn, th = self.zeros(N, N), self.zeros(N, N), self.zeros(N)
        self._f("split")(_ptr(P), _ptr(Nn), _ptr(th), _ptr(Qd), C.c_int(N))
        return P, Nn, th

    def update_y2(self, Y, Qdp_theta, Qdn_theta, Fd):
        Y, P, Nn, Fd = map(self.arr, (Y, Qdp_theta, Qdn_theta, Fd))
        Fdp, Fdn = np.maximum(Fd, 0).astype(self.dtype), np.maximum(-Fd, 0).astype(self.dtype)
        Yn = self.zeros(Y.size)
        self._f("update_y2")(_ptr(Yn), _ptr(Y), _ptr(P), _ptr(Nn), _ptr(Fd), _ptr(Fdp), _ptr(Fdn), C.c_int(Y.size))
        return Yn

    def solve_fixed(self, Qd, Fd, K):
        Qd, Fd = self.arr(Qd), self.arr(Fd)
        Y = self.zeros(Fd.size)
        self._f("iterate")(_ptr(Y), _ptr(Qd), _ptr(Fd), C.c_int(Fd.size), C.c_long(K), C.c_int(1))
        return Y

    def iterate(self, Y, Qd, Fd, K):
        Y, Qd, Fd = self.arr(Y).copy(), self.arr(Qd), self.arr(Fd)
        self._f("iterate")(_ptr(Y), _ptr(Qd), _ptr(Fd), C.c_int(Fd.size), C.c_long(K), C.c_int(0))
        return Y

    def solve_converge(self, Qd, Fd, Md, Qp, Qp_inv, Fp, Mp, Gp, Kp):
        Qd, Fd, Qp, Qp_inv, Fp, Gp, Kp = map(self.arr, (Qd, Fd, Qp, Qp_inv, Fp, Gp, Kp))
        N, M = Gp.shape
        Y, U, md, mp = self.zeros(N), self.zeros(M), self.arr([Md]), self.arr([Mp])
        h = self._f("solve_converge", C.c_long)(_ptr(Y), _ptr(Qd), _ptr(Fd), _ptr(md), _ptr(U), _ptr(Qp),
                                                _ptr(Qp_inv), _ptr(Fp), _ptr(mp), _ptr(Gp), _ptr(Kp),
                                                C.c_int(N), C.c_int(M))
        return Y, U, int(h)

    def main_stdout(self, parent_dir):
        buf = C.create_string_buffer(1 << 16)
        rc = self._f("main", C.c_int)(parent_dir.encode(), buf, C.c_int(len(buf)))
        if rc != 0:
            raise OSError(f"ref_main rc={rc}")
        return buf.value.decode()
