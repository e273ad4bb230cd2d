"""pqp-for-mpc_b200 -- B200-native PQP solver for condensed MPC dual QPs.

Thin ctypes binding of ``libpqp_b200.so`` (C ABI in ``include/pqp.h``).  All computation is in
the hand-written sm_100a kernels under ``csrc/``; there is no CPU or PyTorch fallback here --
if the shared library is missing this module raises, and without a B200 every compute call
returns ``PQP_ERR_NO_DEVICE`` (raised as :class:`PQPError`).

The directory name carries a hyphen (it mirrors the reference's repository name), so import it
through the root-level shim::

    import pqp_for_mpc_b200 as pqp
    prob, dims = pqp.load_example("tests/golden/example")
    with pqp.Solver(dims, prob) as s:
        Y, U, st = s.solve(prob["x"][None, :], iters=312, primal=True)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpqp_b200.so")
COMPAT_PATH = os.path.join(_HERE, "libpqp_compat.so")
CSRC = os.path.join(_HERE, "csrc")

ORDER_FAST, ORDER_STRICT = 0, 1

ERRORS = {0: "PQP_OK", -1: "PQP_ERR_INVALID", -2: "PQP_ERR_NO_DEVICE", -3: "PQP_ERR_CUDA", -4: "PQP_ERR_ALLOC",
          -5: "PQP_ERR_IO", -6: "PQP_ERR_UNSUPPORTED"}

_FP = C.POINTER(C.c_float)


class Dims(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("M", "N", "nState", "nDisH", "pHorizon", "nInput", "nOutput", "nDis")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_PROBLEM_PTRS = ("Qp_inv", "Gp", "Kp", "Fp1", "Fp2", "Fp3", "D", "Mp1", "Mp2", "Mp3", "Mp4", "Mp5", "Mp6", "Fp")


class HostProblem(C.Structure):
    _fields_ = ([(n, _FP) for n in _PROBLEM_PTRS] + [("Mp0", C.c_float)] + [(n, _FP) for n in ("x", "Z", "Theta", "Kx", "Kd")])


class Opts(C.Structure):
    _fields_ = [("theta_floor", C.c_float), ("y_init", C.c_float), ("erc", C.c_float), ("eac", C.c_float),
                ("eaj", C.c_float), ("erj", C.c_float), ("order", C.c_int), ("device", C.c_int),
                ("max_iters", C.c_int), ("check_every", C.c_int), ("batch_capacity", C.c_int),
                ("use_tensor_cores", C.c_int), ("l2_persist", C.c_int), ("exploit_symmetry", C.c_int), ("exploit_structure", C.c_int), ("accelerate", C.c_int)]


class Status(C.Structure):
    _fields_ = [("iters", C.c_int), ("converged", C.c_int), ("min_slack", C.c_float), ("gap", C.c_float),
                ("Jd", C.c_float), ("kkt", C.c_float)]


STATUS_DTYPE = np.dtype([("iters", np.int32), ("converged", np.int32), ("min_slack", np.float32),
                         ("gap", np.float32), ("Jd", np.float32), ("kkt", np.float32)])

# every symbol include/pqp.h declares (tests check the library exports all of them)
ABI_SYMBOLS = (
    "pqp_default_opts", "pqp_dims_mpc", "pqp_strerror", "pqp_last_cuda_error", "pqp_device_count",
    "pqp_load_example", "pqp_load_testfile", "pqp_generate_testproblem", "pqp_write_testfile", "pqp_free_problem",
    "pqp_setup", "pqp_setup_dual", "pqp_destroy", "pqp_solve_batch", "pqp_solve_dual", "pqp_recover_primal",
    "pqp_solve_batch_primal", "pqp_get_dual", "pqp_get_linear_terms", "pqp_get_stream", "pqp_last_solve_ms",
    "pqp_launch_count", "pqp_last_kernel", "pqp_device_qd", "pqp_matmul", "pqp_shift_duals", "pqp_output_offsets", "pqp_update_y2",
    "pqp_solve_dual_full", "pqp_set_constraint_bounds", "pqp_compute_fp", "pqp_compute_cost", "pqp_compute_md", "pqp_compute_u_from_y", "pqp_setup_gemm_ms", "pqp_setup_mirrored",
)
MM_STRICT, MM_SIMT, MM_TENSOR = 0, 1, 2


class PQPError(RuntimeError):
    def __init__(self, code, where=""):
        self.code = code
        detail = ""
        try:
            detail = lib().pqp_last_cuda_error().decode()
        except Exception:
            pass
        super().__init__(f"{where}: {ERRORS.get(code, code)}" + (f" [{detail}]" if detail and code == -3 else ""))


def build(verbose: bool = False) -> str:
    """Compile libpqp_b200.so + libpqp_compat.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-C", CSRC, "all"], check=True, stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """The loaded product library.  Fails loudly when it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(this package has no CPU / PyTorch fallback)")
        L = C.CDLL(LIB_PATH, mode=os.RTLD_GLOBAL)
        L.pqp_strerror.restype = C.c_char_p
        L.pqp_last_cuda_error.restype = C.c_char_p
        L.pqp_last_kernel.restype = C.c_char_p
        L.pqp_last_kernel.argtypes = [C.c_void_p]
        L.pqp_get_stream.restype = C.c_void_p
        L.pqp_get_stream.argtypes = [C.c_void_p]
        L.pqp_last_solve_ms.restype = C.c_float
        L.pqp_last_solve_ms.argtypes = [C.c_void_p]
        L.pqp_setup_gemm_ms.restype = C.c_float
        L.pqp_setup_gemm_ms.argtypes = [C.c_void_p]
        L.pqp_setup_mirrored.restype = C.c_int
        L.pqp_setup_mirrored.argtypes = [C.c_void_p]
        L.pqp_launch_count.restype = C.c_longlong
        L.pqp_launch_count.argtypes = [C.c_void_p]
        L.pqp_device_qd.restype = C.c_void_p
        L.pqp_device_qd.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.pqp_destroy.argtypes = [C.c_void_p]
        L.pqp_destroy.restype = None
        L.pqp_free_problem.restype = None
        L.pqp_default_opts.restype = None
        L.pqp_dims_mpc.restype = None
        L.pqp_generate_testproblem.argtypes = [C.c_ulonglong, C.c_int, C.c_int, C.POINTER(Dims), C.POINTER(HostProblem)]
        _lib = L
    return _lib


def device_count() -> int:
    return int(lib().pqp_device_count())


def default_opts(**kw) -> Opts:
    o = Opts()
    lib().pqp_default_opts(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown option {k}")
        setattr(o, k, v)
    return o


def dims_mpc(pHorizon, nState, nInput, nOutput, nDis) -> Dims:
    d = Dims()
    lib().pqp_dims_mpc(C.byref(d), pHorizon, nState, nInput, nOutput, nDis)
    return d


def dims_plain(M, N) -> Dims:
    d = Dims()
    d.M, d.N = M, N
    return d


EXAMPLE_DIMS = dict(pHorizon=1, nState=29, nInput=7, nOutput=7, nDis=1)  # PQP_CPU.c:13-17


def _shapes(d: Dims):
    M, N, nS, nd = d.M, d.N, d.nState, d.nDisH
    no = d.nOutput * d.pHorizon
    return dict(Qp_inv=(M, M), Gp=(N, M), Kp=(N,), Fp1=(M, nd), Fp2=(M, nS), Fp3=(M,), D=(nd,), Mp1=(nS, nS),
                Mp2=(nd, nS), Mp3=(nd, nd), Mp4=(nS,), Mp5=(nd,), Mp6=(1,), Fp=(M,), x=(nS,), Z=(no, nS), Theta=(no, nd))


def _problem_to_numpy(hp: HostProblem, d: Dims) -> dict:
    out = {}
    for name, shape in _shapes(d).items():
        p = getattr(hp, name)
        if p and int(np.prod(shape)) > 0:
            out[name] = np.ctypeslib.as_array(p, shape=(int(np.prod(shape)),)).reshape(shape).copy()
    out["Mp0"] = float(hp.Mp0)
    return out


def load_example(directory: str, dims: Dims | None = None):
    """pqp_load_example: the reader that replaces input() (PQP_CPU.c:757-930).  Returns (dict of arrays, Dims)."""
    d = dims or dims_mpc(**EXAMPLE_DIMS)
    hp = HostProblem()
    rc = lib().pqp_load_example(directory.encode(), C.byref(d), C.byref(hp))
    if rc:
        raise PQPError(rc, f"pqp_load_example({directory})")
    out = _problem_to_numpy(hp, d)
    lib().pqp_free_problem(C.byref(hp))
    return out, d


def load_testfile(path: str):
    """pqp_load_testfile: testing/ format (test_generator.c:936-987).  Returns (dict, Dims)."""
    d, hp = Dims(), HostProblem()
    rc = lib().pqp_load_testfile(path.encode(), C.byref(d), C.byref(hp))
    if rc:
        raise PQPError(rc, f"pqp_load_testfile({path})")
    out = _problem_to_numpy(hp, d)
    lib().pqp_free_problem(C.byref(hp))
    return out, d


def generate_testproblem(seed: int, M: int, N: int):
    """Seeded instance with the distribution of testing/test_generator.c.  Returns (dict, Dims)."""
    d, hp = Dims(), HostProblem()
    rc = lib().pqp_generate_testproblem(C.c_ulonglong(seed), M, N, C.byref(d), C.byref(hp))
    if rc:
        raise PQPError(rc, "pqp_generate_testproblem")
    out = _problem_to_numpy(hp, d)
    lib().pqp_free_problem(C.byref(hp))
    return out, d


def write_testfile(path: str, prob: dict, d: Dims):
    hp, keep = _numpy_to_problem(prob)
    rc = lib().pqp_write_testfile(path.encode(), C.byref(d), C.byref(hp))
    if rc:
        raise PQPError(rc, f"pqp_write_testfile({path})")


def _numpy_to_problem(prob: dict):
    hp, keep = HostProblem(), []
    for name in _PROBLEM_PTRS + ("x", "Z", "Theta", "Kx", "Kd"):
        v = prob.get(name)
        if v is not None and np.size(v) > 0:
            a = np.ascontiguousarray(v, dtype=np.float32)
            keep.append(a)
            setattr(hp, name, a.ctypes.data_as(_FP))
    hp.Mp0 = float(prob.get("Mp0", 0.0))
    return hp, keep


def update_y2(Y, Qdp_theta, Qdn_theta, Fdp, Fdn, device=-1):
    """pqp_update_y2: one reference-order update from the dense split operands (updateY2 + updY, PQP_CPU.c:603-618, 590-596)."""
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    Y, Qp, Qn, Fdp, Fdn = f(Y), f(Qdp_theta), f(Qdn_theta), f(Fdp), f(Fdn)
    out = np.empty_like(Y)
    rc = lib().pqp_update_y2(_as_ptr(out), _as_ptr(Y), _as_ptr(Qp), _as_ptr(Qn), _as_ptr(Fdp), _as_ptr(Fdn), int(Y.size), int(device))
    if rc:
        raise PQPError(rc, "pqp_update_y2")
    return out


def output_offsets(d: Dims, Z, Theta):
    """pqp_output_offsets: (Kx [N x nState], Kd [N x nDisH]) for prob["Kx"], prob["Kd"] -- Kp(x, D) = Kp + Kx x + Kd D."""
    Z = None if Z is None else np.ascontiguousarray(Z, dtype=np.float32)
    Theta = None if Theta is None else np.ascontiguousarray(Theta, dtype=np.float32)
    Kx = np.zeros((d.N, max(d.nState, 0)), np.float32)
    Kd = np.zeros((d.N, max(d.nDisH, 0)), np.float32)
    rc = lib().pqp_output_offsets(C.byref(d), _as_ptr(Z), _as_ptr(Theta), _as_ptr(Kx), _as_ptr(Kd))
    if rc:
        raise PQPError(rc, "pqp_output_offsets")
    return Kx, Kd


def _as_ptr(a):
    """numpy array -> host pointer; int -> raw (device) pointer; None -> NULL."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    return a.ctypes.data_as(C.c_void_p)


def matmul(A, B, tA=False, tB=False, engine=MM_TENSOR, device=-1):
    """pqp_matmul: the reference's matrixMultiply (PQP_CPU.c:84-147) on the GPU.  A, B as stored (see tA/tB)."""
    A = np.ascontiguousarray(A, np.float32)
    B = np.ascontiguousarray(B, np.float32)
    a, b = (A.shape[1], A.shape[0]) if tA else A.shape
    c = B.shape[0] if tB else B.shape[1]
    assert (B.shape[1] if tB else B.shape[0]) == b
    out = np.empty((a, c), np.float32)
    rc = lib().pqp_matmul(_as_ptr(out), _as_ptr(A), int(tA), _as_ptr(B), int(tB), a, b, c, int(engine), int(device))
    if rc:
        raise PQPError(rc, "pqp_matmul")
    return out


class Solver:
    """One pqp_handle: the x-independent setup of a problem family on one GPU."""

    def __init__(self, dims: Dims | None = None, prob: dict | None = None, *, Qd=None, Gp=None, Qp_inv=None, **opts):
        self._h = C.c_void_p()
        self.opts = default_opts(**opts)
        if Qd is not None:
            Qd = np.ascontiguousarray(Qd, dtype=np.float32)
            N = Qd.shape[0]
            M = 0
            gp = qi = None
            if Gp is not None and Qp_inv is not None:
                gp, qi = np.ascontiguousarray(Gp, np.float32), np.ascontiguousarray(Qp_inv, np.float32)
                M = gp.shape[1]
            d = Dims()
            d.N, d.M = N, M
            self.dims = d
            rc = lib().pqp_setup_dual(C.byref(self._h), N, _as_ptr(Qd), M, _as_ptr(gp), _as_ptr(qi), C.byref(self.opts))
            where = "pqp_setup_dual"
        else:
            self.dims = dims
            hp, keep = _numpy_to_problem(prob)
            rc = lib().pqp_setup(C.byref(self._h), C.byref(dims), C.byref(hp), C.byref(self.opts))
            where = "pqp_setup"
        if rc:
            raise PQPError(rc, where)

    # -- context management -------------------------------------------------------------------
    def close(self):
        if self._h:
            lib().pqp_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- solve ------------------------------------------------------------------------------------
    def solve(self, X=None, iters=1000, *, D=None, Y0=None, primal=False, status=True, Fd=None):
        """pqp_solve_batch / pqp_solve_dual (+ pqp_recover_primal).  X [B x nState]; returns (Y, U|None, status|None)."""
        N, M = self.dims.N, self.dims.M
        f32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)
        X, D, Y0, Fd = f32(X), f32(D), f32(Y0), f32(Fd)
        if Fd is not None:
            B = Fd.reshape(-1, N).shape[0]
        elif X is not None and self.dims.nState > 0:
            B = X.reshape(-1, self.dims.nState).shape[0]
        else:
            B = 1 if Y0 is None else Y0.reshape(-1, N).shape[0]
        Y = np.empty((B, N), np.float32)
        st = np.zeros(B, STATUS_DTYPE) if status else None
        U = np.empty((B, M), np.float32) if primal else None
        if Fd is not None:
            rc = lib().pqp_solve_dual(self._h, _as_ptr(Fd), B, int(iters), _as_ptr(Y0), _as_ptr(Y), _as_ptr(st))
            where = "pqp_solve_dual"
            if rc == 0 and primal:
                raise ValueError("primal recovery after solve(Fd=...) needs recover(Y, Fp=...)")
        elif primal:
            rc = lib().pqp_solve_batch_primal(self._h, _as_ptr(X), _as_ptr(D), B, int(iters), _as_ptr(Y0), _as_ptr(Y),
                                              _as_ptr(U), _as_ptr(st))
            where = "pqp_solve_batch_primal"
        else:
            rc = lib().pqp_solve_batch(self._h, _as_ptr(X), _as_ptr(D), B, int(iters), _as_ptr(Y0), _as_ptr(Y), _as_ptr(st))
            where = "pqp_solve_batch"
        if rc:
            raise PQPError(rc, where)
        return Y, U, st

    def shift_duals(self, Y, y_floor=0.0):
        """pqp_shift_duals: the receding-horizon warm start for the next control period (each of the four constraint blocks
        moves one horizon step forward, the last step is held, values below y_floor are raised to it)."""
        Y = np.ascontiguousarray(Y, np.float32).reshape(-1, self.dims.N)
        out = np.empty_like(Y)
        rc = lib().pqp_shift_duals(self._h, _as_ptr(Y), Y.shape[0], C.c_float(y_floor), _as_ptr(out))
        if rc:
            raise PQPError(rc, "pqp_shift_duals")
        return out

    def recover(self, Y, Fp=None):
        """pqp_recover_primal: U = -Qp_inv (Gp' Y + Fp)  (computeUfromY, PQP_CPU.c:352)."""
        Y = np.ascontiguousarray(Y, np.float32).reshape(-1, self.dims.N)
        B = Y.shape[0]
        Fp = None if Fp is None else np.ascontiguousarray(Fp, np.float32)
        U = np.empty((B, self.dims.M), np.float32)
        rc = lib().pqp_recover_primal(self._h, _as_ptr(Y), _as_ptr(Fp), B, _as_ptr(U))
        if rc:
            raise PQPError(rc, "pqp_recover_primal")
        return U

    # -- introspection --------------------------------------------------------------------------
    def dual(self, want_gq=True):
        N, M = self.dims.N, self.dims.M
        Qd, th = np.empty((N, N), np.float32), np.empty(N, np.float32)
        GQ = np.empty((N, M), np.float32) if (want_gq and M > 0) else None
        rc = lib().pqp_get_dual(self._h, _as_ptr(Qd), _as_ptr(th), _as_ptr(GQ))
        if rc:
            raise PQPError(rc, "pqp_get_dual")
        return Qd, th, GQ

    def linear_terms(self, B=1, want_fp=True):
        Fd = np.empty((B, self.dims.N), np.float32)
        Fp = np.empty((B, self.dims.M), np.float32) if want_fp else None
        rc = lib().pqp_get_linear_terms(self._h, B, _as_ptr(Fd), _as_ptr(Fp))
        if rc:
            raise PQPError(rc, "pqp_get_linear_terms")
        return Fd, Fp

    @property
    def stream(self) -> int:
        return int(lib().pqp_get_stream(self._h) or 0)

    @property
    def last_solve_ms(self) -> float:
        return float(lib().pqp_last_solve_ms(self._h))

    @property
    def setup_gemm_ms(self) -> float:
        return float(lib().pqp_setup_gemm_ms(self._h))

    @property
    def setup_mirrored(self) -> bool:
        """Qd was built from the tiles of its upper triangle (symmetric Qp_inv) and mirrored."""
        return bool(lib().pqp_setup_mirrored(self._h))

    @property
    def launch_count(self) -> int:
        return int(lib().pqp_launch_count(self._h))

    @property
    def last_kernel(self) -> str:
        return lib().pqp_last_kernel(self._h).decode()

    @property
    def handle(self):
        return self._h
