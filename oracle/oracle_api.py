"""ORACLE (test infrastructure, not product code).

ctypes binding of ``oracle/liboracle_nfft4gp.so`` (oracle_path.c + nfft3_shim): the
self-contained CPU restatement of the reference's hot path.  Only tests/,
``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference`` legs
may import this module.  The NFFT3 fastsum boundary is a restatement, pinned on the reference's shipped real-NFFT3 output
(TESTS/TEST2/foo_0_nfft.txt; tests/test_real_nfft3_pin.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "liboracle_nfft4gp.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
SYMMATVEC = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_double, c_double_p, C.c_double, c_double_p)
SOLVE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, c_double_p, c_double_p)

_lib = None


def build(force=False):
    """Compile the C restatement (gcc, a second or two).  Building the checker is not using it."""
    if force or not os.path.exists(SO):
        subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)
    return SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(SO)
        L.oracle_additive_create.restype = C.c_void_p
        L.oracle_additive_create.argtypes = [c_double_p, C.c_int, C.c_int, C.c_int, c_int_p, C.c_int, C.c_int]
        L.oracle_additive_build.restype = C.c_int
        L.oracle_additive_build.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double]
        for nm in ("oracle_additive_matvec", "oracle_additive_grad_matvec"):
            getattr(L, nm).restype = C.c_int
            getattr(L, nm).argtypes = [C.c_void_p, C.c_int, C.c_double, c_double_p, C.c_double, c_double_p]
        L.oracle_additive_free.restype = None
        L.oracle_additive_free.argtypes = [C.c_void_p]
        L.oracle_additive_nwindows.argtypes = [C.c_void_p]
        L.oracle_additive_window_dim.argtypes = [C.c_void_p, C.c_int]
        L.oracle_additive_window_scale.restype = C.c_double
        L.oracle_additive_window_scale.argtypes = [C.c_void_p, C.c_int]
        L.oracle_additive_window_points.restype = c_double_p
        L.oracle_additive_window_points.argtypes = [C.c_void_p, C.c_int]
        L.oracle_additive_window_bhat.restype = c_double_p
        L.oracle_additive_window_bhat.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.oracle_pcg.restype = C.c_int
        L.oracle_pcg.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, c_double_p, c_double_p,
                                 C.c_int, C.c_int, C.c_double, c_double_p, c_double_p, c_int_p]
        for nm in ("oracle_nys_solve", "oracle_fsai_solve", "oracle_afn_solve"):
            getattr(L, nm).restype = C.c_int
            getattr(L, nm).argtypes = [C.c_void_p, C.c_int, c_double_p, c_double_p]
        L.oracle_dense_additive.restype = C.c_int
        L.oracle_dense_additive.argtypes = [c_double_p, C.c_int, C.c_int, C.c_int, c_int_p, C.c_int, C.c_int,
                                            C.c_int, C.c_double, C.c_double, C.c_double, c_double_p]
        L.oracle_dense_symv.restype = C.c_int
        L.oracle_dense_symv.argtypes = [c_double_p, C.c_int, C.c_double, c_double_p, C.c_double, c_double_p]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_num_threads.restype = None
        L.oracle_set_num_threads.argtypes = [C.c_int]
        L.shim_kb_phi.restype = C.c_double
        L.shim_kb_phi.argtypes = [C.c_double, C.c_int, C.c_int]
        L.shim_kb_phi_hut.restype = C.c_double
        L.shim_kb_phi_hut.argtypes = [C.c_int, C.c_int, C.c_int]
        _lib = L
    return _lib


def dptr(a):
    return a.ctypes.data_as(c_double_p)


def iptr(a):
    return a.ctypes.data_as(c_int_p)


KERNELS = {"gaussian": 0, "matern12": 1, "matern": 1, 0: 0, 1: 1}


class OracleNys(C.Structure):
    _fields_ = [("n", C.c_int), ("k", C.c_int), ("perm", c_int_p), ("U", c_double_p), ("s", c_double_p),
                ("eta", C.c_double), ("work", c_double_p)]


class OracleFsai(C.Structure):
    _fields_ = [("n", C.c_int), ("ia", c_int_p), ("ja", c_int_p), ("a", c_double_p), ("work", c_double_p)]


class OracleAfn(C.Structure):
    _fields_ = [("n", C.c_int), ("k", C.c_int), ("perm", c_int_p), ("L", c_double_p), ("K12", c_double_p),
                ("schur", OracleFsai), ("work", c_double_p)]


class OracleAdditiveNFFT:
    """X feature-major (d, n); windows flat int32 (nwindows*dwindows, -1 pads the last window)."""

    def __init__(self, X, windows, nwindows, dwindows, kernel="gaussian"):
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.d, self.n = self.X.shape
        self.windows = np.ascontiguousarray(windows, dtype=np.int32)
        self.kernel = KERNELS[kernel]
        self.h = lib().oracle_additive_create(dptr(self.X), self.n, self.n, self.d, iptr(self.windows), nwindows, dwindows)

    def build(self, f, l, mu):
        assert lib().oracle_additive_build(self.h, self.kernel, f, l, mu) == 0
        return self

    def matvec(self, x, alpha=1.0, beta=0.0, y=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.n) if y is None else np.array(y, dtype=np.float64)
        lib().oracle_additive_matvec(self.h, self.n, alpha, dptr(x), beta, dptr(y))
        return y

    def grad_matvec(self, x, alpha=1.0, beta=0.0, y=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(3 * self.n) if y is None else np.array(y, dtype=np.float64)
        lib().oracle_additive_grad_matvec(self.h, self.n, alpha, dptr(x), beta, dptr(y))
        return y

    def window_scale(self, w):
        return lib().oracle_additive_window_scale(self.h, w)

    def window_points(self, w):
        dw = lib().oracle_additive_window_dim(self.h, w)
        return np.ctypeslib.as_array(lib().oracle_additive_window_points(self.h, w), shape=(self.n, dw)).copy()

    def window_bhat(self, w, which=0, N=32):
        dw = lib().oracle_additive_window_dim(self.h, w)
        raw = np.ctypeslib.as_array(lib().oracle_additive_window_bhat(self.h, w, which), shape=(N ** dw, 2)).copy()
        return (raw[:, 0] + 1j * raw[:, 1]).reshape((N,) * dw)

    def matvec_cb(self):
        """func_symmatvec-shaped callback + opaque handle for oracle_pcg."""
        return C.cast(lib().oracle_additive_matvec, C.c_void_p), self.h

    def free(self):
        if self.h:
            lib().oracle_additive_free(self.h)
            self.h = None


def dense_additive(X, windows, nwindows, dwindows, kernel, f, l, mu):
    X = np.ascontiguousarray(X, dtype=np.float64)
    d, n = X.shape
    windows = np.ascontiguousarray(windows, dtype=np.int32)
    K = np.zeros((n, n))
    lib().oracle_dense_additive(dptr(X), n, n, d, iptr(windows), nwindows, dwindows, KERNELS[kernel], f, l, mu, dptr(K))
    return K


def pcg(mat, n, matvec_ptr, prec, solve_ptr, x0, rhs, maxits, atol, tol):
    """oracle_pcg: returns (x, rel_res, rel_res_v, iters)."""
    x = np.array(x0, dtype=np.float64)
    rhs = np.ascontiguousarray(rhs, dtype=np.float64)
    hist = np.zeros(min(maxits, n) + 1)
    rel = C.c_double(0.0)
    it = C.c_int(0)
    lib().oracle_pcg(mat, n, matvec_ptr, prec, solve_ptr, dptr(x), dptr(rhs), maxits, atol, tol,
                     C.byref(rel), dptr(hist), C.byref(it))
    return x, rel.value, hist, it.value
