"""ORACLE (test infrastructure, not product code).

ctypes binding of ``oracle/_ref/libnfft4gp_ref.so`` = the reference's own C sources
(compiled unmodified from /root/reference by ``oracle/Makefile``) on top of the NFFT3
fastsum shim.  Used to (a) generate the golden fixtures under ``tests/golden`` and
(b) validate the self-contained restatement ``oracle/oracle_path.c``.

Struct layouts mirror INC/_linearalg.h:420-450 (nfft4gp_kernel).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libnfft4gp_ref.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)

FUNC_KERNEL = C.CFUNCTYPE(C.c_int, C.c_void_p, c_double_p, C.c_int, C.c_int, C.c_int, c_int_p, C.c_int,
                          c_int_p, C.c_int, C.POINTER(c_double_p), C.POINTER(c_double_p))
FUNC_SYMMATVEC = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_double, c_double_p, C.c_double, c_double_p)
FUNC_SOLVE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, c_double_p, c_double_p)


class Nfft4GPKernel(C.Structure):
    """INC/_linearalg.h:420-450"""
    _fields_ = [
        ("_params", C.c_double * 5),
        ("_iparams", C.c_int * 5),
        ("_max_n", C.c_int),
        ("_omp", C.c_int),
        ("_noise_level", C.c_double),
        ("_own_buffer", C.c_int),
        ("_buffer", c_double_p),
        ("_own_dbuffer", C.c_int),
        ("_dbuffer", c_double_p),
        ("_fkernel_buffer", C.c_void_p),
        ("_ibufferp", C.c_void_p),
        ("_libufferp", C.c_void_p),
        ("_own_fkernel_buffer_params", C.c_int),
        ("_fkernel_buffer_params", C.c_void_p),
        ("_ldwork", C.c_size_t),
        ("_dwork", c_double_p),
        ("_external", C.c_void_p),
    ]


def available():
    return os.path.exists(REF_SO)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libnfft4gp_ref.so missing: run `make -C oracle ref` where /root/reference exists")
        _lib = C.CDLL(REF_SO, mode=C.RTLD_GLOBAL)
        L = _lib
        L.Nfft4GPNFFTAdditiveKernelParamCreate.restype = C.c_void_p
        L.Nfft4GPNFFTAdditiveKernelParamCreate.argtypes = [c_double_p, C.c_int, C.c_int, C.c_int, c_int_p, C.c_int, C.c_int]
        L.Nfft4GPKernelAdditiveKernelParamCreate.restype = C.c_void_p
        L.Nfft4GPKernelAdditiveKernelParamCreate.argtypes = [c_double_p, C.c_int, C.c_int, C.c_int, c_int_p, C.c_int, C.c_int, C.c_void_p]
        kargs = [C.c_void_p, c_double_p, C.c_int, C.c_int, C.c_int, c_int_p, C.c_int, c_int_p, C.c_int,
                 C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        for nm in ("Nfft4GPNFFTAdditiveKernelGaussianKernel", "Nfft4GPNFFTAdditiveKernelMatern12Kernel",
                   "Nfft4GPKernelAdditiveKernel", "Nfft4GPKernelGaussianKernel", "Nfft4GPKernelMatern12Kernel"):
            getattr(L, nm).restype = C.c_int
            getattr(L, nm).argtypes = kargs
        margs = [C.c_void_p, C.c_int, C.c_double, c_double_p, C.c_double, c_double_p]
        for nm in ("Nfft4GPAdditiveNFFTMatSymv", "Nfft4GPAdditiveNFFTGradMatSymv", "Nfft4GPDenseMatSymv",
                   "Nfft4GPDenseGradMatSymv", "Nfft4GPNFFTMatSymv", "Nfft4GPNFFTGradMatSymv"):
            getattr(L, nm).restype = C.c_int
            getattr(L, nm).argtypes = margs
        L.Nfft4GPAdditiveNFFTKernelFree.argtypes = [C.c_void_p]
        L.Nfft4GPAdditiveNFFTKernelFree.restype = None
        L.Nfft4GPKernelParamFree.argtypes = [C.c_void_p]
        L.Nfft4GPKernelParamFree.restype = None
        L.Nfft4GPSolverPcg.restype = C.c_int
        L.Nfft4GPSolverPcg.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, c_double_p, c_double_p,
                                       C.c_int, C.c_int, C.c_double, c_double_p, C.POINTER(c_double_p), c_int_p, C.c_int]
        for nm in ("Nfft4GPPrecondNysCreate", "Nfft4GPPrecondFsaiCreate", "Nfft4GPPrecondCholCreate"):
            getattr(L, nm).restype = C.c_void_p
            getattr(L, nm).argtypes = []
        for nm in ("Nfft4GPPrecondNysFree", "Nfft4GPPrecondFsaiFree", "Nfft4GPPrecondCholFree"):
            getattr(L, nm).restype = None
            getattr(L, nm).argtypes = [C.c_void_p]
        L.Nfft4GPPrecondNysSetRank.argtypes = [C.c_void_p, C.c_int]
        L.Nfft4GPPrecondNysSetRank.restype = None
        L.Nfft4GPPrecondNysSetPerm.argtypes = [C.c_void_p, c_int_p, C.c_int]
        L.Nfft4GPPrecondNysSetPerm.restype = None
        L.Nfft4GPPrecondFsaiSetLfil.argtypes = [C.c_void_p, C.c_int]
        L.Nfft4GPPrecondFsaiSetLfil.restype = None
        sargs = [c_double_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        for nm in ("Nfft4GPPrecondNysSetupWithKernel", "Nfft4GPPrecondFsaiSetupWithKernel", "Nfft4GPPrecondCholSetupWithKernel"):
            getattr(L, nm).restype = C.c_int
            getattr(L, nm).argtypes = sargs
        for nm in ("Nfft4GPPrecondNysSolve", "Nfft4GPPrecondFsaiSolve", "Nfft4GPPrecondCholSolve"):
            getattr(L, nm).restype = C.c_int
            getattr(L, nm).argtypes = [C.c_void_p, C.c_int, c_double_p, c_double_p]
        L.Nfft4GPDatasetUniformRandom.restype = c_double_p
        L.Nfft4GPDatasetUniformRandom.argtypes = [C.c_int, C.c_int]
        L.free.argtypes = [C.c_void_p]
        L.free.restype = None
    return _lib


def dptr(a):
    return a.ctypes.data_as(c_double_p)


def iptr(a):
    return a.ctypes.data_as(c_int_p)


def set_hyper(handle, f, l, mu):
    k = Nfft4GPKernel.from_address(handle)
    k._params[0] = f
    k._params[1] = l
    k._noise_level = mu


class RefAdditiveNFFT:
    """The reference call sequence of TESTS/TEST1/foo.cpp:209-251 for the NFFT additive kernel.
    X is feature-major (d rows of n), windows a flat int32 array (nwindows*dwindows, -1 pads)."""

    def __init__(self, X, windows, nwindows, dwindows, kernel="gaussian"):
        L = lib()
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.d, self.n = self.X.shape
        self.windows = np.ascontiguousarray(windows, dtype=np.int32)
        self.kernel = kernel
        self.h = L.Nfft4GPNFFTAdditiveKernelParamCreate(dptr(self.X), self.n, self.n, self.d, iptr(self.windows), nwindows, dwindows)
        self.K = C.c_void_p()
        self.dK = C.c_void_p()

    def build(self, f, l, mu):
        L = lib()
        set_hyper(self.h, f, l, mu)
        fn = L.Nfft4GPNFFTAdditiveKernelGaussianKernel if self.kernel == "gaussian" else L.Nfft4GPNFFTAdditiveKernelMatern12Kernel
        self.K = C.c_void_p()
        self.dK = C.c_void_p()
        rc = fn(self.h, dptr(self.X), self.n, self.n, self.d, None, 0, None, 0, C.byref(self.K), C.byref(self.dK))
        assert rc == 0
        return self

    def matvec(self, x, alpha=1.0, beta=0.0, y=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.n) if y is None else np.array(y, dtype=np.float64)
        rc = lib().Nfft4GPAdditiveNFFTMatSymv(self.K, self.n, alpha, dptr(x), beta, dptr(y))
        assert rc == 0
        return y

    def grad_matvec(self, x, alpha=1.0, beta=0.0, y=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(3 * self.n) if y is None else np.array(y, dtype=np.float64)
        rc = lib().Nfft4GPAdditiveNFFTGradMatSymv(self.dK, self.n, alpha, dptr(x), beta, dptr(y))
        assert rc == 0
        return y

    def free(self):
        if self.h:
            lib().Nfft4GPAdditiveNFFTKernelFree(self.h)
            self.h = None


class RefAdditiveDense:
    """Dense additive kernel of SRC/linearalg/kernels.c:3046-3494 (the arbiter)."""

    def __init__(self, X, windows, nwindows, dwindows, kernel="gaussian"):
        L = lib()
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.d, self.n = self.X.shape
        self.windows = np.ascontiguousarray(windows, dtype=np.int32)
        base = L.Nfft4GPKernelGaussianKernel if kernel == "gaussian" else L.Nfft4GPKernelMatern12Kernel
        self.h = L.Nfft4GPKernelAdditiveKernelParamCreate(dptr(self.X), self.n, self.n, self.d, iptr(self.windows), nwindows, dwindows,
                                                         C.cast(base, C.c_void_p))
        self.K = C.c_void_p()
        self.dK = C.c_void_p()

    def build(self, f, l, mu, grad=True):
        L = lib()
        set_hyper(self.h, f, l, mu)
        self._free_mats()
        rc = L.Nfft4GPKernelAdditiveKernel(self.h, dptr(self.X), self.n, self.n, self.d, None, 0, None, 0,
                                          C.byref(self.K), C.byref(self.dK) if grad else None)
        assert rc == 0
        return self

    def matvec(self, x, alpha=1.0, beta=0.0, y=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.n) if y is None else np.array(y, dtype=np.float64)
        lib().Nfft4GPDenseMatSymv(self.K, self.n, alpha, dptr(x), beta, dptr(y))
        return y

    def grad_matvec(self, x, alpha=1.0, beta=0.0, y=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(3 * self.n) if y is None else np.array(y, dtype=np.float64)
        lib().Nfft4GPDenseGradMatSymv(self.dK, self.n, alpha, dptr(x), beta, dptr(y))
        return y

    def matrix(self):
        """Symmetrised copy of the n x n kernel matrix (the reference fills the lower triangle)."""
        a = np.ctypeslib.as_array(C.cast(self.K, c_double_p), shape=(self.n, self.n)).T.copy()  # column-major
        low = np.tril(a)
        return low + np.tril(a, -1).T

    def _free_mats(self):
        L = lib()
        if self.K:
            L.free(self.K)
        if self.dK:
            L.free(self.dK)
        self.K = C.c_void_p()
        self.dK = C.c_void_p()


def ref_pcg(mat, n, matvec_fn, prec, prec_fn, x0, rhs, maxits, atol, tol, print_level=0):
    """Nfft4GPSolverPcg (SRC/solvers/pcg.c:3-206).  Returns (x, rel_res, rel_res_v, iters)."""
    L = lib()
    x = np.array(x0, dtype=np.float64)
    rhs = np.ascontiguousarray(rhs, dtype=np.float64)
    rel = C.c_double(0.0)
    relv = c_double_p()
    it = C.c_int(0)
    L.Nfft4GPSolverPcg(mat, n, C.cast(matvec_fn, C.c_void_p), prec, C.cast(prec_fn, C.c_void_p) if prec_fn else None,
                       dptr(x), dptr(rhs), maxits, atol, tol, C.byref(rel), C.byref(relv), C.byref(it), print_level)
    m = min(maxits, n) + 1
    hist = np.ctypeslib.as_array(relv, shape=(m,)).copy() if relv else np.zeros(0)
    return x, rel.value, hist, it.value
