"""ctypes mirror of include/nfft4gp_b200.h -- same names, argument meaning and error behaviour as the reference's
C interface (INC/_external.h, INC/_solvers.h, INC/_preconds.h), so that the parity tests read like the reference's
own drivers (TESTS/TEST1/foo.cpp:209-251)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NFFT4GP_B200_LIB") or os.path.join(HERE, "libnfft4gp_b200.so")  # override: A/B builds of the kernels

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
FUNC_SYMMATVEC = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_double, c_double_p, C.c_double, c_double_p)
FUNC_SOLVE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, c_double_p, c_double_p)


class Nfft4GPKernel(C.Structure):
    """nfft4gp_kernel (include/nfft4gp_b200.h; reference INC/_linearalg.h:420-450)."""
    _fields_ = [
        ("_params", C.c_double * 5), ("_iparams", C.c_int * 5), ("_max_n", C.c_int), ("_omp", C.c_int),
        ("_noise_level", C.c_double), ("_own_buffer", C.c_int), ("_buffer", c_double_p), ("_own_dbuffer", C.c_int),
        ("_dbuffer", c_double_p), ("_fkernel_buffer", C.c_void_p), ("_ibufferp", C.c_void_p), ("_libufferp", C.c_void_p),
        ("_own_fkernel_buffer_params", C.c_int), ("_fkernel_buffer_params", C.c_void_p), ("_ldwork", C.c_size_t),
        ("_dwork", c_double_p), ("_external", C.c_void_p),
    ]


def build_library(force=False):
    """Compile libnfft4gp_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(HERE, "csrc"), "-j8"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    """Load the C-ABI library.  Fails loudly when it is missing -- there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a).  nfft4gp_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i, d = C.c_void_p, C.c_int, C.c_double
    kargs = [vp, c_double_p, i, i, i, c_int_p, i, c_int_p, i, C.POINTER(vp), C.POINTER(vp)]
    margs = [vp, i, d, c_double_p, d, c_double_p]
    sig = {
        "Nfft4GPNFFTAdditiveKernelParamCreate": (vp, [c_double_p, i, i, i, c_int_p, i, i]),
        "Nfft4GPNFFTAdditiveKernelGaussianKernel": (i, kargs),
        "Nfft4GPNFFTAdditiveKernelMatern12Kernel": (i, kargs),
        "Nfft4GPAdditiveNFFTMatSymv": (i, margs),
        "Nfft4GPAdditiveNFFTGradMatSymv": (i, margs),
        "Nfft4GPAdditiveNFFTKernelFree": (None, [vp]),
        "Nfft4GPNFFTKernelFree": (None, [vp]),
        "Nfft4GPNFFTKernelParamCreate": (vp, [i, i]),
        "Nfft4GPNFFTKernelParamFree": (None, [vp]),
        "Nfft4GPNFFTKernelGaussianKernel": (i, kargs),
        "Nfft4GPNFFTKernelMatern12Kernel": (i, kargs),
        "Nfft4GPNFFTMatSymv": (i, margs),
        "Nfft4GPNFFTGradMatSymv": (i, margs),
        "Nfft4GPSolverPcg": (i, [vp, i, vp, vp, vp, c_double_p, c_double_p, i, i, d, c_double_p, C.POINTER(c_double_p), c_int_p, i]),
        "Nfft4GPSolverFgmres": (i, [vp, i, vp, vp, vp, c_double_p, c_double_p, i, i, i, d, c_double_p, C.POINTER(c_double_p), c_int_p, i]),
        "Nfft4GPB200SolverFgmresDevice": (i, [vp, i, vp, vp, vp, vp, vp, i, i, i, d, c_double_p, C.POINTER(c_double_p), c_int_p, i]),
        "Nfft4GPSolverLanczos": (i, [vp, i, vp, vp, vp, c_double_p, c_double_p, i, i, i, d, c_double_p, C.POINTER(c_double_p), c_int_p, c_int_p,
                                     C.POINTER(c_double_p), C.POINTER(c_double_p), i]),
        "Nfft4GPLanczosQuadratureLogdet": (i, [vp, vp, i, vp, vp, vp, vp, vp, vp, vp, i, i, c_double_p, i, c_double_p, C.POINTER(c_double_p)]),
        "Nfft4GPNFFTAppendData": (c_double_p, [c_double_p, i, i, i, c_double_p, i, i]),
        "Nfft4GPNFFTKernelParamRemovePoints": (i, [vp]), "Nfft4GPNFFTKernelParamFreeNFFTKernel": (i, [vp]),
        "Nfft4GPAdditiveNFFTGpPredict": (i, [c_double_p, c_double_p, c_double_p, i, i, i, c_double_p, i, i, c_double_p, vp, vp, vp, vp, vp, vp, vp,
                                             vp, vp, vp, vp, i, d, i, i, i, c_double_p, C.POINTER(c_double_p), C.POINTER(c_double_p)]),
        "Nfft4GPPrecondNysCreate": (vp, []), "Nfft4GPPrecondNysFree": (None, [vp]), "Nfft4GPPrecondNysReset": (None, [vp]),
        "Nfft4GPPrecondNysSetRank": (None, [vp, i]), "Nfft4GPPrecondNysSetPerm": (None, [vp, c_int_p, i]),
        "Nfft4GPPrecondNysSolve": (i, [vp, i, c_double_p, c_double_p]),
        "Nfft4GPPrecondNysSetupWithKernel": (i, [c_double_p, i, i, i, vp, vp, i, vp]),
        "Nfft4GPPrecondNysDvp": (i, [vp, i, c_int_p, c_double_p, C.POINTER(c_double_p)]),
        "Nfft4GPPrecondNysTrace": (i, [vp, C.POINTER(c_double_p)]),
        "Nfft4GPPrecondNysLogdet": (d, [vp]),
        "Nfft4GPPrecondFsaiCreate": (vp, []), "Nfft4GPPrecondFsaiFree": (None, [vp]), "Nfft4GPPrecondFsaiReset": (None, [vp]),
        "Nfft4GPPrecondFsaiSetLfil": (None, [vp, i]),
        "Nfft4GPPrecondFsaiSolve": (i, [vp, i, c_double_p, c_double_p]),
        "Nfft4GPPrecondFsaiLogdet": (d, [vp]),
        "Nfft4GPPrecondFsaiDvp": (i, [vp, i, c_int_p, c_double_p, C.POINTER(c_double_p)]),
        "Nfft4GPPrecondFsaiTrace": (i, [vp, C.POINTER(c_double_p)]),
        "Nfft4GPPrecondFsaiSetupWithKernelPattern": (i, [vp, vp, c_double_p, i, i, i, vp, vp, i, vp]),
        "Nfft4GPPrecondFsaiInvL": (i, [vp, i, c_double_p, c_double_p]),
        "Nfft4GPPrecondFsaiInvLT": (i, [vp, i, c_double_p, c_double_p]),
        "Nfft4GPB200PrecondFsaiGetGradFactors": (i, [vp, c_double_p]),
        "Nfft4GPPrecondAFNLogdet": (d, [vp]),
        "Nfft4GPPrecondAFNTrace": (i, [vp, C.POINTER(c_double_p)]),
        "Nfft4GPPrecondFsaiSetupWithKernel": (i, [c_double_p, i, i, i, vp, vp, i, vp]),
        "Nfft4GPPrecondCholCreate": (vp, []), "Nfft4GPPrecondCholFree": (None, [vp]), "Nfft4GPPrecondCholReset": (None, [vp]),
        "Nfft4GPPrecondCholSetStable": (None, [vp, i]),
        "Nfft4GPPrecondCholSolve": (i, [vp, i, c_double_p, c_double_p]),
        "Nfft4GPPrecondCholDvp": (i, [vp, i, c_int_p, c_double_p, C.POINTER(c_double_p)]),
        "Nfft4GPPrecondCholTrace": (i, [vp, C.POINTER(c_double_p)]),
        "Nfft4GPPrecondCholLogdet": (d, [vp]),
        "Nfft4GPPrecondCholSetupWithKernel": (i, [c_double_p, i, i, i, vp, vp, i, vp]),
        "Nfft4GPPrecondAFNCreate": (vp, []), "Nfft4GPPrecondAFNFree": (None, [vp]),
        "Nfft4GPPrecondAFNSetRank": (None, [vp, i]), "Nfft4GPPrecondAFNSetLfil": (None, [vp, i]),
        "Nfft4GPPrecondAFNSolve": (i, [vp, i, c_double_p, c_double_p]),
        "Nfft4GPPrecondAFNSetupWithKernel": (i, [c_double_p, i, i, i, vp, vp, i, vp]),
        "Nfft4GPKernelAdditiveKernelParamCreate": (vp, [c_double_p, i, i, i, c_int_p, i, i, vp]),
        "Nfft4GPKernelAdditiveKernel": (i, kargs), "Nfft4GPKernelGaussianKernel": (i, kargs),
        "Nfft4GPKernelMatern12Kernel": (i, kargs), "Nfft4GPKernelParamFree": (None, [vp]),
        "Nfft4GPKernelParamCreate": (vp, [i, i]), "Nfft4GPKernelFree": (None, [vp]),
        "Nfft4GPB200PrecondAFNGetFactors": (i, [vp, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p, c_double_p, c_double_p, c_int_p,
                                                c_int_p, c_int_p, c_double_p]),
        "Nfft4GPB200PrecondAFNSetRankEstimation": (i, [vp, i, i, i, C.c_ulonglong]),
        "Nfft4GPB200PrecondAFNSetK12Storage": (i, [vp, i]),
        "Nfft4GPB200PrecondAFNSetRankMethod": (i, [vp, i]),
        "Nfft4GPB200PrecondAFNSetK12Operator": (i, [vp, vp, i]),
        "Nfft4GPB200RankestNysScaledSample": (i, [c_double_p, i, i, i, vp, vp, c_int_p, i, i]),
        "Nfft4GPB200DebugPartitionChunks": (i, [c_int_p, i, i, i, c_double_p, d, c_int_p]),
        "Nfft4GPB200DebugPoly1dBlockSort": (i, [c_double_p, i, i, c_double_p, C.POINTER(C.c_ushort), C.POINTER(C.c_ushort)]),
        "Nfft4GPB200HostPotrf": (i, [i, c_double_p]), "Nfft4GPB200HostSyev": (i, [i, c_double_p, c_double_p, i]),
        "Nfft4GPB200HostStev": (i, [i, c_double_p, c_double_p, c_double_p]),
        "Nfft4GPB200Dgemm": (i, [C.c_char, C.c_char, i, i, i, d, c_double_p, i, c_double_p, i, d, c_double_p, i]),
        "Nfft4GPB200Fps": (i, [c_double_p, i, i, i, i, c_int_p]),
        "Nfft4GPB200KnnPattern": (i, [c_double_p, i, i, i, i, c_int_p, c_int_p]),
        "Nfft4GPB200Init": (i, [i]), "Nfft4GPB200DeviceCount": (i, []), "Nfft4GPB200Version": (C.c_char_p, []),
        "Nfft4GPB200LaunchCount": (C.c_longlong, []), "Nfft4GPB200LaunchCountReset": (None, []),
        "Nfft4GPB200HostAlloc": (vp, [C.c_size_t]), "Nfft4GPB200HostFree": (None, [vp]),
        "Nfft4GPB200DeviceAlloc": (vp, [C.c_size_t]), "Nfft4GPB200DeviceFree": (None, [vp]),
        "Nfft4GPB200MemcpyH2D": (i, [vp, vp, C.c_size_t]), "Nfft4GPB200MemcpyD2H": (i, [vp, vp, C.c_size_t]),
        "Nfft4GPB200DeviceSynchronize": (i, []),
        "Nfft4GPB200EventCreate": (vp, []), "Nfft4GPB200EventDestroy": (None, [vp]), "Nfft4GPB200EventRecord": (i, [vp]),
        "Nfft4GPB200EventElapsedMs": (C.c_float, [vp, vp]),
        "Nfft4GPB200AdditiveNFFTMatSymvDevice": (i, [vp, i, d, vp, d, vp]),
        "Nfft4GPB200AdditiveNFFTGradMatSymvDevice": (i, [vp, i, d, vp, d, vp]),
        "Nfft4GPB200PrecondSolveDevice": (i, [vp, vp, i, vp, vp]),
        "Nfft4GPB200SolverPcgDevice": (i, [vp, i, vp, vp, vp, vp, vp, i, i, d, c_double_p, C.POINTER(c_double_p), c_int_p, i]),
        "Nfft4GPB200PrecondNysSetFactors": (i, [vp, i, i, c_int_p, c_double_p, c_double_p, d]),
        "Nfft4GPB200PrecondFsaiSetFactors": (i, [vp, i, c_int_p, c_int_p, c_double_p]),
        "Nfft4GPB200PrecondAFNSetFactors": (i, [vp, i, i, c_int_p, c_double_p, c_double_p, c_int_p, c_int_p, c_double_p]),
        "Nfft4GPB200PrecondNysGetFactors": (i, [vp, c_int_p, c_int_p, c_int_p, c_double_p, c_double_p, c_double_p]),
        "Nfft4GPB200PrecondFsaiGetFactors": (i, [vp, c_int_p, c_int_p, c_int_p, c_int_p, c_double_p]),
        "Nfft4GPB200CommGetUniqueId": (i, [vp]), "Nfft4GPB200CommInit": (i, [i, i, vp]), "Nfft4GPB200CommFinalize": (i, []),
        "Nfft4GPB200SetSharding": (i, [vp, i]),
        "Nfft4GPB200SetProbeSharding": (i, [i]),
        "Nfft4GPB200DebugWindowInfo": (i, [vp, i, c_int_p, c_double_p, c_int_p]),
        "Nfft4GPB200DebugSpread": (i, [vp, i, d, c_double_p, c_double_p]),
        "Nfft4GPB200DebugCoeff": (i, [vp, i, i, c_double_p]),
        "Nfft4GPB200DebugGridAfterFFT": (i, [vp, i, d, c_double_p, c_double_p]),
        "Nfft4GPB200SetVariant": (i, [C.c_char_p, i]),
        "Nfft4GPB200MeasureFp64TensorPeak": (d, []),
        "Nfft4GPB200ProfileStart": (None, []),
        "Nfft4GPB200ProfileStop": (i, [c_double_p, C.POINTER(C.c_longlong), i]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here = the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    L._signatures = sig
    L.free = C.CDLL(None).free
    L.free.argtypes = [vp]
    L.free.restype = None
    _lib = L
    return L


def dptr(a):
    return a.ctypes.data_as(c_double_p)


def iptr(a):
    return a.ctypes.data_as(c_int_p)


class AdditiveNFFTKernel:
    """The reference call sequence for the additive NFFT kernel (TESTS/TEST1/foo.cpp:216-251):
    ParamCreate -> write f, l, mu into the handle -> {Gaussian,Matern12}Kernel -> MatSymv / GradMatSymv -> Free.
    X is feature-major (d, n) like the reference's data files; windows flat int32 with -1 padding the last window."""

    def __init__(self, X, windows, nwindows, dwindows, kernel="gaussian"):
        L = lib()
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.d, self.n = self.X.shape
        self.windows = np.ascontiguousarray(windows, dtype=np.int32)
        self.kernel = kernel
        self.h = L.Nfft4GPNFFTAdditiveKernelParamCreate(dptr(self.X), self.n, self.n, self.d, iptr(self.windows), nwindows, dwindows)
        if not self.h:
            raise RuntimeError("Nfft4GPNFFTAdditiveKernelParamCreate failed")
        self.K = C.c_void_p()
        self.dK = C.c_void_p()

    def build(self, f, l, mu):
        L = lib()
        k = Nfft4GPKernel.from_address(self.h)
        k._params[0] = f
        k._params[1] = l
        k._noise_level = mu
        fn = L.Nfft4GPNFFTAdditiveKernelGaussianKernel if self.kernel == "gaussian" else L.Nfft4GPNFFTAdditiveKernelMatern12Kernel
        self.K, self.dK = C.c_void_p(), C.c_void_p()
        rc = fn(self.h, dptr(self.X), self.n, self.n, self.d, None, 0, None, 0, C.byref(self.K), C.byref(self.dK))
        if rc != 0:
            raise RuntimeError("NFFT additive kernel build failed (no CUDA device?)")
        return self

    def matvec(self, x, alpha=1.0, beta=0.0, y=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.n) if y is None else np.array(y, dtype=np.float64)
        if lib().Nfft4GPAdditiveNFFTMatSymv(self.K, self.n, alpha, dptr(x), beta, dptr(y)) != 0:
            raise RuntimeError("Nfft4GPAdditiveNFFTMatSymv failed")
        return y

    def grad_matvec(self, x, alpha=1.0, beta=0.0, y=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(3 * self.n) if y is None else np.array(y, dtype=np.float64)
        if lib().Nfft4GPAdditiveNFFTGradMatSymv(self.dK, self.n, alpha, dptr(x), beta, dptr(y)) != 0:
            raise RuntimeError("Nfft4GPAdditiveNFFTGradMatSymv failed")
        return y

    def matvec_device(self, d_x, d_y, alpha=1.0, beta=0.0, n_local=None):
        """x, y are raw device pointers (ints), e.g. torch_tensor.data_ptr(); n_local = length of this rank's block under
        point-block sharding."""
        if lib().Nfft4GPB200AdditiveNFFTMatSymvDevice(self.K, self.n if n_local is None else n_local, alpha, d_x, beta, d_y) != 0:
            raise RuntimeError("Nfft4GPB200AdditiveNFFTMatSymvDevice failed")

    def window_info(self, w):
        dim, scale, nseg = C.c_int(), C.c_double(), C.c_int()
        assert lib().Nfft4GPB200DebugWindowInfo(self.h, w, C.byref(dim), C.byref(scale), C.byref(nseg)) == 0
        return dim.value, scale.value, nseg.value

    def debug_spread(self, w, x, alpha=1.0):
        dim, _, _ = self.window_info(w)
        g = np.zeros(64 ** dim)
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert lib().Nfft4GPB200DebugSpread(self.h, w, alpha, dptr(x), dptr(g)) == 0
        return g.reshape((64,) * dim)

    def debug_grid_after_fft(self, w, x, alpha=1.0):
        dim, _, _ = self.window_info(w)
        g = np.zeros(64 ** dim)
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert lib().Nfft4GPB200DebugGridAfterFFT(self.h, w, alpha, dptr(x), dptr(g)) == 0
        return g.reshape((64,) * dim)

    def debug_coeff(self, w, which=0):
        dim, _, _ = self.window_info(w)
        c = np.zeros(2 * 32 ** dim)
        assert lib().Nfft4GPB200DebugCoeff(self.h, w, which, dptr(c)) == 0
        c = c.reshape(-1, 2)
        return (c[:, 0] + 1j * c[:, 1]).reshape((32,) * dim)

    def matvec_fn(self):
        return C.cast(lib().Nfft4GPAdditiveNFFTMatSymv, C.c_void_p)

    def free(self):
        if self.h:
            lib().Nfft4GPAdditiveNFFTKernelFree(self.h)
            self.h = None


def _set_hyper(handle, f, l, mu):
    k = Nfft4GPKernel.from_address(handle)
    k._params[0] = f
    k._params[1] = l
    k._noise_level = mu


class DenseAdditiveKernel:
    """Dense additive kernel handle (reference call sequence TESTS/TEST3/foo3.cpp:251-256): the func_kernel the
    preconditioner setups are given.  X feature-major (d, n)."""

    def __init__(self, X, windows, nwindows, dwindows, kernel="gaussian"):
        L = lib()
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.d, self.n = self.X.shape
        self.windows = np.ascontiguousarray(windows, dtype=np.int32)
        base = L.Nfft4GPKernelGaussianKernel if kernel == "gaussian" else L.Nfft4GPKernelMatern12Kernel
        self.h = L.Nfft4GPKernelAdditiveKernelParamCreate(dptr(self.X), self.n, self.n, self.d, iptr(self.windows), nwindows, dwindows,
                                                         C.cast(base, C.c_void_p))
        if not self.h:
            raise RuntimeError("Nfft4GPKernelAdditiveKernelParamCreate failed")

    def set_hyper(self, f, l, mu):
        _set_hyper(self.h, f, l, mu)
        return self

    def fn(self):
        return C.cast(lib().Nfft4GPKernelAdditiveKernel, C.c_void_p)

    def block(self, permr=None, permc=None, grad=False):
        return _kernel_block(lib().Nfft4GPKernelAdditiveKernel, self.h, self.X, permr, permc, grad)

    def free(self):
        if self.h:
            lib().Nfft4GPKernelParamFree(self.h)
            self.h = None


class DensePlainKernel:
    """Plain Gaussian / Matern-1/2 kernel over all features of X (kernels.c:680 / 2390)."""

    def __init__(self, X, kernel="gaussian"):
        L = lib()
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.d, self.n = self.X.shape
        self._fn = L.Nfft4GPKernelGaussianKernel if kernel == "gaussian" else L.Nfft4GPKernelMatern12Kernel
        self.h = L.Nfft4GPKernelParamCreate(self.n, 1)

    def set_hyper(self, f, l, mu):
        _set_hyper(self.h, f, l, mu)
        return self

    def fn(self):
        return C.cast(self._fn, C.c_void_p)

    def block(self, permr=None, permc=None, grad=False):
        return _kernel_block(self._fn, self.h, self.X, permr, permc, grad)

    def free(self):
        if self.h:
            lib().Nfft4GPKernelParamFree(self.h)
            self.h = None


def _kernel_block(fn, h, X, permr, permc, grad=False):
    """func_kernel call shapes of kernels.c:701-991; returns the column-major block as an (nr, nc) array, with grad=True
    also the stacked gradient blocks [d/df, d/dl, d/dmu] as a (3, nr, nc) array."""
    d, n = X.shape
    K = C.c_void_p()
    dK = C.c_void_p()
    pr = None if permr is None else np.ascontiguousarray(permr, dtype=np.int32)
    pc = None if permc is None else np.ascontiguousarray(permc, dtype=np.int32)
    rc = fn(h, dptr(X), n, n, d, iptr(pr) if pr is not None else None, 0 if pr is None else pr.size,
            iptr(pc) if pc is not None else None, 0 if pc is None else pc.size, C.byref(K), C.byref(dK) if grad else None)
    if rc != 0:
        raise RuntimeError("dense kernel evaluation failed")
    nr = n if pr is None else pr.size
    nc = nr if pc is None else pc.size
    out = np.ctypeslib.as_array(C.cast(K, c_double_p), shape=(nc, nr)).T.copy()
    lib().Nfft4GPKernelFree(K)
    if grad:
        dout = np.ctypeslib.as_array(C.cast(dK, c_double_p), shape=(3, nc, nr)).transpose(0, 2, 1).copy()
        lib().Nfft4GPKernelFree(dK)
        return out, dout
    return out


class CholPrecond:
    """Dense Cholesky preconditioner (SRC/preconds/chol.c)."""

    def __init__(self, stable=0):
        self.h = lib().Nfft4GPPrecondCholCreate()
        lib().Nfft4GPPrecondCholSetStable(self.h, stable)

    def setup(self, kernel, require_grad=0):
        rc = lib().Nfft4GPPrecondCholSetupWithKernel(dptr(kernel.X), kernel.n, kernel.n, kernel.d, kernel.fn(), kernel.h, require_grad, self.h)
        if rc != 0:
            raise RuntimeError("Nfft4GPPrecondCholSetupWithKernel failed")
        self.n = kernel.n
        return self

    def solve(self, rhs):
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        x = np.zeros_like(rhs)
        if lib().Nfft4GPPrecondCholSolve(self.h, rhs.size, dptr(x), dptr(rhs)) != 0:
            raise RuntimeError("Nfft4GPPrecondCholSolve failed")
        return x

    def logdet(self):
        return lib().Nfft4GPPrecondCholLogdet(self.h)

    def trace(self):
        tr = np.zeros(3)
        p = dptr(tr)
        if lib().Nfft4GPPrecondCholTrace(self.h, C.byref(p)) != 0:
            raise RuntimeError("Nfft4GPPrecondCholTrace failed")
        return tr

    def dvp(self, x, mask=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(3 * x.size)
        p = dptr(y)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.int32)
        if lib().Nfft4GPPrecondCholDvp(self.h, x.size, iptr(m) if m is not None else None, dptr(x), C.byref(p)) != 0:
            raise RuntimeError("Nfft4GPPrecondCholDvp failed")
        return y.reshape(3, x.size)

    def solve_fn(self):
        return C.cast(lib().Nfft4GPPrecondCholSolve, C.c_void_p)

    def free(self):
        if self.h:
            lib().Nfft4GPPrecondCholFree(self.h)
            self.h = None


class NysPrecond:
    def __init__(self):
        self.h = lib().Nfft4GPPrecondNysCreate()
        self._keep = []

    def logdet(self):
        return lib().Nfft4GPPrecondNysLogdet(self.h)

    def trace(self):
        tr = np.zeros(3)
        p = dptr(tr)
        if lib().Nfft4GPPrecondNysTrace(self.h, C.byref(p)) != 0:
            raise RuntimeError("Nfft4GPPrecondNysTrace failed")
        return tr

    def dvp(self, x, mask=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(3 * x.size)
        p = dptr(y)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.int32)
        if lib().Nfft4GPPrecondNysDvp(self.h, x.size, iptr(m) if m is not None else None, dptr(x), C.byref(p)) != 0:
            raise RuntimeError("Nfft4GPPrecondNysDvp failed")
        return y.reshape(3, x.size)

    def setup(self, kernel, perm, k, require_grad=0):
        """Nfft4GPPrecondNysSetPerm / SetRank / SetupWithKernel (TESTS/TEST3/foo3.cpp:262-265, nys.c:518-660)."""
        L = lib()
        self._perm = None if perm is None else np.ascontiguousarray(perm, dtype=np.int32)
        if self._perm is not None:
            L.Nfft4GPPrecondNysSetPerm(self.h, iptr(self._perm), 0)
        L.Nfft4GPPrecondNysSetRank(self.h, k)
        rc = L.Nfft4GPPrecondNysSetupWithKernel(dptr(kernel.X), kernel.n, kernel.n, kernel.d, kernel.fn(), kernel.h, require_grad, self.h)
        if rc != 0:
            raise RuntimeError("Nfft4GPPrecondNysSetupWithKernel failed")
        self.n = kernel.n
        return self

    def get_factors(self):
        L = lib()
        n, k, eta = C.c_int(), C.c_int(), C.c_double()
        assert L.Nfft4GPB200PrecondNysGetFactors(self.h, C.byref(n), C.byref(k), None, None, None, C.byref(eta)) == 0
        perm = np.zeros(n.value, dtype=np.int32)
        U = np.zeros((k.value, n.value))
        s = np.zeros(k.value)
        assert L.Nfft4GPB200PrecondNysGetFactors(self.h, None, None, iptr(perm), dptr(U), dptr(s), None) == 0
        return perm, U.T.copy(), s, eta.value

    def set_factors(self, perm, U, s, eta):
        n, k = U.shape
        Uf = np.asfortranarray(U, dtype=np.float64)
        perm = None if perm is None else np.ascontiguousarray(perm, dtype=np.int32)
        s = np.ascontiguousarray(s, dtype=np.float64)
        rc = lib().Nfft4GPB200PrecondNysSetFactors(self.h, n, k, iptr(perm) if perm is not None else None,
                                                   Uf.ctypes.data_as(c_double_p), dptr(s), eta)
        if rc != 0:
            raise RuntimeError("Nfft4GPB200PrecondNysSetFactors failed")
        self.n = n

    def solve(self, rhs):
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        x = np.zeros_like(rhs)
        if lib().Nfft4GPPrecondNysSolve(self.h, rhs.size, dptr(x), dptr(rhs)) != 0:
            raise RuntimeError("Nfft4GPPrecondNysSolve failed")
        return x

    def solve_fn(self):
        return C.cast(lib().Nfft4GPPrecondNysSolve, C.c_void_p)

    def free(self):
        if self.h:
            lib().Nfft4GPPrecondNysFree(self.h)
            self.h = None


class FsaiPrecond:
    def __init__(self):
        self.h = lib().Nfft4GPPrecondFsaiCreate()

    def setup(self, kernel, lfil, require_grad=0, pattern=None):
        """Nfft4GPPrecondFsaiSetLfil / SetupWithKernel (fsai.c:100-104, 302-673); pattern = (ia, ja): SetupWithKernelPattern (:314)."""
        L = lib()
        L.Nfft4GPPrecondFsaiSetLfil(self.h, lfil)
        if pattern is None:
            rc = L.Nfft4GPPrecondFsaiSetupWithKernel(dptr(kernel.X), kernel.n, kernel.n, kernel.d, kernel.fn(), kernel.h, require_grad, self.h)
        else:
            # the callee takes ownership of the two arrays (the reference keeps them as its CSR pattern): hand it malloc'd copies
            libc = C.CDLL(None)
            libc.malloc.restype = C.c_void_p
            libc.malloc.argtypes = [C.c_size_t]
            ia = np.ascontiguousarray(pattern[0], dtype=np.int32)
            ja = np.ascontiguousarray(pattern[1], dtype=np.int32)
            pi, pj = libc.malloc(ia.nbytes), libc.malloc(max(ja.nbytes, 4))
            C.memmove(pi, ia.ctypes.data, ia.nbytes)
            C.memmove(pj, ja.ctypes.data, ja.nbytes)
            rc = L.Nfft4GPPrecondFsaiSetupWithKernelPattern(pi, pj, dptr(kernel.X), kernel.n, kernel.n, kernel.d, kernel.fn(), kernel.h,
                                                            require_grad, self.h)
        if rc != 0:
            raise RuntimeError("FSAI setup failed")
        return self

    def grad_factors(self):
        ia, ja, a = self.get_factors()
        da = np.zeros((3, a.size))
        assert lib().Nfft4GPB200PrecondFsaiGetGradFactors(self.h, dptr(da)) == 0
        return da

    def trace(self):
        tr = np.zeros(3)
        p = dptr(tr)
        if lib().Nfft4GPPrecondFsaiTrace(self.h, C.byref(p)) != 0:
            raise RuntimeError("Nfft4GPPrecondFsaiTrace failed")
        return tr

    def dvp(self, x, mask=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(3 * x.size)
        p = dptr(y)
        m = None if mask is None else iptr(np.ascontiguousarray(mask, dtype=np.int32))
        if lib().Nfft4GPPrecondFsaiDvp(self.h, x.size, m, dptr(x), C.byref(p)) != 0:
            raise RuntimeError("Nfft4GPPrecondFsaiDvp failed")
        return y

    def inv_l(self, rhs, transpose=False):
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        x = np.zeros_like(rhs)
        fn = lib().Nfft4GPPrecondFsaiInvLT if transpose else lib().Nfft4GPPrecondFsaiInvL
        if fn(self.h, rhs.size, dptr(x), dptr(rhs)) != 0:
            raise RuntimeError("FSAI triangular solve failed")
        return x

    def get_factors(self):
        L = lib()
        n, nnz = C.c_int(), C.c_int()
        assert L.Nfft4GPB200PrecondFsaiGetFactors(self.h, C.byref(n), C.byref(nnz), None, None, None) == 0
        ia = np.zeros(n.value + 1, dtype=np.int32)
        ja = np.zeros(nnz.value, dtype=np.int32)
        a = np.zeros(nnz.value)
        assert L.Nfft4GPB200PrecondFsaiGetFactors(self.h, None, None, iptr(ia), iptr(ja), dptr(a)) == 0
        return ia, ja, a

    def set_factors(self, ia, ja, a):
        ia = np.ascontiguousarray(ia, dtype=np.int32)
        ja = np.ascontiguousarray(ja, dtype=np.int32)
        a = np.ascontiguousarray(a, dtype=np.float64)
        if lib().Nfft4GPB200PrecondFsaiSetFactors(self.h, ia.size - 1, iptr(ia), iptr(ja), dptr(a)) != 0:
            raise RuntimeError("Nfft4GPB200PrecondFsaiSetFactors failed")

    def solve(self, rhs):
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        x = np.zeros_like(rhs)
        if lib().Nfft4GPPrecondFsaiSolve(self.h, rhs.size, dptr(x), dptr(rhs)) != 0:
            raise RuntimeError("Nfft4GPPrecondFsaiSolve failed")
        return x

    def solve_fn(self):
        return C.cast(lib().Nfft4GPPrecondFsaiSolve, C.c_void_p)

    def free(self):
        if self.h:
            lib().Nfft4GPPrecondFsaiFree(self.h)
            self.h = None


class AfnPrecond:
    def __init__(self):
        self.h = lib().Nfft4GPPrecondAFNCreate()

    def set_k12_storage(self, bits):
        """Nfft4GPB200PrecondAFNSetK12Storage: 32 = keep the K12 block in fp32 (halves the HBM traffic of the apply)."""
        if lib().Nfft4GPB200PrecondAFNSetK12Storage(self.h, bits) != 0:
            raise RuntimeError("Nfft4GPB200PrecondAFNSetK12Storage failed")
        return self

    def set_k12_operator(self, nfft_kernel, release_block=False):
        """Nfft4GPB200PrecondAFNSetK12Operator: the K12 products of the apply through the NFFT operator `nfft_kernel`
        (an AdditiveNFFTKernel built with the hyper-parameters of the setup) instead of the dense block; None = dense block."""
        if lib().Nfft4GPB200PrecondAFNSetK12Operator(self.h, nfft_kernel.K if nfft_kernel is not None else None, int(release_block)) != 0:
            raise RuntimeError("Nfft4GPB200PrecondAFNSetK12Operator failed")
        return self

    def setup(self, kernel, max_k, lfil, est_m=500, numests=5, noran=0, seed=906, rank_method=0):
        """Nfft4GPPrecondAFNSetRank / SetLfil / SetupWithKernel (afn.h:65-132; MATLAB afn_setup.m:30-109).
        rank_method 1 = first stage of the rank estimation as the reference's compiled Nfft4GPRankestNysScaled."""
        L = lib()
        assert L.Nfft4GPB200PrecondAFNSetRankMethod(self.h, rank_method) == 0
        L.Nfft4GPPrecondAFNSetRank(self.h, max_k)
        L.Nfft4GPPrecondAFNSetLfil(self.h, lfil)
        assert L.Nfft4GPB200PrecondAFNSetRankEstimation(self.h, est_m, numests, noran, seed) == 0
        rc = L.Nfft4GPPrecondAFNSetupWithKernel(dptr(kernel.X), kernel.n, kernel.n, kernel.d, kernel.fn(), kernel.h, 0, self.h)
        if rc != 0:
            raise RuntimeError("Nfft4GPPrecondAFNSetupWithKernel failed")
        return self

    def get_factors(self):
        """dict(n, k, k_est, is_nys, perm[, Linv, K12, ia, ja, a])"""
        L = lib()
        n, k, ke, isn, nnz = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        assert L.Nfft4GPB200PrecondAFNGetFactors(self.h, C.byref(n), C.byref(k), C.byref(ke), C.byref(isn), None, None, None,
                                                 C.byref(nnz), None, None, None) == 0
        out = dict(n=n.value, k=k.value, k_est=ke.value, is_nys=bool(isn.value))
        perm = np.zeros(n.value, dtype=np.int32)
        if isn.value:
            assert L.Nfft4GPB200PrecondAFNGetFactors(self.h, None, None, None, None, iptr(perm), None, None, None, None, None, None) == 0
            out["perm"] = perm
            return out
        n2 = n.value - k.value
        Linv = np.zeros((k.value, k.value))
        K12 = np.zeros((n2, k.value))
        ia = np.zeros(n2 + 1, dtype=np.int32)
        ja = np.zeros(nnz.value, dtype=np.int32)
        a = np.zeros(nnz.value)
        assert L.Nfft4GPB200PrecondAFNGetFactors(self.h, None, None, None, None, iptr(perm), dptr(Linv), dptr(K12), None,
                                                 iptr(ia), iptr(ja), dptr(a)) == 0
        out.update(perm=perm, Linv=Linv.T.copy(), K12=K12.T.copy(), ia=ia, ja=ja, a=a)
        return out

    def set_factors(self, perm, L, K12, ia, ja, a):
        k = L.shape[0]
        n = perm.size
        perm = np.ascontiguousarray(perm, dtype=np.int32)
        Lf = np.asfortranarray(L, dtype=np.float64)
        Kf = np.asfortranarray(K12, dtype=np.float64)
        ia = np.ascontiguousarray(ia, dtype=np.int32)
        ja = np.ascontiguousarray(ja, dtype=np.int32)
        a = np.ascontiguousarray(a, dtype=np.float64)
        rc = lib().Nfft4GPB200PrecondAFNSetFactors(self.h, n, k, iptr(perm), Lf.ctypes.data_as(c_double_p),
                                                   Kf.ctypes.data_as(c_double_p), iptr(ia), iptr(ja), dptr(a))
        if rc != 0:
            raise RuntimeError("Nfft4GPB200PrecondAFNSetFactors failed")

    def solve(self, rhs):
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        x = np.zeros_like(rhs)
        if lib().Nfft4GPPrecondAFNSolve(self.h, rhs.size, dptr(x), dptr(rhs)) != 0:
            raise RuntimeError("Nfft4GPPrecondAFNSolve failed")
        return x

    def solve_fn(self):
        return C.cast(lib().Nfft4GPPrecondAFNSolve, C.c_void_p)

    def logdet(self):
        return float(lib().Nfft4GPPrecondAFNLogdet(self.h))

    def free(self):
        if self.h:
            lib().Nfft4GPPrecondAFNFree(self.h)
            self.h = None


def rankest_nys_scaled_sample(kernel, rows, max_rank):
    """Nfft4GPB200RankestNysScaledSample: one sample of the reference's Nfft4GPRankestNysScaledEstimateRank (rankest.c:248-352)
    on the given sample rows, for one of this library's dense kernel handles (DensePlainKernel / DenseAdditiveKernel)."""
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    return int(lib().Nfft4GPB200RankestNysScaledSample(dptr(kernel.X), kernel.n, kernel.n, kernel.d, kernel.fn(), kernel.h, iptr(rows),
                                                       rows.size, max_rank))


def solver_pcg(mat, n, matvec_fn, prec, solve_fn, x0, rhs, maxits, atol, tol, print_level=0):
    """Nfft4GPSolverPcg with host vectors.  Returns (x, rel_res, rel_res_v, iters)."""
    L = lib()
    x = np.array(x0, dtype=np.float64)
    rhs = np.ascontiguousarray(rhs, dtype=np.float64)
    rel = C.c_double(0.0)
    relv = c_double_p()
    it = C.c_int(0)
    rc = L.Nfft4GPSolverPcg(mat, n, matvec_fn, prec, solve_fn, dptr(x), dptr(rhs), maxits, atol, tol,
                            C.byref(rel), C.byref(relv), C.byref(it), print_level)
    if rc != 0:
        raise RuntimeError("Nfft4GPSolverPcg failed")
    m = min(maxits, n) + 1
    if it.value == 0 and not atol and rel.value <= tol:
        m = 1  # early exits of pcg.c:32-41 / 70-84 allocate a single entry
    hist = np.ctypeslib.as_array(relv, shape=(m,)).copy() if relv else np.zeros(0)
    if relv:
        L.free(relv)
    return x, rel.value, hist, it.value


def solver_fgmres(mat, n, matvec_fn, prec, solve_fn, x0, rhs, kdim, maxits, atol, tol, print_level=1):
    """Nfft4GPSolverFgmres with host vectors.  Returns (x, rel_res, rel_res_v, iters).  print_level defaults to 1 because the
    reference prints its end-of-cycle message at level 0 (fgmres.c:198-201); use -1 for silence."""
    L = lib()
    x = np.array(x0, dtype=np.float64)
    rhs = np.ascontiguousarray(rhs, dtype=np.float64)
    rel = C.c_double(0.0)
    relv = c_double_p()
    it = C.c_int(0)
    rc = L.Nfft4GPSolverFgmres(mat, n, matvec_fn, prec, solve_fn, dptr(x), dptr(rhs), kdim, maxits, atol, tol,
                               C.byref(rel), C.byref(relv), C.byref(it), print_level)
    if rc != 0:
        raise RuntimeError("Nfft4GPSolverFgmres failed")
    m = maxits + 1 if it.value > 0 else 1
    hist = np.ctypeslib.as_array(relv, shape=(m,)).copy() if relv else np.zeros(0)
    if relv:
        L.free(relv)
    return x, rel.value, hist, it.value


def gp_predict(x, X1, y, X2, windows, nwindows, dwindows, kernel="gaussian", precond=None, dense=None, tol=1e-8, maxits=200,
               transform=0, want_std=True, print_level=1):
    """Nfft4GPAdditiveNFFTGpPredict as TESTS/TEST3/foo3.cpp and TEST4/foo.cpp call it: NFFT handle on the training points, NFFT
    handle on [train; test], optional preconditioner (object with .h / .solve_fn() and a setup function pointer) built from the dense
    kernel handle `dense`.  Returns (mean, std or None)."""
    L = lib()
    X1 = np.ascontiguousarray(X1, dtype=np.float64)
    X2 = np.ascontiguousarray(X2, dtype=np.float64)
    d, n = X1.shape
    n2 = X2.shape[1]
    xall_p = L.Nfft4GPNFFTAppendData(dptr(X1), n, n, d, dptr(X2), n2, n2)
    Xall = np.ctypeslib.as_array(xall_p, shape=(d, n + n2))
    win = np.ascontiguousarray(windows, dtype=np.int32)
    A = AdditiveNFFTKernel(X1, win, nwindows, dwindows, kernel)
    Al = AdditiveNFFTKernel(Xall, win, nwindows, dwindows, kernel)
    fk = L.Nfft4GPNFFTAdditiveKernelGaussianKernel if kernel == "gaussian" else L.Nfft4GPNFFTAdditiveKernelMatern12Kernel
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    lab, std = c_double_p(), c_double_p()
    pargs = [None, None, None, None, None, None]
    if precond is not None:
        setup_fn, P = precond
        pargs = [dense.fn(), dense.h, C.cast(L.Nfft4GPKernelParamFree, C.c_void_p), C.cast(setup_fn, C.c_void_p), P.solve_fn(), P.h]
    rc = L.Nfft4GPAdditiveNFFTGpPredict(dptr(x), dptr(X1), dptr(y), n, n, d, dptr(X2), n2, n2, xall_p, C.cast(fk, C.c_void_p), A.h, Al.h,
                                        C.cast(L.Nfft4GPNFFTKernelFree, C.c_void_p), C.cast(L.Nfft4GPAdditiveNFFTMatSymv, C.c_void_p), *pargs,
                                        0, tol, maxits, transform, print_level, None, C.byref(lab), C.byref(std) if want_std else None)
    if rc != 0:
        raise RuntimeError("Nfft4GPAdditiveNFFTGpPredict failed")
    mean = np.ctypeslib.as_array(lab, shape=(n2,)).copy()
    sd = np.ctypeslib.as_array(std, shape=(n2,)).copy() if want_std else None
    L.free(lab)
    if want_std:
        L.free(std)
    A.free()
    Al.free()
    L.free(xall_p)
    return mean, sd


def solver_lanczos(mat, n, matvec_fn, prec, solve_fn, x0, rhs, wsize, maxits, atol, tol, print_level=1):
    """Nfft4GPSolverLanczos with host vectors.  Returns (x, rel_res, rel_res_v, iters, TD, TE) with T of tsize rows."""
    L = lib()
    x = np.array(x0, dtype=np.float64)
    rhs = np.ascontiguousarray(rhs, dtype=np.float64)
    rel = C.c_double(0.0)
    relv, TD, TE = c_double_p(), c_double_p(), c_double_p()
    it, ts = C.c_int(0), C.c_int(0)
    rc = L.Nfft4GPSolverLanczos(mat, n, matvec_fn, prec, solve_fn, dptr(x), dptr(rhs), wsize, maxits, atol, tol, C.byref(rel), C.byref(relv),
                                C.byref(it), C.byref(ts), C.byref(TD), C.byref(TE), print_level)
    if rc != 0:
        raise RuntimeError("Nfft4GPSolverLanczos failed")
    hist = np.ctypeslib.as_array(relv, shape=(maxits + 1,)).copy() if it.value > 0 else np.zeros(1)
    td = np.ctypeslib.as_array(TD, shape=(ts.value,)).copy() if TD else np.zeros(0)
    te = np.ctypeslib.as_array(TE, shape=(max(ts.value - 1, 0),)).copy() if TE else np.zeros(0)
    for ptr in (relv, TD, TE):
        if ptr:
            L.free(ptr)
    return x, rel.value, hist, it.value, td, te


def lanczos_quadrature_logdet(mat, dmat, n, matvec_fn, dmatvec_fn, maxits, probes, print_level=1):
    """Nfft4GPLanczosQuadratureLogdet without a preconditioner, with given Rademacher probes (nvecs, n).  Returns (logdet / n,
    dlogdet[3] / n)."""
    L = lib()
    probes = np.ascontiguousarray(probes, dtype=np.float64)
    nvecs = probes.shape[0]
    val = C.c_double(0.0)
    dval = c_double_p()
    rc = L.Nfft4GPLanczosQuadratureLogdet(mat, dmat, n, matvec_fn, dmatvec_fn, None, None, None, None, None, maxits, nvecs, dptr(probes),
                                          print_level, C.byref(val), C.byref(dval))
    if rc != 0:
        raise RuntimeError("Nfft4GPLanczosQuadratureLogdet failed")
    out = np.ctypeslib.as_array(dval, shape=(3,)).copy()
    L.free(dval)
    return val.value, out
