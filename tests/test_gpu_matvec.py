"""GPU parity of the NFFT additive matvec (rows M1, M2, M3, S1, S2 of SURVEY.md section 8(a)) against the CPU oracle,
called through the C ABI exactly like the reference drivers call nfft4gp (TESTS/TEST1/foo.cpp:216-251).
Tolerance: north_star -- <= 1e-10 relative l2 in fp64 (we observe ~1e-14)."""
import ctypes as C

import numpy as np
import pytest

from common import MATVEC_CASES, gaussian_cloud, rand_vec, rel_l2, uniform_cloud
from oracle import oracle_api as O

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _pair(gpu, case, cloud=uniform_cloud):
    name, n, d, win, nw, dw, kern, f, l, mu = case
    X = cloud(n, d)
    win = np.array(win, dtype=np.int32)
    A = gpu.AdditiveNFFTKernel(X, win, nw, dw, kern).build(f, l, mu)
    B = O.OracleAdditiveNFFT(X, win, nw, dw, kern).build(f, l, mu)
    return A, B, n


@pytest.mark.parametrize("case", MATVEC_CASES, ids=[c[0] for c in MATVEC_CASES])
@pytest.mark.parametrize("variant", [-1, 5, 1, 0], ids=["auto", "dmma", "dfma", "generic"])
def test_matvec_matches_oracle(gpu, case, variant):
    L = gpu.lib()
    L.Nfft4GPB200SetVariant(b"spread3d", variant)
    L.Nfft4GPB200SetVariant(b"gather3d", variant)
    try:
        A, B, n = _pair(gpu, case)
        x = rand_vec(n)
        y0 = rand_vec(n, seed=5)
        for alpha, beta in [(1.0, 0.0), (-1.0, 1.0), (0.7, 0.3)]:
            yg = A.matvec(x, alpha, beta, y0)
            yo = B.matvec(x, alpha, beta, y0)
            assert rel_l2(yg, yo) <= TOL, (case[0], alpha, beta, rel_l2(yg, yo))
        A.free()
        B.free()
    finally:
        L.Nfft4GPB200SetVariant(b"spread3d", -1)
        L.Nfft4GPB200SetVariant(b"gather3d", -1)


@pytest.mark.parametrize("cloud", [uniform_cloud, gaussian_cloud], ids=["uniform", "gaussian"])
def test_dense_windows_fused_tensor_core_path(gpu, cloud):
    """n large enough (>= 24 points per occupied cell) for the automatic choice to take the fused fp64 tensor-core launches
    (all three windows in one persistent wave): matvec and gradient matvec (shared adjoint, second gather channel)."""
    case = ("dense", 250000, 9, list(range(9)), 3, 3, "gaussian", 1.0, 0.25 * 250000 ** (1.0 / 9), 0.01)
    A, B, n = _pair(gpu, case, cloud=cloud)
    for w in range(3):
        dim, scale, nseg = A.window_info(w)
        assert n / nseg >= 24, "case is not dense enough to exercise the fused path"
    x = rand_vec(n)
    y0 = rand_vec(n, seed=5)
    for alpha, beta in [(1.0, 0.0), (0.7, 0.3)]:
        assert rel_l2(A.matvec(x, alpha, beta, y0), B.matvec(x, alpha, beta, y0)) <= TOL
    yg, yo = A.grad_matvec(x), B.grad_matvec(x)
    for blk in range(3):
        sl = slice(blk * n, (blk + 1) * n)
        assert rel_l2(yg[sl], yo[sl]) <= TOL, blk
    # same handle, explicit per-window variants must agree with the fused result
    L = gpu.lib()
    y_auto = A.matvec(x)
    try:
        for v in (5, 1):
            L.Nfft4GPB200SetVariant(b"spread3d", v)
            L.Nfft4GPB200SetVariant(b"gather3d", v)
            assert rel_l2(A.matvec(x), y_auto) <= 1e-13
    finally:
        L.Nfft4GPB200SetVariant(b"spread3d", -1)
        L.Nfft4GPB200SetVariant(b"gather3d", -1)
    A.free()
    B.free()


@pytest.mark.parametrize("case", MATVEC_CASES[:7], ids=[c[0] for c in MATVEC_CASES[:7]])
def test_grad_matvec_matches_oracle(gpu, case):
    A, B, n = _pair(gpu, case)
    x = rand_vec(n)
    y0 = rand_vec(3 * n, seed=6)
    for alpha, beta in [(1.0, 0.0), (0.5, 1.0), (-0.3, 0.4)]:
        yg = A.grad_matvec(x, alpha, beta, y0)
        yo = B.grad_matvec(x, alpha, beta, y0)
        for blk in range(3):
            sl = slice(blk * n, (blk + 1) * n)
            assert rel_l2(yg[sl], yo[sl]) <= TOL, (case[0], blk, alpha, beta)
    A.free()
    B.free()


def test_gaussian_cloud_and_rebuild(gpu):
    """Clustered points (heavy cells -> split segments) and a hyper-parameter rebuild on the same handle."""
    case = ("gauss_cloud", 20000, 9, list(range(9)), 3, 3, "gaussian", 1.0, 1.0, 0.01)
    A, B, n = _pair(gpu, case, cloud=gaussian_cloud)
    x = rand_vec(n)
    assert rel_l2(A.matvec(x), B.matvec(x)) <= TOL
    for (f, l, mu) in [(0.7, 0.5, 0.2), (1.5, 3.0, 1e-3)]:
        A.build(f, l, mu)
        B.build(f, l, mu)
        assert rel_l2(A.matvec(x), B.matvec(x)) <= TOL
    A.free()
    B.free()


def test_window_scale_is_bit_identical(gpu):
    """The per-window scale factor (nfft_interface.c:162-196) is computed with the reference's arithmetic."""
    case = MATVEC_CASES[2]
    A, B, n = _pair(gpu, case)
    for w in range(3):
        dim, scale, nseg = A.window_info(w)
        assert scale == B.window_scale(w)
        assert dim == 3 and nseg > 0
    A.free()
    B.free()


def test_stages_against_numpy(gpu):
    """Stage taps: spread grid vs a direct numpy spread; Fourier coefficient table vs the oracle's b_k / phi_hat^2."""
    n, d = 800, 3
    X = uniform_cloud(n, d)
    win = np.arange(3, dtype=np.int32)
    A = gpu.AdditiveNFFTKernel(X, win, 1, 3, "gaussian").build(1.0, 1.0, 0.01)
    B = O.OracleAdditiveNFFT(X, win, 1, 3, "gaussian").build(1.0, 1.0, 0.01)
    x = rand_vec(n)
    g = A.debug_spread(0, x)
    pts = B.window_points(0)
    L = O.lib()
    ref = np.zeros((64, 64, 64))
    for j in range(n):
        c = np.floor(pts[j] * 64).astype(int)
        psi = [[L.shim_kb_phi(pts[j, t] - (c[t] - 4 + lj) / 64.0, 64, 4) for lj in range(10)] for t in range(3)]
        i0 = [(c[0] - 4 + lj) % 64 for lj in range(10)]
        i1 = [(c[1] - 4 + lj) % 64 for lj in range(10)]
        i2 = [(c[2] - 4 + lj) % 64 for lj in range(10)]
        ref[np.ix_(i0, i1, i2)] += x[j] * np.einsum("a,b,c->abc", psi[0], psi[1], psi[2])
    assert rel_l2(g, ref) <= 1e-12
    # coefficient table
    bh = B.window_bhat(0, 0)
    ph = np.array([1.0 / L.shim_kb_phi_hut(k - 16, 64, 4) for k in range(32)])
    scale = np.einsum("a,b,c->abc", ph, ph, ph) ** 2
    assert rel_l2(A.debug_coeff(0, 0), bh * scale) <= 1e-12
    # grid after scaling + inverse FFT vs numpy FFTs of the device's own spread grid
    h = A.debug_grid_after_fft(0, x)
    G = np.fft.ifftn(g) * g.size  # e^{+}
    k = np.r_[0:16, 48:64]
    kk = np.r_[16:32, 0:16]
    band = G[np.ix_(k, k, k)]
    coef = A.debug_coeff(0, 0)[np.ix_(kk, kk, kk)]
    Hh = np.zeros((64, 64, 64), dtype=complex)
    Hh[np.ix_(k, k, k)] = band * coef
    href = np.real(np.fft.fftn(Hh))
    assert rel_l2(h, href) <= 1e-12
    A.free()
    B.free()


# ------------------------------------------------------------------------------------------------ single-window entry points
def _single_build(L, capi, h, X, kernel, f, l, mu):
    k = capi.Nfft4GPKernel.from_address(h)
    k._params[0], k._params[1], k._noise_level = f, l, mu
    K, dK = C.c_void_p(), C.c_void_p()
    fn = L.Nfft4GPNFFTKernelGaussianKernel if kernel == "gaussian" else L.Nfft4GPNFFTKernelMatern12Kernel
    d, n = X.shape
    assert fn(h, capi.dptr(X), n, n, d, None, 0, None, 0, C.byref(K), C.byref(dK)) == 0
    return K, dK


@pytest.mark.parametrize("dim,kernel,n", [(3, "gaussian", 2500), (2, "matern12", 1800), (1, "gaussian", 1500), (3, "matern12", 40000)])
def test_single_window_entry_points(gpu, dim, kernel, n):
    """Rows M1 / M3 through their own entry points (nfft_interface.c:3-161, 400-620): Nfft4GPNFFTKernelParamCreate,
    Nfft4GPNFFTKernel{Gaussian,Matern12}Kernel, Nfft4GPNFFTMatSymv, Nfft4GPNFFTGradMatSymv against the oracle (an additive
    kernel with ONE window is the same operator), for all (alpha, beta) branches of :428-494; then
    Nfft4GPNFFTKernelParamRemovePoints + a rebuild on OTHER data of another size (nfft_interface.c:101-128)."""
    L, capi = gpu.lib(), gpu.capi
    f, l, mu = 1.1, 0.3 * n ** (1.0 / dim) / 4.0 + 0.5, 0.02
    h = L.Nfft4GPNFFTKernelParamCreate(2 * n, dim)
    win = np.arange(dim, dtype=np.int32)
    for rebuild, nn in enumerate((n, n + n // 3)):
        X = uniform_cloud(nn, dim, seed=40 + rebuild)
        K, dK = _single_build(L, capi, h, X, kernel, f, l, mu)
        B = O.OracleAdditiveNFFT(X, win, 1, dim, kernel).build(f, l, mu)
        x = rand_vec(nn, 5)
        y0, g0 = rand_vec(nn, 6), rand_vec(3 * nn, 7)
        for al, be in ((1.0, 0.0), (-1.0, 1.0), (0.7, 0.3)):
            y = y0.copy()
            assert L.Nfft4GPNFFTMatSymv(K, nn, al, capi.dptr(x), be, capi.dptr(y)) == 0
            assert rel_l2(y, B.matvec(x, al, be, y0)) <= 1e-10, (dim, kernel, al, be)
            g = g0.copy()
            assert L.Nfft4GPNFFTGradMatSymv(dK, nn, al, capi.dptr(x), be, capi.dptr(g)) == 0
            assert rel_l2(g, B.grad_matvec(x, al, be, g0)) <= 1e-10, (dim, kernel, al, be)
        # a new length scale on the same points: only the coefficient tables are rebuilt
        K, dK = _single_build(L, capi, h, X, kernel, f, 1.7 * l, mu)
        B.build(f, 1.7 * l, mu)
        y = np.zeros(nn)
        assert L.Nfft4GPNFFTMatSymv(K, nn, 1.0, capi.dptr(x), 0.0, capi.dptr(y)) == 0
        assert rel_l2(y, B.matvec(x)) <= 1e-10
        # a vector of the wrong length is refused (the handle was built for nn points)
        assert L.Nfft4GPNFFTMatSymv(K, nn - 1, 1.0, capi.dptr(x), 0.0, capi.dptr(y)) == -1
        B.free()
        assert L.Nfft4GPNFFTKernelParamRemovePoints(h) == 0
    L.Nfft4GPNFFTKernelParamFree(h)


# ------------------------------------------------------------------------------------------------ 1-D windows: table-free path
@pytest.mark.parametrize("n,d,kernel,force", [(40000, 5, "gaussian", False), (50000, 3, "matern12", False), (3000, 4, "gaussian", True),
                                              (700, 2, "matern12", True), (8192 * 3 + 1, 2, "gaussian", True)])
def test_poly1d_path_matches_oracle(gpu, n, d, kernel, force):
    """BASELINE config C4's kernel class -- an additive kernel of 1-D windows -- on the polynomial path (nfft1d_poly.cu: moments per
    cell in the spread, one degree-14 Horner evaluation per point in the gather, points sorted by cell inside blocks of the original
    order): matvec and gradient matvec against the oracle for all (alpha, beta) branches, and against the tap-table kernels."""
    import os
    X = uniform_cloud(n, d, seed=50 + d)
    win = np.arange(d, dtype=np.int32)
    f, l, mu = 1.2, 0.4 * n ** (1.0 / d), 0.03
    if force:
        os.environ["NFFT4GP_B200_POLY1D"] = "1"
    try:
        A = gpu.AdditiveNFFTKernel(X, win, d, 1, kernel).build(f, l, mu)
    finally:
        os.environ.pop("NFFT4GP_B200_POLY1D", None)
    assert gpu.lib().Nfft4GPB200DebugSpread(A.h, 0, 1.0, gpu.capi.dptr(np.zeros(n)), gpu.capi.dptr(np.zeros(64))) == -1   # no tap tables on this path
    B = O.OracleAdditiveNFFT(X, win, d, 1, kernel).build(f, l, mu)
    x = rand_vec(n, 5)
    y0, g0 = rand_vec(n, 6), rand_vec(3 * n, 7)
    for al, be in ((1.0, 0.0), (-1.0, 1.0), (0.7, 0.3)):
        assert rel_l2(A.matvec(x, al, be, y0), B.matvec(x, al, be, y0)) <= TOL, (al, be)
        assert rel_l2(A.grad_matvec(x, al, be, g0), B.grad_matvec(x, al, be, g0)) <= TOL, (al, be)
    # new hyper-parameters on the same points
    A.build(0.8, 1.9 * l, 0.2)
    B.build(0.8, 1.9 * l, 0.2)
    assert rel_l2(A.matvec(x), B.matvec(x)) <= TOL
    os.environ["NFFT4GP_B200_POLY1D"] = "0"
    try:
        T = gpu.AdditiveNFFTKernel(X, win, d, 1, kernel).build(0.8, 1.9 * l, 0.2)
    finally:
        os.environ.pop("NFFT4GP_B200_POLY1D", None)
    assert rel_l2(A.matvec(x), T.matvec(x)) <= 1e-12
    assert T.debug_spread(0, x).shape == (64,)      # the table path keeps its stage taps
    for h in (A, B, T):
        h.free()


def test_poly1d_upper_end_point_and_direct_kernels(gpu):
    """(1) A point whose scaled coordinate is exactly +1/4 (cell key 48: outside the 32 home cells whose moments the staged spread keeps in
    registers, nfft1d_poly.cu) in every window and in the last, odd-sized block.  (2) The direct-load kernels -- the fallback of the
    staged ones for an unaligned vector, and the gradient gather -- in a fresh process (the switch is read once per process)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    n, d = 3 * 1024 + 1, 3
    rng = np.random.default_rng(77)
    a = rng.random((d, (n - 1) // 2)) * 0.49
    X = np.zeros((d, n))
    X[:, 0:n - 1:2], X[:, 1:n - 1:2] = a, -a            # mean exactly 0 (the last point is 0)
    X[:, 0], X[:, 1] = 0.5, -0.5                        # radius 1/2 -> scale 1/2 -> scaled coordinates exactly +-1/4
    X[:, n - 3], X[:, n - 2] = 0.5, -0.5                # ... also next to the last block
    win = np.arange(d, dtype=np.int32)
    os.environ["NFFT4GP_B200_POLY1D"] = "1"
    try:
        A = gpu.AdditiveNFFTKernel(X, win, d, 1, "gaussian").build(1.1, 0.3, 0.02)
    finally:
        os.environ.pop("NFFT4GP_B200_POLY1D", None)
    B = O.OracleAdditiveNFFT(X, win, d, 1, "gaussian").build(1.1, 0.3, 0.02)
    x = rand_vec(n, 8)
    assert rel_l2(A.matvec(x), B.matvec(x)) <= TOL
    assert rel_l2(A.grad_matvec(x), B.grad_matvec(x)) <= TOL
    A.free()
    B.free()
    code = (
        "import numpy as np, sys\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as gpu\n"
        "from oracle import oracle_api as O\n"
        "assert gpu.lib().Nfft4GPB200Init(0) == 0\n"
        "rng = np.random.default_rng(5); n, d = 40000, 4\n"
        "X = rng.random((d, n)) * 3.0; win = np.arange(d, dtype=np.int32); x = rng.standard_normal(n)\n"
        "A = gpu.AdditiveNFFTKernel(X, win, d, 1, 'matern12').build(1.3, 1.1, 0.05)\n"
        "B = O.OracleAdditiveNFFT(X, win, d, 1, 'matern12').build(1.3, 1.1, 0.05)\n"
        "e = np.linalg.norm(A.matvec(x) - B.matvec(x)) / np.linalg.norm(B.matvec(x))\n"
        "print('REL', e); assert e <= 1e-10\n"
    ) % (root, os.path.join(root, "tests"))
    env = dict(os.environ, NFFT4GP_B200_POLY1D_ASYNC="0")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "REL" in r.stdout, r.stdout + r.stderr


# ------------------------------------------------------------------------------------------------ 2-D windows: table-free path
@pytest.mark.parametrize("n,d,kernel,force,cloud", [(70000, 4, "gaussian", False, "uniform"), (3000, 2, "matern12", True, "uniform"),
                                                    (100000, 6, "matern12", False, "gaussian"), (257, 4, "gaussian", True, "uniform")])
def test_poly2d_path_matches_oracle(gpu, n, d, kernel, force, cloud):
    """Additive kernels of 2-D windows on the table-free path (nfft2d_poly.cu: bivariate degree-14 polynomial per grid cell, moments and
    evaluation as Vandermonde GEMMs on the fp64 tensor cores; 28 B per point instead of 172 B of taps): matvec and gradient matvec
    against the oracle for all (alpha, beta) branches, and against the tap-table kernels on the same data."""
    import os
    X = (uniform_cloud if cloud == "uniform" else gaussian_cloud)(n, d, seed=60 + d)
    win = np.arange(d, dtype=np.int32)
    f, l, mu = 1.3, 0.5 * n ** (1.0 / d) if cloud == "uniform" else 0.8, 0.02
    if force:
        os.environ["NFFT4GP_B200_POLY2D"] = "1"
    try:
        A = gpu.AdditiveNFFTKernel(X, win, d // 2, 2, kernel).build(f, l, mu)
    finally:
        os.environ.pop("NFFT4GP_B200_POLY2D", None)
    B = O.OracleAdditiveNFFT(X, win, d // 2, 2, kernel).build(f, l, mu)
    x = rand_vec(n, 5)
    y0, g0 = rand_vec(n, 6), rand_vec(3 * n, 7)
    for al, be in ((1.0, 0.0), (-1.0, 1.0), (0.7, 0.3)):
        assert rel_l2(A.matvec(x, al, be, y0), B.matvec(x, al, be, y0)) <= TOL, (al, be)
        assert rel_l2(A.grad_matvec(x, al, be, g0), B.grad_matvec(x, al, be, g0)) <= TOL, (al, be)
    A.build(0.8, 1.7 * l, 0.2)                       # new hyper-parameters on the same points
    B.build(0.8, 1.7 * l, 0.2)
    assert rel_l2(A.matvec(x), B.matvec(x)) <= TOL
    os.environ["NFFT4GP_B200_POLY2D"] = "0"
    try:
        T = gpu.AdditiveNFFTKernel(X, win, d // 2, 2, kernel).build(0.8, 1.7 * l, 0.2)
    finally:
        os.environ.pop("NFFT4GP_B200_POLY2D", None)
    assert rel_l2(A.matvec(x), T.matvec(x)) <= 1e-12
    for h in (A, B, T):
        h.free()
