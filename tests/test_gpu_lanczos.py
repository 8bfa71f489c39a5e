"""GPU parity of the preconditioned Lanczos solver and the stochastic Lanczos quadrature of the log-determinant
(SRC/solvers/lanczos.c, the routines behind the log det term of the GP loss) against the reference's own compiled code
(tests/golden/lanczos_dense300.npz, scripts/make_golden.py)."""
import ctypes as C
import os

import numpy as np
import pytest

from common import rel_l2

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_lanczos_solver_matches_reference_golden(gpu):
    g = np.load(os.path.join(GOLD, "lanczos_dense300.npz"))
    z = np.load(os.path.join(GOLD, "nys_gauss300.npz"))
    K, rhs = np.ascontiguousarray(z["K"]), z["rhs"]
    n = K.shape[0]

    def dense_mv(_h, nn, alpha, xp, beta, yp):
        x = np.ctypeslib.as_array(xp, shape=(nn,))
        y = np.ctypeslib.as_array(yp, shape=(nn,))
        y[:] = alpha * (K @ x) + (beta * y if beta != 0.0 else 0.0)
        return 0

    cb = gpu.capi.FUNC_SYMMATVEC(dense_mv)
    mv = C.cast(cb, C.c_void_p)
    dummy = C.c_int(0)
    N = gpu.NysPrecond()
    N.set_factors(z["perm"], z["U"], z["s"], float(z["eta"]))
    for tag, (ph, pf) in {"plain": (None, None), "nys": (N.h, N.solve_fn())}.items():
        maxits = int(g[f"{tag}_maxits"])
        x, rel, hist, it, TD, TE = gpu.solver_lanczos(C.addressof(dummy), n, mv, ph, pf, np.zeros(n), rhs, maxits, maxits, 0, float(g[f"{tag}_tol"]))
        assert abs(it - int(g[f"{tag}_it"])) <= 1, (tag, it, int(g[f"{tag}_it"]))
        assert TD.size == int(g[f"{tag}_tsize"])
        assert rel_l2(x, g[f"{tag}_x"]) <= 1e-6, tag
        m = min(it, int(g[f"{tag}_it"]))
        assert np.abs(hist[:m] / g[f"{tag}_hist"][:m] - 1.0).max() <= 1e-3, tag
        # the tridiagonal matrix: with full re-orthogonalisation its leading part is reproducible to rounding amplified by the
        # (moderate) loss of orthogonality; the spectrum -- what the quadrature uses -- must agree
        assert np.abs(TD[:20] - g[f"{tag}_TD"][:20]).max() <= 1e-8 * np.abs(g[f"{tag}_TD"][:20]).max(), tag
        assert np.abs(np.abs(TE[:20]) - np.abs(g[f"{tag}_TE"][:20])).max() <= 1e-8 * np.abs(g[f"{tag}_TE"][:20]).max(), tag
        T = np.diag(TD) + np.diag(TE, 1) + np.diag(TE, -1)
        Tg = np.diag(g[f"{tag}_TD"]) + np.diag(g[f"{tag}_TE"], 1) + np.diag(g[f"{tag}_TE"], -1)
        ev, evg = np.linalg.eigvalsh(T), np.linalg.eigvalsh(Tg)
        assert abs(ev[-1] / evg[-1] - 1.0) <= 1e-8 and abs(ev[0] / evg[0] - 1.0) <= 1e-4, tag
    N.free()


def test_lanczos_quadrature_logdet_matches_reference_golden(gpu):
    g = np.load(os.path.join(GOLD, "lanczos_dense300.npz"))
    X, win = g["slq_X"], g["slq_windows"]
    n = X.shape[1]
    L = gpu.lib()
    A = gpu.AdditiveNFFTKernel(X, win, 2, 3, "gaussian").build(float(g["slq_f"]), float(g["slq_l"]), float(g["slq_mu"]))
    val, dval = gpu.lanczos_quadrature_logdet(A.K, A.dK, n, A.matvec_fn(), C.cast(L.Nfft4GPAdditiveNFFTGradMatSymv, C.c_void_p),
                                              int(g["slq_maxits"]), g["slq_probes"])
    assert abs(val - float(g["slq_logdet"])) <= 1e-6 * abs(float(g["slq_logdet"]))
    assert np.abs(dval - g["slq_dlogdet"]).max() <= 1e-6 * np.abs(g["slq_dlogdet"]).max()
    # sanity against the exact log-determinant of the dense kernel (4 probes: a few percent)
    from oracle import np_setup as S
    Kd = S.dense_matrix(X, float(g["slq_f"]), float(g["slq_l"]), float(g["slq_mu"]), "gaussian", S.window_dims(win, 2, 3))
    exact = np.linalg.slogdet(Kd)[1] / n
    assert abs(val - exact) <= 0.15 * abs(exact)
    A.free()
