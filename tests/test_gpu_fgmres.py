"""GPU parity of the flexible GMRES (SRC/solvers/fgmres.c, the solver the reference's optimizer and drivers call) through
the C ABI: against the reference's own compiled solver on its golden problem (foreign CPU matvec through the function-pointer
plug point, Nystrom factors uploaded) and against the NumPy restatement driven by the CPU oracle's NFFT matvec."""
import ctypes as C
import os

import numpy as np
import pytest

from common import rand_vec, rel_l2, uniform_cloud
from oracle import np_setup as S
from oracle import np_solvers as NS
from oracle import oracle_api as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_fgmres_matches_reference_golden(gpu):
    g = np.load(os.path.join(GOLD, "fgmres_dense300.npz"))
    z = np.load(os.path.join(GOLD, "nys_gauss300.npz"))
    K, rhs = np.ascontiguousarray(z["K"]), z["rhs"]
    n = K.shape[0]

    def dense_mv(_h, nn, alpha, xp, beta, yp):          # func_symmatvec of a caller's own CPU matrix
        x = np.ctypeslib.as_array(xp, shape=(nn,))
        y = np.ctypeslib.as_array(yp, shape=(nn,))
        y[:] = alpha * (K @ x) + (beta * y if beta != 0.0 else 0.0)
        return 0

    cb = gpu.capi.FUNC_SYMMATVEC(dense_mv)
    mv = C.cast(cb, C.c_void_p)
    dummy = C.c_int(0)
    N = gpu.NysPrecond()
    N.set_factors(z["perm"], z["U"], z["s"], float(z["eta"]))
    for tag, (ph, pf) in {"plain": (None, None), "nys": (N.h, N.solve_fn()), "cap": (None, None)}.items():
        x, rel, hist, it = gpu.solver_fgmres(C.addressof(dummy), n, mv, ph, pf, np.zeros(n), rhs, int(g[f"{tag}_kdim"]),
                                             int(g[f"{tag}_maxits"]), 0, float(g[f"{tag}_tol"]))
        # see tests/test_oracle_setup.py::test_numpy_fgmres_matches_reference_golden for the tolerances
        assert abs(it - int(g[f"{tag}_it"])) <= 2, (tag, it, int(g[f"{tag}_it"]))
        m = min(it, int(g[f"{tag}_it"]))
        assert np.abs(hist[:12] - g[f"{tag}_hist"][:12]).max() <= 1e-10, tag
        assert np.abs(hist[:m] / g[f"{tag}_hist"][:m] - 1.0).max() <= 0.5, tag
        if tag == "cap":   # stopped at the iteration limit: compare the TRUE residuals
            assert abs(np.linalg.norm(rhs - K @ x) / np.linalg.norm(rhs - K @ g[f"{tag}_x"]) - 1.0) <= 0.3
        else:
            assert rel_l2(x, g[f"{tag}_x"]) <= 1e-5, tag
        assert abs(rel / float(g[f"{tag}_rel"]) - 1.0) <= 0.5, tag
    N.free()


def test_fgmres_on_nfft_matvec_with_afn(gpu):
    """The reference's solve as gp_loss.c:194-213 runs it (FGMRES, kdim = maxits) on the NFFT additive matvec with the AFN
    preconditioner; iteration count +-1 and solution against the NumPy FGMRES on the CPU oracle's matvec."""
    n, d, k, lfil = 3000, 6, 300, 100
    X = uniform_cloud(n, d)
    win = np.arange(6, dtype=np.int32)
    wins = S.window_dims(win, 2, 3)
    f, l, mu = 1.0, 1.0, 0.01
    D = gpu.DenseAdditiveKernel(X, win, 2, 3, "gaussian").set_hyper(f, l, mu)
    Af = gpu.AfnPrecond().setup(D, k, lfil, numests=0, noran=1)
    A = gpu.AdditiveNFFTKernel(X, win, 2, 3, "gaussian").build(f, l, mu)
    rhs = rand_vec(n, 4)
    x1, rel1, h1, it1 = gpu.solver_fgmres(A.K, n, A.matvec_fn(), Af.h, Af.solve_fn(), np.zeros(n), rhs, 100, 100, 0, 1e-6)
    assert 0 < it1 < 100 and rel1 <= 1e-6
    assert rel_l2(A.matvec(x1), rhs) <= 2e-6
    Fo = S.afn_setup(X, k, lfil, f, l, mu, "gaussian", wins)
    B = O.OracleAdditiveNFFT(X, win, 2, 3, "gaussian").build(f, l, mu)
    xo, relo, ho, ito = NS.fgmres(lambda v: B.matvec(v), lambda v: S.afn_apply(Fo, v), np.zeros(n), rhs, 100, 100, 0, 1e-6)
    assert abs(it1 - ito) <= 1, (it1, ito)
    assert rel_l2(x1, xo) <= 1e-5
    assert np.abs(h1[:10] - ho[:10]).max() <= 1e-8 * h1[0]
    # restarts: kdim 10
    x2, rel2, h2, it2 = gpu.solver_fgmres(A.K, n, A.matvec_fn(), Af.h, Af.solve_fn(), np.zeros(n), rhs, 10, 400, 0, 1e-6)
    xo2, relo2, ho2, ito2 = NS.fgmres(lambda v: B.matvec(v), lambda v: S.afn_apply(Fo, v), np.zeros(n), rhs, 10, 400, 0, 1e-6)
    assert abs(it2 - ito2) <= max(1, ito2 // 50), (it2, ito2)
    assert rel_l2(A.matvec(x2), rhs) <= 2e-6
    A.free()
    B.free()
    Af.free()
    D.free()


def test_fgmres_early_exits(gpu):
    """fgmres.c:43-54 (zero right-hand side: x = 0, one history entry) and :75-92 (the initial guess already solves it)."""
    n, d = 1200, 3
    X = uniform_cloud(n, d)
    win = np.arange(3, dtype=np.int32)
    A = gpu.AdditiveNFFTKernel(X, win, 1, 3, "gaussian").build(1.0, 1.0, 0.5)
    x, rel, hist, it = gpu.solver_fgmres(A.K, n, A.matvec_fn(), None, None, np.ones(n), np.zeros(n), 10, 50, 0, 1e-8)
    assert it == 0 and rel == 0.0 and hist.size == 1 and np.all(x == 0.0)
    rhs = rand_vec(n, 3)
    xs, rels, hs, its = gpu.solver_fgmres(A.K, n, A.matvec_fn(), None, None, np.zeros(n), rhs, 30, 200, 0, 1e-10)
    assert its > 0 and rels <= 1e-10
    x2, rel2, h2, it2 = gpu.solver_fgmres(A.K, n, A.matvec_fn(), None, None, xs, rhs, 30, 200, 0, 1e-6)
    assert it2 <= 1 and rel2 <= 1e-6
    A.free()
