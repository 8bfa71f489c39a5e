"""Gradient side of the Nystrom preconditioner (nys.c:175-516: log det, traces, derivative-vector products) against the
reference's compiled code (tests/golden/nys_grad_gauss300.npz, scripts/make_golden.py nys_grad_case) and against dense
NumPy algebra on the same factors."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b))


@pytest.fixture(scope="module")
def setup():
    import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as pkg
    z = np.load(os.path.join(GOLD, "nys_gauss300.npz"))
    g = np.load(os.path.join(GOLD, "nys_grad_gauss300.npz"))
    K = pkg.DensePlainKernel(z["X"], "gaussian").set_hyper(float(z["f"]), float(z["l"]), float(z["mu"]))
    N = pkg.NysPrecond().setup(K, z["perm"], int(z["k"]), require_grad=1)
    yield pkg, z, g, K, N
    N.free()
    K.free()


def test_nys_logdet_trace_dvp_match_reference(setup):
    pkg, z, g, K, N = setup
    assert rel(N.solve(z["rhs"]), g["solve"]) < 1e-6          # conditioning of eig(U'U), as in the setup test
    assert abs(N.logdet() - float(g["logdet"])) < 1e-8 * abs(float(g["logdet"]))
    # the traces are differences of O(n f^2 / eta) terms: relative to that scale
    scale = z["X"].shape[1] * float(z["f"]) ** 2 / (float(z["mu"]) * float(z["f"]) ** 2)
    assert np.abs(N.trace() - g["trace"]).max() < 1e-6 * scale
    assert rel(N.trace(), g["trace"]) < 1e-5
    y = N.dvp(z["rhs"])
    assert rel(y.ravel(), g["dvp"]) < 1e-5
    ym = N.dvp(z["rhs"], mask=[0, 0, 1])
    assert np.all(ym[:2] == 0) and rel(ym[2], y[2]) < 1e-15


def test_nys_gradient_against_dense_algebra(setup):
    """M = K1 K11^-1 K1' + eta I in the permuted order; dM by the formulas of nys.c:246-276 in NumPy (dense inverses)."""
    pkg, z, g, K, N = setup
    X, f, l, mu, perm, k = z["X"], float(z["f"]), float(z["l"]), float(z["mu"]), z["perm"], int(z["k"])
    n = X.shape[1]
    Xp = X[:, perm]
    D2 = ((Xp[:, :, None] - Xp[:, None, :k]) ** 2).sum(0)
    K1 = f * f * np.exp(-D2 / (2 * l * l))
    dK1 = [2 / f * K1, f * f * D2 / l ** 3 * np.exp(-D2 / (2 * l * l))]
    K11 = K1[:k].copy()
    fro = np.linalg.norm(K11)
    K11s = K11 + np.sqrt(k) * np.spacing(fro) * np.eye(k)
    Ki = np.linalg.inv(K11s)
    eta = mu * f * f
    M = K1 @ Ki @ K1.T + eta * np.eye(n)
    Mi = np.linalg.inv(M)
    dM = [dK1[i] @ Ki @ K1.T + K1 @ Ki @ dK1[i].T - K1 @ Ki @ dK1[i][:k] @ Ki @ K1.T for i in range(2)] + [f * f * np.eye(n)]
    tr = np.array([np.trace(Mi @ dM[i]) for i in range(3)])
    assert rel(N.trace(), tr) < 1e-5
    x = z["rhs"]
    y = N.dvp(x)
    for i in range(3):
        ref = np.zeros(n)
        ref[perm] = Mi @ (dM[i] @ x[perm])
        assert rel(y[i], ref) < 1e-5, i
    sign, ld = np.linalg.slogdet(M)
    assert abs(N.logdet() - ld) < 1e-6 * abs(ld)


def test_nys_without_gradient_refuses(setup):
    pkg, z, g, K, N = setup
    N0 = pkg.NysPrecond().setup(K, z["perm"], int(z["k"]))
    with pytest.raises(RuntimeError):
        N0.trace()
    with pytest.raises(RuntimeError):
        N0.dvp(z["rhs"])
    assert abs(N0.logdet() - float(g["logdet"])) < 1e-8 * abs(float(g["logdet"]))
    N0.free()
