"""Dense Cholesky preconditioner (SRC/preconds/chol.c) and the gradient blocks of the dense kernels against the reference's
compiled code (tests/golden/chol_gauss300.npz, scripts/make_golden.py chol_case)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "chol_gauss300.npz")


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b))


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def kernel(gold):
    import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as pkg
    k = pkg.DensePlainKernel(gold["X"], "gaussian").set_hyper(float(gold["f"]), float(gold["l"]), float(gold["mu"]))
    yield k
    k.free()


def test_dense_kernel_gradient_blocks(gold, kernel):
    K, dK = kernel.block(gold["rows"], gold["cols"], grad=True)
    assert rel(K, gold["cross_K"]) < 1e-13
    assert rel(dK, gold["cross_dK"]) < 1e-12
    assert np.all(dK[2] == 0.0)
    K, dK = kernel.block(gold["rows"], None, grad=True)
    assert rel(K, gold["sym_K"]) < 1e-13
    assert rel(dK, gold["sym_dK"]) < 1e-12
    assert np.all(np.triu(dK[1], 1) == 0.0)  # lower triangle only, as the reference


def test_additive_dense_gradient_is_mean_of_windows(gold):
    import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as pkg
    X = gold["X"]
    f, l, mu = float(gold["f"]), float(gold["l"]), float(gold["mu"])
    win = np.arange(9, dtype=np.int32)
    A = pkg.DenseAdditiveKernel(X, win, 3, 3, "matern12").set_hyper(f, l, mu)
    rows, cols = gold["rows"], gold["cols"]
    K, dK = A.block(rows, cols, grad=True)
    acc = np.zeros_like(K)
    dacc = np.zeros_like(K)
    for w in range(3):
        D = X[3 * w:3 * w + 3][:, rows].T[:, None, :] - X[3 * w:3 * w + 3][:, cols].T[None, :, :]
        r = np.sqrt((D * D).sum(-1))
        acc += f * f * np.exp(-r / l) / 3
        dacc += f * f * r / (l * l) * np.exp(-r / l) / 3
    assert rel(K, acc) < 1e-13 and rel(dK[1], dacc) < 1e-13 and rel(dK[0], 2 / f * acc) < 1e-13
    A.free()


@pytest.mark.parametrize("stable", [0, 1])
def test_chol_solve_logdet_trace_dvp(gold, kernel, stable):
    import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as pkg
    ch = pkg.CholPrecond(stable).setup(kernel, require_grad=1)
    x = ch.solve(gold["rhs"])
    # cond(K) ~ 1e4 at mu = 0.05: the solve is reproduced to well below the 1e-10 bar
    assert rel(x, gold[f"solve{stable}"]) < 1e-11
    assert abs(ch.logdet() - float(gold[f"logdet{stable}"])) < 1e-9 * abs(float(gold[f"logdet{stable}"]))
    assert rel(ch.trace(), gold[f"trace{stable}"]) < 1e-10
    y = ch.dvp(gold["rhs"])
    assert rel(y.ravel(), gold[f"dvp{stable}"]) < 1e-10
    ym = ch.dvp(gold["rhs"], mask=[0, 1, 0])
    assert np.all(ym[0] == 0) and np.all(ym[2] == 0) and rel(ym[1], y[1]) < 1e-15
    ch.free()


def test_chol_without_gradient_refuses_trace(gold, kernel, capfd):
    import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as pkg
    ch = pkg.CholPrecond(0).setup(kernel, require_grad=0)
    with pytest.raises(RuntimeError):
        ch.trace()
    assert rel(ch.solve(gold["rhs"]), gold["solve0"]) < 1e-11
    ch.free()


def test_pcg_with_exact_dense_factorisation(gold, kernel):
    """PCG on the NFFT matvec of the SAME kernel, preconditioned by its exact dense factorisation, stays device-resident
    (b200_lookup_dev_solve knows the Cholesky solve) and converges in a handful of iterations (as many as the NFFT approximation error of the
    operator leaves: unpreconditioned CG needs > 100 here)."""
    import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as pkg
    X = gold["X"][:3]
    f, l, mu = float(gold["f"]), float(gold["l"]), float(gold["mu"])
    n = X.shape[1]
    win = np.arange(3, dtype=np.int32)
    A = pkg.AdditiveNFFTKernel(X, win, 1, 3, "gaussian").build(f, l, mu)
    D = pkg.DenseAdditiveKernel(X, win, 1, 3, "gaussian").set_hyper(f, l, mu)
    ch = pkg.CholPrecond(1).setup(D)
    x, relres, hist, its = pkg.solver_pcg(A.h, n, A.matvec_fn(), ch.h, ch.solve_fn(), np.zeros(n), gold["rhs"], 50, 0, 1e-10)
    assert 0 < its <= 15
    _, _, _, its0 = pkg.solver_pcg(A.h, n, A.matvec_fn(), None, None, np.zeros(n), gold["rhs"], 300, 0, 1e-10)
    assert its0 == 0 or its0 > 3 * its
    assert rel(A.matvec(x), gold["rhs"]) < 1e-9
    ch.free(), D.free(), A.free()
