"""Where oracle/_ref exists (this container: the reference's own C sources compiled unmodified on the NFFT3 restatement), the
self-contained C restatement oracle_path.c is compared with it DIRECTLY on fresh random inputs -- not only through the stored
fixtures: additive matvec and gradient matvec for every window shape, both kernels, and the PCG of the reference."""
import numpy as np
import pytest

from common import rand_vec, rel_l2, uniform_cloud, gaussian_cloud

CASES = [
    ("3x3d_gauss", 2500, 9, list(range(9)), 3, 3, "gaussian", 1.1, 1.3, 0.02),
    ("permuted_3d_matern", 1800, 13, [1, 8, 6, 9, 4, 2, 5, 3, 11], 3, 3, "matern12", 0.8, 0.9, 0.05),
    ("padded_last", 1500, 5, [0, 1, 3, 2, 4, -1], 2, 3, "gaussian", 1.2, 0.7, 0.01),
    ("2d_matern", 2000, 4, [0, 1, 2, 3], 2, 2, "matern12", 1.0, 0.6, 0.03),
    ("1d_gauss", 1200, 3, [0, 1, 2], 3, 1, "gaussian", 0.9, 0.4, 0.02),
]


def _ref():
    from oracle import refapi as R
    try:
        R.lib()
    except Exception:
        pytest.skip("oracle/_ref is not built here (it needs /root/reference)")
    return R


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("cloud", [uniform_cloud, gaussian_cloud], ids=["uniform", "gaussian"])
def test_restatement_equals_reference_wrapper(case, cloud):
    R = _ref()
    from oracle import oracle_api as O
    name, n, d, win, nw, dw, kern, f, l, mu = case
    X = cloud(n, d, seed=len(name))
    win = np.array(win, dtype=np.int32)
    A = R.RefAdditiveNFFT(X, win, nw, dw, kern).build(f, l, mu)
    B = O.OracleAdditiveNFFT(X, win, nw, dw, kern).build(f, l, mu)
    x, y0 = rand_vec(n, 3), rand_vec(n, 4)
    assert rel_l2(B.matvec(x), A.matvec(x)) <= 1e-13
    assert rel_l2(B.matvec(x, alpha=0.3, beta=-1.5, y=y0.copy()), A.matvec(x, alpha=0.3, beta=-1.5, y=y0.copy())) <= 1e-13
    if hasattr(A, "grad_matvec") and hasattr(B, "grad_matvec"):
        assert rel_l2(B.grad_matvec(x), A.grad_matvec(x)) <= 1e-12
    B.free()
