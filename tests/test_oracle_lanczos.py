"""CPU: the NumPy restatement of the reference's Lanczos solver (oracle/np_solvers.py:lanczos, SRC/solvers/lanczos.c:3-300)
against the run of the reference's own compiled code stored in tests/golden/lanczos_dense300.npz (scripts/make_golden.py),
without and with the reference's Nystrom preconditioner."""
import os

import numpy as np

from common import rel_l2
from oracle import np_solvers as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_numpy_lanczos_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "lanczos_dense300.npz"))
    z = np.load(os.path.join(GOLD, "nys_gauss300.npz"))
    K, rhs, perm, U, s, eta = z["K"], z["rhs"], z["perm"], z["U"], z["s"], float(z["eta"])
    n = K.shape[0]

    def nys(r):                                   # nys.c:115-173 with the reference's factors
        y = r[perm]
        c = U.T @ y
        out = np.zeros(n)
        out[perm] = (y - U @ c) / eta + U @ (s * c)
        return out

    for tag, prec in (("plain", None), ("nys", nys)):
        maxits = int(g[f"{tag}_maxits"])
        x, rel, hist, it, TD, TE = S.lanczos(lambda v: K @ v, n, np.zeros(n), rhs, maxits, maxits, 0, float(g[f"{tag}_tol"]), prec)
        assert abs(it - int(g[f"{tag}_it"])) <= 1, (tag, it)
        assert rel_l2(x, g[f"{tag}_x"]) <= 1e-6, tag
        m = min(it, int(g[f"{tag}_it"]))
        assert np.abs(hist[:m] / g[f"{tag}_hist"][:m] - 1.0).max() <= 1e-3, tag
        lead = min(m, 20)                       # the leading entries of T before rounding differences accumulate
        assert np.abs(TD[:lead] - g[f"{tag}_TD"][:lead]).max() <= 1e-8 * np.abs(g[f"{tag}_TD"][:lead]).max(), tag
        assert np.abs(TE[:lead - 1] - g[f"{tag}_TE"][:lead - 1]).max() <= 1e-8 * np.abs(g[f"{tag}_TD"][:lead]).max(), tag


def test_numpy_slq_matches_reference_golden():
    """log det / n and its gradient by stochastic Lanczos quadrature on the NFFT additive kernel (CPU oracle matvecs) with the
    probe vectors of the reference run (lanczos.c:421-610)."""
    from oracle import oracle_api as O
    g = np.load(os.path.join(GOLD, "lanczos_dense300.npz"))
    X, win = g["slq_X"], g["slq_windows"]
    n = X.shape[1]
    B = O.OracleAdditiveNFFT(X, win, 2, 3, "gaussian").build(float(g["slq_f"]), float(g["slq_l"]), float(g["slq_mu"]))
    val, dval = S.slq_logdet(B.matvec, B.grad_matvec, n, int(g["slq_maxits"]), g["slq_probes"])
    assert abs(val - float(g["slq_logdet"])) <= 1e-6 * abs(float(g["slq_logdet"]))
    assert np.abs(dval - g["slq_dlogdet"]).max() <= 1e-6 * np.abs(g["slq_dlogdet"]).max()
    B.free()
