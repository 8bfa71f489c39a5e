"""CPU: the NumPy restatement of the preconditioner setups (oracle/np_setup.py) against the golden vectors produced by the
reference's own compiled code (scripts/make_golden.py), and the host-side dense helpers of the product against NumPy."""
import ctypes as C
import os

import numpy as np

from oracle import np_setup as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def test_dense_kernels_match_reference_matrices():
    z = _load("solver_dense300.npz")
    wins = S.window_dims(z["windows"], 3, 3)
    K = S.dense_matrix(z["X"], float(z["f"]), float(z["l"]), float(z["mu"]), "gaussian", wins)
    assert np.abs(K - z["K"]).max() <= 1e-13 * np.abs(z["K"]).max()
    z = _load("fsai_gauss300.npz")
    K = S.dense_matrix(z["X"], float(z["f"]), float(z["l"]), float(z["mu"]), "gaussian", None)
    assert np.abs(K - z["K"]).max() <= 1e-13 * np.abs(z["K"]).max()


def test_fsai_setup_matches_reference():
    z = _load("fsai_gauss300.npz")
    n = z["X"].shape[1]
    lfil = int(z["fsai_lfil"])
    pattern = S.knn_pattern(z["X"], lfil)
    K = z["K"]
    vals = S.fsai_rows(lambda J: K[np.ix_(J, J)], pattern)
    G = S.rows_to_dense(n, pattern, vals)
    Gref = S.csr_to_dense(n, z["fsai_ia"], z["fsai_ja"], z["fsai_a"])
    assert ((G != 0) == (Gref != 0)).all(), "KNN pattern differs from the reference"
    assert np.abs(G - Gref).max() <= 1e-10 * np.abs(Gref).max()
    zf = G.T @ (G @ z["rhs"])
    assert np.linalg.norm(zf - z["fsai_z"]) <= 1e-10 * np.linalg.norm(z["fsai_z"])


def test_nystrom_setup_matches_reference():
    z = _load("nys_gauss300.npz")
    U, s, eta = S.nys_setup(z["X"], z["perm"], int(z["k"]), float(z["f"]), float(z["l"]), float(z["mu"]))
    assert abs(eta - float(z["eta"])) <= 1e-15
    # singular values: the reference stores s = 1 / (sigma^2 + eta); the small ones sit at the rounding level of U'U
    big = z["s"] < 0.5 / eta
    assert np.abs(s[big] - z["s"][big]).max() <= 1e-6 * np.abs(z["s"][big]).max()
    x = S.nys_apply(z["perm"], U, s, eta, z["rhs"])
    assert np.linalg.norm(x - z["z"]) <= 1e-6 * np.linalg.norm(z["z"])


def test_rank_estimate_runs_and_fps_is_a_permutation():
    rng = np.random.default_rng(5)
    X = rng.random((3, 400)) * 400 ** (1 / 3)
    p = S.fps(X, 400)
    assert sorted(p.tolist()) == list(range(400))
    k_lo = S.rank_estimate(X, 1.0, 50.0, 0.01, 100, m=200, numests=2)    # very smooth kernel: low rank
    k_hi = S.rank_estimate(X, 1.0, 0.05, 0.01, 100, m=200, numests=2)    # nearly diagonal kernel: full rank
    assert k_lo < 100 < k_hi


def test_host_dense_helpers():
    import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as pkg
    L = pkg.lib()
    rng = np.random.default_rng(0)
    for n in (1, 2, 7, 64, 200):
        A = rng.standard_normal((n, n))
        A = A @ A.T + 0.1 * np.eye(n)
        w0 = np.linalg.eigvalsh(A)
        for vec in (1, 0):
            Af = np.asfortranarray(A.copy())
            w = np.zeros(n)
            assert L.Nfft4GPB200HostSyev(n, pkg.capi.dptr(Af), pkg.capi.dptr(w), vec) == 0
            assert np.abs(w - w0).max() <= 1e-12 * np.abs(w0).max()
            if vec:
                assert np.abs(Af @ np.diag(w) @ Af.T - A).max() <= 1e-11 * np.abs(A).max()
                assert np.abs(Af.T @ Af - np.eye(n)).max() <= 1e-12
        Af = np.asfortranarray(A.copy())
        assert L.Nfft4GPB200HostPotrf(n, pkg.capi.dptr(Af)) == 0
        assert np.abs(np.tril(Af) - np.linalg.cholesky(A)).max() <= 1e-11 * np.abs(A).max()
    bad = np.asfortranarray(np.array([[1.0, 2.0], [2.0, 1.0]]))
    assert L.Nfft4GPB200HostPotrf(2, pkg.capi.dptr(bad)) == 2


def test_numpy_fgmres_matches_reference_golden():
    """oracle/np_solvers.fgmres against the reference's compiled Nfft4GPSolverFgmres (restarts, Nystrom, iteration cap)."""
    from oracle import np_solvers as NS
    g = _load("fgmres_dense300.npz")
    z = _load("nys_gauss300.npz")
    K, rhs, perm, U, s, eta = z["K"], z["rhs"], z["perm"], z["U"], z["s"], float(z["eta"])
    mv = lambda v: K @ v
    prec = lambda v: S.nys_apply(perm, U, s, eta, v)
    for tag, pf in (("plain", None), ("nys", prec), ("cap", None)):
        x, rel, hist, it = NS.fgmres(mv, pf, np.zeros(K.shape[0]), rhs, int(g[f"{tag}_kdim"]), int(g[f"{tag}_maxits"]), 0, float(g[f"{tag}_tol"]))
        # The residual ESTIMATES of a GMRES cycle amplify rounding differences (dot-product summation order) by many orders
        # of magnitude while the Krylov basis loses orthogonality and recover afterwards (4e-13 -> 2.5e-3 -> 1e-13 inside one
        # cycle of this problem), so: tight on the first steps, a band on the whole history, iteration count +-2.
        assert abs(it - int(g[f"{tag}_it"])) <= 2, (tag, it, int(g[f"{tag}_it"]))
        m = min(it, int(g[f"{tag}_it"]))
        assert np.abs(hist[:12] - g[f"{tag}_hist"][:12]).max() <= 1e-10
        assert np.abs(hist[:m] / g[f"{tag}_hist"][:m] - 1.0).max() <= 0.5
        if tag == "cap":   # stopped at the iteration limit: x is as sensitive as the estimates; compare the TRUE residuals
            r_here = np.linalg.norm(rhs - K @ x)
            r_gold = np.linalg.norm(rhs - K @ g[f"{tag}_x"])
            assert abs(r_here / r_gold - 1.0) <= 0.3
        else:
            assert np.linalg.norm(x - g[f"{tag}_x"]) <= 1e-5 * np.linalg.norm(g[f"{tag}_x"])
