"""GPU parity of the preconditioner SETUPS (rows K1, N2, G2, A2 of SURVEY.md section 8(a)) through the C ABI, against the
golden vectors of the reference's compiled code and the NumPy restatement oracle/np_setup.py."""
import ctypes as C
import os

import numpy as np
import pytest

from common import rand_vec, rel_l2, uniform_cloud
from oracle import np_setup as S
from oracle import oracle_api as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


# ------------------------------------------------------------------------------------------------ building blocks
@pytest.mark.parametrize("ta,tb,M,N,K", [("N", "N", 37, 21, 19), ("N", "T", 300, 40, 40), ("T", "N", 40, 40, 5000),
                                         ("T", "N", 130, 70, 20000), ("N", "N", 500, 1000, 500), ("T", "T", 65, 129, 33)])
def test_dmma_gemm(gpu, ta, tb, M, N, K):
    L = gpu.lib()
    rng = np.random.default_rng(M * 7 + N)
    A = rng.standard_normal((M, K) if ta == "N" else (K, M))
    B = rng.standard_normal((K, N) if tb == "N" else (N, K))
    Cm = rng.standard_normal((M, N))
    Af, Bf, Cf = np.asfortranarray(A), np.asfortranarray(B), np.asfortranarray(Cm.copy())
    for alpha, beta in ((1.0, 0.0), (-0.5, 2.0)):
        Cf = np.asfortranarray(Cm.copy())
        rc = L.Nfft4GPB200Dgemm(ta.encode(), tb.encode(), M, N, K, alpha, gpu.capi.dptr(Af), Af.shape[0], gpu.capi.dptr(Bf), Bf.shape[0],
                                beta, gpu.capi.dptr(Cf), M)
        assert rc == 0
        ref = alpha * (A if ta == "N" else A.T) @ (B if tb == "N" else B.T) + beta * Cm
        assert np.abs(Cf - ref).max() <= 1e-12 * np.sqrt(K) * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("n,d,k", [(500, 3, 500), (4000, 9, 64), (3000, 20, 50)])
def test_fps_matches_numpy(gpu, n, d, k):
    X = uniform_cloud(n, d, seed=n)
    out = np.zeros(k, dtype=np.int32)
    assert gpu.lib().Nfft4GPB200Fps(gpu.capi.dptr(X), n, n, d, k, gpu.capi.iptr(out)) == 0
    assert (out == S.fps(X, k)).all()


@pytest.mark.parametrize("n,d,lfil", [(50, 3, 64), (700, 3, 12), (1500, 9, 33), (600, 20, 17), (400, 2, 1), (300, 5, 2)])
def test_knn_pattern_matches_numpy(gpu, n, d, lfil):
    X = uniform_cloud(n, d, seed=3 * n + d)
    pattern = S.knn_pattern(X, lfil)
    nnz = sum(len(r) for r in pattern)
    ia = np.zeros(n + 1, dtype=np.int32)
    ja = np.zeros(nnz, dtype=np.int32)
    assert gpu.lib().Nfft4GPB200KnnPattern(gpu.capi.dptr(X), n, n, d, lfil, gpu.capi.iptr(ia), gpu.capi.iptr(ja)) == 0
    assert ia[-1] == nnz
    for i, r in enumerate(pattern):
        assert (ja[ia[i]:ia[i + 1]] == r).all(), i


def test_dense_kernel_blocks(gpu):
    z = _load("solver_dense300.npz")
    X, win = z["X"], z["windows"]
    f, l, mu = float(z["f"]), float(z["l"]), float(z["mu"])
    n = X.shape[1]
    D = gpu.DenseAdditiveKernel(X, win, 3, 3, "gaussian").set_hyper(f, l, mu)
    K = D.block()
    Kref = z["K"]                                        # the reference's own dense additive matrix
    assert np.abs(np.tril(K) - np.tril(Kref)).max() <= 1e-13 * np.abs(Kref).max()
    assert np.abs(np.triu(K, 1)).max() == 0.0            # only the lower triangle is filled (kernels.c: CALLOC + lower loop)
    rng = np.random.default_rng(1)
    pr = rng.permutation(n)[:57].astype(np.int32)
    pc = rng.permutation(n)[:31].astype(np.int32)
    Kb = D.block(pr, pc)
    assert np.abs(Kb - (Kref - mu * f * f * np.eye(n))[np.ix_(pr, pc)]).max() <= 1e-13   # K(permr, permc): no noise term
    Kl = D.block(pr)
    assert np.abs(Kl - np.tril(Kref[np.ix_(pr, pr)])).max() <= 1e-13
    D.free()
    # Matern-1/2, padded last window, plain kernels
    X2 = uniform_cloud(200, 5, seed=9)
    win2 = np.array([0, 1, 3, 2, 4, -1], dtype=np.int32)
    D2 = gpu.DenseAdditiveKernel(X2, win2, 2, 3, "matern12").set_hyper(0.9, 1.3, 0.02)
    ref2 = S.dense_matrix(X2, 0.9, 1.3, 0.02, "matern12", S.window_dims(win2, 2, 3))
    assert np.abs(D2.block() - np.tril(ref2)).max() <= 1e-13
    D2.free()
    P = gpu.DensePlainKernel(X2, "gaussian").set_hyper(1.2, 0.8, 0.1)
    assert np.abs(P.block() - np.tril(S.dense_matrix(X2, 1.2, 0.8, 0.1, "gaussian"))).max() <= 1e-13
    P.free()


# ------------------------------------------------------------------------------------------------ FSAI (G2)
def test_fsai_setup_matches_reference_golden(gpu):
    z = _load("fsai_gauss300.npz")
    X = z["X"]
    n = X.shape[1]
    P = gpu.DensePlainKernel(X, "gaussian").set_hyper(float(z["f"]), float(z["l"]), float(z["mu"]))
    F = gpu.FsaiPrecond().setup(P, int(z["fsai_lfil"]))
    ia, ja, a = F.get_factors()
    G = S.csr_to_dense(n, ia, ja, a)
    Gref = S.csr_to_dense(n, z["fsai_ia"], z["fsai_ja"], z["fsai_a"])
    assert (ia == z["fsai_ia"]).all()
    assert ((G != 0) == (Gref != 0)).all(), "pattern differs from the reference's KNN pattern"
    assert np.abs(G - Gref).max() <= 1e-10 * np.abs(Gref).max()
    for i in range(n):                                   # diagonal stored last in every row (INC/_preconds.h:271)
        assert ja[ia[i + 1] - 1] == i
    x = F.solve(z["rhs"])
    assert rel_l2(x, z["fsai_z"]) <= 1e-10
    # log det of (G'G)^{-1} (fsai.c:278-300) from the reference's factor values
    ld_ref = -2.0 * np.log(z["fsai_a"][z["fsai_ia"][1:] - 1]).sum()
    ld = gpu.lib().Nfft4GPPrecondFsaiLogdet(F.h)
    assert abs(ld - ld_ref) <= 1e-10 * abs(ld_ref)
    F.free()
    P.free()


def test_fsai_gradient_side_matches_reference_golden(gpu):
    """Rows f2: FSAI set up with require_grad against the reference's compiled fsai.c (tests/golden/fsai_grad_gauss300.npz):
    derivative factor values (fsai.c:495-560), Nfft4GPPrecondFsaiTrace (:221), Nfft4GPPrecondFsaiDvp (:125), the sparse triangular
    solves Nfft4GPPrecondFsaiInvL / InvLT (:675-721, synchronisation-free on the GPU), and Nfft4GPPrecondFsaiSetupWithKernelPattern
    (:314) fed with the reference's own pattern."""
    z, g = _load("fsai_gauss300.npz"), _load("fsai_grad_gauss300.npz")
    X = z["X"]
    n = X.shape[1]
    P = gpu.DensePlainKernel(X, "gaussian").set_hyper(float(z["f"]), float(z["l"]), float(z["mu"]))
    for pattern in (None, (z["fsai_ia"], z["fsai_ja"])):
        F = gpu.FsaiPrecond().setup(P, int(z["fsai_lfil"]), require_grad=1, pattern=pattern)
        ia, ja, a = F.get_factors()
        da = F.grad_factors()
        Gref = S.csr_to_dense(n, z["fsai_ia"], z["fsai_ja"], z["fsai_a"])
        assert np.abs(S.csr_to_dense(n, ia, ja, a) - Gref).max() <= 1e-10 * np.abs(Gref).max()
        for t in range(3):
            dG, dGref = S.csr_to_dense(n, ia, ja, da[t]), S.csr_to_dense(n, z["fsai_ia"], z["fsai_ja"], g["dL"][t])
            assert np.abs(dG - dGref).max() <= 1e-8 * max(np.abs(dGref).max(), 1.0), t
        assert np.abs(F.trace() - g["trace"]).max() <= 1e-9 * np.abs(g["trace"]).max()
        assert rel_l2(F.inv_l(z["rhs"]), g["invl"]) <= 1e-10
        assert rel_l2(F.inv_l(z["rhs"], transpose=True), g["invlt"]) <= 1e-10
        y = F.dvp(z["rhs"])
        for t in range(3):
            assert rel_l2(y[t * n:(t + 1) * n], g["dvp"][t * n:(t + 1) * n]) <= 1e-8, t
        ym = F.dvp(z["rhs"], mask=[0, 1, 0])
        assert rel_l2(ym[n:2 * n], g["dvp"][n:2 * n]) <= 1e-8 and not ym[:n].any() and not ym[2 * n:].any()
        F.free()
    # without require_grad the gradient side refuses, as the reference does (fsai.c:141-145)
    F = gpu.FsaiPrecond().setup(P, int(z["fsai_lfil"]))
    with pytest.raises(RuntimeError):
        F.trace()
    F.free()
    P.free()


def test_fsai_triangular_solves_at_scale(gpu):
    """The synchronisation-free sparse triangular solves on a pattern with long dependency chains (n = 20000, lfil 50, additive
    kernel): G (G^{-1} b) = b and G' (G^{-T} b) = b through the CSR products of the apply."""
    n, d, lfil = 20000, 6, 50
    X = uniform_cloud(n, d, seed=77)
    D = gpu.DenseAdditiveKernel(X, np.arange(6, dtype=np.int32), 2, 3, "gaussian").set_hyper(1.0, 0.8, 0.05)
    F = gpu.FsaiPrecond().setup(D, lfil, require_grad=1)
    ia, ja, a = F.get_factors()
    import scipy.sparse as sp
    G = sp.csr_matrix((a, ja, ia), shape=(n, n))
    b = rand_vec(n, 3)
    assert rel_l2(G @ F.inv_l(b), b) <= 1e-11
    assert rel_l2(G.T @ F.inv_l(b, transpose=True), b) <= 1e-11
    tr = F.trace()
    assert abs(tr[0] + 2.0 * n / 1.0) <= 1e-9 * n          # d/df: dG = -G / f  =>  trace = -2 n / f
    F.free()
    D.free()


def test_afn_logdet(gpu):
    """Nfft4GPPrecondAFNLogdet (afn.h:99, MATLAB afn_logdet.m): 2 (sum log L_ii + sum log 1 / G_ii) from the factors of the handle."""
    n, d, k, lfil = 2000, 6, 150, 40
    X = uniform_cloud(n, d, seed=31)
    D = gpu.DenseAdditiveKernel(X, np.arange(6, dtype=np.int32), 2, 3, "gaussian").set_hyper(1.2, 0.8, 0.05)
    Af = gpu.AfnPrecond().setup(D, k, lfil, numests=0, noran=1)
    F = Af.get_factors()
    diagG = F["a"][F["ia"][1:] - 1]
    ref = -2.0 * np.log(np.diag(F["Linv"])).sum() - 2.0 * np.log(diagG).sum()
    assert abs(Af.logdet() - ref) <= 1e-10 * abs(ref)
    # against the log det of the explicit preconditioner matrix M = (apply)^{-1} on a small problem
    n2, k2 = 300, 40
    X2 = uniform_cloud(n2, d, seed=32)
    D2 = gpu.DenseAdditiveKernel(X2, np.arange(6, dtype=np.int32), 2, 3, "gaussian").set_hyper(1.2, 0.8, 0.05)
    A2 = gpu.AfnPrecond().setup(D2, k2, 20, numests=0, noran=1)
    Minv = np.column_stack([A2.solve(e) for e in np.eye(n2)])
    assert abs(A2.logdet() + np.linalg.slogdet(Minv)[1]) <= 1e-8 * abs(A2.logdet())
    for h in (Af, D, A2, D2):
        h.free()


def test_fsai_setup_additive_kernel(gpu):
    n, d, lfil = 2500, 9, 40
    X = uniform_cloud(n, d, seed=77)
    win = np.arange(9, dtype=np.int32)
    f, l, mu = 1.1, 0.9, 0.05
    D = gpu.DenseAdditiveKernel(X, win, 3, 3, "gaussian").set_hyper(f, l, mu)
    F = gpu.FsaiPrecond().setup(D, lfil)
    ia, ja, a = F.get_factors()
    wins = S.window_dims(win, 3, 3)
    pattern = S.knn_pattern(X, lfil)
    vals = S.fsai_rows(lambda J: S.dense_kernel(X, J, J, f, l, "gaussian", wins) + mu * f * f * np.eye(len(J)), pattern)
    for i in range(n):
        assert (ja[ia[i]:ia[i + 1]] == pattern[i]).all()
        assert np.abs(a[ia[i]:ia[i + 1]] - vals[i]).max() <= 1e-9 * np.abs(vals[i]).max()
    r = rand_vec(n, 3)
    G = S.rows_to_dense(n, pattern, vals)
    assert rel_l2(F.solve(r), G.T @ (G @ r)) <= 1e-10
    F.free()
    D.free()


# ------------------------------------------------------------------------------------------------ Nystrom (N2)
def test_nystrom_setup_matches_reference_golden(gpu):
    z = _load("nys_gauss300.npz")
    X = z["X"]
    n, k = X.shape[1], int(z["k"])
    P = gpu.DensePlainKernel(X, "gaussian").set_hyper(float(z["f"]), float(z["l"]), float(z["mu"]))
    N = gpu.NysPrecond().setup(P, z["perm"], k)
    perm, U, s, eta = N.get_factors()
    assert (perm == z["perm"]).all() and abs(eta - float(z["eta"])) <= 1e-15
    assert np.abs(U.T @ U - np.eye(k)).max() <= 1e-6     # orthonormal up to the conditioning of eig(U'U)
    big = z["s"] < 0.5 / eta
    assert np.abs(s[big] - z["s"][big]).max() <= 1e-6 * np.abs(z["s"][big]).max()
    x = N.solve(z["rhs"])
    assert rel_l2(x, z["z"]) <= 1e-6                     # reference's own apply of its own factors
    Uo, so, _ = S.nys_setup(X, z["perm"], k, float(z["f"]), float(z["l"]), float(z["mu"]))
    assert rel_l2(x, S.nys_apply(z["perm"], Uo, so, eta, z["rhs"])) <= 1e-6
    N.free()
    P.free()


def test_nystrom_setup_additive_and_pcg(gpu):
    """TESTS/TEST3/foo3.cpp:262-291: NFFT additive matvec + Nystrom from the dense additive kernel, random permutation."""
    n, d, k = 4000, 9, 300
    X = uniform_cloud(n, d)
    win = np.arange(9, dtype=np.int32)
    f, l, mu = 1.0, 0.3 * n ** (1.0 / d), 0.01
    wins = S.window_dims(win, 3, 3)
    D = gpu.DenseAdditiveKernel(X, win, 3, 3, "gaussian").set_hyper(f, l, mu)
    perm = np.random.default_rng(807).permutation(n).astype(np.int32)
    N = gpu.NysPrecond().setup(D, perm, k)
    Uo, so, eta = S.nys_setup(X, perm, k, f, l, mu, "gaussian", wins)
    r = rand_vec(n, 5)
    assert rel_l2(N.solve(r), S.nys_apply(perm, Uo, so, eta, r)) <= 1e-6
    A = gpu.AdditiveNFFTKernel(X, win, 3, 3, "gaussian").build(f, l, mu)
    x0, rel0, h0, it0 = gpu.solver_pcg(A.K, n, A.matvec_fn(), None, None, np.zeros(n), r, 500, 0, 1e-6)
    x1, rel1, h1, it1 = gpu.solver_pcg(A.K, n, A.matvec_fn(), N.h, N.solve_fn(), np.zeros(n), r, 500, 0, 1e-6)
    # pcg.c quirk: *piter = 0 when the iteration limit is hit without convergence (plain CG needs > 500 here)
    assert 0 < it1 and (it0 == 0 or it1 < it0), (it1, it0)
    y = A.matvec(x1)
    assert rel_l2(y, r) <= 2e-6
    A.free()
    N.free()
    D.free()


# ------------------------------------------------------------------------------------------------ AFN (A2)
@pytest.mark.parametrize("kernel,n,k,lfil", [("gaussian", 1500, 60, 24), ("matern12", 1200, 40, 16)])
def test_afn_setup_matches_numpy(gpu, kernel, n, k, lfil):
    d = 6
    X = uniform_cloud(n, d, seed=31)
    win = np.arange(6, dtype=np.int32)
    wins = S.window_dims(win, 2, 3)
    f, l, mu = 1.2, 0.8, 0.05
    D = gpu.DenseAdditiveKernel(X, win, 2, 3, kernel).set_hyper(f, l, mu)
    Af = gpu.AfnPrecond().setup(D, k, lfil, numests=0, noran=1)     # fixed rank: no estimation, always the AFN form
    F = Af.get_factors()
    Fo = S.afn_setup(X, k, lfil, f, l, mu, kernel, wins)
    assert F["k"] == k and not F["is_nys"]
    assert (F["perm"] == Fo["perm"]).all()
    assert np.abs(F["Linv"] - np.linalg.inv(Fo["L"])).max() <= 1e-9 * np.abs(np.linalg.inv(Fo["L"])).max()
    assert np.abs(F["K12"] - Fo["K12"]).max() <= 1e-13
    n2 = n - k
    G = S.csr_to_dense(n2, F["ia"], F["ja"], F["a"])
    assert ((G != 0) == (Fo["G"] != 0)).all()
    assert np.abs(G - Fo["G"]).max() <= 1e-8 * np.abs(Fo["G"]).max()
    r = rand_vec(n, 8)
    assert rel_l2(Af.solve(r), S.afn_apply(Fo, r)) <= 1e-8
    Af.free()
    D.free()


def test_afn_setup_matches_reference_compiled_parts(gpu):
    """Row A2 pinned on the reference's COMPILED code: the GPU AFN setup against the chain Nfft4GPSortFps (ordering.c:713) ->
    Nfft4GPExpandPerm -> Nfft4GPPrecondCholSetupWithKernel (stable) -> Nfft4GPKernelSchurCombineKernel (kernels.c:3496-4009) ->
    Nfft4GPPrecondFsaiSetupWithKernel, and its apply against afn.c:109-136 composed from the reference's Chol / FSAI solves
    (tests/golden/afn_ref_parts400.npz, scripts/make_golden.py:afn_case)."""
    z = np.load(os.path.join(GOLD, "afn_ref_parts400.npz"))
    X, f, l, mu, k, lfil = z["X"], float(z["f"]), float(z["l"]), float(z["mu"]), int(z["k"]), int(z["lfil"])
    n = X.shape[1]
    P = gpu.DensePlainKernel(X, "gaussian").set_hyper(f, l, mu)
    Af = gpu.AfnPrecond().setup(P, k, lfil, numests=0, noran=1)
    F = Af.get_factors()
    assert F["k"] == k and not F["is_nys"]
    assert (F["perm"][:k] == z["fps_sel"]).all(), "landmarks differ from Nfft4GPSortFps"
    assert (F["perm"] == z["perm"]).all()
    Linv = np.linalg.inv(z["L"])
    assert np.abs(F["Linv"] - Linv).max() <= 1e-9 * np.abs(Linv).max()
    assert np.abs(F["K12"] - z["K12"]).max() <= 1e-13
    n2 = n - k
    G, Gref = S.csr_to_dense(n2, F["ia"], F["ja"], F["a"]), S.csr_to_dense(n2, z["fsai_ia"], z["fsai_ja"], z["fsai_a"])
    assert (F["ia"] == z["fsai_ia"]).all()
    assert ((G != 0) == (Gref != 0)).all(), "Schur FSAI pattern differs from the reference"    # the order inside a row is free ...
    assert (F["ja"][F["ia"][1:] - 1] == np.arange(n2)).all()                                  # ... except the diagonal, stored last
    assert np.abs(G - Gref).max() <= 1e-8 * np.abs(Gref).max()
    assert rel_l2(Af.solve(z["rhs"]), z["z"]) <= 1e-8
    Af.free()
    P.free()


def test_rank_estimate_matches_reference_nys_scaled(gpu):
    """Nfft4GPB200RankestNysScaledSample (FPS and kernel block on the GPU, one elimination sweep) against the reference's
    compiled Nfft4GPRankestNysScaled (rankest.c:354) on the sample rows its rand() stream drew."""
    z = np.load(os.path.join(GOLD, "afn_ref_parts400.npz"))
    for cfg, rows, rk in zip(z["rankest_cfg"], z["rankest_rows"], z["rankest_rank"]):
        nn, dd, ll, m, maxr = int(cfg[0]), int(cfg[1]), float(cfg[2]), int(cfg[3]), int(cfg[4])
        X = np.random.default_rng(int(z["rankest_seed"])).random((dd, nn)) * nn ** (1.0 / dd)
        P = gpu.DensePlainKernel(X, "gaussian").set_hyper(1.0, ll, 0.01)
        got = gpu.rankest_nys_scaled_sample(P, rows[:m], maxr)
        if got <= int(np.floor(max(m // 50, 1) * nn / m)):
            got = 0                       # rankest.c:378-383
        assert got == int(rk), (cfg, got, int(rk))
        P.free()
    # the same first stage inside the AFN setup (rank_method = 1), against the NumPy restatement with the product's sampler
    n, d = 3000, 3
    X = uniform_cloud(n, d, seed=12)
    for l, expect_nys in ((40.0, True), (0.4, False)):
        P = gpu.DensePlainKernel(X, "gaussian").set_hyper(1.0, l, 0.01)
        Af = gpu.AfnPrecond().setup(P, 80, 20, est_m=300, numests=2, noran=0, seed=906, rank_method=1)
        F = Af.get_factors()
        assert F["is_nys"] == expect_nys, (l, F["k_est"])
        assert np.isfinite(Af.solve(rand_vec(n, 2))).all()
        Af.free()
        P.free()


def test_afn_k12_fp32_storage(gpu):
    """Optional fp32 storage of the K12 block (Nfft4GPB200PrecondAFNSetK12Storage): the apply agrees with the fp64 one to
    fp32 rounding, stays symmetric, returns the rounded block through GetFactors, and PCG needs the same iterations."""
    n, d, k, lfil = 3000, 6, 300, 100
    X = uniform_cloud(n, d)
    win = np.arange(6, dtype=np.int32)
    f, l, mu = 1.0, 1.0, 0.01
    D = gpu.DenseAdditiveKernel(X, win, 2, 3, "gaussian").set_hyper(f, l, mu)
    A64 = gpu.AfnPrecond().setup(D, k, lfil, numests=0, noran=1)
    A32 = gpu.AfnPrecond().set_k12_storage(32).setup(D, k, lfil, numests=0, noran=1)
    F64, F32 = A64.get_factors(), A32.get_factors()
    assert (F32["K12"] == F64["K12"].astype(np.float32).astype(np.float64)).all()
    r, q = rand_vec(n, 8), rand_vec(n, 9)
    z64, z32 = A64.solve(r), A32.solve(r)
    assert 0 < rel_l2(z32, z64) <= 1e-5
    assert abs(q @ z32 - r @ A32.solve(q)) <= 1e-10 * np.linalg.norm(q) * np.linalg.norm(z32)    # symmetric apply
    A = gpu.AdditiveNFFTKernel(X, win, 2, 3, "gaussian").build(f, l, mu)
    rhs = rand_vec(n, 4)
    x1, rel1, _, it1 = gpu.solver_pcg(A.K, n, A.matvec_fn(), A64.h, A64.solve_fn(), np.zeros(n), rhs, 800, 0, 1e-6)
    x2, rel2, _, it2 = gpu.solver_pcg(A.K, n, A.matvec_fn(), A32.h, A32.solve_fn(), np.zeros(n), rhs, 800, 0, 1e-6)
    assert it1 > 0 and abs(it2 - it1) <= 1 and rel2 <= 1e-6
    assert rel_l2(x2, x1) <= 1e-5
    # switching an existing fp64 handle converts in place; going back needs a new setup
    A64.set_k12_storage(32)
    assert rel_l2(A64.solve(r), z32) <= 1e-12
    with pytest.raises(RuntimeError):
        A64.set_k12_storage(64)
    for h in (A, A64, A32, D):
        h.free()


def test_nystrom_reference_quirk_mode(gpu, monkeypatch):
    """NFFT4GP_B200_REFERENCE_QUIRKS=1: the Nystrom setup on the dense additive kernel reproduces the landmark block the REFERENCE
    factors (tests/golden/nys_additive_quirk300.npz, compiled nys.c + kernels.c); the default (intended rows) differs from it."""
    z = _load("nys_additive_quirk300.npz")
    X, f, l, mu, k, perm = z["X"], float(z["f"]), float(z["l"]), float(z["mu"]), int(z["k"]), z["perm"]
    D = gpu.DenseAdditiveKernel(X, z["windows"], 2, 3, "gaussian").set_hyper(f, l, mu)
    monkeypatch.setenv("NFFT4GP_B200_REFERENCE_QUIRKS", "1")
    Nq = gpu.NysPrecond().setup(D, perm, k)
    assert rel_l2(Nq.solve(z["rhs"]), z["z"]) <= 1e-6          # eig(U'U) conditioning, as test_nystrom_setup_matches_reference_golden
    monkeypatch.delenv("NFFT4GP_B200_REFERENCE_QUIRKS")
    Ni = gpu.NysPrecond().setup(D, perm, k)
    assert rel_l2(Ni.solve(z["rhs"]), z["z"]) >= 1e-3
    for h in (Nq, Ni, D):
        h.free()


def test_afn_k12_through_the_nfft_operator(gpu):
    """Nfft4GPB200PrecondAFNSetK12Operator: the two K12 products of the AFN apply as NFFT matvecs (rows of K applied to vectors that
    vanish on the other block).  The apply agrees with the dense-block apply to the NFFT accuracy, stays symmetric, PCG needs the
    same number of iterations and reaches the same solution; the dense block can be released afterwards."""
    n, d, k, lfil = 3000, 6, 300, 100
    X = uniform_cloud(n, d)
    win = np.arange(6, dtype=np.int32)
    f, l, mu = 1.0, 1.0, 0.01
    D = gpu.DenseAdditiveKernel(X, win, 2, 3, "gaussian").set_hyper(f, l, mu)
    A = gpu.AdditiveNFFTKernel(X, win, 2, 3, "gaussian").build(f, l, mu)
    P = gpu.AfnPrecond().setup(D, k, lfil, numests=0, noran=1)
    r, q = rand_vec(n, 8), rand_vec(n, 9)
    z_dense = P.solve(r)
    rhs = rand_vec(n, 4)
    x1, rel1, _, it1 = gpu.solver_pcg(A.K, n, A.matvec_fn(), P.h, P.solve_fn(), np.zeros(n), rhs, 800, 0, 1e-6)
    P.set_k12_operator(A)
    z_op = P.solve(r)
    err = rel_l2(z_op, z_dense)
    assert 0 < err <= 1e-5, err                     # the NFFT approximation of the kernel block (measured below 1e-6 at this length scale)
    assert abs(q @ z_op - r @ P.solve(q)) <= 1e-9 * np.linalg.norm(q) * np.linalg.norm(z_op)    # symmetric apply
    x2, rel2, _, it2 = gpu.solver_pcg(A.K, n, A.matvec_fn(), P.h, P.solve_fn(), np.zeros(n), rhs, 800, 0, 1e-6)
    assert it1 > 0 and abs(it2 - it1) <= 1 and rel2 <= 1e-6, (it1, it2)
    assert rel_l2(x2, x1) <= 1e-5
    # back to the dense block, then release it for good
    P.set_k12_operator(None)
    assert rel_l2(P.solve(r), z_dense) <= 1e-14
    P.set_k12_operator(A, release_block=True)
    assert rel_l2(P.solve(r), z_op) <= 1e-12
    with pytest.raises(RuntimeError):
        P.set_k12_operator(None)
    for h in (A, P, D):
        h.free()


def test_afn_rank_estimate_and_branches(gpu):
    n, d = 3000, 3
    X = uniform_cloud(n, d, seed=12)
    for l, expect_nys in ((40.0, True), (0.4, False)):
        P = gpu.DensePlainKernel(X, "gaussian").set_hyper(1.0, l, 0.01)
        Af = gpu.AfnPrecond().setup(P, 80, 20, est_m=300, numests=2, noran=0, seed=906)
        F = Af.get_factors()
        k_or = S.rank_estimate(X, 1.0, l, 0.01, 80, m=300, numests=2, seed=906)
        assert F["k_est"] == k_or, (F["k_est"], k_or)
        assert F["is_nys"] == expect_nys
        assert F["k"] == min(max(k_or, 1), 80)
        r = rand_vec(n, 2)
        x = Af.solve(r)
        assert np.isfinite(x).all()
        Af.free()
        P.free()


def test_afn_pcg_on_nfft_matvec(gpu):
    """BASELINE config C3 in small: AFN-preconditioned CG on the NFFT additive matvec converges to 1e-6 in fewer
    iterations than plain CG and to the same solution; iteration count +-1 against the oracle's PCG driven by the oracle
    matvec and the NumPy restatement of the same AFN factors."""
    # the setting of MATLAB test_solve_additive.m:33-47 (n = 3000, two 3-D windows, maxrank 300, lfil 100); NumPy probe on
    # the dense kernel: plain CG 523 iterations, AFN(300, 100) 49, at l = 1.0
    n, d, k, lfil = 3000, 6, 300, 100
    X = uniform_cloud(n, d)
    win = np.arange(6, dtype=np.int32)
    wins = S.window_dims(win, 2, 3)
    f, l, mu = 1.0, 1.0, 0.01
    D = gpu.DenseAdditiveKernel(X, win, 2, 3, "gaussian").set_hyper(f, l, mu)
    Af = gpu.AfnPrecond().setup(D, k, lfil, numests=0, noran=1)
    A = gpu.AdditiveNFFTKernel(X, win, 2, 3, "gaussian").build(f, l, mu)
    rhs = rand_vec(n, 4)
    x0, rel0, h0, it0 = gpu.solver_pcg(A.K, n, A.matvec_fn(), None, None, np.zeros(n), rhs, 800, 0, 1e-6)
    x1, rel1, h1, it1 = gpu.solver_pcg(A.K, n, A.matvec_fn(), Af.h, Af.solve_fn(), np.zeros(n), rhs, 800, 0, 1e-6)
    assert 0 < it1 and (it0 == 0 or 4 * it1 < it0), (it1, it0)
    assert rel1 <= 1e-6 and rel_l2(A.matvec(x1), rhs) <= 2e-6
    # oracle side: same algorithm on the CPU
    Fo = S.afn_setup(X, k, lfil, f, l, mu, "gaussian", wins)
    B = O.OracleAdditiveNFFT(X, win, 2, 3, "gaussian").build(f, l, mu)
    mv, h = B.matvec_cb()
    keep = []

    def prec(_p, nn, xp, rp):
        r = np.ctypeslib.as_array(rp, shape=(nn,))
        np.ctypeslib.as_array(xp, shape=(nn,))[:] = S.afn_apply(Fo, r.copy())
        return 0

    cb = O.SOLVE(prec)
    keep.append(cb)
    dummy = C.c_int(0)   # pcg.c:104 applies the preconditioner only when prec_data != NULL
    xo, relo, histo, ito = O.pcg(h, n, mv, C.addressof(dummy), C.cast(cb, C.c_void_p), np.zeros(n), rhs, 800, 0, 1e-6)
    assert abs(it1 - ito) <= 1, (it1, ito)
    assert rel_l2(x1, xo) <= 1e-5
    A.free()
    B.free()
    Af.free()
    D.free()


# ------------------------------------------------------------------------------------------------ edge cases
def test_setup_edge_cases(gpu):
    L = gpu.lib()
    X = uniform_cloud(90, 3, seed=5)
    P = gpu.DensePlainKernel(X, "gaussian").set_hyper(1.0, 0.9, 0.05)
    K = S.dense_matrix(X, 1.0, 0.9, 0.05)
    r = rand_vec(90, 1)
    # FSAI with lfil >= n: dense lower pattern (kernels.c:135-146) -> G'G = K^{-1} exactly
    F = gpu.FsaiPrecond().setup(P, 128)
    assert rel_l2(F.solve(r), np.linalg.solve(K, r)) <= 1e-9
    ia, ja, a = F.get_factors()
    assert ia[-1] == 90 * 91 // 2
    F.free()
    # lfil = 1: diagonal scaling 1 / sqrt(K_ii)
    F = gpu.FsaiPrecond().setup(P, 1)
    assert rel_l2(F.solve(r), r / np.diag(K)) <= 1e-13
    F.free()
    # lfil beyond the shared-memory row block is refused loudly, never silently truncated
    F = gpu.FsaiPrecond()
    L.Nfft4GPPrecondFsaiSetLfil(F.h, 200)
    X2 = uniform_cloud(400, 3, seed=6)
    P2 = gpu.DensePlainKernel(X2, "gaussian").set_hyper(1.0, 0.9, 0.05)
    rc = L.Nfft4GPPrecondFsaiSetupWithKernel(gpu.capi.dptr(P2.X), 400, 400, 3, P2.fn(), P2.h, 0, F.h)
    assert rc == -1
    # require_grad is accepted (the reference's own prediction passes 1, nfft_interface.c:936); the factors are the same
    L.Nfft4GPPrecondFsaiSetLfil(F.h, 20)
    assert L.Nfft4GPPrecondFsaiSetupWithKernel(gpu.capi.dptr(P2.X), 400, 400, 3, P2.fn(), P2.h, 1, F.h) == 0
    F.free()
    P2.free()
    # Nystrom with k = n: exact inverse of K (up to the stabilising shift); AFN with k = n: pure Cholesky block
    N = gpu.NysPrecond().setup(P, np.arange(90, dtype=np.int32), 90)
    assert rel_l2(N.solve(r), np.linalg.solve(K, r)) <= 1e-6
    N.free()
    Af = gpu.AfnPrecond().setup(P, 90, 10, numests=0, noran=1)
    assert rel_l2(Af.solve(r), np.linalg.solve(K, r)) <= 1e-8
    Af.free()
    P.free()


def test_nystrom_setup_with_foreign_kernel(gpu):
    """A caller's own CPU func_kernel (here: NumPy through a ctypes callback) is honoured by the Nystrom setup exactly as the
    reference calls it (nys.c:597: K(perm, perm[0:k]) without noise); the factorisation still runs on the GPU."""
    n, d, k = 500, 4, 40
    X = uniform_cloud(n, d, seed=8)
    f, l, mu = 1.2, 0.8, 0.03
    calls = []
    keep = []

    def fkernel(_str, data, nn, ldim, dd, permr, kr, permc, kc, Kp, dKp):
        pr = np.ctypeslib.as_array(permr, shape=(kr,)).copy()
        pc = np.ctypeslib.as_array(permc, shape=(kc,)).copy()
        calls.append((kr, kc))
        blk = np.asfortranarray(S.dense_kernel(X, pr, pc, f, l))
        import ctypes as C
        libc = C.CDLL(None)                                     # *Kp == NULL: the callee allocates with malloc, the library frees it
        libc.malloc.restype = C.c_void_p
        libc.malloc.argtypes = [C.c_size_t]
        mem = libc.malloc(8 * kr * kc)
        C.memmove(mem, blk.ctypes.data, 8 * kr * kc)
        Kp[0] = C.cast(mem, gpu.capi.c_double_p)
        return 0

    import ctypes as C
    FK = C.CFUNCTYPE(C.c_int, C.c_void_p, gpu.capi.c_double_p, C.c_int, C.c_int, C.c_int, gpu.capi.c_int_p, C.c_int, gpu.capi.c_int_p, C.c_int,
                     C.POINTER(gpu.capi.c_double_p), C.POINTER(gpu.capi.c_double_p))
    cb = FK(fkernel)
    keep.append(cb)
    L = gpu.lib()
    kh = L.Nfft4GPKernelParamCreate(n, 1)
    gpu.capi._set_hyper(kh, f, l, mu)
    perm = np.random.default_rng(9).permutation(n).astype(np.int32)
    N = gpu.NysPrecond()
    L.Nfft4GPPrecondNysSetPerm(N.h, gpu.capi.iptr(perm), 0)
    L.Nfft4GPPrecondNysSetRank(N.h, k)
    assert L.Nfft4GPPrecondNysSetupWithKernel(gpu.capi.dptr(X), n, n, d, C.cast(cb, C.c_void_p), kh, 0, N.h) == 0
    assert calls == [(n, k)]
    Uo, so, eta = S.nys_setup(X, perm, k, f, l, mu)
    r = rand_vec(n, 2)
    assert rel_l2(N.solve(r), S.nys_apply(perm, Uo, so, eta, r)) <= 1e-6
    N.free()
    L.Nfft4GPKernelParamFree(kh)
