"""ORACLE (test infrastructure, not product code): NumPy restatement of the preconditioner SETUPS of the hot path.

Rows N2 / G2 / A2 / K1 of SURVEY.md section 8(a).  Each function cites the reference lines it follows.  Pinned against the
reference's own compiled code where the reference has one (tests/golden/nys_gauss300.npz, fsai_gauss300.npz,
solver_dense300.npz were produced by oracle/_ref = the unmodified reference sources).  SRC/preconds/afn.c itself is not
built (SURVEY.md finding 2), but every part it calls is: afn_setup / afn_apply / fps / expand_perm are pinned on the chain
Nfft4GPSortFps -> Nfft4GPExpandPerm -> Nfft4GPPrecondCholSetupWithKernel -> Nfft4GPKernelSchurCombineKernel ->
Nfft4GPPrecondFsaiSetupWithKernel run by scripts/make_golden.py:afn_case (tests/golden/afn_ref_parts400.npz: landmarks and
permutation identical, Cholesky factor 3e-15, FSAI pattern identical and values 5e-13, apply 1e-12), and
rank_estimate_nys_scaled on the compiled Nfft4GPRankestNysScaled with its rand() sample rows replayed (identical ranks).
rank_estimate (MATLAB afn_setup.m:111-293, a bisection in the 2-norm) has no compiled counterpart.
Only tests/ may import this module.
"""
import numpy as np


# ---------------------------------------------------------------------------------------------------- dense kernels (K1)
def window_dims(windows, nwindows, dwindows):
    w = np.asarray(windows).reshape(nwindows, dwindows)
    return [[int(c) for c in row if c >= 0] for row in w]


def dense_kernel(X, rows, cols, f, l, kernel="gaussian", windows=None):
    """K(rows, cols) WITHOUT the noise term.  X feature-major (d, n).
    Gaussian f^2 exp(-r^2 / (2 l^2)) (kernels.c:495,523,549); Matern-1/2 f^2 exp(-r / l) (kernels.c:2218-2222);
    additive = mean over the windows (kernels.c:3108,3187).  windows: list of feature-id lists or None (all features)."""
    X = np.asarray(X, dtype=np.float64)
    wins = [list(range(X.shape[0]))] if windows is None else windows
    out = np.zeros((len(rows), len(cols)))
    for feats in wins:
        A = X[np.ix_(feats, rows)].T
        B = X[np.ix_(feats, cols)].T
        r2 = ((A[:, None, :] - B[None, :, :]) ** 2).sum(-1)
        out += np.exp(-r2 / (2.0 * l * l)) if kernel == "gaussian" else np.exp(-np.sqrt(r2) / l)
    return f * f * out / len(wins)


def dense_matrix(X, f, l, mu, kernel="gaussian", windows=None):
    n = X.shape[1]
    idx = np.arange(n)
    return dense_kernel(X, idx, idx, f, l, kernel, windows) + f * f * mu * np.eye(n)


# ---------------------------------------------------------------------------------------------------- shared pieces
def chol_shift(A):
    """sqrt(n) * ulp(||A||_F)  (chol.c:448-455)"""
    fro = np.linalg.norm(A, "fro")
    return np.sqrt(A.shape[0]) * (np.nextafter(fro, fro + 1.0) - fro)


def fps(X, k):
    """Farthest point sampling: first point = nearest to the centroid, then arg max of the distance to the selected set,
    ties to the lowest index (ordering.c:422-711, MATLAB fps_par.m)."""
    Xp = np.asarray(X, dtype=np.float64).T
    c = Xp.mean(axis=0)
    sel = [int(np.argmin(((Xp - c) ** 2).sum(1)))]
    mind = ((Xp - Xp[sel[0]]) ** 2).sum(1)
    for _ in range(1, k):
        j = int(np.argmax(mind))
        sel.append(j)
        mind = np.minimum(mind, ((Xp - Xp[j]) ** 2).sum(1))
    return np.array(sel, dtype=np.int32)


def expand_perm(sel, n):
    """Nfft4GPExpandPerm (utils.c:208): selected first, the rest ascending."""
    mark = np.zeros(n, dtype=bool)
    mark[sel] = True
    return np.concatenate([np.asarray(sel, dtype=np.int64), np.nonzero(~mark)[0]]).astype(np.int32)


def knn_pattern(X, lfil):
    """Nfft4GPDistanceEuclidKnn (kernels.c:121-277): rows i < lfil dense lower; row i >= lfil = lfil-1 nearest among 0..i-1 (all
    features) + the diagonal last.  Returns a list of index arrays (neighbours ascending, diagonal last)."""
    Xp = np.asarray(X, dtype=np.float64).T
    n = Xp.shape[0]
    rows = []
    for i in range(n):
        if n <= lfil or i < lfil:
            rows.append(np.arange(i + 1))
            continue
        d2 = ((Xp[:i] - Xp[i]) ** 2).sum(1)
        nb = np.argpartition(d2, lfil - 2)[: lfil - 1] if lfil > 1 else np.array([], dtype=int)
        rows.append(np.concatenate([np.sort(nb), [i]]))
    return rows


def fsai_rows(Kfun, pattern):
    """fsai.c:353-399: per row, g = K(J,J)^{-1} e_last / sqrt(e_last' K(J,J)^{-1} e_last).  Kfun(J) returns the block."""
    vals = []
    for J in pattern:
        Kb = Kfun(np.asarray(J))
        e = np.zeros(len(J))
        e[-1] = 1.0
        g = np.linalg.solve(Kb, e)
        vals.append(g / np.sqrt(g[-1]))
    return vals


def dense_kernel_dl(X, rows, cols, f, l, kernel="gaussian", windows=None):
    """dK/dl of dense_kernel: Gaussian f^2 exp(-r^2 / 2 l^2) r^2 / l^3 (kernels.c:522), Matern-1/2 f^2 exp(-r / l) r / l^2 (:2220)."""
    X = np.asarray(X, dtype=np.float64)
    wins = [list(range(X.shape[0]))] if windows is None else windows
    out = np.zeros((len(rows), len(cols)))
    for feats in wins:
        A = X[np.ix_(feats, rows)].T
        B = X[np.ix_(feats, cols)].T
        r2 = ((A[:, None, :] - B[None, :, :]) ** 2).sum(-1)
        out += np.exp(-r2 / (2.0 * l * l)) * r2 / l ** 3 if kernel == "gaussian" else np.exp(-np.sqrt(r2) / l) * np.sqrt(r2) / l ** 2
    return f * f * out / len(wins)


def fsai_rows_grad(Kfun, dKfuns, pattern):
    """fsai.c:495-560 (require_grad): per row g = K^{-1} e / sqrt(e' K^{-1} e) and, for every hyper-parameter, w = -K^{-1} dK g,
    dg = w - 0.5 w_last dd g with dd = 1 / sqrt(e' K^{-1} e).  dKfuns: three callables J -> dK_theta(J, J).  Returns (vals, dvals[3])."""
    vals, dvals = [], [[], [], []]
    for J in pattern:
        J = np.asarray(J)
        Kb = Kfun(J)
        e = np.zeros(len(J))
        e[-1] = 1.0
        u = np.linalg.solve(Kb, e)
        dd = 1.0 / np.sqrt(u[-1])
        g = u * dd
        vals.append(g)
        for t in range(3):
            w = -np.linalg.solve(Kb, dKfuns[t](J) @ g)
            dvals[t].append(w - 0.5 * w[-1] * dd * g)
    return vals, dvals


def fsai_trace(vals, dvals):
    """fsai.c:221-276: 2 sum_i dG_ii / G_ii per hyper-parameter (the diagonal is the last entry of a row)."""
    return np.array([2.0 * sum(dg[-1] / g[-1] for g, dg in zip(vals, dvals[t])) for t in range(3)])


def fsai_dvp(G, dG, x):
    """fsai.c:125-219 for one hyper-parameter: y = G' ( G^{-T} dG' G^{-T} x + dG G^{-1} G^{-T} x )."""
    w1 = np.linalg.solve(G.T, x)
    w2 = np.linalg.solve(G.T, dG.T @ w1) + dG @ np.linalg.solve(G, w1)
    return G.T @ w2


def csr_to_dense(n, ia, ja, a):
    G = np.zeros((n, n))
    for i in range(n):
        G[i, ja[ia[i]:ia[i + 1]]] = a[ia[i]:ia[i + 1]]
    return G


def rows_to_dense(n, pattern, vals):
    G = np.zeros((n, n))
    for i, (J, g) in enumerate(zip(pattern, vals)):
        G[i, J] = g
    return G


# ---------------------------------------------------------------------------------------------------- Nystrom (N2)
def nys_setup(X, perm, k, f, l, mu, kernel="gaussian", windows=None):
    """nys.c:518-660 (intended semantics: K11 = K(perm[:k], perm[:k]) without noise + shift).  Returns (U, s, eta) with
    orthonormal U (n x k, rows in perm order), s = 1 / (sigma^2 + eta), eta = mu f^2."""
    K1 = dense_kernel(X, perm, perm[:k], f, l, kernel, windows)
    K11 = K1[:k, :k].copy()
    K11 += chol_shift(K11) * np.eye(k)
    Lc = np.linalg.cholesky(K11)
    B = np.linalg.solve(Lc, K1.T).T
    U, sig, _ = np.linalg.svd(B, full_matrices=False)
    eta = mu * f * f
    return U, 1.0 / (sig ** 2 + eta), eta


def nys_apply(perm, U, s, eta, r):
    """nys.c:115-173"""
    y = r[perm]
    z = U.T @ y
    out = (y - U @ z) / eta + U @ (s * z)
    x = np.zeros_like(r)
    x[perm] = out
    return x


# ---------------------------------------------------------------------------------------------------- AFN (A2, A1)
def splitmix64(state):
    state = (state + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return state, z ^ (z >> 31)


def sample_without_replacement(pool, m, state):
    pool = list(pool)
    n = len(pool)
    for i in range(m):
        state, r = splitmix64(state)
        j = i + r % (n - i)
        pool[i], pool[j] = pool[j], pool[i]
    return pool[:m], pool, state


def rank_estimate(X, f, l, mu, max_k, m=500, numests=5, seed=906, kernel="gaussian", windows=None):
    """MATLAB afn_setup.m:111-293 with the sub-sampling generator of the product (splitmix64 Fisher-Yates)."""
    d, n = X.shape
    m = min(m, n)
    state = seed
    r = 0
    for _ in range(numests):
        rows, _, state = sample_without_replacement(range(n), m, state)
        Xm = X[:, rows] * (m / n) ** (1.0 / d)
        p = fps(Xm, m)
        Km = dense_kernel(Xm, p, p, 1.0, l, kernel, windows)
        n2 = np.linalg.norm(Km, 2)
        Km = Km + np.sqrt(m) * (np.nextafter(n2, n2 + 1.0) - n2) * np.eye(m)
        nrm = np.linalg.norm(Km, "fro")
        rs, re = 1, m
        while rs < re:
            rc = (rs + re) // 2
            K11 = Km[:rc, :rc]
            K1 = Km[:, :rc]
            err = np.linalg.norm(Km - K1 @ np.linalg.solve(K11, K1.T), 2)
            if err / nrm < 0.1:
                re = rc
            else:
                rs = rc + 1
        r = max(r, rs)
    k_scaled = int(np.ceil(r * n / m))
    if k_scaled > max_k:
        return k_scaled
    mk = max(min(max_k, n), m)
    thr = 1.1 * mu / (f * f)
    r = 0
    for _ in range(numests):
        rows, _, state = sample_without_replacement(range(n), mk, state)
        scale = 1.0
        if mk > m:
            rows, _, state = sample_without_replacement(rows, m, state)
            scale = (m / mk) ** (1.0 / d)
        Xm = X[:, rows] * scale
        K = dense_kernel(Xm, np.arange(m), np.arange(m), f, l, kernel, windows) + mu * f * f * np.eye(m)
        sv = np.linalg.eigvalsh(K)
        r = max(r, int((sv > thr).sum()))
    return int(np.floor(r * mk / m))


def rank_estimate_nys_scaled(X, rows, f, l, max_rank, kernel="gaussian", windows=None):
    """One sample of Nfft4GPRankestNysScaledEstimateRank (rankest.c:248-352) on the given sample rows (the reference draws them
    with rand()): scale the sample to the density of the full set, order it by farthest point sampling, K without noise plus
    the shift nu = sqrt(n) ulp(||K||_F), then scan k = 0, ngap, 2 ngap, ... (ngap = n1 / 50) for the first scaled Nystrom error
    ||K1' K11^{-1} K1 - (K + nu I)||_F / ||K + nu I||_F < 0.1 (K11 unshifted, rankest.c:196-246), stopping early once
    (i - 1) ngap n / n1 > 2 max_rank; the result is scaled back by n / n1."""
    d, n = X.shape
    n1 = len(rows)
    small = X[:, rows] * (n1 / n) ** (1.0 / d)
    p = fps(small, n1)
    K = dense_kernel(small, p, p, f, l, kernel, windows)
    fro = np.linalg.norm(K, "fro")
    nu = np.sqrt(n) * (np.nextafter(fro, fro + 1.0) - fro)
    Ks = K + nu * np.eye(n1)
    fro = np.linalg.norm(Ks, "fro")
    ngap = max(n1 // 50, 1)
    rank = n1
    for i in range(50):
        if i * ngap >= n1 or int(np.floor(((i - 1) * ngap) * n / n1)) > 2 * max_rank:
            rank = (i - 1) * ngap
            break
        kk = i * ngap
        if kk == 0:
            err = 1.0
        else:
            B = np.linalg.solve(np.linalg.cholesky(K[:kk, :kk]), K[:kk, :])
            err = np.linalg.norm(B.T @ B - Ks, "fro") / fro
        if err < 0.1:
            rank = kk
            break
    return int(np.floor(rank * n / n1))


def afn_setup(X, k, lfil, f, l, mu, kernel="gaussian", windows=None):
    """afn_setup.m:61-101 for a given rank k (AFN branch): FPS landmarks, shifted Cholesky of K11 (with noise), K12,
    Schur complement FSAI (schurCombinedKernel.m + fsai_setup.m; lfil entries per row incl. the diagonal as fsai.c)."""
    d, n = X.shape
    sel = fps(X, k)
    perm = expand_perm(sel, n)
    p1, p2 = perm[:k], perm[k:]
    K11 = dense_kernel(X, p1, p1, f, l, kernel, windows) + mu * f * f * np.eye(k)
    K11 += chol_shift(K11) * np.eye(k)
    Lc = np.linalg.cholesky(K11)
    K12 = dense_kernel(X, p1, p2, f, l, kernel, windows)
    GK12 = np.linalg.solve(Lc, K12)
    n2 = n - k
    S = dense_kernel(X, p2, p2, f, l, kernel, windows) + mu * f * f * np.eye(n2) - GK12.T @ GK12
    pattern = knn_pattern(X[:, p2], lfil)
    vals = fsai_rows(lambda J: S[np.ix_(J, J)], pattern)
    return dict(perm=perm, L=Lc, K12=K12, pattern=pattern, vals=vals, G=rows_to_dense(n2, pattern, vals))


def afn_apply(F, r):
    """afn_solve.m:39-69"""
    perm, Lc, K12, G = F["perm"], F["L"], F["K12"], F["G"]
    k = Lc.shape[0]
    rp = r[perm]
    a11 = lambda v: np.linalg.solve(Lc.T, np.linalg.solve(Lc, v))
    xu, xl = rp[:k], rp[k:]
    zl = xl - K12.T @ a11(xu)
    yl = G.T @ (G @ zl)
    yu = a11(xu - K12 @ yl)
    out = np.zeros_like(r)
    out[perm] = np.concatenate([yu, yl])
    return out
