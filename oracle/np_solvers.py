"""ORACLE (test infrastructure, not product code): NumPy restatement of the reference's flexible GMRES,
SRC/solvers/fgmres.c:3-252 with Nfft4GPModifiedGS (SRC/linearalg/matops.c:274-346, tol_reorth = -1), statement by statement
including the quirks (restart scaled by the estimated residual norm; Givens breakdown leaves without updating x).
Pinned against the reference's own compiled Nfft4GPSolverFgmres through tests/golden/fgmres_dense300.npz
(scripts/make_golden.py).  Only tests/ may import this module."""
import numpy as np


def fgmres(matvec, precond, x0, rhs, kdim, maxits, atol, tol):
    """matvec(v) -> A v; precond(v) -> M^{-1} v or None.  Returns (x, rel_res, rel_res_v, iters)."""
    eps = np.finfo(np.float64).eps
    n = rhs.size
    x = np.array(x0, dtype=np.float64)
    normb = np.linalg.norm(rhs)
    if n == 0 or normb < eps:
        return np.zeros(n), 0.0, np.zeros(1), 0
    V = np.zeros((kdim + 1, n))
    Z = np.zeros((kdim + 1, n))
    H = np.zeros((kdim, kdim + 1))
    c = np.zeros(kdim)
    s = np.zeros(kdim)
    rs = np.zeros(kdim + 1)
    V[0] = rhs - matvec(x)
    normr = np.linalg.norm(V[0])
    if normr < eps:
        return x, 0.0, np.zeros(1), 0
    tolr = tol if atol else tol * normb
    hist = np.zeros(maxits + 1)
    hist[0] = normr / normb
    it = 0
    vi = 0
    while it < maxits:
        rs[0] = normr
        V[0] *= 1.0 / normr
        i = 0
        broke = False
        while i < kdim and it < maxits:
            i += 1
            it += 1
            v = V[i - 1]
            Z[i - 1] = precond(v) if precond is not None else v
            w = matvec(Z[i - 1])
            for j in range(i):
                t1 = float(w @ V[j])
                H[i - 1, j] = t1
                w = w - t1 * V[j]
            t = np.linalg.norm(w)
            H[i - 1, i] = t
            V[i] = w / t
            for j in range(1, i):
                hii = H[i - 1, j - 1]
                H[i - 1, j - 1] = c[j - 1] * hii + s[j - 1] * H[i - 1, j]
                H[i - 1, j] = -s[j - 1] * hii + c[j - 1] * H[i - 1, j]
            hii, hii1 = H[i - 1, i - 1], H[i - 1, i]
            gam = np.sqrt(hii * hii + hii1 * hii1)
            if abs(gam) < eps:
                broke = True
                break
            c[i - 1] = hii / gam
            s[i - 1] = hii1 / gam
            rs[i] = -s[i - 1] * rs[i - 1]
            rs[i - 1] = c[i - 1] * rs[i - 1]
            H[i - 1, i - 1] = c[i - 1] * hii + s[i - 1] * hii1
            normr = abs(rs[i])
            hist[it] = normr / normb
            if normr <= tolr:
                break
        if broke:
            break
        rs[i - 1] /= H[i - 1, i - 1]
        for k in range(i - 2, -1, -1):
            for j in range(k + 1, i):
                rs[k] -= H[j, k] * rs[j]
            rs[k] /= H[k, k]
        for j in range(i):
            x = x + rs[j] * Z[j]
        if normr <= tolr:
            break
        V[0] = rhs - matvec(x)
    return x, normr / normb, hist, it


def lanczos(matvec, n, x0, rhs, wsize, maxits, atol, tol, prec=None):
    """NumPy restatement of the solve part of the reference's Nfft4GPSolverLanczos (SRC/solvers/lanczos.c:3-300): preconditioned
    Lanczos with the orthogonalisation Nfft4GPModifiedGS2 (matops.c:348-434: coefficients <z, v_i>, updates with z_i, one
    re-orthogonalisation pass while ||z|| drops below 0.7071 of its previous value), the incremental Cholesky factor of T for
    the residual estimate and the two bidiagonal solves for x.  Returns (x, rel_res, history, iterations, TD, TE) with T's
    diagonal / sub-diagonal up to the iteration reached (the reference then keeps extending T to `wsize` rows, :305-394).
    TEST INFRASTRUCTURE (oracle)."""
    eps = np.finfo(np.float64).eps
    if wsize <= 0 or wsize > maxits:
        wsize = maxits
    x = np.array(x0, dtype=np.float64)
    normb = np.linalg.norm(rhs)
    if normb < eps:
        return np.zeros(n), 0.0, np.zeros(1), 0, np.zeros(0), np.zeros(0)
    apply_prec = (lambda r: prec(r)) if prec is not None else (lambda r: r.copy())
    V = np.zeros((maxits + 1, n))
    Z = np.zeros((maxits + 1, n))
    TD, TE = np.zeros(maxits + 1), np.zeros(maxits + 1)
    TLD, TLE = np.zeros(maxits + 1), np.zeros(maxits + 1)
    Z[0] = rhs - matvec(x)
    V[0] = apply_prec(Z[0])
    normr = np.linalg.norm(Z[0])
    beta = np.sqrt(V[0] @ Z[0])
    if beta < eps:
        return x, 0.0, np.zeros(1), 0, np.zeros(0), np.zeros(0)
    tolr = tol / beta if atol else tol
    hist = np.zeros(maxits + 1)
    hist[0] = normr / normb
    V[0] /= beta
    Z[0] /= beta
    it, chol_size, ls, dotvz = 0, 0, 0.0, 0.0
    while it < maxits:
        it += 1
        z = matvec(V[it - 1])
        k = min(it - 1, wsize)
        # Nfft4GPModifiedGS2 against V[0..k] / Z[0..k]; the last two coefficients are T's new sub-diagonal / diagonal entries
        normw = np.linalg.norm(z)
        first = True
        while True:
            h = np.zeros(k + 1)
            for i in range(k + 1):
                h[i] = z @ V[i]
                z = z - h[i] * Z[i]
            t = np.linalg.norm(z)
            if k >= 1:
                TE[it - 2] = h[k - 1] if first else TE[it - 2] + h[k - 1]
            TD[it - 1] = h[k] if first else TD[it - 1] + h[k]
            first = False
            if not (t < normw * 0.7071 and t >= eps):
                break
            normw = t
        Z[it] = z
        if t < eps:
            break
        V[it] = apply_prec(Z[it])
        dotvz = np.sqrt(V[it] @ Z[it])
        if dotvz < eps:
            break
        V[it] /= dotvz
        Z[it] /= dotvz
        normz = np.linalg.norm(Z[it])
        if it != 1:
            TLE[it - 2] = TE[it - 2] / TLD[it - 2]
            TLD[it - 1] = np.sqrt(TD[it - 1] - TLE[it - 2] ** 2)
            le = 1.0 / TLD[it - 1]
            ls = -ls * TLE[it - 2] * le
            normr = abs(le * ls) * dotvz * beta * normz
        else:
            TLD[0] = np.sqrt(TD[0])
            ls = 1.0 / TLD[0]
            normr = dotvz / TD[0] * beta * normz
        chol_size += 1
        hist[it] = normr / normb
        if normr <= tolr:
            break
    if chol_size > 0:
        y = np.zeros(chol_size)
        y[0] = beta / TLD[0]
        for k in range(1, chol_size):
            y[k] = (-y[k - 1] * TLE[k - 1]) / TLD[k]
        y[chol_size - 1] /= TLD[chol_size - 1]
        for k in range(chol_size - 2, -1, -1):
            y[k] = (y[k] - TLE[k] * y[k + 1]) / TLD[k]
        x = x + y @ V[:chol_size]
    return x, normr / normb, hist[:it + 1], it, TD[:it], TE[:max(it - 1, 0)]


def slq_logdet(matvec, grad_matvec, n, maxits, probes):
    """NumPy restatement of Nfft4GPLanczosQuadratureLogdet (SRC/solvers/lanczos.c:421-610) without a preconditioner:
    logdet(K) / n ~ mean_z e1' log(T_z) e1 and its three derivatives ~ mean_z (dK_j z)' K^{-1} z / n, with K^{-1} z taken from
    the same Lanczos run.  `grad_matvec(z)` returns the three stacked blocks.  TEST INFRASTRUCTURE (oracle)."""
    eps = np.finfo(np.float64).eps
    val, dval = 0.0, np.zeros(3)
    for z in probes:
        dAz = grad_matvec(z).reshape(3, n)
        x, _, _, it, TD, TE = lanczos(matvec, n, np.zeros(n), z, maxits, maxits, 0, eps)
        T = np.diag(TD) + np.diag(TE, 1) + np.diag(TE, -1)
        w, Q = np.linalg.eigh(T)
        val += float(np.sum(Q[0] ** 2 * np.log(np.abs(w))))
        dval += dAz @ x
    return val / len(probes), dval / len(probes) / n
