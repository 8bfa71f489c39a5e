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
