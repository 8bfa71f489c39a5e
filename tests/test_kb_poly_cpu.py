"""CPU checks of the algorithm behind the table-free 1-D / 2-D paths (nfft1d_poly.cu, nfft2d_poly.cu), in NumPy:
the degree-14 polynomials of csrc/kb_window_coeffs.h against the Kaiser-Bessel window of the reference's NFFT3 (m = 4, sigma = 2;
the same formula as k_psi in nfft_kernels.cu and the oracle's shim), and the two identities the kernels rest on --
   spread:  sum_j a_j psi(x_j) (taps)  ==  P applied to the per-cell moments  M_c[k] = sum_{j in c} a_j v_j^k
   gather:  sum_t h[c - 4 + t] psi_t(x_j)  ==  C_c(v_j),  C_c = sum_t h[c - 4 + t] p_t
in one and two dimensions (2-D: M_cell = sum a v1^k1 v2^k2, g += P M P';  C_cell = P' H P)."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200", "csrc", "kb_window_coeffs.h")
M, TAPS, NOS = 4, 10, 64


def load_table():
    txt = open(HDR).read()
    deg = int(re.search(r"#define KB_POLY_DEG (\d+)", txt).group(1))
    vals = [float.fromhex(t) for t in re.findall(r"-?0x[0-9a-fA-F.]+p[+-]?\d+", txt)]
    P = np.array(vals).reshape(TAPS, deg + 1)
    return P


def kb_phi(t):
    """phi(t) of the Kaiser-Bessel window, t in grid units (nfft_kernels.cu:k_psi; NFFT3 PHI with b = pi (2 - 1/sigma))."""
    b = np.pi * (2.0 - 0.5)
    s = M * M - t * t
    out = np.full_like(t, b / np.pi)
    pos, neg = s > 0, s < 0
    r = np.sqrt(np.abs(s))
    out[pos] = np.sinh(b * r[pos]) / (np.pi * r[pos])
    out[neg] = np.sin(b * r[neg]) / (np.pi * r[neg])
    return out


def taps_of(x):
    """cell key field c + 32, taps psi_t = phi(64 x - (c - 4 + t)), t = 0..9, and the in-cell offset v = 2 u - 1."""
    t64 = x * NOS
    c = np.floor(t64)
    u = t64 - c
    psi = np.stack([kb_phi(u + M - t) for t in range(TAPS)], axis=-1)
    return (c.astype(int) + 32) & 63, psi, 2.0 * u - 1.0


def test_polynomials_match_the_window():
    P = load_table()
    assert P.shape == (10, 15)
    phi0 = kb_phi(np.zeros(1))[0]
    # double Horner evaluation against NumPy's window: the argument of sinh is ~ 19, so NumPy's own value carries ~ 19 ulp of error
    u = np.random.default_rng(0).random(20000)
    v = 2.0 * u - 1.0
    for t in range(TAPS):
        err = np.abs(np.polyval(P[t][::-1], v) - kb_phi(u + M - t)).max() / phi0
        assert err <= 1e-14, (t, err)
    # ... and against the window in 40-digit arithmetic where mpmath is available: the header's figure (4.6e-16 phi(0))
    try:
        import mpmath as mp
    except ImportError:
        return
    mp.mp.dps = 40
    b = mp.pi * (2 - mp.mpf(1) / 2)

    def phi(t):
        s_ = M * M - t * t
        if s_ > 0:
            return mp.sinh(b * mp.sqrt(s_)) / (mp.pi * mp.sqrt(s_))
        if s_ < 0:
            return mp.sin(b * mp.sqrt(-s_)) / (mp.pi * mp.sqrt(-s_))
        return b / mp.pi
    worst = 0.0
    for uu in np.random.default_rng(1).random(150):
        vv = 2.0 * uu - 1.0
        for t in range(TAPS):
            worst = max(worst, abs(float(mp.mpf(float(np.polyval(P[t][::-1], vv))) - phi(mp.mpf(float(uu)) + M - t)) / phi0))
    assert worst <= 1e-15, worst


def test_moment_spread_and_coefficient_gather_1d():
    P = load_table()
    rng = np.random.default_rng(1)
    n = 3000
    x = (rng.random(n) - 0.5) * 0.5
    x[0], x[1] = 0.25, -0.25                    # both ends of the scaled range
    a = rng.standard_normal(n)
    key, psi, v = taps_of(x)
    # taps: g[(key + 28 + t) & 63] += a psi_t
    g_ref = np.zeros(NOS)
    for t in range(TAPS):
        np.add.at(g_ref, (key + 28 + t) & 63, a * psi[:, t])
    # moments per cell, then the 15 -> 10 transform
    Mom = np.zeros((NOS, 15))
    np.add.at(Mom, key, a[:, None] * v[:, None] ** np.arange(15)[None, :])
    g = np.zeros(NOS)
    for c in range(NOS):
        for t in range(TAPS):
            g[(c + 28 + t) & 63] += P[t] @ Mom[c]
    assert np.abs(g - g_ref).max() <= 1e-13 * np.abs(g_ref).max()
    # gather: one polynomial per cell
    h = rng.standard_normal(NOS)
    f_ref = sum(h[(key + 28 + t) & 63] * psi[:, t] for t in range(TAPS))
    C = np.stack([sum(h[(c + 28 + t) & 63] * P[t] for t in range(TAPS)) for c in range(NOS)])
    f = np.einsum("jk,jk->j", C[key], v[:, None] ** np.arange(15)[None, :])
    assert np.abs(f - f_ref).max() <= 1e-13 * np.abs(f_ref).max()


def test_bivariate_moments_and_coefficients_2d():
    P = load_table()
    rng = np.random.default_rng(2)
    n = 2000
    X = (rng.random((n, 2)) - 0.5) * 0.5
    a = rng.standard_normal(n)
    k1, psi1, v1 = taps_of(X[:, 0])
    k2, psi2, v2 = taps_of(X[:, 1])
    pw = np.arange(15)
    # taps
    g_ref = np.zeros((NOS, NOS))
    for t1 in range(TAPS):
        for t2 in range(TAPS):
            np.add.at(g_ref, ((k1 + 28 + t1) & 63, (k2 + 28 + t2) & 63), a * psi1[:, t1] * psi2[:, t2])
    # bivariate moments per occupied cell (what the Vandermonde GEMM of k_spread2_poly accumulates), then g += P M P'
    cell = k1 * NOS + k2
    g = np.zeros((NOS, NOS))
    for c in np.unique(cell):
        sel = cell == c
        A = (a[sel, None] * v1[sel, None] ** pw[None, :]).T          # [k1][j]
        B = v2[sel, None] ** pw[None, :]                             # [j][k2]
        Q = P @ (A @ B) @ P.T                                        # 10 x 10 footprint of the cell
        i1 = ((c // NOS) + 28 + np.arange(TAPS)) & 63
        i2 = ((c % NOS) + 28 + np.arange(TAPS)) & 63
        g[np.ix_(i1, i2)] += Q
    assert np.abs(g - g_ref).max() <= 1e-13 * np.abs(g_ref).max()
    # gather: C_cell = P' H_cell P, f_j = sum_{k1,k2} C[k1][k2] v1^k1 v2^k2
    h = rng.standard_normal((NOS, NOS))
    f_ref = np.zeros(n)
    for t1 in range(TAPS):
        for t2 in range(TAPS):
            f_ref += h[(k1 + 28 + t1) & 63, (k2 + 28 + t2) & 63] * psi1[:, t1] * psi2[:, t2]
    f = np.zeros(n)
    for c in np.unique(cell):
        sel = cell == c
        i1 = ((c // NOS) + 28 + np.arange(TAPS)) & 63
        i2 = ((c % NOS) + 28 + np.arange(TAPS)) & 63
        C = P.T @ h[np.ix_(i1, i2)] @ P
        T = (v1[sel, None] ** pw[None, :]) @ C                       # the tensor-core product of k_gather2_poly
        f[sel] = np.einsum("jk,jk->j", T, v2[sel, None] ** pw[None, :])
    assert np.abs(f - f_ref).max() <= 1e-13 * np.abs(f_ref).max()
