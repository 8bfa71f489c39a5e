"""Independent NumPy restatement of the NFFT-based fast summation (NFFT3 fastsum with eps_I = eps_B = 0, Kaiser-Bessel
window, m = 4, sigma = 2, N = 32) used to cross-check the C oracle (tests only; small n).  Written against the published
algorithm, with numpy's FFT instead of the oracle's own radix-2 code."""
import numpy as np
from scipy.special import i0

N, NOS, M = 32, 64, 4
B = np.pi * (2.0 - 0.5)


def kb_phi(x):
    t = x * NOS
    d2 = M * M - t * t
    out = np.empty_like(t)
    pos, neg = d2 > 0, d2 < 0
    s = np.sqrt(np.abs(d2))
    out[pos] = np.sinh(B * s[pos]) / (np.pi * s[pos])
    out[neg] = np.sin(B * s[neg]) / (np.pi * s[neg])
    out[~pos & ~neg] = B / np.pi
    return out


def kb_phi_hat(k):
    return i0(M * np.sqrt(B * B - (2.0 * np.pi * k / NOS) ** 2))


def radial(kind, which, r, c):
    q = (r * r / (c * c)) if kind == "gaussian" else (np.abs(r) / c)
    return q * np.exp(-q) if which else np.exp(-q)


def scale_points(cols):
    """cols: (d, n) window columns -> scaled point-major (n, d) and the scale (nfft_interface.c:150-213)."""
    x = cols - cols.mean(axis=1, keepdims=True)
    r = np.sqrt((x * x).sum(axis=0)).max()
    s = 0.25 / r if (r > 0.25 or r < 0.125) else 1.0
    return (x * s).T.copy(), s


def bhat(d, kind, which, c):
    g = np.arange(N) / N - 0.5
    grids = np.meshgrid(*([g] * d), indexing="ij")
    r = np.sqrt(sum(gg * gg for gg in grids))
    r = np.minimum(r, 0.5)
    Bk = radial(kind, which, r, c) / N ** d
    return np.fft.fftshift(np.fft.fftn(np.fft.ifftshift(Bk)))


def fastsum(pts, alpha, b):
    """f_i = sum_j alpha_j K_RF(x_i - x_j) through adjoint NFFT, multiplication by b, NFFT (complex result)."""
    n, d = pts.shape
    c = np.floor(pts * NOS).astype(int)
    lj = np.arange(2 * M + 2)
    psi = kb_phi(pts[:, :, None] - (c[:, :, None] - M + lj[None, None, :]) / NOS)  # (n, d, 10)
    idx = (c[:, :, None] - M + lj[None, None, :]) % NOS
    g = np.zeros((NOS,) * d, dtype=complex)
    for j in range(n):
        w = psi[j, 0]
        for t in range(1, d):
            w = np.multiply.outer(w, psi[j, t])
        np.add.at(g, np.ix_(*[idx[j, t] for t in range(d)]), alpha[j] * w)
    ghat = np.fft.ifftn(g) * g.size  # e^{+}
    k = np.arange(-N // 2, N // 2)
    sel = np.ix_(*([k % NOS] * d))
    ph = 1.0 / kb_phi_hat(k)
    dec = ph
    for t in range(1, d):
        dec = np.multiply.outer(dec, ph)
    fhat = ghat[sel] * dec * b
    hh = np.zeros((NOS,) * d, dtype=complex)
    hh[sel] = fhat * dec
    h = np.fft.fftn(hh)  # e^{-}
    f = np.zeros(n, dtype=complex)
    for j in range(n):
        w = psi[j, 0]
        for t in range(1, d):
            w = np.multiply.outer(w, psi[j, t])
        f[j] = (h[np.ix_(*[idx[j, t] for t in range(d)])] * w).sum()
    return f


def additive_matvec(X, windows, nw, dw, kind, f, l, mu, x, alpha=1.0, beta=0.0, y=None):
    n = X.shape[1]
    win = [w for w in np.asarray(windows).reshape(nw, dw)]
    acc = np.zeros(n)
    s = alpha / nw
    for w in win:
        feats = [int(t) for t in w if t >= 0]
        pts, sc = scale_points(X[feats])
        c = l * sc * np.sqrt(2.0) if kind == "gaussian" else l * sc
        fw = fastsum(pts, s * x, bhat(len(feats), kind, 0, c))
        acc += f * f * (fw.real + mu * s * x)
    return (beta * y if (y is not None and beta != 0.0) else 0.0) + acc
