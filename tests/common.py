"""Shared helpers of the test-suite: reference-style synthetic data."""
import numpy as np


def uniform_cloud(n, d, seed=906):
    """Nfft4GPDatasetUniformRandom (SRC/utils/utils.c:266-286): n^(1/d) * U(0,1) per coordinate, feature-major (d, n)."""
    rng = np.random.default_rng(seed)
    return rng.random((d, n)) * n ** (1.0 / d)


def gaussian_cloud(n, d, seed=907):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((d, n)) * n ** (1.0 / d) / np.sqrt(12.0)


def rand_vec(n, seed=1):
    """x = U(0,1) - 0.5 (TESTS/TEST1/foo.cpp:243-247)."""
    return np.random.default_rng(seed).random(n) - 0.5


def windows_consecutive(nwindows, dwindows):
    import numpy as np
    return np.arange(nwindows * dwindows, dtype=np.int32)


def rel_l2(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))


# (name, n, d, windows, nwindows, dwindows, kernel, f, l, mu)
MATVEC_CASES = [
    ("c1_3x3d_gauss", 10000, 9, list(range(9)), 3, 3, "gaussian", 1.2, 2.0, 0.01),
    ("small_3x3d_gauss", 1500, 9, list(range(9)), 3, 3, "gaussian", 1.0, 1.0, 0.05),
    ("bike_pattern", 2000, 13, [1, 8, 6, 9, 4, 2, 5, 3, 11], 3, 3, "gaussian", 0.9, 1.5, 0.01),
    ("padded_last", 1200, 5, [0, 1, 3, 2, 4, -1], 2, 3, "gaussian", 1.1, 1.0, 0.02),
    ("2d_windows", 3000, 4, [0, 1, 2, 3], 2, 2, "gaussian", 1.0, 1.0, 0.01),
    ("1d_windows", 2500, 3, [0, 1, 2], 3, 1, "gaussian", 1.3, 3.0, 0.01),
    ("matern_3d", 2000, 6, list(range(6)), 2, 3, "matern12", 1.0, 2.0, 0.01),
    ("matern_1d2d", 2000, 3, [0, 1, 2, -1], 2, 2, "matern12", 0.8, 1.0, 0.1),
    ("tiny", 7, 3, [0, 1, 2], 1, 3, "gaussian", 1.0, 1.0, 0.01),
]
