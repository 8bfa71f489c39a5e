"""Parity on REAL data: 3000 sampled points of the reference's `bike` and `poletele` data sets (the inputs of its NFFT-vs-dense
experiment, TESTS/TEST1) with the shipped window files -- permuted feature columns, and for poletele a padded last window --
against the additive matvec of the reference's own wrapper (tests/golden/real_*3000.npz, scripts/make_realdata_fixture.py).
Real features are discrete and clustered: many points per grid cell, most cells empty."""
import os

import numpy as np
import pytest

from common import rel_l2

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [("bike", "g", "gaussian"), ("bike", "m", "matern12"), ("poletele", "g", "gaussian"), ("poletele", "m", "matern12")]


def _load(name, tag):
    z = np.load(os.path.join(GOLD, f"real_{name}3000.npz"))
    return z, z[f"{tag}_windows"], int(z[f"{tag}_nw"]), int(z[f"{tag}_dw"])


@pytest.mark.parametrize("name,tag,kern", CASES)
def test_oracle_matches_reference_wrapper_on_real_data(name, tag, kern):
    from oracle import oracle_api as O
    z, win, nw, dw = _load(name, tag)
    B = O.OracleAdditiveNFFT(z["X"], win, nw, dw, kern).build(float(z["f"]), float(z["l"]), float(z["mu"]))
    assert rel_l2(B.matvec(z["x"]), z[f"{tag}_y"]) <= 1e-13
    y0 = np.cos(np.arange(3000.0))
    assert rel_l2(B.matvec(z["x"], alpha=-0.7, beta=0.4, y=y0.copy()), z[f"{tag}_y_ab"]) <= 1e-13
    B.free()


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [-1, 5, 1, 0], ids=["auto", "dmma", "dfma", "generic"])
@pytest.mark.parametrize("name,tag,kern", CASES)
def test_gpu_matches_reference_wrapper_on_real_data(gpu, name, tag, kern, variant):
    z, win, nw, dw = _load(name, tag)
    L = gpu.lib()
    L.Nfft4GPB200SetVariant(b"spread3d", variant)
    L.Nfft4GPB200SetVariant(b"gather3d", variant)
    try:
        A = gpu.AdditiveNFFTKernel(z["X"], win, nw, dw, kern).build(float(z["f"]), float(z["l"]), float(z["mu"]))
        assert rel_l2(A.matvec(z["x"]), z[f"{tag}_y"]) <= 1e-10
        y0 = np.cos(np.arange(3000.0))
        assert rel_l2(A.matvec(z["x"], alpha=-0.7, beta=0.4, y=y0.copy()), z[f"{tag}_y_ab"]) <= 1e-10
        g = A.grad_matvec(z["x"])
        assert rel_l2(g[:3000], 2.0 / float(z["f"]) * (A.matvec(z["x"]))) <= 1e-12
        A.free()
    finally:
        L.Nfft4GPB200SetVariant(b"spread3d", -1)
        L.Nfft4GPB200SetVariant(b"gather3d", -1)
