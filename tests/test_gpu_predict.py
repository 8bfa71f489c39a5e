"""GPU parity of the GP prediction entry point (SRC/external/nfft_interface.c:873-1068) against the reference's own compiled
function (tests/golden/predict_nfft300.npz, scripts/make_golden.py): predictive mean and standard deviation."""
import os

import numpy as np
import pytest

from common import rel_l2

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_gp_predict_matches_reference_golden(gpu):
    g = np.load(os.path.join(GOLD, "predict_nfft300.npz"))
    mean, std = gpu.gp_predict(g["x"], g["X1"], g["y"], g["X2"], g["windows"], 2, 3, "gaussian", tol=float(g["tol"]),
                               maxits=int(g["maxits"]), transform=0, want_std=True)
    assert rel_l2(mean, g["label_predict"]) <= 1e-6
    assert rel_l2(std, g["std_predict"]) <= 1e-6


def test_gp_predict_with_preconditioner(gpu):
    """The call sequence of TESTS/TEST3/foo3.cpp:262-316 (Nystrom from the dense additive kernel, random permutation): same mean
    as without the preconditioner."""
    g = np.load(os.path.join(GOLD, "predict_nfft300.npz"))
    L = gpu.lib()
    X1 = g["X1"]
    n = X1.shape[1]
    # the prediction routine writes the hyper-parameters into the two NFFT handles only (nfft_interface.c:913-920); the dense
    # handle of the preconditioner carries the values of the last loss evaluation (gp_loss.c:145-150) -- here: the same ones
    f, l, mu = np.log1p(np.exp(g["x"]))
    D = gpu.DenseAdditiveKernel(X1, g["windows"], 2, 3, "gaussian").set_hyper(float(f), float(l), float(mu))
    N = gpu.NysPrecond()
    perm = np.random.default_rng(807).permutation(n).astype(np.int32)
    L.Nfft4GPPrecondNysSetPerm(N.h, gpu.capi.iptr(perm), 0)
    L.Nfft4GPPrecondNysSetRank(N.h, 60)
    mean, _ = gpu.gp_predict(g["x"], X1, g["y"], g["X2"], g["windows"], 2, 3, "gaussian", precond=(L.Nfft4GPPrecondNysSetupWithKernel, N),
                             dense=D, tol=1e-10, maxits=200, transform=0, want_std=False)
    assert rel_l2(mean, g["label_predict"]) <= 1e-6
    N.free()
    D.free()
