"""Real-data fixtures for the matvec parity tests: 3000 sampled points of the reference's `bike` and `poletele` data sets
(TESTS/TEST1/data, the inputs of the reference's NFFT-vs-dense experiment, TESTS/TEST1/foo.cpp: `foo.ex ./data/<name>
./data/<name>.<g|m>.window 3000 <kernel> f l mu 906 1`) with their shipped window files -- three 3-D windows of permuted
columns for bike, two windows with a PADDED last window (-1) for poletele -- and the additive NFFT matvec of the reference's
own wrapper (oracle/_ref: nfft_interface.c unmodified on the NFFT3 restatement) for both kernels.  Real features are far from
uniform (discrete columns, duplicates, clusters): many points share a grid cell, many cells are empty.
Run here (needs /root/reference): make -C oracle ref && python scripts/make_realdata_fixture.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refapi as R  # noqa: E402

T1 = "/root/reference/TESTS/TEST1/data"
OUT = os.path.join(ROOT, "tests", "golden")


def read_feature(path):
    with open(path) as fh:
        n, d = [int(t) for t in fh.readline().split()]
        vals = np.array(fh.read().split(), dtype=np.float64)
    return np.ascontiguousarray(vals.reshape(d, n))


def read_window(path):
    with open(path) as fh:
        nw, dw = [int(t) for t in fh.readline().split()]
        vals = np.array(fh.read().split(), dtype=np.int32)
    return nw, dw, vals[:nw * dw].copy()


for name, hyper in (("bike", (1.0, 0.3, 0.01)), ("poletele", (1.2, 0.5, 0.02))):
    X = read_feature(os.path.join(T1, name + ".train.feature"))
    rng = np.random.default_rng(906)
    sel = np.sort(rng.choice(X.shape[1], 3000, replace=False))
    Xs = np.ascontiguousarray(X[:, sel])
    x = rng.random(3000) - 0.5
    out = {"X": Xs, "x": x, "f": hyper[0], "l": hyper[1], "mu": hyper[2]}
    for tag, kern in (("g", "gaussian"), ("m", "matern12")):
        nw, dw, win = read_window(os.path.join(T1, f"{name}.{tag}.window"))
        A = R.RefAdditiveNFFT(Xs, win, nw, dw, kern).build(*hyper)
        out[f"{tag}_windows"], out[f"{tag}_nw"], out[f"{tag}_dw"] = win, nw, dw
        out[f"{tag}_y"] = A.matvec(x)
        out[f"{tag}_y_ab"] = A.matvec(x, alpha=-0.7, beta=0.4, y=np.cos(np.arange(3000.0)))
        print(name, kern, "windows", win, "|y|", np.linalg.norm(out[f"{tag}_y"]))
    np.savez_compressed(os.path.join(OUT, f"real_{name}3000.npz"), **out)
    print("wrote", name, os.path.getsize(os.path.join(OUT, f"real_{name}3000.npz")) // 1024, "KiB")
