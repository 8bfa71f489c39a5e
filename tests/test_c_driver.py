"""The pure-C driver examples/test3_solve.c (the reference's host language, TESTS/TEST1 + TEST3 call sequences): it must
compile against include/nfft4gp_b200.h with a C compiler and link against the library alone; on a GPU it must solve."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "test3_solve")


def _build():
    import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as pkg
    pkg.build_library()   # the driver links against the in-tree library: make sure it exists (no-op when it does)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples")], stdout=subprocess.DEVNULL)
    assert os.path.exists(EXE)


def test_c_driver_builds_and_fails_loudly_without_gpu():
    _build()
    import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as pkg
    if pkg.lib().Nfft4GPB200DeviceCount() > 0:
        pytest.skip("a CUDA device is present: covered by the gpu test")
    p = subprocess.run([EXE, "1000"], capture_output=True, text=True, timeout=120)
    assert p.returncode == 2 and "no CPU fallback" in p.stdout


@pytest.mark.gpu
def test_c_driver_solves_on_gpu():
    _build()
    p = subprocess.run([EXE, "20000", "6", "3", "0", "1.0", "0.25", "0.01", "300", "100"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    m = re.search(r"RESULT (.*)", p.stdout)
    assert m, p.stdout[-2000:]
    r = dict(kv.split("=") for kv in m.group(1).split())
    its, its0 = int(r["iters_afn"]), int(r["iters_plain"])
    assert its > 0 and (its0 == 0 or its < its0), (its, its0)
    assert float(r["true_rel_afn"]) <= 2e-6
    assert int(r["launches"]) > 0


# ---- TEST1-style driver: reads the reference's text formats, NFFT vs dense errors (TESTS/TEST1/foo.cpp) ----
EXE1 = os.path.join(ROOT, "examples", "test1_matvec")


def _write_reference_format(tmp_path):
    """the 3000-point bike sample of tests/golden in the reference's file formats (TEST4/foo.cpp:9-120)"""
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", "real_bike3000.npz"))
    X, win, nw, dw = z["X"], z["g_windows"], int(z["g_nw"]), int(z["g_dw"])
    prefix = os.path.join(str(tmp_path), "bike")
    with open(prefix + ".train.feature", "w") as fh:
        fh.write(f"{X.shape[1]} {X.shape[0]}\n")
        for row in X:
            fh.write(" ".join(repr(float(v)) for v in row) + " \n")
    with open(prefix + ".g.window", "w") as fh:
        fh.write(f"{nw} {dw}\n")
        for w in range(nw):
            fh.write(" ".join(str(int(v)) for v in win[w * dw:(w + 1) * dw]) + "\n")
    return prefix, X.shape


def test_test1_driver_reads_reference_formats(tmp_path):
    _build()
    import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as pkg
    if pkg.lib().Nfft4GPB200DeviceCount() > 0:
        pytest.skip("a CUDA device is present: covered by the gpu test")
    prefix, shape = _write_reference_format(tmp_path)
    p = subprocess.run([EXE1, prefix, prefix + ".g.window", "2000", "0", "1.0", "0.3", "0.01", "906"], capture_output=True, text=True, timeout=120)
    assert p.returncode == 2
    assert f"n_sample = 2000 of {shape[1]}, d = {shape[0]}, kernel = 0" in p.stdout and "nwindows = 3, dwindows = 3" in p.stdout


@pytest.mark.gpu
def test_test1_driver_nfft_vs_dense_on_gpu(tmp_path):
    _build()
    prefix, shape = _write_reference_format(tmp_path)
    p = subprocess.run([EXE1, prefix, prefix + ".g.window", "-1", "0", "1.0", "0.3", "0.01", "906"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    m = re.search(r"RESULT (.*)", p.stdout)
    assert m, p.stdout[-2000:]
    r = {k: float(v) for k, v in (kv.split("=") for kv in m.group(1).split())}
    assert int(r["n"]) == shape[1]
    assert r["rel_l2"] < 1e-3                                     # the NFFT approximation error at this length scale
    assert abs(r["rel_l2_df"] - r["rel_l2"]) <= 1e-6 * max(r["rel_l2"], 1e-12) + 1e-15     # dK/df = (2/f) K on both sides
    assert r["rel_l2_dmu"] < 1e-14                                # f^2 x exactly on both sides
    assert r["rel_l2_dl"] < 1e-2
    assert "L2 Gradient Error 3." in p.stdout and "Linf Error." in p.stdout
