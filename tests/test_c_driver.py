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
