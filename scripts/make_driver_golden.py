"""Expected outputs of the reference's OWN TESTS drivers, unmodified, on the reference's OWN CPU code (scripts/run_reference_driver_cpu.sh:
/root/reference/SRC compiled unmodified on the NFFT3 fastsum restatement): loss trace, final hyper-parameters, predictions and the
printed prediction error, stored as tests/golden/<test>_driver_reference_cpu.npz.  The same driver linked against libnfft4gp_b200.so
(scripts/run_reference_driver.sh) is run on the GPU by tests/test_reference_driver_gpu.py and compared at the printed precision.
  python scripts/make_driver_golden.py TEST4/foo.cpp test4 "poletele ./data/poletele ./data/poletele.g.window -1 -1 0 0 0 0 3 10 10 10 1"
"""
import os
import re
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_table(path):
    """TEST2 / TEST3 write a header line, TEST4 does not"""
    try:
        return np.loadtxt(path)
    except ValueError:
        return np.loadtxt(path, skiprows=1)


drv, tag, args = sys.argv[1], sys.argv[2], sys.argv[3]
subprocess.check_call([os.path.join(ROOT, "scripts", "run_reference_driver_cpu.sh"), drv])
work = os.path.join("/tmp/ref_cpu", os.path.dirname(drv))
exe = os.path.join(work, os.path.basename(drv)[:-4] + ".ex")
name = args.split()[0]
t0 = time.time()
p = subprocess.run([exe] + args.split(), capture_output=True, text=True, cwd=work, env=dict(os.environ, OMP_NUM_THREADS="1"))
assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
m = re.search(r"Final parameters: ([-0-9.eE]+), ([-0-9.eE]+), ([-0-9.eE]+)", p.stdout)
e = re.search(r"Prediction error: ([-0-9.eE]+)", p.stdout)
out = {"args": args, "final_params": np.array([float(m.group(i)) for i in (1, 2, 3)]), "cpu_seconds": time.time() - t0,
       "pred": load_table(os.path.join(work, f"{name}_0_nfft.txt")), "loss": np.atleast_1d(np.loadtxt(os.path.join(work, f"{name}_0_loss_nfft.txt")))}
if e:
    out["prediction_error"] = float(e.group(1))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"{tag}_driver_reference_cpu.npz"), **out)
print("wrote", tag, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
