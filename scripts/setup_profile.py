"""Workload for the ncu captures of the SETUP-side kernels: AFN setup (farthest point sampling, landmark GEMMs on the
fp64 tensor cores, brute-force KNN pattern, Schur FSAI rows), a Nystrom setup, one AFN apply (CSR products, K12 GEMVs) and the sparse
triangular solves of the FSAI gradient side.  python scripts/setup_profile.py [n] [rank] [lfil]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as pkg  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
lfil = int(sys.argv[3]) if len(sys.argv) > 3 else 100
L = pkg.lib()
assert L.Nfft4GPB200Init(0) == 0
rng = np.random.default_rng(906)
d = 9
X = rng.random((d, n)) * n ** (1.0 / d)
win = np.arange(9, dtype=np.int32)
D = pkg.DenseAdditiveKernel(X, win, 3, 3, "gaussian").set_hyper(1.0, 0.25 * n ** (1.0 / d), 0.01)
t0 = time.perf_counter()
P = pkg.AfnPrecond().setup(D, k, lfil, numests=0, noran=1)
L.Nfft4GPB200DeviceSynchronize()
print(f"AFN setup n={n} rank={k} lfil={lfil}: {time.perf_counter() - t0:.2f} s")
r = rng.random(n) - 0.5
z = P.solve(r)
t0 = time.perf_counter()
N = pkg.NysPrecond().setup(D, rng.permutation(n).astype(np.int32), min(k, 500))
L.Nfft4GPB200DeviceSynchronize()
print(f"Nystrom setup rank {min(k, 500)}: {time.perf_counter() - t0:.2f} s")
m = min(n, 50000)
Xs = np.ascontiguousarray(X[:, :m])
Ds = pkg.DenseAdditiveKernel(Xs, win, 3, 3, "gaussian").set_hyper(1.0, 0.25 * n ** (1.0 / d), 0.01)
t0 = time.perf_counter()
F = pkg.FsaiPrecond().setup(Ds, 50, require_grad=1)
L.Nfft4GPB200DeviceSynchronize()
print(f"FSAI setup with gradient n={m} lfil=50: {time.perf_counter() - t0:.2f} s")
t0 = time.perf_counter()
y = F.dvp(r[:m])
print(f"FSAI Dvp (6 sparse triangular solves, 9 CSR products): {time.perf_counter() - t0:.3f} s; trace {F.trace()}")
for h in (P, N, F, D, Ds):
    h.free()
