"""BASELINE config C5: gradient matvec (dK/df, dK/dl, dK/dmu applied to a vector, nfft_interface.c:819-840) at n = 4 10^6, d = 9,
for 32 Rademacher probe vectors (vecops.c:27-45) -- the inner loop of the stochastic trace estimators of the GP loss gradient.
Device-resident timing; python scripts/c5_grad.py [n] [probes]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as pkg  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4000000
probes = int(sys.argv[2]) if len(sys.argv) > 2 else 32
L = pkg.lib()
assert L.Nfft4GPB200Init(0) == 0
rng = np.random.default_rng(906)
X = rng.random((9, n)) * n ** (1.0 / 9)
t0 = time.perf_counter()
A = pkg.AdditiveNFFTKernel(X, np.arange(9, dtype=np.int32), 3, 3, "gaussian").build(1.0, 0.25 * n ** (1.0 / 9), 0.01)
L.Nfft4GPB200DeviceSynchronize()
print(f"plan build {time.perf_counter() - t0:.2f} s")
Z = np.where(rng.random((probes, n)) < 0.5, -1.0, 1.0)
dz = L.Nfft4GPB200DeviceAlloc(8 * n * probes)
dy = L.Nfft4GPB200DeviceAlloc(8 * 3 * n)
L.Nfft4GPB200MemcpyH2D(dz, Z.ctypes.data, 8 * n * probes)
e0, e1 = L.Nfft4GPB200EventCreate(), L.Nfft4GPB200EventCreate()
gm = L.Nfft4GPB200AdditiveNFFTGradMatSymvDevice
for p in range(3):
    assert gm(A.K, n, 1.0, dz + 8 * n * p, 0.0, dy) == 0
L.Nfft4GPB200DeviceSynchronize()
L.Nfft4GPB200EventRecord(e0)
for p in range(probes):
    gm(A.K, n, 1.0, dz + 8 * n * p, 0.0, dy)
L.Nfft4GPB200EventRecord(e1)
L.Nfft4GPB200DeviceSynchronize()
ms = L.Nfft4GPB200EventElapsedMs(e0, e1)
# property check on the last probe: block 0 = (2/f) K z, block 2 = f^2 z
g = np.zeros(3 * n)
L.Nfft4GPB200MemcpyD2H(g.ctypes.data, dy, 8 * 3 * n)
dk = L.Nfft4GPB200DeviceAlloc(8 * n)
A.matvec_device(dz + 8 * n * (probes - 1), dk)
Kz = np.zeros(n)
L.Nfft4GPB200MemcpyD2H(Kz.ctypes.data, dk, 8 * n)
e0_ = np.linalg.norm(g[:n] - 2.0 * Kz) / np.linalg.norm(Kz)
e2_ = np.linalg.norm(g[2 * n:] - Z[-1]) / np.linalg.norm(Z[-1])
print(f"C5: n={n}, {probes} Rademacher probes: {ms:.2f} ms total, {ms / probes:.3f} ms per gradient matvec, {1000 * probes / ms:.1f} gradient matvecs/s; "
      f"block checks: dK/df {e0_:.1e}, dK/dmu {e2_:.1e}")
