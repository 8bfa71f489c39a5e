"""BASELINE config C4 in small: additive Matern-1/2 kernel with many 1-D / 2-D windows; device-resident matvec timing and
achieved algorithmic HBM rate ((16 d_w + 16) n bytes per window).  python scripts/c4_matvec.py [n] [d]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as pkg  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 16
L = pkg.lib()
assert L.Nfft4GPB200Init(0) == 0
rng = np.random.default_rng(906)
X = rng.random((d, n)) * n ** (1.0 / d)
x = rng.random(n) - 0.5
for dw in (2, 1):
    P = d // dw
    win = np.arange(d, dtype=np.int32)
    A = pkg.AdditiveNFFTKernel(X, win, P, dw, "matern12").build(1.0, 0.5 * n ** (1.0 / d), 0.01)
    dx = L.Nfft4GPB200DeviceAlloc(8 * n)
    dy = L.Nfft4GPB200DeviceAlloc(8 * n)
    L.Nfft4GPB200MemcpyH2D(dx, x.ctypes.data, 8 * n)
    e0, e1 = L.Nfft4GPB200EventCreate(), L.Nfft4GPB200EventCreate()
    for _ in range(3):
        A.matvec_device(dx, dy)
    L.Nfft4GPB200DeviceSynchronize()
    reps = 10
    L.Nfft4GPB200EventRecord(e0)
    for _ in range(reps):
        A.matvec_device(dx, dy)
    L.Nfft4GPB200EventRecord(e1)
    ms = L.Nfft4GPB200EventElapsedMs(e0, e1) / reps
    L.Nfft4GPB200ProfileStart()
    for _ in range(reps):
        A.matvec_device(dx, dy)
    names = ["prep", "spread", "fft_fwd", "fft_inv", "gather", "combine", "vec", "nys", "csr", "gemv", "other"]
    sums = (C.c_double * 11)()
    cnts = (C.c_longlong * 11)()
    L.Nfft4GPB200ProfileStop(sums, cnts, 11)
    alg = (16 * dw + 16) * n * P
    table = (2 * dw * 80 + 16 + 16 + 8) * n * P   # tap rows read twice, a written + read, f written + read (combine), x
    cls = " ".join(f"{names[i]}={sums[i]/reps*1000:.0f}us" for i in range(11) if cnts[i])
    print(f"n={n} d={d}: {P} windows of dimension {dw}: {ms:.3f} ms/matvec, {1000/ms:.1f} matvecs/s; algorithmic {alg/1e6:.0f} MB -> "
          f"{alg/ms/1e6:.0f} GB/s; with the tap tables {table/1e6:.0f} MB -> {table/ms/1e6:.0f} GB/s | {cls}")
    A.free()
    L.Nfft4GPB200DeviceFree(dx)
    L.Nfft4GPB200DeviceFree(dy)
