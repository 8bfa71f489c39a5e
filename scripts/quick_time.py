"""Quick device timing of the additive matvec at the headline size (development aid, not the bench)."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as pkg
from common import uniform_cloud, gaussian_cloud, rand_vec

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
cloud = gaussian_cloud if (len(sys.argv) > 2 and sys.argv[2] == "gauss") else uniform_cloud
variants = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 3, 4, 2]
L = pkg.lib()
assert L.Nfft4GPB200Init(0) == 0
X = cloud(n, 9)
win = np.arange(9, dtype=np.int32)
t0 = time.time()
A = pkg.AdditiveNFFTKernel(X, win, 3, 3, "gaussian").build(1.0, 0.25 * n ** (1 / 9), 0.01)
print("setup s", time.time() - t0, "nseg", [A.window_info(w)[2] for w in range(3)])
x = rand_vec(n)
dx = L.Nfft4GPB200DeviceAlloc(8 * n)
dy = L.Nfft4GPB200DeviceAlloc(8 * n)
L.Nfft4GPB200MemcpyH2D(dx, x.ctypes.data, 8 * n)
e0, e1 = L.Nfft4GPB200EventCreate(), L.Nfft4GPB200EventCreate()
names = ["prep", "spread", "fft_fwd", "fft_inv", "gather", "combine", "vec", "nys", "csr", "gemv", "other"]
yref = None
for variant in variants:
    L.Nfft4GPB200SetVariant(b"spread3d", variant)
    L.Nfft4GPB200SetVariant(b"gather3d", variant)
    for _ in range(3):
        A.matvec_device(dx, dy)
    L.Nfft4GPB200DeviceSynchronize()
    reps = 20 if variant else 3
    L.Nfft4GPB200EventRecord(e0)
    for _ in range(reps):
        A.matvec_device(dx, dy)
    L.Nfft4GPB200EventRecord(e1)
    ms = L.Nfft4GPB200EventElapsedMs(e0, e1) / reps
    L.Nfft4GPB200ProfileStart()
    for _ in range(reps):
        A.matvec_device(dx, dy)
    sums = (C.c_double * 11)()
    cnts = (C.c_longlong * 11)()
    L.Nfft4GPB200ProfileStop(sums, cnts, 11)
    y = np.zeros(n)
    L.Nfft4GPB200MemcpyD2H(y.ctypes.data, dy, 8 * n)
    if yref is None:
        yref = y
    cls = " ".join(f"{names[i]}={sums[i]/reps*1000:.0f}us" for i in range(11) if cnts[i])
    print(f"variant {variant}: {ms:.3f} ms/matvec -> {1000.0/ms:.1f} matvecs/s | {cls} | rel diff vs first {np.linalg.norm(y-yref)/np.linalg.norm(yref):.1e}")
