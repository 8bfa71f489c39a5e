"""Development aid: ONE fused matvec at n (default 1e6) -- for kernels built with in-kernel instrumentation (printf)."""
import sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as pkg
from common import uniform_cloud, rand_vec
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
L = pkg.lib()
assert L.Nfft4GPB200Init(0) == 0
X = uniform_cloud(n, 9)
A = pkg.AdditiveNFFTKernel(X, np.arange(9, dtype=np.int32), 3, 3, "gaussian").build(1.0, 0.25 * n ** (1 / 9), 0.01)
x = rand_vec(n)
dx = L.Nfft4GPB200DeviceAlloc(8 * n)
dy = L.Nfft4GPB200DeviceAlloc(8 * n)
L.Nfft4GPB200MemcpyH2D(dx, x.ctypes.data, 8 * n)
for i in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    print("---- matvec", i, flush=True)
    A.matvec_device(dx, dy)
    L.Nfft4GPB200DeviceSynchronize()
