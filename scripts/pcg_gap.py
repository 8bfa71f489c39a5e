"""Development aid: per-iteration wall time of the device PCG with the AFN preconditioner (rank k, fp64 / fp32 K12), repeated,
to separate one-time costs from per-iteration ones."""
import sys
import time
import numpy as np
sys.path.insert(0, ".")
import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as pkg
n, k, its = int(float(sys.argv[1])), int(sys.argv[2]), int(sys.argv[3])
L = pkg.lib()
assert L.Nfft4GPB200Init(0) == 0
rng = np.random.default_rng(906)
X = rng.random((9, n)) * n ** (1.0 / 9)
rhs = rng.random(n) - 0.5
win = np.arange(9, dtype=np.int32)
f, l, mu = 1.0, 0.25 * n ** (1.0 / 9), 0.01
A = pkg.AdditiveNFFTKernel(X, win, 3, 3, "gaussian").build(f, l, mu)
D = pkg.DenseAdditiveKernel(X, win, 3, 3, "gaussian").set_hyper(f, l, mu)
P = pkg.AfnPrecond().setup(D, k, 100, numests=0, noran=1)
for bits in (64, 32):
    if bits == 32:
        P.set_k12_storage(32)
    for rep in range(3):
        L.Nfft4GPB200DeviceSynchronize()
        t0 = time.perf_counter()
        pkg.solver_pcg(A.K, n, A.matvec_fn(), P.h, P.solve_fn(), np.zeros(n), rhs, its, 0, 1e-30)
        L.Nfft4GPB200DeviceSynchronize()
        dt = time.perf_counter() - t0
        print(f"bits {bits} rep {rep}: {its} iterations {dt:.3f} s -> {1000*dt/its:.3f} ms/iteration", flush=True)
