import ctypes as C, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as pkg
L = pkg.lib(); assert L.Nfft4GPB200Init(0) == 0
n = 4000000
rng = np.random.default_rng(906)
X = rng.random((9, n)) * n ** (1.0 / 9)
Z = np.where(rng.random((2, n)) < 0.5, -1.0, 1.0)
A = pkg.AdditiveNFFTKernel(X, np.arange(9, dtype=np.int32), 3, 3, "gaussian").build(1.0, 0.25 * n ** (1.0 / 9), 0.01)
gfn = C.cast(L.Nfft4GPAdditiveNFFTGradMatSymv, C.c_void_p)
pkg.lanczos_quadrature_logdet(A.K, A.dK, n, A.matvec_fn(), gfn, 20, Z[:1], print_level=0)
names = ["prep", "spread", "fft_fwd", "fft_inv", "gather", "combine", "vec", "nys", "csr", "gemv", "other"]
t0 = time.perf_counter()
pkg.lanczos_quadrature_logdet(A.K, A.dK, n, A.matvec_fn(), gfn, 20, Z, print_level=0)
L.Nfft4GPB200DeviceSynchronize()
print("2 probes wall", time.perf_counter() - t0)
L.Nfft4GPB200ProfileStart()
t0 = time.perf_counter()
pkg.lanczos_quadrature_logdet(A.K, A.dK, n, A.matvec_fn(), gfn, 20, Z, print_level=0)
L.Nfft4GPB200DeviceSynchronize()
w = time.perf_counter() - t0
sums = (C.c_double * 11)(); cnts = (C.c_longlong * 11)()
L.Nfft4GPB200ProfileStop(sums, cnts, 11)
print("profiled 2 probes wall", w, "kernel ms total", sum(sums), {names[i]: (round(sums[i], 2), cnts[i]) for i in range(11) if cnts[i]})
