"""BASELINE config C3: AFN-preconditioned CG (landmark rank k + FSAI of the Schur complement) on the NFFT additive
Gaussian-kernel matvec, to 1e-6 relative residual.  Prints one JSON line per size with setup and solve times (CUDA-synchronised
wall clock around the C-ABI calls), iteration counts with and without the preconditioner, and the true residual."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as pkg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--k", type=int, default=500)
    ap.add_argument("--lfil", type=int, default=100)
    ap.add_argument("--l-ratio", type=float, default=0.25)
    ap.add_argument("--mu", type=float, default=0.01)
    ap.add_argument("--numests", type=int, default=5)
    ap.add_argument("--noran", type=int, default=0)
    ap.add_argument("--maxits", type=int, default=1000)
    ap.add_argument("--plain", type=int, default=1)
    ap.add_argument("--precond", default="afn")
    args = ap.parse_args()
    L = pkg.lib()
    assert L.Nfft4GPB200Init(0) == 0
    n, d = args.n, 9
    rng = np.random.default_rng(906)
    X = rng.random((d, n)) * n ** (1.0 / d)
    rhs = rng.random(n) - 0.5
    win = np.arange(9, dtype=np.int32)
    f, l, mu = 1.0, args.l_ratio * n ** (1.0 / d), args.mu
    out = dict(n=n, k=args.k, lfil=args.lfil, l=l, mu=mu, precond=args.precond)
    t0 = time.perf_counter()
    A = pkg.AdditiveNFFTKernel(X, win, 3, 3, "gaussian").build(f, l, mu)
    L.Nfft4GPB200DeviceSynchronize()
    out["nfft_plan_s"] = time.perf_counter() - t0
    D = pkg.DenseAdditiveKernel(X, win, 3, 3, "gaussian").set_hyper(f, l, mu)
    t0 = time.perf_counter()
    if args.precond == "afn":
        P = pkg.AfnPrecond().setup(D, args.k, args.lfil, numests=args.numests, noran=args.noran)
        F = P.get_factors() if n <= 20000 else None
        nn, kk, ke, isn = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        L.Nfft4GPB200PrecondAFNGetFactors(P.h, C.byref(nn), C.byref(kk), C.byref(ke), C.byref(isn), None, None, None, None, None, None, None)
        out.update(k_used=kk.value, k_est=ke.value, is_nys=bool(isn.value))
    elif args.precond == "nys":
        perm = np.random.default_rng(807).permutation(n).astype(np.int32)
        P = pkg.NysPrecond().setup(D, perm, args.k)
    elif args.precond == "fsai":
        P = pkg.FsaiPrecond().setup(D, args.lfil)
    L.Nfft4GPB200DeviceSynchronize()
    out["setup_s"] = time.perf_counter() - t0
    # one apply, timed
    r = rng.random(n) - 0.5
    P.solve(r)
    t0 = time.perf_counter()
    P.solve(r)
    out["apply_host_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    x1, rel1, h1, it1 = pkg.solver_pcg(A.K, n, A.matvec_fn(), P.h, P.solve_fn(), np.zeros(n), rhs, args.maxits, 0, 1e-6)
    out["solve_s"] = time.perf_counter() - t0
    out.update(iters=it1, rel_res=rel1)
    y = A.matvec(x1)
    out["true_rel_res"] = float(np.linalg.norm(y - rhs) / np.linalg.norm(rhs))
    if args.plain:
        t0 = time.perf_counter()
        x0, rel0, h0, it0 = pkg.solver_pcg(A.K, n, A.matvec_fn(), None, None, np.zeros(n), rhs, args.maxits, 0, 1e-6)
        out["plain_solve_s"] = time.perf_counter() - t0
        out.update(plain_iters=it0, plain_rel_res=rel0)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
