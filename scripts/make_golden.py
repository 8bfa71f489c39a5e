"""Generate the golden fixtures of tests/golden/ by running the REFERENCE'S OWN C sources (oracle/_ref =
/root/reference/SRC compiled unmodified on the NFFT3 fastsum shim) in this container.  The fixtures travel to the GPU box;
/root/reference does not.  Run:  make -C oracle ref && python scripts/make_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import refapi as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
L = R.lib()


def cloud(n, d, seed):
    rng = np.random.default_rng(seed)
    return rng.random((d, n)) * n ** (1.0 / d)


def nfft_case(name, n, d, win, nw, dw, kern, f, l, mu, seed):
    X = cloud(n, d, seed)
    win = np.array(win, dtype=np.int32)
    rng = np.random.default_rng(seed + 1)
    x = rng.random(n) - 0.5
    y0 = rng.random(n) - 0.5
    g0 = rng.random(3 * n) - 0.5
    A = R.RefAdditiveNFFT(X, win, nw, dw, "gaussian" if kern == "gaussian" else "matern").build(f, l, mu)
    out = {"X": X, "windows": win, "nw": nw, "dw": dw, "kernel": kern, "f": f, "l": l, "mu": mu, "x": x, "y0": y0, "g0": g0}
    for tag, (al, be) in {"a1b0": (1.0, 0.0), "am1b1": (-1.0, 1.0), "a07b03": (0.7, 0.3)}.items():
        out["y_" + tag] = A.matvec(x, al, be, y0)
        out["g_" + tag] = A.grad_matvec(x, al, be, g0)
    # dense arbiter (unpadded windows only: the reference's dense skip-window detection is broken for padded ones)
    if -1 not in win.tolist():
        D = R.RefAdditiveDense(X, win, nw, dw, "gaussian" if kern == "gaussian" else "matern").build(f, l, mu)
        out["y_dense"] = D.matvec(x)
        out["g_dense"] = D.grad_matvec(x)
    np.savez_compressed(os.path.join(OUT, f"nfft_{name}.npz"), **out)
    A.free()
    print("wrote", name)


def solver_case():
    """Reference PCG / Nystrom / FSAI on the reference's dense additive kernel."""
    n, d = 300, 9
    X = cloud(n, d, 41)
    win = np.arange(9, dtype=np.int32)
    f, l, mu = 1.1, 0.5, 0.05
    D = R.RefAdditiveDense(X, win, 3, 3, "gaussian").build(f, l, mu, grad=False)
    rng = np.random.default_rng(42)
    rhs = rng.random(n) - 0.5
    mv = L.Nfft4GPDenseMatSymv
    out = {"X": X, "windows": win, "f": f, "l": l, "mu": mu, "rhs": rhs, "K": D.matrix()}
    x, rel, hist, it = R.ref_pcg(D.K, n, mv, None, None, np.zeros(n), rhs, 400, 0, 1e-8)
    out.update(pcg_x=x, pcg_rel=rel, pcg_hist=hist, pcg_it=it)
    x2, rel2, hist2, it2 = R.ref_pcg(D.K, n, mv, None, None, np.zeros(n), rhs, 7, 0, 1e-12)  # non-convergence: iter = 0
    out.update(pcg7_x=x2, pcg7_rel=rel2, pcg7_hist=hist2, pcg7_it=it2)
    # Nystrom with a fixed permutation, rank 40
    perm = np.random.default_rng(43).permutation(n).astype(np.int32)
    nys = L.Nfft4GPPrecondNysCreate()
    L.Nfft4GPPrecondNysSetRank(nys, 40)
    L.Nfft4GPPrecondNysSetPerm(nys, R.iptr(perm), 0)
    assert L.Nfft4GPPrecondNysSetupWithKernel(R.dptr(D.X), n, n, d, C.cast(L.Nfft4GPKernelAdditiveKernel, C.c_void_p), D.h, 0, nys) == 0
    z = np.zeros(n)
    L.Nfft4GPPrecondNysSolve(nys, n, R.dptr(z), R.dptr(rhs))
    xn, reln, histn, itn = R.ref_pcg(D.K, n, mv, nys, L.Nfft4GPPrecondNysSolve, np.zeros(n), rhs, 400, 0, 1e-8)
    out.update(nys_perm=perm, nys_k=40, nys_z=z, nys_pcg_it=itn, nys_pcg_rel=reln, nys_pcg_hist=histn)

    # read the factors back from the reference struct (INC/_preconds.h:320-351)
    class PNys(C.Structure):
        _fields_ = [("_k_setup", C.c_int), ("_own_perm", C.c_int), ("_perm", C.c_void_p), ("_n", C.c_int), ("_tits", C.c_int),
                    ("_titt", C.c_double), ("_tset", C.c_double), ("_tlogdet", C.c_double), ("_tdvp", C.c_double),
                    ("_nys_opt", C.c_int), ("_k", C.c_int), ("_eta", C.c_double), ("_f2", C.c_double),
                    ("_U", R.c_double_p), ("_s", R.c_double_p), ("_work", R.c_double_p)]

    ps = PNys.from_address(nys)
    out["nys_U"] = np.ctypeslib.as_array(ps._U, shape=(40, n)).T.copy()  # column-major n x k
    out["nys_s"] = np.ctypeslib.as_array(ps._s, shape=(40,)).copy()
    out["nys_eta"] = ps._eta
    np.savez_compressed(os.path.join(OUT, "solver_dense300.npz"), **out)
    print("wrote solver case: pcg it", it, "nys-pcg it", itn)


def fsai_case():
    """Reference FSAI on the reference's plain dense Gaussian kernel (3 features).  NB: Nfft4GPPrecondFsaiSetupWithKernel
    segfaults in the reference when it is handed the dense ADDITIVE kernel (observed here, gcc 13.3, also single-threaded),
    so the golden vector for the FSAI row uses the kernel the reference's own drivers combine it with."""
    n, d = 300, 3
    X = cloud(n, d, 51)
    f, l, mu = 1.0, 0.8, 0.05
    L.Nfft4GPKernelParamCreate.restype = C.c_void_p
    L.Nfft4GPKernelParamCreate.argtypes = [C.c_int, C.c_int]
    kh = L.Nfft4GPKernelParamCreate(n, 1)
    R.set_hyper(kh, f, l, mu)
    K = C.c_void_p()
    assert L.Nfft4GPKernelGaussianKernel(kh, R.dptr(X), n, n, d, None, 0, None, 0, C.byref(K), None) == 0
    Km = np.ctypeslib.as_array(C.cast(K, R.c_double_p), shape=(n, n)).T.copy()
    Km = np.tril(Km) + np.tril(Km, -1).T
    rng = np.random.default_rng(52)
    rhs = rng.random(n) - 0.5
    out = {"X": X, "f": f, "l": l, "mu": mu, "rhs": rhs, "K": Km}
    fs = L.Nfft4GPPrecondFsaiCreate()
    L.Nfft4GPPrecondFsaiSetLfil(fs, 12)
    assert L.Nfft4GPPrecondFsaiSetupWithKernel(R.dptr(X), n, n, d, C.cast(L.Nfft4GPKernelGaussianKernel, C.c_void_p), kh, 0, fs) == 0
    zf = np.zeros(n)
    L.Nfft4GPPrecondFsaiSolve(fs, n, R.dptr(zf), R.dptr(rhs))

    class PFsai(C.Structure):
        _fields_ = [("_lfil", C.c_int), ("_n", C.c_int), ("_tits", C.c_int), ("_titt", C.c_double), ("_tset", C.c_double),
                    ("_tlogdet", C.c_double), ("_tdvp", C.c_double), ("_L_i", R.c_int_p), ("_L_j", R.c_int_p),
                    ("_L_a", R.c_double_p), ("_dL_a", R.c_double_p), ("_work", R.c_double_p)]

    pf = PFsai.from_address(fs)
    ia = np.ctypeslib.as_array(pf._L_i, shape=(n + 1,)).copy()
    nnz = int(ia[-1])
    out.update(fsai_lfil=12, fsai_z=zf, fsai_ia=ia, fsai_ja=np.ctypeslib.as_array(pf._L_j, shape=(nnz,)).copy(),
               fsai_a=np.ctypeslib.as_array(pf._L_a, shape=(nnz,)).copy())
    xf, relf, histf, itf = R.ref_pcg(K, n, L.Nfft4GPDenseMatSymv, fs, L.Nfft4GPPrecondFsaiSolve, np.zeros(n), rhs, 400, 0, 1e-8)
    x0, rel0, hist0, it0 = R.ref_pcg(K, n, L.Nfft4GPDenseMatSymv, None, None, np.zeros(n), rhs, 400, 0, 1e-8)
    out.update(fsai_pcg_it=itf, fsai_pcg_rel=relf, fsai_pcg_x=xf, pcg_it=it0)
    np.savez_compressed(os.path.join(OUT, "fsai_gauss300.npz"), **out)
    print("wrote fsai case: nnz", nnz, "fsai-pcg it", itf, "plain pcg it", it0)


def nys_plain_case():
    """Reference Nystrom setup + apply on the reference's PLAIN Gaussian kernel (9 features).  With the dense ADDITIVE kernel
    the reference's setup reads the wrong entries of the handle's feature copy (kernels.c:3160 indexes it with the caller's
    n = k while nys.c:549 passes sub-sampled data), so the setup row is pinned on the kernel where the reference is right."""
    n, d, k = 300, 9, 40
    X = cloud(n, d, 61)
    f, l, mu = 1.1, 1.4, 0.05
    L.Nfft4GPKernelParamCreate.restype = C.c_void_p
    L.Nfft4GPKernelParamCreate.argtypes = [C.c_int, C.c_int]
    kh = L.Nfft4GPKernelParamCreate(n, 1)
    R.set_hyper(kh, f, l, mu)
    K = C.c_void_p()
    assert L.Nfft4GPKernelGaussianKernel(kh, R.dptr(X), n, n, d, None, 0, None, 0, C.byref(K), None) == 0
    Km = np.ctypeslib.as_array(C.cast(K, R.c_double_p), shape=(n, n)).T.copy()
    Km = np.tril(Km) + np.tril(Km, -1).T
    rhs = np.random.default_rng(62).random(n) - 0.5
    perm = np.random.default_rng(63).permutation(n).astype(np.int32)
    nys = L.Nfft4GPPrecondNysCreate()
    L.Nfft4GPPrecondNysSetRank(nys, k)
    L.Nfft4GPPrecondNysSetPerm(nys, R.iptr(perm), 0)
    assert L.Nfft4GPPrecondNysSetupWithKernel(R.dptr(X), n, n, d, C.cast(L.Nfft4GPKernelGaussianKernel, C.c_void_p), kh, 0, nys) == 0
    z = np.zeros(n)
    L.Nfft4GPPrecondNysSolve(nys, n, R.dptr(z), R.dptr(rhs))
    xn, reln, histn, itn = R.ref_pcg(K, n, L.Nfft4GPDenseMatSymv, nys, L.Nfft4GPPrecondNysSolve, np.zeros(n), rhs, 400, 0, 1e-8)
    x0, rel0, hist0, it0 = R.ref_pcg(K, n, L.Nfft4GPDenseMatSymv, None, None, np.zeros(n), rhs, 400, 0, 1e-8)

    class PNys(C.Structure):
        _fields_ = [("_k_setup", C.c_int), ("_own_perm", C.c_int), ("_perm", C.c_void_p), ("_n", C.c_int), ("_tits", C.c_int),
                    ("_titt", C.c_double), ("_tset", C.c_double), ("_tlogdet", C.c_double), ("_tdvp", C.c_double),
                    ("_nys_opt", C.c_int), ("_k", C.c_int), ("_eta", C.c_double), ("_f2", C.c_double),
                    ("_U", R.c_double_p), ("_s", R.c_double_p), ("_work", R.c_double_p)]

    ps = PNys.from_address(nys)
    out = {"X": X, "f": f, "l": l, "mu": mu, "rhs": rhs, "K": Km, "perm": perm, "k": k, "z": z,
           "U": np.ctypeslib.as_array(ps._U, shape=(k, n)).T.copy(), "s": np.ctypeslib.as_array(ps._s, shape=(k,)).copy(),
           "eta": ps._eta, "nys_pcg_it": itn, "nys_pcg_x": xn, "pcg_it": it0}
    np.savez_compressed(os.path.join(OUT, "nys_gauss300.npz"), **out)
    print("wrote nys plain case: nys-pcg it", itn, "plain pcg it", it0)


def chol_case():
    """Reference dense Cholesky preconditioner (SRC/preconds/chol.c) with gradients on the reference's plain Gaussian kernel:
    kernel value + gradient blocks, solve, log det, traces tr(K^-1 dK_i) and derivative-vector products K^-1 dK_i x."""
    z = np.load(os.path.join(OUT, "nys_gauss300.npz"))
    X, f, l, mu, rhs = z["X"], float(z["f"]), float(z["l"]), float(z["mu"]), z["rhs"]
    n, d = X.shape[1], X.shape[0]
    L.Nfft4GPKernelParamCreate.restype = C.c_void_p
    L.Nfft4GPKernelParamCreate.argtypes = [C.c_int, C.c_int]
    kh = L.Nfft4GPKernelParamCreate(n, 1)
    R.set_hyper(kh, f, l, mu)
    fk = C.cast(L.Nfft4GPKernelGaussianKernel, C.c_void_p)
    out = {}
    # gradient blocks of the three call shapes
    rows = np.random.default_rng(71).permutation(n)[:25].astype(np.int32)
    cols = np.random.default_rng(72).permutation(n)[:17].astype(np.int32)
    K, dK = C.c_void_p(), C.c_void_p()
    assert L.Nfft4GPKernelGaussianKernel(kh, R.dptr(X), n, n, d, R.iptr(rows), 25, R.iptr(cols), 17, C.byref(K), C.byref(dK)) == 0
    out["cross_K"] = np.ctypeslib.as_array(C.cast(K, R.c_double_p), shape=(17, 25)).T.copy()
    out["cross_dK"] = np.ctypeslib.as_array(C.cast(dK, R.c_double_p), shape=(3, 17, 25)).transpose(0, 2, 1).copy()
    K, dK = C.c_void_p(), C.c_void_p()
    assert L.Nfft4GPKernelGaussianKernel(kh, R.dptr(X), n, n, d, R.iptr(rows), 25, None, 0, C.byref(K), C.byref(dK)) == 0
    out["sym_K"] = np.ctypeslib.as_array(C.cast(K, R.c_double_p), shape=(25, 25)).T.copy()
    out["sym_dK"] = np.ctypeslib.as_array(C.cast(dK, R.c_double_p), shape=(3, 25, 25)).transpose(0, 2, 1).copy()
    out["rows"], out["cols"] = rows, cols
    for stable in (0, 1):
        ch = L.Nfft4GPPrecondCholCreate()
        L.Nfft4GPPrecondCholSetStable(ch, stable)
        assert L.Nfft4GPPrecondCholSetupWithKernel(R.dptr(X), n, n, d, fk, kh, 1, ch) == 0
        x = np.zeros(n)
        L.Nfft4GPPrecondCholSolve(ch, n, R.dptr(x), R.dptr(rhs))
        L.Nfft4GPPrecondCholLogdet.restype = C.c_double
        L.Nfft4GPPrecondCholLogdet.argtypes = [C.c_void_p]
        ld = L.Nfft4GPPrecondCholLogdet(ch)
        tr = np.zeros(3)
        trp = R.dptr(tr)
        L.Nfft4GPPrecondCholTrace.argtypes = [C.c_void_p, C.POINTER(R.c_double_p)]
        assert L.Nfft4GPPrecondCholTrace(ch, C.byref(trp)) == 0
        y = np.zeros(3 * n)
        yp = R.dptr(y)
        L.Nfft4GPPrecondCholDvp.argtypes = [C.c_void_p, C.c_int, C.c_void_p, R.c_double_p, C.POINTER(R.c_double_p)]
        assert L.Nfft4GPPrecondCholDvp(ch, n, None, R.dptr(rhs), C.byref(yp)) == 0
        out[f"solve{stable}"], out[f"logdet{stable}"], out[f"trace{stable}"], out[f"dvp{stable}"] = x, ld, tr, y
        print("chol stable", stable, "logdet", ld, "traces", tr)
    out.update({"X": X, "f": f, "l": l, "mu": mu, "rhs": rhs})
    np.savez_compressed(os.path.join(OUT, "chol_gauss300.npz"), **out)


def nys_grad_case():
    """Reference Nystrom preconditioner set up with require_grad (nys.c:518-660) on the reference's plain Gaussian kernel:
    log det, traces tr(M^-1 dM_i) and derivative-vector products M^-1 dM_i x (nys.c:175-516)."""
    z = np.load(os.path.join(OUT, "nys_gauss300.npz"))
    X, f, l, mu, rhs, perm, k = z["X"], float(z["f"]), float(z["l"]), float(z["mu"]), z["rhs"], z["perm"], int(z["k"])
    n, d = X.shape[1], X.shape[0]
    L.Nfft4GPKernelParamCreate.restype = C.c_void_p
    L.Nfft4GPKernelParamCreate.argtypes = [C.c_int, C.c_int]
    kh = L.Nfft4GPKernelParamCreate(n, 1)
    R.set_hyper(kh, f, l, mu)
    nys = L.Nfft4GPPrecondNysCreate()
    L.Nfft4GPPrecondNysSetRank(nys, k)
    L.Nfft4GPPrecondNysSetPerm(nys, R.iptr(perm), 0)
    assert L.Nfft4GPPrecondNysSetupWithKernel(R.dptr(X), n, n, d, C.cast(L.Nfft4GPKernelGaussianKernel, C.c_void_p), kh, 1, nys) == 0
    L.Nfft4GPPrecondNysLogdet.restype = C.c_double
    L.Nfft4GPPrecondNysLogdet.argtypes = [C.c_void_p]
    ld = L.Nfft4GPPrecondNysLogdet(nys)
    tr = np.zeros(3)
    trp = R.dptr(tr)
    L.Nfft4GPPrecondNysTrace.argtypes = [C.c_void_p, C.POINTER(R.c_double_p)]
    assert L.Nfft4GPPrecondNysTrace(nys, C.byref(trp)) == 0
    y = np.zeros(3 * n)
    yp = R.dptr(y)
    L.Nfft4GPPrecondNysDvp.argtypes = [C.c_void_p, C.c_int, C.c_void_p, R.c_double_p, C.POINTER(R.c_double_p)]
    assert L.Nfft4GPPrecondNysDvp(nys, n, None, R.dptr(rhs), C.byref(yp)) == 0
    zs = np.zeros(n)
    L.Nfft4GPPrecondNysSolve(nys, n, R.dptr(zs), R.dptr(rhs))
    print("nys grad: logdet", ld, "traces", tr)
    np.savez_compressed(os.path.join(OUT, "nys_grad_gauss300.npz"), logdet=ld, trace=tr, dvp=y, solve=zs)


def fgmres_case():
    """Reference Nfft4GPSolverFgmres (SRC/solvers/fgmres.c) on the reference's dense Gaussian kernel, with restarts (kdim 15),
    without and with the reference's Nystrom preconditioner, plus a run that stops at the iteration limit."""
    z = np.load(os.path.join(OUT, "nys_gauss300.npz"))
    X, f, l, mu, rhs, perm, k = z["X"], float(z["f"]), float(z["l"]), float(z["mu"]), z["rhs"], z["perm"], int(z["k"])
    n, d = X.shape[1], X.shape[0]
    L.Nfft4GPKernelParamCreate.restype = C.c_void_p
    L.Nfft4GPKernelParamCreate.argtypes = [C.c_int, C.c_int]
    kh = L.Nfft4GPKernelParamCreate(n, 1)
    R.set_hyper(kh, f, l, mu)
    K = C.c_void_p()
    assert L.Nfft4GPKernelGaussianKernel(kh, R.dptr(X), n, n, d, None, 0, None, 0, C.byref(K), None) == 0
    nys = L.Nfft4GPPrecondNysCreate()
    L.Nfft4GPPrecondNysSetRank(nys, k)
    L.Nfft4GPPrecondNysSetPerm(nys, R.iptr(perm), 0)
    assert L.Nfft4GPPrecondNysSetupWithKernel(R.dptr(X), n, n, d, C.cast(L.Nfft4GPKernelGaussianKernel, C.c_void_p), kh, 0, nys) == 0
    L.Nfft4GPSolverFgmres.restype = C.c_int
    L.Nfft4GPSolverFgmres.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, R.c_double_p, R.c_double_p, C.c_int,
                                      C.c_int, C.c_int, C.c_double, R.c_double_p, C.POINTER(R.c_double_p), R.c_int_p, C.c_int]

    def run(prec, pf, kdim, maxits, tol):
        x = np.zeros(n)
        rel = C.c_double(0.0)
        relv = R.c_double_p()
        it = C.c_int(0)
        L.Nfft4GPSolverFgmres(K, n, C.cast(L.Nfft4GPDenseMatSymv, C.c_void_p), prec, C.cast(pf, C.c_void_p) if pf else None, R.dptr(x),
                              R.dptr(rhs), kdim, maxits, 0, tol, C.byref(rel), C.byref(relv), C.byref(it), 1)
        hist = np.ctypeslib.as_array(relv, shape=(maxits + 1,)).copy()
        return x, rel.value, hist, it.value

    out = {}
    for tag, (prec, pf, kdim, maxits, tol) in {"plain": (None, None, 15, 300, 1e-8), "nys": (nys, L.Nfft4GPPrecondNysSolve, 15, 300, 1e-8),
                                                "cap": (None, None, 10, 25, 1e-12)}.items():
        x, rel, hist, it = run(prec, pf, kdim, maxits, tol)
        out.update({f"{tag}_x": x, f"{tag}_rel": rel, f"{tag}_hist": hist, f"{tag}_it": it, f"{tag}_kdim": kdim, f"{tag}_maxits": maxits,
                    f"{tag}_tol": tol})
        print("fgmres", tag, "iterations", it, "rel", rel)
    np.savez_compressed(os.path.join(OUT, "fgmres_dense300.npz"), **out)


def predict_case():
    """Reference Nfft4GPAdditiveNFFTGpPredict (SRC/external/nfft_interface.c:873-1068, compiled unmodified on the fastsum shim):
    mean and standard deviation at the test points, softplus-transformed hyper-parameters, no preconditioner."""
    n, n2, d = 300, 12, 6
    X1 = cloud(n, d, 71)
    X2 = cloud(n2, d, 72) * (n2 / n) ** (-1.0 / d)   # same box as the training points
    rng = np.random.default_rng(73)
    y = np.sin(X1[0]) + 0.5 * np.cos(X1[3]) + 0.05 * rng.standard_normal(n)
    win = np.arange(6, dtype=np.int32)
    Xall = np.ascontiguousarray(np.concatenate([X1, X2], axis=1))
    A = R.RefAdditiveNFFT(X1, win, 2, 3, "gaussian")
    Al = R.RefAdditiveNFFT(Xall, win, 2, 3, "gaussian")
    x = np.array([0.3, 0.9, -3.0])                   # softplus -> f = 0.854, l = 1.24, mu = 0.0486
    L.Nfft4GPAdditiveNFFTGpPredict.restype = C.c_int
    vp, i_, dd, dp = C.c_void_p, C.c_int, C.c_double, R.c_double_p
    L.Nfft4GPAdditiveNFFTGpPredict.argtypes = [dp, dp, dp, i_, i_, i_, dp, i_, i_, dp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i_, dd, i_, i_, i_,
                                               dp, C.POINTER(dp), C.POINTER(dp)]
    lab = dp()
    std = dp()
    rc = L.Nfft4GPAdditiveNFFTGpPredict(R.dptr(x), R.dptr(X1), R.dptr(y), n, n, d, R.dptr(X2), n2, n2, R.dptr(Xall),
                                        C.cast(L.Nfft4GPNFFTAdditiveKernelGaussianKernel, vp), A.h, Al.h, C.cast(L.Nfft4GPNFFTKernelFree, vp),
                                        C.cast(L.Nfft4GPAdditiveNFFTMatSymv, vp), None, None, None, None, None, None, 0, 1e-10, 200, 0, 1,
                                        None, C.byref(lab), C.byref(std))
    assert rc == 0
    out = {"X1": X1, "X2": X2, "y": y, "windows": win, "x": x, "tol": 1e-10, "maxits": 200,
           "label_predict": np.ctypeslib.as_array(lab, shape=(n2,)).copy(), "std_predict": np.ctypeslib.as_array(std, shape=(n2,)).copy()}
    np.savez_compressed(os.path.join(OUT, "predict_nfft300.npz"), **out)
    print("predict: mean", out["label_predict"][:4], "std", out["std_predict"][:4])


def lanczos_case():
    """Reference Nfft4GPSolverLanczos (SRC/solvers/lanczos.c:3-419) on the reference's dense Gaussian kernel, without and with the
    reference's Nystrom preconditioner, and Nfft4GPLanczosQuadratureLogdet (:421-610) on the NFFT additive kernel with given
    Rademacher probes (no preconditioner)."""
    z = np.load(os.path.join(OUT, "nys_gauss300.npz"))
    X, f, l, mu, rhs, perm, k = z["X"], float(z["f"]), float(z["l"]), float(z["mu"]), z["rhs"], z["perm"], int(z["k"])
    n, d = X.shape[1], X.shape[0]
    L.Nfft4GPKernelParamCreate.restype = C.c_void_p
    L.Nfft4GPKernelParamCreate.argtypes = [C.c_int, C.c_int]
    kh = L.Nfft4GPKernelParamCreate(n, 1)
    R.set_hyper(kh, f, l, mu)
    K = C.c_void_p()
    assert L.Nfft4GPKernelGaussianKernel(kh, R.dptr(X), n, n, d, None, 0, None, 0, C.byref(K), None) == 0
    nys = L.Nfft4GPPrecondNysCreate()
    L.Nfft4GPPrecondNysSetRank(nys, k)
    L.Nfft4GPPrecondNysSetPerm(nys, R.iptr(perm), 0)
    assert L.Nfft4GPPrecondNysSetupWithKernel(R.dptr(X), n, n, d, C.cast(L.Nfft4GPKernelGaussianKernel, C.c_void_p), kh, 0, nys) == 0
    dp = R.c_double_p
    L.Nfft4GPSolverLanczos.restype = C.c_int
    L.Nfft4GPSolverLanczos.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, dp, dp, C.c_int, C.c_int, C.c_int, C.c_double, dp,
                                       C.POINTER(dp), R.c_int_p, R.c_int_p, C.POINTER(dp), C.POINTER(dp), C.c_int]
    out = {}
    for tag, (prec, pf, maxits, tol) in {"plain": (None, None, 120, 1e-8), "nys": (nys, L.Nfft4GPPrecondNysSolve, 80, 1e-8)}.items():
        x = np.zeros(n)
        rel = C.c_double(0.0)
        relv, TD, TE = dp(), dp(), dp()
        it, ts = C.c_int(0), C.c_int(0)
        L.Nfft4GPSolverLanczos(K, n, C.cast(L.Nfft4GPDenseMatSymv, C.c_void_p), prec, C.cast(pf, C.c_void_p) if pf else None, R.dptr(x), R.dptr(rhs),
                               maxits, maxits, 0, tol, C.byref(rel), C.byref(relv), C.byref(it), C.byref(ts), C.byref(TD), C.byref(TE), 1)
        out.update({f"{tag}_x": x, f"{tag}_rel": rel.value, f"{tag}_it": it.value, f"{tag}_tsize": ts.value, f"{tag}_maxits": maxits, f"{tag}_tol": tol,
                    f"{tag}_hist": np.ctypeslib.as_array(relv, shape=(maxits + 1,)).copy(),
                    f"{tag}_TD": np.ctypeslib.as_array(TD, shape=(ts.value,)).copy(),
                    f"{tag}_TE": np.ctypeslib.as_array(TE, shape=(ts.value - 1,)).copy()})
        print("lanczos", tag, "iterations", it.value, "tsize", ts.value, "rel", rel.value)
    # stochastic Lanczos quadrature on the NFFT additive kernel
    n2, d2 = 300, 6
    X2 = cloud(n2, d2, 81)
    win = np.arange(6, dtype=np.int32)
    f2, l2, mu2 = 1.1, 0.9, 0.05
    A = R.RefAdditiveNFFT(X2, win, 2, 3, "gaussian").build(f2, l2, mu2)
    probes = np.where(np.random.default_rng(82).random((4, n2)) < 0.5, -1.0, 1.0)
    L.Nfft4GPLanczosQuadratureLogdet.restype = C.c_int
    L.Nfft4GPLanczosQuadratureLogdet.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.c_void_p, C.c_int, C.c_int, dp, C.c_int, dp, C.POINTER(dp)]
    val = C.c_double(0.0)
    dval = dp()
    rc = L.Nfft4GPLanczosQuadratureLogdet(A.K, A.dK, n2, C.cast(L.Nfft4GPAdditiveNFFTMatSymv, C.c_void_p), C.cast(L.Nfft4GPAdditiveNFFTGradMatSymv, C.c_void_p),
                                          None, None, None, None, None, 60, 4, R.dptr(np.ascontiguousarray(probes)), 1, C.byref(val), C.byref(dval))
    assert rc == 0
    out.update(slq_X=X2, slq_windows=win, slq_f=f2, slq_l=l2, slq_mu=mu2, slq_probes=probes, slq_maxits=60, slq_logdet=val.value,
               slq_dlogdet=np.ctypeslib.as_array(dval, shape=(3,)).copy())
    print("slq logdet/n", val.value, "dlogdet/n", out["slq_dlogdet"])
    np.savez_compressed(os.path.join(OUT, "lanczos_dense300.npz"), **out)


def afn_case():
    """The AFN setup assembled from the reference's COMPILED parts in the order of SRC/preconds/afn.c:161-420 (afn.c itself is
    not built, SURVEY.md finding 2): Nfft4GPSortFps (ordering.c:713, parallel algorithm :422) -> Nfft4GPExpandPerm (utils.c:208) ->
    Nfft4GPPrecondCholSetupWithKernel on the landmark rows (chol.c:428, stable shift :439-466) ->
    Nfft4GPKernelSchurCombineKernelParamCreate / Nfft4GPKernelSchurCombineKernel (kernels.c:3496-4009) ->
    Nfft4GPPrecondFsaiSetupWithKernel on the non-landmark points (fsai.c:302), and the 5-step apply of afn.c:109-136 composed
    from Nfft4GPPrecondCholSolve, the reference kernel's K12 block and Nfft4GPPrecondFsaiSolve.  Plain Gaussian kernel (the
    reference's FSAI setup segfaults on its dense ADDITIVE kernel, see fsai_case).  Also: Nfft4GPRankestNysScaled
    (rankest.c:248-391) with the sample rows replayed through srand()."""
    libc = C.CDLL(None)
    n, d, k, lfil = 400, 3, 60, 20
    X = cloud(n, d, 91)
    f, l, mu = 1.1, 1.0, 0.02
    L.Nfft4GPKernelParamCreate.restype = C.c_void_p
    L.Nfft4GPKernelParamCreate.argtypes = [C.c_int, C.c_int]
    kh = L.Nfft4GPKernelParamCreate(n, 1)
    R.set_hyper(kh, f, l, mu)
    fk = C.cast(L.Nfft4GPKernelGaussianKernel, C.c_void_p)

    class OrdFps(C.Structure):   # INC/_linearalg.h:820-841
        _fields_ = [("_algorithm", C.c_int), ("_tol", C.c_double), ("_rho", C.c_double), ("_fdist", C.c_void_p), ("_fdist_params", C.c_void_p),
                    ("_dist", C.c_void_p), ("_build_pattern", C.c_int), ("_pattern_lfil", C.c_int), ("_pattern_opt", C.c_int),
                    ("_S_i", C.c_void_p), ("_S_j", C.c_void_p)]

    L.Nfft4GPOrdFpsCreate.restype = C.c_void_p
    L.Nfft4GPSortFps.argtypes = [C.c_void_p, R.c_double_p, C.c_int, C.c_int, C.c_int, R.c_int_p, C.POINTER(R.c_int_p)]
    fps = L.Nfft4GPOrdFpsCreate()
    o = OrdFps.from_address(fps)
    o._algorithm, o._tol, o._build_pattern = 0, 0.0, 0      # afn.c:201-205
    kk, pp = C.c_int(k), R.c_int_p()
    assert L.Nfft4GPSortFps(fps, R.dptr(X), n, n, d, C.byref(kk), C.byref(pp)) == 0 and kk.value == k
    sel = np.ctypeslib.as_array(pp, shape=(k,)).copy()
    L.Nfft4GPExpandPerm.restype = R.c_int_p
    L.Nfft4GPExpandPerm.argtypes = [R.c_int_p, C.c_int, C.c_int]
    L.Nfft4GPSubData.restype = R.c_double_p
    L.Nfft4GPSubData.argtypes = [R.c_double_p, C.c_int, C.c_int, C.c_int, R.c_int_p, C.c_int]
    permp = L.Nfft4GPExpandPerm(pp, k, n)
    perm = np.ctypeslib.as_array(permp, shape=(n,)).copy()
    perm2p = C.cast(C.addressof(permp.contents) + 4 * k, R.c_int_p)
    n2 = n - k
    data1 = L.Nfft4GPSubData(R.dptr(X), n, n, d, permp, k)
    ch = L.Nfft4GPPrecondCholCreate()
    L.Nfft4GPPrecondCholSetStable.argtypes = [C.c_void_p, C.c_int]
    L.Nfft4GPPrecondCholSetStable(ch, 1)
    assert L.Nfft4GPPrecondCholSetupWithKernel(data1, k, k, d, fk, kh, 0, ch) == 0

    class PChol(C.Structure):    # INC/_preconds.h:48-71
        _fields_ = [("_stable", C.c_int), ("_chol_data", R.c_double_p), ("_GdKG_data", R.c_double_p), ("_dchol_data", R.c_double_p),
                    ("_dK_data", R.c_double_p), ("_n", C.c_int)]

    pc = PChol.from_address(ch)
    Lc = np.tril(np.ctypeslib.as_array(pc._chol_data, shape=(k, k)).T.copy())
    L.Nfft4GPKernelSchurCombineKernelParamCreate.restype = C.c_void_p
    L.Nfft4GPKernelSchurCombineKernelParamCreate.argtypes = [R.c_double_p, C.c_int, C.c_int, C.c_int, R.c_int_p, C.c_int, R.c_double_p,
                                                             R.c_double_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    sk = L.Nfft4GPKernelSchurCombineKernelParamCreate(R.dptr(X), n, n, d, permp, k, pc._chol_data, None, fk, kh, 1, 0)
    data2 = L.Nfft4GPSubData(R.dptr(X), n, n, d, perm2p, n2)
    fs = L.Nfft4GPPrecondFsaiCreate()
    L.Nfft4GPPrecondFsaiSetLfil(fs, lfil)
    assert L.Nfft4GPPrecondFsaiSetupWithKernel(data2, n2, n2, d, C.cast(L.Nfft4GPKernelSchurCombineKernel, C.c_void_p), sk, 0, fs) == 0

    class PFsai(C.Structure):
        _fields_ = [("_lfil", C.c_int), ("_n", C.c_int), ("_tits", C.c_int), ("_titt", C.c_double), ("_tset", C.c_double),
                    ("_tlogdet", C.c_double), ("_tdvp", C.c_double), ("_L_i", R.c_int_p), ("_L_j", R.c_int_p),
                    ("_L_a", R.c_double_p), ("_dL_a", R.c_double_p), ("_work", R.c_double_p)]

    pf = PFsai.from_address(fs)
    ia = np.ctypeslib.as_array(pf._L_i, shape=(n2 + 1,)).copy()
    nnz = int(ia[-1])
    ja = np.ctypeslib.as_array(pf._L_j, shape=(nnz,)).copy()
    a = np.ctypeslib.as_array(pf._L_a, shape=(nnz,)).copy()
    # one Schur block straight from the Schur kernel (rows 5..24 of the non-landmark points)
    blk = np.arange(5, 25, dtype=np.int32)
    gidx = np.ascontiguousarray(blk)                     # the Schur kernel looks its rows up through _ibufferp / _libufferp (fsai.c:377-381)
    ks = R.Nfft4GPKernel.from_address(sk)
    ipp = C.cast(ks._ibufferp, C.POINTER(R.c_int_p))
    lpp = C.cast(ks._libufferp, C.POINTER(C.c_size_t))
    ipp[0] = R.iptr(gidx)
    lpp[0] = len(blk)
    sub = L.Nfft4GPSubData(data2, n2, n2, d, R.iptr(gidx), len(blk))
    Sb = C.c_void_p()
    assert L.Nfft4GPKernelSchurCombineKernel(sk, sub, len(blk), len(blk), d, None, 0, None, 0, C.byref(Sb), None) == 0
    Sblk = np.ctypeslib.as_array(C.cast(Sb, R.c_double_p), shape=(len(blk), len(blk))).T.copy()
    Sblk = np.tril(Sblk) + np.tril(Sblk, -1).T
    # apply (afn.c:109-136): rp = r(perm); y1 = A11^-1 rp1; rp2 -= A21 y1; y2 = FSAI(rp2); rp1 -= A12 y2; y1 = A11^-1 rp1; scatter
    K12p = C.c_void_p()
    assert L.Nfft4GPKernelGaussianKernel(kh, R.dptr(X), n, n, d, permp, k, perm2p, n2, C.byref(K12p), None) == 0
    K12 = np.ctypeslib.as_array(C.cast(K12p, R.c_double_p), shape=(n2, k)).T.copy()
    rhs = np.random.default_rng(92).random(n) - 0.5
    rp = rhs[perm]
    y1 = np.zeros(k)
    r1 = np.ascontiguousarray(rp[:k])
    L.Nfft4GPPrecondCholSolve(ch, k, R.dptr(y1), R.dptr(r1))
    r2 = np.ascontiguousarray(rp[k:] - K12.T @ y1)
    y2 = np.zeros(n2)
    L.Nfft4GPPrecondFsaiSolve(fs, n2, R.dptr(y2), R.dptr(r2))
    r1 = np.ascontiguousarray(rp[:k] - K12 @ y2)
    L.Nfft4GPPrecondCholSolve(ch, k, R.dptr(y1), R.dptr(r1))
    z = np.zeros(n)
    z[perm] = np.concatenate([y1, y2])
    out = {"X": X, "f": f, "l": l, "mu": mu, "k": k, "lfil": lfil, "fps_sel": sel, "perm": perm, "L": Lc, "K12": K12, "fsai_ia": ia,
           "fsai_ja": ja, "fsai_a": a, "schur_rows": blk, "schur_block": Sblk, "rhs": rhs, "z": z}

    # ---- rank estimation (rankest.c:248-391), one sample per call, rows replayed through srand()
    class Rankest(C.Structure):  # INC/_linearalg.h:913-923
        _fields_ = [("_nsample", C.c_int), ("_nsample_r", C.c_int), ("_max_rank", C.c_int), ("_full_tol", C.c_double),
                    ("_kernel_func", C.c_void_p), ("_kernel_str", C.c_void_p), ("_ordering_str", C.c_void_p), ("_perm", C.c_void_p)]

    L.Nfft4GPRankestStrCreate.restype = C.c_void_p
    L.Nfft4GPRankestNysScaled.argtypes = [C.c_void_p, R.c_double_p, C.c_int, C.c_int, C.c_int]
    L.Nfft4GPRandPerm.restype = R.c_int_p
    L.Nfft4GPRandPerm.argtypes = [C.c_int, C.c_int]
    cases = [(2000, 3, 1.0, 200, 300), (2000, 3, 4.0, 200, 300), (2000, 6, 1.5, 200, 2000), (1500, 3, 0.3, 150, 2000), (1200, 3, 2.0, 300, 500)]
    rk_cfg, rk_rows, rk_out = [], [], []
    for ci, (nn, dd, ll, m, maxr) in enumerate(cases):
        Xr = cloud(nn, dd, 95)
        kh2 = L.Nfft4GPKernelParamCreate(nn, 1)
        R.set_hyper(kh2, 1.0, ll, 0.01)
        rs = L.Nfft4GPRankestStrCreate()
        r = Rankest.from_address(rs)
        r._nsample, r._nsample_r, r._max_rank, r._kernel_str = m, 1, maxr, kh2
        libc.srand(777 + ci)
        rows = np.ctypeslib.as_array(L.Nfft4GPRandPerm(nn, m), shape=(m,)).copy()
        libc.srand(777 + ci)
        rk = L.Nfft4GPRankestNysScaled(rs, R.dptr(Xr), nn, nn, dd)
        rk_cfg.append([nn, dd, ll, m, maxr])
        rk_rows.append(np.pad(rows, (0, 300 - m), constant_values=-1))
        rk_out.append(rk)
    out.update(rankest_cfg=np.array(rk_cfg), rankest_rows=np.array(rk_rows, dtype=np.int32), rankest_rank=np.array(rk_out), rankest_seed=95)
    np.savez_compressed(os.path.join(OUT, "afn_ref_parts400.npz"), **out)
    print("wrote afn case: nnz", nnz, "ranks", rk_out)


def fsai_grad_case():
    """Reference FSAI set up with require_grad (fsai.c:302-673 gradient branch :495-560) on the reference's plain Gaussian kernel, the
    problem of fsai_case: derivative factor values dG/df, dG/dl, dG/dmu, Nfft4GPPrecondFsaiTrace (:221-276), Nfft4GPPrecondFsaiDvp
    (:125-219) and the two sparse triangular solves Nfft4GPPrecondFsaiInvL / InvLT (:675-721)."""
    z = np.load(os.path.join(OUT, "fsai_gauss300.npz"))
    X, f, l, mu, rhs, lfil = z["X"], float(z["f"]), float(z["l"]), float(z["mu"]), z["rhs"], int(z["fsai_lfil"])
    n, d = X.shape[1], X.shape[0]
    L.Nfft4GPKernelParamCreate.restype = C.c_void_p
    L.Nfft4GPKernelParamCreate.argtypes = [C.c_int, C.c_int]
    kh = L.Nfft4GPKernelParamCreate(n, 1)
    R.set_hyper(kh, f, l, mu)
    fs = L.Nfft4GPPrecondFsaiCreate()
    L.Nfft4GPPrecondFsaiSetLfil(fs, lfil)
    assert L.Nfft4GPPrecondFsaiSetupWithKernel(R.dptr(X), n, n, d, C.cast(L.Nfft4GPKernelGaussianKernel, C.c_void_p), kh, 1, fs) == 0

    class PFsai(C.Structure):
        _fields_ = [("_lfil", C.c_int), ("_n", C.c_int), ("_tits", C.c_int), ("_titt", C.c_double), ("_tset", C.c_double),
                    ("_tlogdet", C.c_double), ("_tdvp", C.c_double), ("_L_i", R.c_int_p), ("_L_j", R.c_int_p),
                    ("_L_a", R.c_double_p), ("_dL_a", R.c_double_p), ("_work", R.c_double_p)]

    pf = PFsai.from_address(fs)
    ia = np.ctypeslib.as_array(pf._L_i, shape=(n + 1,)).copy()
    nnz = int(ia[-1])
    assert (ia == z["fsai_ia"]).all()
    dL = np.ctypeslib.as_array(pf._dL_a, shape=(3, nnz)).copy()
    tr = np.zeros(3)
    trp = R.dptr(tr)
    L.Nfft4GPPrecondFsaiTrace.argtypes = [C.c_void_p, C.POINTER(R.c_double_p)]
    assert L.Nfft4GPPrecondFsaiTrace(fs, C.byref(trp)) == 0
    y = np.zeros(3 * n)
    yp = R.dptr(y)
    L.Nfft4GPPrecondFsaiDvp.argtypes = [C.c_void_p, C.c_int, C.c_void_p, R.c_double_p, C.POINTER(R.c_double_p)]
    assert L.Nfft4GPPrecondFsaiDvp(fs, n, None, R.dptr(rhs), C.byref(yp)) == 0
    xl, xlt = np.zeros(n), np.zeros(n)
    for nm in ("Nfft4GPPrecondFsaiInvL", "Nfft4GPPrecondFsaiInvLT"):
        getattr(L, nm).argtypes = [C.c_void_p, C.c_int, R.c_double_p, R.c_double_p]
    L.Nfft4GPPrecondFsaiInvL(fs, n, R.dptr(xl), R.dptr(rhs))
    L.Nfft4GPPrecondFsaiInvLT(fs, n, R.dptr(xlt), R.dptr(rhs))
    np.savez_compressed(os.path.join(OUT, "fsai_grad_gauss300.npz"), dL=dL, trace=tr, dvp=y, invl=xl, invlt=xlt)
    print("wrote fsai grad case: traces", tr)


def nys_additive_quirk_case():
    """The reference's Nystrom setup handed the dense ADDITIVE kernel (what TESTS/TEST3/foo3.cpp does): nys.c:549-566 forms K11 by calling
    the kernel on k sub-sampled points with n = k, and kernels.c:3160 strides the handle's feature copy with that n, so the factored
    landmark block belongs to the first P dw k values of the copy read as a [P dw] x k table.  Recorded so that the library's
    NFFT4GP_B200_REFERENCE_QUIRKS=1 mode is pinned on the compiled reference (its default uses the intended block)."""
    n, d, k = 300, 6, 40
    X = np.random.default_rng(3).random((d, n)) * n ** (1.0 / d)
    win = np.arange(6, dtype=np.int32)
    f, l, mu = 1.1, 0.9, 0.05
    D = R.RefAdditiveDense(X, win, 2, 3, "gaussian")
    R.set_hyper(D.h, f, l, mu)
    perm = np.random.default_rng(4).permutation(n).astype(np.int32)
    nys = L.Nfft4GPPrecondNysCreate()
    L.Nfft4GPPrecondNysSetRank(nys, k)
    L.Nfft4GPPrecondNysSetPerm(nys, R.iptr(perm), 0)
    assert L.Nfft4GPPrecondNysSetupWithKernel(R.dptr(D.X), n, n, d, C.cast(L.Nfft4GPKernelAdditiveKernel, C.c_void_p), D.h, 0, nys) == 0
    rhs = np.random.default_rng(5).random(n) - 0.5
    z = np.zeros(n)
    L.Nfft4GPPrecondNysSolve(nys, n, R.dptr(z), R.dptr(rhs))
    np.savez_compressed(os.path.join(OUT, "nys_additive_quirk300.npz"), X=X, windows=win, f=f, l=l, mu=mu, k=k, perm=perm, rhs=rhs, z=z)
    print("wrote nys additive quirk case")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "quirk":
        nys_additive_quirk_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "fsaigrad":
        fsai_grad_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "afn":
        afn_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "nysgrad":
        nys_grad_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "chol":
        chol_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "lanczos":
        lanczos_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "predict":
        predict_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "fgmres":
        fgmres_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "nys":
        nys_plain_case()
        sys.exit(0)
    nfft_case("gauss_3x3d", 400, 9, list(range(9)), 3, 3, "gaussian", 1.2, 0.5, 0.01, 11)
    nfft_case("bike_pattern", 350, 13, [1, 8, 6, 9, 4, 2, 5, 3, 11], 3, 3, "gaussian", 0.9, 0.45, 0.02, 12)
    nfft_case("padded_last", 300, 5, [0, 1, 3, 2, 4, -1], 2, 3, "gaussian", 1.0, 0.6, 0.01, 13)
    nfft_case("matern_2d", 300, 4, [0, 1, 2, 3], 2, 2, "matern12", 1.1, 1.5, 0.05, 14)
    nfft_case("gauss_1d", 250, 3, [0, 1, 2], 3, 1, "gaussian", 1.3, 1.2, 0.01, 15)
    solver_case()
    fsai_case()
    nys_plain_case()
    fgmres_case()
    predict_case()
    lanczos_case()
    afn_case()
    fsai_grad_case()
    nys_additive_quirk_case()
