"""BASELINE config C4: additive Matern-1/2 kernel with many 1-D / 2-D windows (d = 64 features, n = 4 10^6), windows sharded across
the GPUs with ONE ncclAllReduce of y per matvec (SURVEY.md 8(e) mode 1).  Device-resident timing (CUDA events on the library
stream, barrier on both sides, max over ranks), per-class kernel times, achieved HBM rate on the ALGORITHMIC bytes
((16 d_w + 16) n per window) against the measured copy bandwidth, and a parity figure of the N-GPU result against
the single-process oracle on a sub-sampled problem.  One JSON line per window shape.

  python scripts/c4_matvec.py [n] [d] [dims: 1,2]                     one GPU
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/c4_matvec.py 4e6 64 1,2
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as pkg  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dims = [int(t) for t in sys.argv[3].split(",")] if len(sys.argv) > 3 else [2, 1]
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
L = pkg.lib()
assert L.Nfft4GPB200Init(local_rank) == 0
dist = torch = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_char * 128)()
        assert L.Nfft4GPB200CommGetUniqueId(buf) == 0
        uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    uid = uid.cuda()
    dist.broadcast(uid, 0)
    assert L.Nfft4GPB200CommInit(rank, world, bytes(uid.cpu().numpy().tobytes())) == 0


def barrier():
    L.Nfft4GPB200DeviceSynchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()


def max_over_ranks(v):
    if dist is None:
        return v
    t = torch.tensor([v], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


hbm_peak = 6536.7
try:
    hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
rng = np.random.default_rng(906)
X = rng.random((d, n)) * n ** (1.0 / d)
x = rng.random(n) - 0.5
f, l, mu = 1.0, 0.5 * n ** (1.0 / d), 0.01
names = ["prep", "spread", "fft_fwd", "fft_inv", "gather", "combine", "vec", "nys", "csr", "gemv", "other"]
for dw in dims:
    P = d // dw
    win = np.arange(d, dtype=np.int32)
    import time
    t0 = time.perf_counter()
    A = pkg.AdditiveNFFTKernel(X, win, P, dw, "matern12")
    if world > 1:
        assert L.Nfft4GPB200SetSharding(A.h, 1) == 0
    A.build(f, l, mu)
    barrier()
    plan_s = max_over_ranks(time.perf_counter() - t0)
    dx = L.Nfft4GPB200DeviceAlloc(8 * n)
    dy = L.Nfft4GPB200DeviceAlloc(8 * n)
    L.Nfft4GPB200MemcpyH2D(dx, x.ctypes.data, 8 * n)
    e0, e1 = L.Nfft4GPB200EventCreate(), L.Nfft4GPB200EventCreate()
    for _ in range(3):
        A.matvec_device(dx, dy)
    reps = 10
    barrier()
    L.Nfft4GPB200EventRecord(e0)
    for _ in range(reps):
        A.matvec_device(dx, dy)
    L.Nfft4GPB200EventRecord(e1)
    barrier()
    ms = max_over_ranks(L.Nfft4GPB200EventElapsedMs(e0, e1) / reps)
    L.Nfft4GPB200ProfileStart()
    for _ in range(reps):
        A.matvec_device(dx, dy)
    sums = (C.c_double * 11)()
    cnts = (C.c_longlong * 11)()
    L.Nfft4GPB200ProfileStop(sums, cnts, 11)
    y = np.zeros(n)
    L.Nfft4GPB200MemcpyD2H(y.ctypes.data, dy, 8 * n)
    # parity: rows of the N-GPU result against the dense Matern sums are the approximation's bar, so compare with the CPU oracle
    # on the same operator restricted to a sub-sample instead: build both on the first m points
    m = min(n, 20000)
    poly2d = dw == 2 and n >= 65536 and os.environ.get("NFFT4GP_B200_POLY2D") != "0" and not os.environ.get("NFFT4GP_B200_WINDOW_EVAL")
    rel = None
    if rank == 0:
        from oracle import oracle_api as O
        Xs = np.ascontiguousarray(X[:, :m])
        os.environ["NFFT4GP_B200_POLY1D"] = "1" if dw == 1 else "0"
        keep2 = os.environ.get("NFFT4GP_B200_POLY2D")
        if poly2d:
            os.environ["NFFT4GP_B200_POLY2D"] = "1"      # the sub-sample is below the automatic threshold: same path as the timed run
        As = pkg.AdditiveNFFTKernel(Xs, win, P, dw, "matern12")
        As.build(f, l, mu)
        os.environ.pop("NFFT4GP_B200_POLY1D", None)
        os.environ.pop("NFFT4GP_B200_POLY2D", None)
        if keep2 is not None:
            os.environ["NFFT4GP_B200_POLY2D"] = keep2
        Bs = O.OracleAdditiveNFFT(Xs, win, P, dw, "matern12").build(f, l, mu)
        ys, yo = As.matvec(x[:m]), Bs.matvec(x[:m])
        rel = float(np.linalg.norm(ys - yo) / np.linalg.norm(yo))
        As.free()
        Bs.free()
    rel_1gpu = None
    if rank == 0 and world > 1:
        # the sharded result against the same matvec computed by this rank alone (unsharded handle on the full data)
        A1 = pkg.AdditiveNFFTKernel(X, win, P, dw, "matern12").build(f, l, mu)
        y1 = A1.matvec(x)
        rel_1gpu = float(np.linalg.norm(y - y1) / np.linalg.norm(y1))
        A1.free()
    alg = (16 * dw + 16) * n * P
    line = {"config": f"C4: additive Matern-1/2 NFFT matvec, n={n}, d={d}, {P} windows of dimension {dw}", "n_gpus": world,
            "parallelism": "single" if world == 1 else f"window-sharded x{world} + 1 allreduce(y, {8 * n / 1e6:.0f} MB) per matvec",
            "ms_per_matvec": ms, "matvecs_per_s": 1000.0 / ms, "plan_build_s": plan_s,
            "algorithmic_bytes": alg, "algorithmic_GBs": alg / ms / 1e6, "hbm_peak_GBs_per_gpu": hbm_peak,
            "hbm_frac_algorithmic": alg / ms / 1e6 / (hbm_peak * world),   # against the aggregate roof of the GPUs used
            "per_class_us_rank0": {names[i]: round(sums[i] / reps * 1000, 1) for i in range(11) if cnts[i]},
            "launches_per_matvec_rank0": int(sum(cnts[i] for i in range(11)) / reps),
            "path": ("table-free polynomial path (nfft1d_poly.cu)" if dw == 1 and n >= 32768 and os.environ.get("NFFT4GP_B200_POLY1D") != "0" else
                     "table-free bivariate polynomial path on the fp64 tensor cores (nfft2d_poly.cu)" if poly2d else
                     "tap tables, fp64 tensor-core kernels (nfft2d_mma.cu)" if dw == 2 else "tap tables, streaming kernels (nfft12_stream.cu)"),
            "parity_rel_l2_vs_oracle_on_first_20000_points": rel, "parity_rel_l2_sharded_vs_one_gpu": rel_1gpu}
    if rank == 0:
        print(json.dumps(line), flush=True)
    A.free()
    L.Nfft4GPB200DeviceFree(dx)
    L.Nfft4GPB200DeviceFree(dy)
if dist is not None:
    dist.barrier()
    L.Nfft4GPB200CommFinalize()
    dist.destroy_process_group()
