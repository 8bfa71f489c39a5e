"""BASELINE config C5: GP hyper-parameter loss + gradient terms by stochastic Lanczos quadrature -- per Rademacher probe one derivative-kernel
(gradient) NFFT matvec and one Lanczos run (Nfft4GPLanczosQuadratureLogdet, SRC/solvers/lanczos.c:421-610) -- for 32 probes at
n = 4 10^6, d = 9, with the probes distributed over the GPUs (Nfft4GPB200SetProbeSharding: probe i on rank i mod R, every rank holds
the whole operator, ONE allreduce of four doubles at the end; SURVEY.md 8(e) "independent probe vectors are an additional
embarrassingly-parallel axis").  Wall time of the whole call, max over ranks; parity: the sharded estimate against the
single-process estimate with the same probes at a smaller n (where one rank can run all of them in seconds).

  python scripts/c5_slq.py [n] [probes] [maxits]
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 scripts/c5_slq.py 4e6 32 20
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as pkg  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4000000
probes = int(sys.argv[2]) if len(sys.argv) > 2 else 32
maxits = int(sys.argv[3]) if len(sys.argv) > 3 else 20
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


def problem(m, seed=906):
    rng = np.random.default_rng(seed)
    X = rng.random((9, m)) * m ** (1.0 / 9)
    Z = np.where(rng.random((probes, m)) < 0.5, -1.0, 1.0)
    A = pkg.AdditiveNFFTKernel(X, np.arange(9, dtype=np.int32), 3, 3, "gaussian").build(1.0, 0.25 * m ** (1.0 / 9), 0.01)
    return A, Z


def slq(A, Z, m, sharded):
    L.Nfft4GPB200SetProbeSharding(1 if sharded else 0)
    gfn = C.cast(L.Nfft4GPAdditiveNFFTGradMatSymv, C.c_void_p)
    out = pkg.lanczos_quadrature_logdet(A.K, A.dK if hasattr(A, "dK") else A.K, m, A.matvec_fn(), gfn, maxits, Z, print_level=0)
    L.Nfft4GPB200SetProbeSharding(0)
    return out


# ---- parity at a size where one rank runs all probes quickly
m = 100000
As, Zs = problem(m, 907)
val_sh, dval_sh = slq(As, Zs, m, world > 1)
par = None
if rank == 0:
    val_1, dval_1 = slq(As, Zs, m, False)
    par = {"n": m, "logdet_rel_diff": abs(val_sh - val_1) / abs(val_1), "dlogdet_max_rel_diff": float(np.abs(dval_sh - dval_1).max() / np.abs(dval_1).max())}
elif world > 1:
    pass
As.free()
barrier()
# ---- timing at the full size
t0 = time.perf_counter()
A, Z = problem(n)
barrier()
plan_s = max_over_ranks(time.perf_counter() - t0)
slq(A, Z[: max(world, 1)], n, world > 1)     # warm-up: one probe per rank
barrier()
t0 = time.perf_counter()
val, dval = slq(A, Z, n, world > 1)
barrier()
sec = max_over_ranks(time.perf_counter() - t0)
if rank == 0:
    print(json.dumps({"config": f"C5: stochastic Lanczos quadrature of log det K and its gradient, n={n}, d=9 (three 3-D windows), {probes} Rademacher probes, "
                                f"{maxits} Lanczos steps each (1 gradient matvec + {maxits} matvecs per probe)", "n_gpus": world,
                      "parallelism": "single" if world == 1 else f"probes sharded x{world} (probe i on rank i mod {world}), one allreduce of 4 doubles",
                      "seconds": sec, "probes_per_s": probes / sec, "ms_per_probe": 1000.0 * sec / probes, "plan_build_s": plan_s,
                      "logdet_per_n": val, "dlogdet_per_n": [float(v) for v in dval], "parity_sharded_vs_single_process": par}), flush=True)
A.free()
if dist is not None:
    dist.barrier()
    L.Nfft4GPB200CommFinalize()
    dist.destroy_process_group()
