#!/usr/bin/env python
"""bench.py -- headline benchmark of nfft4gp_b200 (contract: see the task statement / DESIGN.md section "Measurement").

Metric  : NFFT additive-kernel matvecs/s (BASELINE.json), config C2: n = 10^6 synthetic points, d = 9 split into three
          3-feature windows, Gaussian kernel, fp64.  A "step" = one full additive matvec  y <- K x.
value   : device-resident vectors (inputs already in HBM), CUDA events on the library stream, max over ranks.
e2e     : the same metric through the reference-facing C-ABI call Nfft4GPAdditiveNFFTMatSymv with HOST (pinned) vectors,
          host->device and device->host copies inside the timed region.
roofline: dominant kernel class of the step, per-launch duration measured with CUDA events on the launching stream
          inside this run; algorithmic bytes per launch from SURVEY.md 8(d).
cpu_baseline / --impl reference: the reference's own CPU implementation of the path (reference wrapper
          nfft_interface.c on the NFFT3 fastsum restatement, oracle/_ref; or the self-contained oracle port) on the box's
          host cores.
"""
import argparse
import ctypes as C
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "nfft_additive_matvecs_per_s"
UNIT = "matvecs/s"
WORKLOAD = "C2: additive Gaussian NFFT matvec n={n} d=9 three 3-D windows N=32 n_os=64 m=4 (BASELINE.json configs[1])"
PROF_CLASSES = ["prep", "spread", "fft_fwd", "fft_inv", "gather", "combine", "vec", "nys", "csr", "gemv", "other"]


CLOUD = "uniform"   # --cloud: "uniform" (the reference's synthetic data set) or "gaussian" (same variance per coordinate)


def synthetic_problem(n, d=9, seed=906):
    """Reference-style synthetic cloud (Nfft4GPDatasetUniformRandom, SRC/utils/utils.c:266-286): n^(1/d) U(0,1),
    feature-major; x = U(0,1) - 0.5 (TESTS/TEST1/foo.cpp:243-247); l resolved by N = 32.
    --cloud gaussian: N(0, 1) n^(1/d) / sqrt(12) per coordinate (the north star's second cloud: dense cells in the centre, empty ones
    towards the bounding radius)."""
    rng = np.random.default_rng(seed)
    if CLOUD == "gaussian":
        X = rng.standard_normal((d, n)) * n ** (1.0 / d) / np.sqrt(12.0)
    else:
        X = rng.random((d, n)) * n ** (1.0 / d)
    x = rng.random(n) - 0.5
    windows = np.arange(d, dtype=np.int32)
    side = n ** (1.0 / d)
    l = 0.25 * side  # resolved by the 32 Fourier modes: NFFT vs dense ~3e-8 rel. (tests/test_oracle.py)
    return X, x, windows, dict(f=1.0, l=l, mu=0.01)


def ncu_traffic(kernel_class, n, world):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, looked up BY KEY in the committed summary of
    the `ncu --set full` capture of this very command (profiles/ncu_traffic.json, written by scripts/ncu_traffic.py from the raw-page
    CSV it names).  None when no capture exists for this kernel / problem size / GPU count."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None, None
    try:
        with open(p) as fh:
            t = json.load(fh)
        e = t.get(kernel_class)
        if not e or int(e.get("n", -1)) != n or int(e.get("n_gpus", 1)) != world:
            return None, None
        return float(e["dram_bytes_read"]) + float(e["dram_bytes_write"]), e.get("source")
    except Exception:
        return None, None


def host_cores():
    """Host cores this process may use (the affinity mask: a container limit shows here, os.cpu_count() does not see it)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def use_all_host_cores():
    """torch.distributed.run exports OMP_NUM_THREADS=1 to every worker, which made the N > 1 reference arm single-threaded in
    round 1.  The CPU legs set their OpenMP team explicitly to the cores of the box and report the number they got."""
    from oracle import oracle_api
    want = host_cores()
    os.environ["OMP_NUM_THREADS"] = str(want)
    L = oracle_api.lib()
    L.oracle_set_num_threads(want)
    return int(L.oracle_num_threads())


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_handle(X, windows, hp):
    """The reference's CPU implementation of the path: oracle/_ref (reference wrapper, unmodified, on the fastsum
    restatement) when it was built here, else the self-contained oracle port."""
    from oracle import refapi
    if refapi.available():
        try:
            A = refapi.RefAdditiveNFFT(X, windows, 3, 3, "gaussian").build(hp["f"], hp["l"], hp["mu"])
            return A, "reference"
        except Exception:
            pass
    from oracle import oracle_api
    A = oracle_api.OracleAdditiveNFFT(X, windows, 3, 3, "gaussian").build(hp["f"], hp["l"], hp["mu"])
    return A, "port"


def time_cpu(A, x, steps, warmup):
    for _ in range(warmup):
        A.matvec(x)
    t0 = time.perf_counter()
    for _ in range(steps):
        y = A.matvec(x)
    dt = time.perf_counter() - t0
    return steps / dt, y


def cpu_cores():
    from oracle import oracle_api
    return int(oracle_api.lib().oracle_num_threads())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = args.n
    cores = use_all_host_cores()
    X, x, windows, hp = synthetic_problem(n)
    A, kind = cpu_reference_handle(X, windows, hp)
    val, _ = time_cpu(A, x, args.steps, args.warmup)
    if cores == 1 and host_cores() > 1:
        print("bench.py: WARNING the CPU reference ran on 1 thread of a multi-core host", file=sys.stderr)
    sample = f"{args.steps} full additive matvecs at n={n}, d=9 (3x3-D windows) after {args.warmup} warm-up, {cores} OpenMP threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 / val, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic" if CLOUD == "uniform" else "synthetic (gaussian cloud)",
        "config": {"workload": WORKLOAD.format(n=n), "cloud": CLOUD, "l": hp["l"], "f": hp["f"], "mu": hp["mu"]},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "host_cores": host_cores(), "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


FP64_TENSOR_PEAK_FALLBACK_TFLOPS = 37.1   # DMMA.8x8x4 rate of this pool's B200 (scripts/fp64_dmma.cu), used only if the in-run measurement fails


def binding_roofline(roofline, dom, fp64_ach_tflops, fp64_peak_tflops=None, peak_measured_in_run=False):
    """The 3-D spread / gather are bound by the fp64 tensor pipe (62 flop per algorithmic byte): report THAT
    roof at the top level (bound "tensor", TFLOP/s of algorithmic fp64 flops against the DMMA rate measured in this run by
    Nfft4GPB200MeasureFp64TensorPeak) and keep the HBM view of the same launch next to it.  Kernels that are HBM-bound keep the
    HBM roof.  Pure function (tests/test_bench_contract.py)."""
    if dom not in ("spread", "gather"):
        return roofline
    peak = fp64_peak_tflops if fp64_peak_tflops and fp64_peak_tflops > 0 else FP64_TENSOR_PEAK_FALLBACK_TFLOPS
    out = dict(roofline)
    out["hbm"] = {k: roofline[k] for k in ("achieved", "peak", "unit", "frac", "peak_source") if k in roofline}
    out.update({"bound": "tensor", "achieved": fp64_ach_tflops, "peak": peak, "unit": "TFLOP/s",
                "frac": fp64_ach_tflops / peak,
                "peak_source": ("fp64 tensor-core (mma.sync.m8n8k4.f64 = DMMA.8x8x4) rate measured IN THIS RUN on this GPU by "
                                "Nfft4GPB200MeasureFp64TensorPeak (register operands, 16 warps/SM, best of 3)" if peak_measured_in_run else
                                "fallback: DMMA.8x8x4 rate measured earlier on this pool's B200 with scripts/fp64_dmma.cu") +
                               "; MEASURED_PEAKS.json has no fp64 figure (its dense bf16 number does not apply to an fp64 kernel)",
                "note": "algorithmic flops: 2 x 10^3 per point, window and direction (SURVEY.md 8(d)); the HBM view of the same launch is under 'hbm'"})
    return out


def run_pcg_section(pkg, L, A, X, rhs, windows, hp, args, hbm_peak, ranks=None):
    """Second half of BASELINE.json's metric: AFN-preconditioned CG on the NFFT additive matvec to 1e-6 relative residual
    (config C3, n = 10^6, d = 9, Gaussian).  Two landmark ranks: the one BASELINE.json names (500; too small for n = 10^6 at
    mu = 0.01 -- reported with an iteration cap) and one that converges.  Setup (GPU: KNN pattern, FSAI rows, landmark GEMMs
    on the fp64 tensor cores) and solve are timed separately with a device synchronise on both sides; a second, profiled
    pass of 50 iterations gives the per-class kernel times of one iteration and the HBM fraction of the AFN apply."""
    n = X.shape[1]
    mu = args.pcg_mu
    rank, world, dist, torch = ranks if ranks else (0, 1, None, None)

    def sync_all():
        L.Nfft4GPB200DeviceSynchronize()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    A.build(hp["f"], hp["l"], mu)
    D = pkg.DenseAdditiveKernel(X, windows, 3, 3, "gaussian").set_hyper(hp["f"], hp["l"], mu)
    out = {"config": f"C3: AFN-PCG to 1e-6 rel. residual, n={n}, d=9, Gaussian, f={hp['f']}, l={hp['l']:.4f}, mu={mu}, lfil=100, rhs=U(0,1)-0.5",
           "unit": "s", "runs": []}
    variants = [(500, 300, 64), (args.pcg_rank, 3000, 64)]
    if world == 1:
        variants.append((args.pcg_rank, 3000, "nfft"))   # labelled extra: K12 products of the apply through the NFFT operator (fp64)
    variants.append((args.pcg_rank, 3000, 32))           # labelled extra: K12 block stored in fp32
    for lrank, maxits, k12_bits in variants:
        sync_all()
        t0 = time.perf_counter()
        if k12_bits == 64:
            P = pkg.AfnPrecond().setup(D, lrank, 100, numests=0, noran=1)   # N > 1: replicated on every rank; the apply is sharded
        elif k12_bits == "nfft":
            P = P_keep.set_k12_operator(A)       # same factors; the dense block stays (the fp32 run below converts it)
        else:
            P_keep.set_k12_operator(None)
            P = P_keep.set_k12_storage(32)       # same factors, K12 block rounded to fp32 in place (conversion time = "setup_s")
        sync_all()
        setup_s = max_over_ranks(time.perf_counter() - t0)
        L.Nfft4GPB200LaunchCountReset()
        sync_all()
        t0 = time.perf_counter()
        xs, rel, hist, its = pkg.solver_pcg(A.K, n, A.matvec_fn(), P.h, P.solve_fn(), np.zeros(n), rhs, maxits, 0, 1e-6)
        sync_all()
        solve_s = max_over_ranks(time.perf_counter() - t0)
        launches = int(L.Nfft4GPB200LaunchCount())
        done = its if its > 0 else maxits
        # profiled pass: 50 iterations with per-launch events
        L.Nfft4GPB200ProfileStart()
        pkg.solver_pcg(A.K, n, A.matvec_fn(), P.h, P.solve_fn(), np.zeros(n), rhs, 50, 0, 1e-30)
        sums = (C.c_double * len(PROF_CLASSES))()
        counts = (C.c_longlong * len(PROF_CLASSES))()
        L.Nfft4GPB200ProfileStop(sums, counts, len(PROF_CLASSES))
        # 50 iterations + the initial residual (one more matvec): matvec classes ran 51 times, everything else belongs to the 50 iterations
        mv_cls = ("prep", "spread", "fft_fwd", "fft_inv", "gather", "combine")
        per_it = {PROF_CLASSES[i]: round(sums[i] / (51.0 if PROF_CLASSES[i] in mv_cls else 50.0), 5) for i in range(len(PROF_CLASSES)) if counts[i] > 0}
        n2 = (n - lrank) / world
        k12_label = k12_bits
        if k12_bits == "nfft":
            k12_bits = 0                                    # no dense block is read
        gemv_bytes = 2.0 * (k12_bits / 8.0) * lrank * n2   # two passes over this rank's columns of K12 per apply
        csr_bytes = 2.0 * 12.0 * 100.0 * n2          # this rank's rows of G and G' (value + column index) per apply
        run = {"landmark_rank": lrank, "baseline_config": bool(lrank == 500 and k12_bits == 64), "iteration_cap": maxits,
               "precond": f"AFN rank {lrank} + FSAI(lfil 100) of the Schur complement" + (", K12 block stored in fp32 (fp64 vectors and sums)" if k12_bits == 32 else "")
               + (", K12 products of the apply through the NFFT operator (two fast matvecs, fp64; Nfft4GPB200PrecondAFNSetK12Operator)" if k12_label == "nfft" else "")
               + (f", apply sharded over {world} GPUs (2 allreduces of k, 2 of n-k doubles)" if world > 1 else ""),
               "k12_storage_bits": 64 if k12_label == "nfft" else k12_bits, "k12_form": "nfft-operator" if k12_label == "nfft" else "dense-block", "setup_s": setup_s, "solve_s": solve_s,
               "iters": its, "converged": bool(its > 0), "rel_res": rel, "ms_per_iter": 1000.0 * solve_s / max(done, 1),
               "gpu_launches": launches, "per_class_ms_per_iter": per_it,
               "afn_apply_hbm": {"gemv_GBs": gemv_bytes / (per_it.get("gemv", float("nan")) * 1e-3) / 1e9,
                                 "spmv_GBs": csr_bytes / (per_it.get("csr", float("nan")) * 1e-3) / 1e9, "peak_GBs": hbm_peak}}
        out["runs"].append(run)
        if lrank == args.pcg_rank and k12_label == 64:
            P_keep = P
        elif k12_label == "nfft":
            pass                                   # P is P_keep
        else:
            P.free()
    # Headline = fp64 storage only.  BASELINE's own rank (500) first; when it does not converge inside its cap the value is the
    # fp64 run at the rank that does (off-config, labelled).  The fp32-K12 run is an extra and never the value.
    base = next((r for r in out["runs"] if r["baseline_config"]), None)
    fp64_conv = [r for r in out["runs"] if r["converged"] and r["k12_storage_bits"] == 64 and r["k12_form"] == "dense-block"]
    pick = base if (base and base["converged"]) else (min(fp64_conv, key=lambda r: r["solve_s"]) if fp64_conv else None)
    out["value"] = pick["solve_s"] if pick else None
    out["value_run"] = (pick["precond"] + ("" if pick.get("baseline_config") else " -- OFF-CONFIG rank: BASELINE's rank 500 does not converge (see runs[0])")) if pick else None
    out["value_storage"] = "fp64"
    out["baseline_config_converged"] = bool(base and base["converged"])
    op = next((r for r in out["runs"] if r["k12_form"] == "nfft-operator"), None)
    if op is not None:
        out["extra_nfft_k12"] = {"solve_s": op["solve_s"], "iters": op["iters"], "converged": op["converged"], "ms_per_iter": op["ms_per_iter"],
                                 "note": "same AFN factors, fp64 throughout; the two K12 products of every apply are NFFT matvecs instead of passes "
                                         "over the dense block (the reference's afn.h keeps A12 behind a function pointer).  Reported for "
                                         "information, never the value"}
    extra = next((r for r in out["runs"] if r["k12_storage_bits"] == 32), None)
    if extra is not None:
        out["extra_fp32_k12"] = {"solve_s": extra["solve_s"], "iters": extra["iters"], "converged": extra["converged"],
                                 "note": "K12 block rounded to fp32 (narrower storage than the reference): reported for information, never the value"}
    D.free()
    A.build(hp["f"], hp["l"], hp["mu"])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--n", type=int, default=1000000)
    ap.add_argument("--cloud", choices=["uniform", "gaussian"], default="uniform", help="synthetic point cloud (default: the reference's uniform one)")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-pcg", action="store_true", help="skip the C3 AFN-PCG solve section")
    ap.add_argument("--pcg-rank", type=int, default=4000, help="landmark rank of the converging AFN configuration")
    ap.add_argument("--pcg-mu", type=float, default=0.01)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    global CLOUD
    CLOUD = args.cloud
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import torch  # plumbing only: torch.distributed rendezvous + barrier

    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as pkg
    L = pkg.lib()
    if L.Nfft4GPB200Init(local_rank) != 0:
        raise SystemExit("nfft4gp_b200: no usable CUDA device (no CPU fallback)")

    n = args.n
    X, x, windows, hp = synthetic_problem(n)
    A = pkg.AdditiveNFFTKernel(X, windows, 3, 3, "gaussian")

    scaling = "strong"
    parallelism = "single"
    mode = 0
    if world > 1:
        # one NCCL communicator for the library; the unique id travels through torch.distributed
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_char * 128)()
            assert L.Nfft4GPB200CommGetUniqueId(buf) == 0
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        raw = bytes(uid.cpu().numpy().tobytes())
        assert L.Nfft4GPB200CommInit(rank, world, raw) == 0
        mode = int(os.environ.get("NFFT4GP_SHARD_MODE", "2"))
        assert L.Nfft4GPB200SetSharding(A.h, mode) == 0
        parallelism = {1: f"window-sharded x{world} + 1 allreduce(y) per matvec",
                       2: f"point-block-sharded x{world} + 1 allreduce(oversampled grids, 6 MiB) per matvec"}[mode]
    A.build(hp["f"], hp["l"], hp["mu"])
    # vectors: full length, or this rank's block of points under point-block sharding (x and y stay sharded)
    lo, hi = (rank * n // world, (rank + 1) * n // world) if mode == 2 else (0, n)
    nl = hi - lo
    xl = np.ascontiguousarray(x[lo:hi])
    dx = L.Nfft4GPB200DeviceAlloc(8 * nl)
    dy = L.Nfft4GPB200DeviceAlloc(8 * nl)
    L.Nfft4GPB200MemcpyH2D(dx, xl.ctypes.data, 8 * nl)
    e0, e1 = L.Nfft4GPB200EventCreate(), L.Nfft4GPB200EventCreate()

    def barrier():
        L.Nfft4GPB200DeviceSynchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        L.Nfft4GPB200DeviceSynchronize()

    def timed_region(steps, fn):
        barrier()
        L.Nfft4GPB200EventRecord(e0)
        for _ in range(steps):
            fn()
        L.Nfft4GPB200EventRecord(e1)
        ms = L.Nfft4GPB200EventElapsedMs(e0, e1)
        barrier()
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    step_dev = lambda: A.matvec_device(dx, dy, n_local=nl)
    for _ in range(args.warmup):
        step_dev()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    L.Nfft4GPB200LaunchCountReset()
    ms_total = timed_region(args.steps, step_dev)
    launches = int(L.Nfft4GPB200LaunchCount())
    # second pass of the same K steps with per-launch CUDA events on the library stream: kernel durations for the roofline
    L.Nfft4GPB200ProfileStart()
    ms_prof_total = timed_region(args.steps, step_dev)
    sums = (C.c_double * len(PROF_CLASSES))()
    counts = (C.c_longlong * len(PROF_CLASSES))()
    L.Nfft4GPB200ProfileStop(sums, counts, len(PROF_CLASSES))
    clocks = sampler.stop() if rank == 0 else None

    ms_per_step = ms_total / args.steps
    value = 1000.0 / ms_per_step

    # ---- e2e through the reference-facing host-pointer call: pinned host vectors (the contract's definition of e2e) and, next to it,
    # PAGEABLE ones (plain malloc'd numpy arrays -- what the reference's callers pass)
    def time_host_calls(xp, yp):
        def step_host():
            assert L.Nfft4GPAdditiveNFFTMatSymv(A.K, n, 1.0, xp, 0.0, yp) == 0
        for _ in range(3):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        L.Nfft4GPB200DeviceSynchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return args.steps / dt

    hx = L.Nfft4GPB200HostAlloc(8 * n)
    hy = L.Nfft4GPB200HostAlloc(8 * n)
    C.memmove(hx, x.ctypes.data, 8 * n)
    hxp = C.cast(hx, C.POINTER(C.c_double))
    hyp = C.cast(hy, C.POINTER(C.c_double))
    e2e_val = time_host_calls(hxp, hyp)
    x_pg = np.array(x, dtype=np.float64, copy=True)
    y_pg = np.zeros(n)
    e2e_pageable = time_host_calls(x_pg.ctypes.data_as(C.POINTER(C.c_double)), y_pg.ctypes.data_as(C.POINTER(C.c_double)))
    y_host = y_pg.copy()
    if dist is not None and mode == 2:
        # every rank filled its block [lo, hi) of y: assemble the full vector (parity of the N-GPU result against the CPU below)
        full = torch.zeros(n, device="cuda", dtype=torch.float64)
        full[lo:hi] = torch.from_numpy(y_host[lo:hi]).cuda()
        dist.all_reduce(full)
        y_host = full.cpu().numpy()

    # fp64 tensor-core roof of THIS GPU, measured in this process (the binding roof of the 3-D kernels)
    fp64_tensor_peak = float(L.Nfft4GPB200MeasureFp64TensorPeak())

    pcg_multi = None
    if world > 1 and mode == 2 and not args.no_pcg:
        pcg_multi = run_pcg_section(pkg, L, A, X, x, windows, hp, args, float(peaks()[0]["hbm_gbs"]), ranks=(rank, world, dist, torch))

    if rank != 0:
        if dist is not None:
            dist.barrier()
            L.Nfft4GPB200CommFinalize()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel class
    pk, pk_kind = peaks()
    per_class = {PROF_CLASSES[i]: {"ms_per_step": sums[i] / args.steps, "launches_per_step": counts[i] / args.steps}
                 for i in range(len(PROF_CLASSES)) if counts[i] > 0}
    dom = max(per_class, key=lambda k: per_class[k]["ms_per_step"])
    dom_launch_ms = per_class[dom]["ms_per_step"] / per_class[dom]["launches_per_step"]
    # algorithmic bytes per launch: per window and direction coordinates 8*d_w*n + vector 8*n = 32 n;
    # the fused kernels process all three windows in one launch (launches_per_step = 1), the per-window ones one each
    n_loc = nl
    win_per_launch = 3.0 / max(per_class[dom]["launches_per_step"], 1e-9) if dom in ("spread", "gather") else 1.0
    alg_bytes = {"spread": 32.0 * n_loc * win_per_launch, "gather": 32.0 * n_loc * win_per_launch}.get(dom, 16.0 * n)
    achieved = alg_bytes / (dom_launch_ms * 1e-3) / 1e9
    peak = float(pk["hbm_gbs"])
    # fp64 view of the same launch: 2 * 10^3 flop per point (1000 FMAs), B200 fp64 FMA pipe = 64 lanes/SM/clk
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    fp64_peak = 148 * 64 * 2 * sm_mhz * 1e6 / 1e12
    fp64_ach = 2000.0 * n_loc * win_per_launch / (dom_launch_ms * 1e-3) / 1e12
    traffic, traffic_src = ncu_traffic(dom, n, world)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({pk_kind})",
                "traffic": traffic, "traffic_source": traffic_src,
                "launch_ms": dom_launch_ms, "algorithmic_bytes_per_launch": alg_bytes, "windows_per_launch": win_per_launch,
                "note": "3-D windows are fp64-pipe-bound (62 flop/B, SURVEY.md finding 4): see fp64",
                "fp64": {"achieved_tflops": fp64_ach, "peak_tflops": fp64_peak, "frac": fp64_ach / fp64_peak,
                         "peak_source": f"148 SM x 64 DFMA/clk x {sm_mhz:.0f} MHz (median SM clock under load)",
                         "dmma_m8n8k4_measured_in_run_tflops": fp64_tensor_peak,
                         "earlier_unit_peaks_tflops": {"dmma_m8n8k4": 37.1, "dfma": 33.0, "source": "scripts/fp64_dmma.cu on this pool's B200: one shared unit, no overlap"}},
                "step_share": per_class[dom]["ms_per_step"] / (ms_prof_total / args.steps),
                "per_class_ms_per_step": {k: round(v["ms_per_step"], 5) for k, v in per_class.items()},
                "profiled_ms_per_step": ms_prof_total / args.steps}
    try:
        roofline = binding_roofline(roofline, dom, fp64_ach, fp64_tensor_peak, fp64_tensor_peak > 0)
    except Exception as exc:  # never lose the line over a reporting detail
        print(f"bench.py: binding_roofline failed: {exc!r}", file=sys.stderr)

    # ---- CPU baseline on the host cores (bounded sample, rank 0 at N = 1) + parity of THIS run's result against the CPU (every N)
    cpu = None
    parity = None
    if not args.no_cpu:
        cores = use_all_host_cores()
        Aref, kind = cpu_reference_handle(X, windows, hp)
        cval, y_cpu = time_cpu(Aref, x, args.cpu_steps if world == 1 else 1, 1)
        rel = float(np.linalg.norm(y_host - y_cpu) / np.linalg.norm(y_cpu))
        parity = {"rel_l2_gpu_vs_cpu": rel, "n_gpus": world, "against": kind, "tolerance": 1e-10,
                  "what": "y of the timed e2e call (all ranks' blocks assembled) vs one CPU matvec on the same inputs"}
        if world == 1:
            cpu = {"value": cval, "unit": UNIT, "cores": cores, "host_cores": host_cores(), "kind": kind,
                   "sample": f"{args.cpu_steps} full additive matvecs at n={n} (same workload) after 1 warm-up, {cores} OpenMP threads",
                   "rel_l2_gpu_vs_cpu": rel}

    pcg = None
    if world == 1 and not args.no_pcg:
        try:
            pcg = run_pcg_section(pkg, L, A, X, x, windows, hp, args, peak)
        except Exception as exc:  # the headline metric of this line is the matvec: report the failure, keep the line
            print(f"bench.py: the C3 (AFN-PCG) section failed: {exc!r}", file=sys.stderr)
            pcg = {"error": repr(exc)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic" if CLOUD == "uniform" else "synthetic (gaussian cloud)",
        "config": {"workload": WORKLOAD.format(n=n), "cloud": CLOUD,
                   "parallelism": parallelism, "l": hp["l"], "f": hp["f"], "mu": hp["mu"],
                   "l2_policy": "inputs larger than L2: window tables streamed per step = 2 x 3 x 240 MB >> 126 MB"},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": 8 * nl * world, "d2h_bytes_per_step": 8 * nl * world,
                "api": "Nfft4GPAdditiveNFFTMatSymv(host x, host y), pinned host memory",
                "pageable_value": e2e_pageable,
                "pageable_note": "same call with plain malloc'd host arrays (what the reference's callers pass): cudaMemcpy stages them through the driver's bounce buffer"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "parity": parity,
        "pcg": pcg if pcg is not None else pcg_multi,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        L.Nfft4GPB200CommFinalize()
        dist.destroy_process_group()
    return 0


def _main_with_clean_stdout():
    """The C libraries print progress lines with printf (the reference does: 'Initializing NFFT kernel ...').  Route
    everything written to fd 1 during the run to stderr so that stdout carries exactly ONE line: the JSON result."""
    import ctypes
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    out = io.StringIO()
    py_stdout, sys.stdout = sys.stdout, out
    try:
        rc = main()
    finally:
        sys.stdout = py_stdout
        try:
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    sys.stdout.write(out.getvalue())
    sys.stdout.flush()
    return rc


if __name__ == "__main__":
    sys.exit(_main_with_clean_stdout())
