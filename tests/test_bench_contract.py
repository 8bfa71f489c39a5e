"""bench.py contract checks that need no GPU: the reference arm prints exactly ONE JSON line with the agreed keys (C-level
prints of the reference go to stderr), W >= 3 warm-up steps are enforced, and the product arm refuses to run without a CUDA
device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=600):
    env = dict(os.environ, OMP_NUM_THREADS="1")                      # what torch.distributed.run exports to its workers
    env.pop("RANK", None)
    env.pop("WORLD_SIZE", None)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT,
                          env=env)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = _run(["--impl", "reference", "--n", "20000", "--steps", "2", "--warmup", "1"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout[:2000]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "nfft_additive_matvecs_per_s" and d["unit"] == "matvecs/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["warmup"] >= 3 and d["steps"] == 2                      # W >= 3 is enforced whatever the caller asks for
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1000.0) < 1e-6 * 1000.0
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["cores"] == cb["host_cores"]                           # all host cores, whatever OMP_NUM_THREADS the launcher exported
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--n", "20000", "--steps", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_product_arm_fails_loudly_without_a_gpu():
    import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as pkg
    if pkg.lib().Nfft4GPB200DeviceCount() > 0:
        pytest.skip("a CUDA device is present")
    p = _run(["--n", "20000", "--steps", "1", "--no-pcg", "--no-cpu"], timeout=300)
    assert p.returncode != 0
    assert not any(l.strip().startswith("{") for l in p.stdout.splitlines())      # no JSON line from a path that did not run


def test_binding_roofline_reports_the_tensor_roof_for_the_3d_kernels():
    sys.path.insert(0, ROOT)
    import bench
    base = {"bound": "hbm", "kernel": "spread", "achieved": 293.0, "peak": 6536.7, "unit": "GB/s", "frac": 293.0 / 6536.7,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)", "traffic": 7.5e8, "launch_ms": 0.327, "fp64": {"frac": 0.49}}
    r = bench.binding_roofline(base, "spread", 18.3)              # no in-run measurement: the recorded fallback peak
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and r["peak"] == bench.FP64_TENSOR_PEAK_FALLBACK_TFLOPS
    assert abs(r["frac"] - 18.3 / bench.FP64_TENSOR_PEAK_FALLBACK_TFLOPS) < 1e-15 and abs(r["achieved"] - 18.3) < 1e-15
    assert "fallback" in r["peak_source"]
    m = bench.binding_roofline(base, "spread", 18.3, 36.4, True)    # the DMMA rate measured in the run is the denominator
    assert m["peak"] == 36.4 and abs(m["frac"] - 18.3 / 36.4) < 1e-15 and "IN THIS RUN" in m["peak_source"]
    assert r["hbm"]["unit"] == "GB/s" and abs(r["hbm"]["frac"] - base["frac"]) < 1e-15 and r["traffic"] == 7.5e8
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r
    assert base["bound"] == "hbm"                                    # input untouched
    same = bench.binding_roofline(dict(base, kernel="combine"), "combine", 0.1)
    assert same["bound"] == "hbm" and "hbm" not in same


def test_both_arms_name_the_same_workload():
    """The driver pairs the two arms by config: identical `config.workload` strings."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count("WORKLOAD.format(n=n)") == 2 and src.count('"workload":') == 2


def test_ncu_traffic_lookup_is_by_key():
    sys.path.insert(0, ROOT)
    import bench
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        pytest.skip("no committed ncu traffic summary")
    t = json.load(open(p))
    for cls, e in t.items():
        val, src = bench.ncu_traffic(cls, int(e["n"]), int(e.get("n_gpus", 1)))
        assert val == float(e["dram_bytes_read"]) + float(e["dram_bytes_write"]) and os.path.exists(os.path.join(ROOT, src))
        assert bench.ncu_traffic(cls, int(e["n"]) + 1, 1) == (None, None)      # another problem size: no figure
