#!/bin/bash
# round 2, call D (8 GPUs): C4 window-sharded, C5 probe-sharded, the bench line at N = 8 (parity key, PCG section)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/r2d_smi.txt
timeout 420 $TR --master-port 29511 scripts/c4_matvec.py 4e6 64 1,2 > gpurun_out/r2d_c4_8gpu.jsonl 2> gpurun_out/r2d_c4.err; echo "c4 rc $?"
cut -c1-900 gpurun_out/r2d_c4_8gpu.jsonl
timeout 420 $TR --master-port 29512 scripts/c5_slq.py 4e6 32 20 > gpurun_out/r2d_c5_8gpu.json 2> gpurun_out/r2d_c5.err; echo "c5 rc $?"
cut -c1-900 gpurun_out/r2d_c5_8gpu.json
timeout 600 $TR --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2d_bench_n8.json 2> gpurun_out/r2d_bench_n8.err; echo "bench rc $?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2d_bench_n8.json"))
    print("N", d["n_gpus"], "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d.get("parity"))
    print(d["roofline"].get("per_class_ms_per_step"))
    for r in (d.get("pcg") or {}).get("runs", []):
        print(r["landmark_rank"], r["k12_form"], r["k12_storage_bits"], "setup", round(r["setup_s"], 2), "solve", round(r["solve_s"], 2), "iters", r["iters"], r["per_class_ms_per_iter"])
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/r2d_bench_n8.err
