#!/bin/bash
# round 2, call J (8 GPUs): C4 with the 2-D tensor-core kernels, C5 with the device-resident probe loop, multi-GPU parity script
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 420 $TR --master-port 29531 scripts/c4_matvec.py 4e6 64 1,2 2>>gpurun_out/r2j_err.log | grep '^{' > gpurun_out/r2j_c4_8gpu.jsonl; echo "c4 rc $?"
cut -c1-420 gpurun_out/r2j_c4_8gpu.jsonl
timeout 420 $TR --master-port 29532 scripts/c5_slq.py 4e6 32 20 2>>gpurun_out/r2j_err.log | grep '^{' > gpurun_out/r2j_c5_8gpu.json; echo "c5 rc $?"
cut -c1-700 gpurun_out/r2j_c5_8gpu.json
timeout 300 $TR --master-port 29533 scripts/check_multigpu.py 100000 2>&1 | grep "mode" > gpurun_out/r2j_check_multigpu_8gpus.txt; cat gpurun_out/r2j_check_multigpu_8gpus.txt
