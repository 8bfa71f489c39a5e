#!/bin/bash
# round 2, call X (8 GPUs): C4 with the final kernels -- staged 1-D polynomial path and table-free 2-D path, window-sharded
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29551 scripts/c4_matvec.py 4e6 64 1,2 2>>gpurun_out/r2x_err.log | grep '^{' > gpurun_out/r2x_c4_8gpu.jsonl; echo "c4 rc $?"
cut -c1-330 gpurun_out/r2x_c4_8gpu.jsonl; tail -2 gpurun_out/r2x_err.log
