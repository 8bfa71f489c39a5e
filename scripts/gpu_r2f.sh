#!/bin/bash
# round 2, call F (2 GPUs): multi-GPU parity script (modes 1 and 2, gradient matvec and PCG under window sharding), full GPU test-suite,
# ncu of the setup-side kernels (without the 1000 launches of the farthest-point steps)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/check_multigpu.py 200000 > gpurun_out/r2f_check_multigpu.log 2>&1; echo "check rc $?"
grep "mode" gpurun_out/r2f_check_multigpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 scripts/check_multigpu_pcg.py > gpurun_out/r2f_check_multigpu_pcg.log 2>&1; echo "check pcg rc $?"
grep -i "iterations\|rel\|ok\|error" gpurun_out/r2f_check_multigpu_pcg.log | tail -6
CUDA_VISIBLE_DEVICES=0 timeout 1500 python -m pytest tests -m gpu -q -x --durations=6 > gpurun_out/r2f_pytest_full.log 2>&1; echo "pytest rc $?"
tail -12 gpurun_out/r2f_pytest_full.log
CUDA_VISIBLE_DEVICES=0 timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_knn|k_fsai_rows|k_dgemm_dmma|k_csr_mv|k_sptrsv|k_gemv|k_chol' -c 40 -f -o gpurun_out/r2f_setup_full python scripts/setup_profile.py 100000 1000 100 > gpurun_out/r2f_ncu1.log 2>&1
ncu -i gpurun_out/r2f_setup_full.ncu-rep --page raw --csv > gpurun_out/r2f_setup_full_raw.csv 2>/dev/null
python scripts/ncu_traffic.py gpurun_out/r2f_setup_full_raw.csv
