#!/bin/bash
# round 2, call I: loads-per-run batching of the 1-D polynomial kernels (A/B: 2 / 4 / 8 points per lane), C5 with reused Krylov bases
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for t in main pu2 pu8; do
   if [ "$t" == "main" ]; then unset NFFT4GP_B200_LIB; else export NFFT4GP_B200_LIB=$PWD/ab/libnfft4gp_b200_$t.so; fi
   echo "== $t"; timeout 300 python scripts/c4_matvec.py 4e6 64 1 2>/dev/null | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['ms_per_matvec'], d['per_class_us_rank0'], d['parity_rel_l2_vs_oracle_on_first_20000_points'])"
done
unset NFFT4GP_B200_LIB
timeout 600 python scripts/c5_slq.py 4e6 32 20 2>/dev/null | grep '^{' > gpurun_out/r2i_c5_1gpu.json; python -c "import json; d=json.load(open('gpurun_out/r2i_c5_1gpu.json')); print('C5 1 GPU', d['seconds'], d['ms_per_probe'], d['parity_sharded_vs_single_process'])"
NFFT4GP_B200_SLQ_TIMING=1 timeout 300 python scripts/c5_profile_probe.py 2>&1 | grep "slq timing\|probes wall" | tail -12
