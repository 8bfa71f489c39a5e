#!/bin/bash
# round 2, call S: table-free 2-D windows (nfft2d_poly.cu: bivariate polynomial per cell, Vandermonde GEMMs on the fp64 tensor cores):
# parity, C4 with 32 two-D windows against the tap-table tensor-core kernels, ncu --set full of the new kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_matvec.py -m gpu -x -q -k "poly2d" 2>&1 | tail -15
for v in 0 auto; do
   if [ "$v" == "0" ]; then export NFFT4GP_B200_POLY2D=0; else unset NFFT4GP_B200_POLY2D; fi
   echo "== POLY2D=$v"; timeout 400 python scripts/c4_matvec.py 4e6 64 2 2>gpurun_out/r2s_err_$v.log | grep '^{' | tee gpurun_out/r2s_c4_2d_poly2d_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['ms_per_matvec'], d['per_class_us_rank0'], d['launches_per_matvec_rank0'], d['parity_rel_l2_vs_oracle_on_first_20000_points'], d['path'])"
   tail -2 gpurun_out/r2s_err_$v.log
done
unset NFFT4GP_B200_POLY2D
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:k_(spread|gather)2_poly|k_poly2d" -c 8 -f -o gpurun_out/r2s_poly2d_full python scripts/c4_matvec.py 4e6 4 2 > gpurun_out/r2s_ncu.log 2>&1
ncu -i gpurun_out/r2s_poly2d_full.ncu-rep --page raw --csv > gpurun_out/r2s_ncu_set_full_poly2d_raw.csv 2>/dev/null
ncu -i gpurun_out/r2s_poly2d_full.ncu-rep --page source --csv > gpurun_out/r2s_ncu_source_poly2d.csv 2>/dev/null
python scripts/ncu_traffic.py gpurun_out/r2s_ncu_set_full_poly2d_raw.csv | tail -n +2
python scripts/ncu_source_summary.py gpurun_out/r2s_ncu_source_poly2d.csv 2>&1 | cut -c1-200 | grep "stall mix\|^==\|^   [A-Z]" | head -60
