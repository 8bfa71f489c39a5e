#!/bin/bash
# round 2, call T: table-free 2-D kernels with the squares hoisted to the staging step and the next step prefetched
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_matvec.py -m gpu -x -q -k "poly2d" 2>&1 | tail -4
echo "== POLY2D=auto"; timeout 400 python scripts/c4_matvec.py 4e6 64 2 2>gpurun_out/r2t_err.log | grep '^{' | tee gpurun_out/r2t_c4_2d_poly2d.json | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['ms_per_matvec'], d['per_class_us_rank0'], d['launches_per_matvec_rank0'], d['parity_rel_l2_vs_oracle_on_first_20000_points'], d['path'])"
tail -2 gpurun_out/r2t_err.log
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:k_(spread|gather)2_poly" -c 4 -f -o gpurun_out/r2t_poly2d_full python scripts/c4_matvec.py 4e6 4 2 > gpurun_out/r2t_ncu.log 2>&1
ncu -i gpurun_out/r2t_poly2d_full.ncu-rep --page raw --csv > gpurun_out/r2t_ncu_set_full_poly2d_raw.csv 2>/dev/null
ncu -i gpurun_out/r2t_poly2d_full.ncu-rep --page source --csv > gpurun_out/r2t_ncu_source_poly2d.csv 2>/dev/null
python scripts/ncu_traffic.py gpurun_out/r2t_ncu_set_full_poly2d_raw.csv | tail -n +2
python scripts/ncu_source_summary.py gpurun_out/r2t_ncu_source_poly2d.csv 2>&1 | cut -c1-200 | grep "stall mix\|^==" | head -8
