#!/bin/bash
# round 2, call P: staged 1-D polynomial kernels with L2 eviction hints on the bulk copies and 128-bit coefficient loads; Estrin evaluation as A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_matvec.py -m gpu -x -q -k "poly1d or c4" 2>&1 | tail -3 || exit 1
run() { # tag lib async
   if [ "$2" == "main" ]; then unset NFFT4GP_B200_LIB; else export NFFT4GP_B200_LIB=$PWD/ab/libnfft4gp_b200_$2.so; fi
   export NFFT4GP_B200_POLY1D_ASYNC=$3
   echo "== $1"; timeout 300 python scripts/c4_matvec.py 4e6 64 1 2>gpurun_out/r2p_err_$1.log | grep '^{' | tee gpurun_out/r2p_c4_1d_$1.json | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['ms_per_matvec'], d['per_class_us_rank0'], d['parity_rel_l2_vs_oracle_on_first_20000_points'])"
   tail -2 gpurun_out/r2p_err_$1.log
}

run staged_w32u2 main 1
run staged_w32u4 w32u4 1
run staged_estrin estrin 1
run staged_estrinu4 estrinu4 1
unset NFFT4GP_B200_LIB NFFT4GP_B200_POLY1D_ASYNC
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:k_poly1d_(spread|gather)_async" -c 2 -f -o gpurun_out/r2p_poly1d_async_full python scripts/c4_matvec.py 4e6 16 1 > gpurun_out/r2p_ncu.log 2>&1
ncu -i gpurun_out/r2p_poly1d_async_full.ncu-rep --page raw --csv > gpurun_out/r2p_ncu_set_full_poly1d_async_raw.csv 2>/dev/null
ncu -i gpurun_out/r2p_poly1d_async_full.ncu-rep --page source --csv > gpurun_out/r2p_ncu_source_poly1d_async.csv 2>/dev/null
python scripts/ncu_traffic.py gpurun_out/r2p_ncu_set_full_poly1d_async_raw.csv | tail -n +2
python scripts/ncu_source_summary.py gpurun_out/r2p_ncu_source_poly1d_async.csv 2>&1 | cut -c1-200 | grep "stall mix\|^=="
