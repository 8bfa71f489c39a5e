#!/bin/bash
# round 2, call K: cp.async double-buffered (staged) 1-D polynomial spread / gather against the direct-load kernels (A/B by env), parity tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for a in 0 1; do
   export NFFT4GP_B200_POLY1D_ASYNC=$a
   echo "== async=$a"; timeout 300 python scripts/c4_matvec.py 4e6 64 1 2>gpurun_out/r2k_err_$a.log | grep '^{' | tee gpurun_out/r2k_c4_1d_async$a.json | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['ms_per_matvec'], d['per_class_us_rank0'], d['parity_rel_l2_vs_oracle_on_first_20000_points'])"
   tail -3 gpurun_out/r2k_err_$a.log
done
unset NFFT4GP_B200_POLY1D_ASYNC
timeout 900 python -m pytest tests -m gpu -x -q -k "poly1d or c4 or single_window or quirk or matvec" 2>&1 | tail -5
timeout 200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_matvec.py -m gpu -x -q -k "poly1d" 2>&1 | tail -6
