#!/bin/bash
# round 2, call Q: full GPU suite, smoke(), bench line (both arms) with the staged 1-D kernels in the library; Estrin A/B of the 1-D gather
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2q_bench_line.json 2> gpurun_out/r2q_bench_err.log; echo "bench rc $?"; cut -c1-600 gpurun_out/r2q_bench_line.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2q_bench_line_reference_arm.json 2>> gpurun_out/r2q_bench_err.log; cut -c1-300 gpurun_out/r2q_bench_line_reference_arm.json
for t in main estrin; do
   if [ "$t" == "main" ]; then unset NFFT4GP_B200_LIB; else export NFFT4GP_B200_LIB=$PWD/ab/libnfft4gp_b200_$t.so; fi
   echo "== $t"; timeout 300 python scripts/c4_matvec.py 4e6 64 1 2>/dev/null | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['ms_per_matvec'], d['per_class_us_rank0'], d['parity_rel_l2_vs_oracle_on_first_20000_points'])"
done
