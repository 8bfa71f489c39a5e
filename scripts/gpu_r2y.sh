#!/bin/bash
# round 2, call Y: final state -- full GPU suite, smoke(), bench line (both arms)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2y_bench_line.json 2> gpurun_out/r2y_bench_err.log; echo "bench rc $?"; cut -c1-250 gpurun_out/r2y_bench_line.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2y_bench_line_reference_arm.json 2>> gpurun_out/r2y_bench_err.log; cut -c1-200 gpurun_out/r2y_bench_line_reference_arm.json
