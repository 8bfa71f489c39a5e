#!/bin/bash
# round 2, call A: GPU test-suite, full bench line (+ reference arm), ncu launch list and --set full capture of the hot kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests -m gpu -q -rA --durations=15 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2a_pytest.log
timeout 600 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc $?"
timeout 300 python bench.py --impl reference --steps 5 > gpurun_out/r2a_bench_ref.json 2> /dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-pcg > gpurun_out/r2a_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_(spread|gather)3_mma_batch|k_fft' -c 14 -f -o gpurun_out/r2a_full python bench.py --steps 2 --warmup 3 --no-cpu --no-pcg > gpurun_out/r2a_ncu2.log 2>&1
ncu -i gpurun_out/r2a_full.ncu-rep --page raw --csv > gpurun_out/r2a_full_raw.csv 2>/dev/null
ncu -i gpurun_out/r2a_full.ncu-rep --page details --csv > gpurun_out/r2a_full_details.csv 2>/dev/null
tail -5 gpurun_out/r2a_pytest.log
