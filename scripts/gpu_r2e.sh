#!/bin/bash
# round 2, call E: FSAI gradient side / AFN logdet tests, ncu of the setup-side kernels and of the 2-D stream kernels, C5 on one GPU
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rA --durations=8 -k "fsai or afn_logdet or afn_k12" > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2e_pytest.log
grep -n "passed\|failed\|FAILED\|Error" gpurun_out/r2e_pytest.log | tail -12
timeout 600 python scripts/setup_profile.py 200000 1000 100 > gpurun_out/r2e_setup.log 2>&1; grep -v "Initializing\|Problem size" gpurun_out/r2e_setup.log | tail -6
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_knn|k_fsai_rows|k_dgemm_dmma|k_fps_step|k_csr_mv|k_sptrsv|k_gemv' -c 40 -f -o gpurun_out/r2e_setup_full python scripts/setup_profile.py 100000 1000 100 > gpurun_out/r2e_ncu1.log 2>&1
ncu -i gpurun_out/r2e_setup_full.ncu-rep --page raw --csv > gpurun_out/r2e_setup_full_raw.csv 2>/dev/null
python scripts/ncu_traffic.py gpurun_out/r2e_setup_full_raw.csv
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_spread_stream|k_gather_stream' -c 4 -f -o gpurun_out/r2e_stream2d_full python scripts/c4_matvec.py 4e6 4 2 > gpurun_out/r2e_ncu2.log 2>&1
ncu -i gpurun_out/r2e_stream2d_full.ncu-rep --page raw --csv > gpurun_out/r2e_stream2d_full_raw.csv 2>/dev/null
ncu -i gpurun_out/r2e_stream2d_full.ncu-rep --page details --csv > gpurun_out/r2e_stream2d_full_details.csv 2>/dev/null
python scripts/ncu_traffic.py gpurun_out/r2e_stream2d_full_raw.csv
timeout 600 python scripts/c5_slq.py 4e6 32 20 > gpurun_out/r2e_c5_1gpu.json 2> gpurun_out/r2e_c5.err; cut -c1-700 gpurun_out/r2e_c5_1gpu.json
