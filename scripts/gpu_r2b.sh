#!/bin/bash
# round 2, call B: new / changed GPU tests, C4 at full size on one GPU (1-D polynomial path, 2-D table path), ncu of the C4 kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA --durations=10 -k "poly1d or c4_style or single_window or afn_setup_matches_reference or rank_estimate_matches or c3_pcg or reference_test3 or reference_test2 or fullsize" > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2b_pytest.log
tail -4 gpurun_out/r2b_pytest.log
timeout 900 python scripts/c4_matvec.py 4e6 64 1,2 > gpurun_out/r2b_c4.jsonl 2> gpurun_out/r2b_c4.err; echo "c4 rc $?"
cat gpurun_out/r2b_c4.jsonl | cut -c1-700
NFFT4GP_B200_POLY1D=0 timeout 600 python scripts/c4_matvec.py 4e6 64 1 > gpurun_out/r2b_c4_tables.jsonl 2> /dev/null
cut -c1-400 gpurun_out/r2b_c4_tables.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_poly1d|k_spread_stream|k_gather_stream' -c 12 -f -o gpurun_out/r2b_c4_full python scripts/c4_matvec.py 4e6 16 1,2 > gpurun_out/r2b_ncu.log 2>&1
ncu -i gpurun_out/r2b_c4_full.ncu-rep --page raw --csv > gpurun_out/r2b_c4_full_raw.csv 2>/dev/null
ncu -i gpurun_out/r2b_c4_full.ncu-rep --page details --csv > gpurun_out/r2b_c4_full_details.csv 2>/dev/null
python scripts/ncu_traffic.py gpurun_out/r2b_c4_full_raw.csv
