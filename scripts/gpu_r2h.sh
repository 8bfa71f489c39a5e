#!/bin/bash
# round 2, call H: 2-D tensor-core kernels (parity tests, C4 2-D timing, ncu), where a C5 probe spends its time
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "matvec_matches_oracle or c4_style or realdata or gradient or poly1d or single_window" > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/r2h_pytest.log
timeout 600 python scripts/c4_matvec.py 4e6 64 2 > gpurun_out/r2h_c4_2d.jsonl 2> gpurun_out/r2h_c4.err; cut -c1-700 gpurun_out/r2h_c4_2d.jsonl
timeout 300 python scripts/c5_profile_probe.py 2>&1 | grep -v "Rel. residual\|Initializing\|Problem size" | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_spread2_mma|k_gather2_mma' -c 4 -f -o gpurun_out/r2h_mma2d_full python scripts/c4_matvec.py 4e6 4 2 > gpurun_out/r2h_ncu.log 2>&1
ncu -i gpurun_out/r2h_mma2d_full.ncu-rep --page raw --csv > gpurun_out/r2h_mma2d_full_raw.csv 2>/dev/null
ncu -i gpurun_out/r2h_mma2d_full.ncu-rep --page details --csv > gpurun_out/r2h_mma2d_full_details.csv 2>/dev/null
python scripts/ncu_traffic.py gpurun_out/r2h_mma2d_full_raw.csv
