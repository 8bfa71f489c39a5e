#!/bin/bash
# round 2, call G: device-resident SLQ probe loop (tests + C5 timing), ncu --set full of the setup-side / preconditioner kernels one family at a time
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "lanczos or slq or reference_test or predict" > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r2g_pytest.log
timeout 600 python scripts/c5_slq.py 4e6 32 20 > gpurun_out/r2g_c5_1gpu.json 2> gpurun_out/r2g_c5.err; cut -c1-500 gpurun_out/r2g_c5_1gpu.json
for fam in k_knn k_fsai_rows k_csr_mv k_sptrsv k_gemv k_dgemm_dmma k_nys; do
   timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$fam" -c 3 -f -o gpurun_out/r2g_${fam}_full python scripts/setup_profile.py 100000 1000 100 > gpurun_out/r2g_ncu_$fam.log 2>&1
   ncu -i gpurun_out/r2g_${fam}_full.ncu-rep --page raw --csv > gpurun_out/r2g_${fam}_full_raw.csv 2>/dev/null
   python scripts/ncu_traffic.py gpurun_out/r2g_${fam}_full_raw.csv | tail -n +2
done
