#!/bin/bash
# round 2, call C: re-run of the changed tests, C4 1-D after the kernel fixes, full bench line with the NFFT-operator K12 extra
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA --durations=8 -k "poly1d or c4_style or reference_test3 or afn_k12 or afn_setup or fsai or nystrom" > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2c_pytest.log
grep -n "TEST3:\|TEST4:\|passed\|failed" gpurun_out/r2c_pytest.log | tail -5
timeout 600 python scripts/c4_matvec.py 4e6 64 1 > gpurun_out/r2c_c4.jsonl 2> gpurun_out/r2c_c4.err; echo "c4 rc $?"
cut -c1-600 gpurun_out/r2c_c4.jsonl
timeout 900 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"])
for r in d["pcg"]["runs"]:
    print(r["landmark_rank"], r["k12_form"], r["k12_storage_bits"], "setup", round(r["setup_s"], 2), "solve", round(r["solve_s"], 2), "iters", r["iters"], r["per_class_ms_per_iter"])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_poly1d_spread|k_poly1d_gather' -c 4 -f -o gpurun_out/r2c_poly_full python scripts/c4_matvec.py 4e6 16 1 > gpurun_out/r2c_ncu.log 2>&1
ncu -i gpurun_out/r2c_poly_full.ncu-rep --page raw --csv > gpurun_out/r2c_poly_full_raw.csv 2>/dev/null
ncu -i gpurun_out/r2c_poly_full.ncu-rep --page details --csv > gpurun_out/r2c_poly_full_details.csv 2>/dev/null
python scripts/ncu_traffic.py gpurun_out/r2c_poly_full_raw.csv
