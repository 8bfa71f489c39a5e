"""Summarise an `ncu --set full` capture (raw page exported with `ncu -i X.ncu-rep --page raw --csv`) for bench.py and the judge:
per kernel the launch duration, dram__bytes_read.sum / dram__bytes_write.sum, the fp64 / tensor pipe activity and the issue rate.
Writes profiles/ncu_traffic.json (looked up BY KEY by bench.py's roofline.traffic) when --json is given.
  python scripts/ncu_traffic.py profiles/r2_a_ncu_set_full_raw.csv --n 1000000 --json
"""
import argparse
import csv
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum"]
CLASS = {"k_spread3_mma_batch": "spread", "k_gather3_mma_batch": "gather"}


def load(path):
    rows = list(csv.reader(open(path)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    for h, i in list(col.items()):          # the --set full sections prefix some metrics ("FBSP.TriageCompute.dram__throughput...")
        m = re.sub(r"^[A-Z_]+\.[A-Za-z]+\.", "", h)
        col.setdefault(m, i)
    out = []
    for r in body:
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
        rec = {"kernel": name, "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]}
        for m in WANT:
            if m in col and r[col[m]] not in ("", "no data", "n/a"):
                try:
                    v = float(r[col[m]].replace(",", ""))
                except ValueError:
                    continue
                rec[m] = v * UNIT.get(units[col[m]], 1.0) if ("bytes" in m or "time" in m) else v
        out.append(rec)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--n", type=int, default=1000000)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--json", action="store_true")
    a = ap.parse_args()
    recs = load(a.csv)
    by = {}
    for r in recs:
        by.setdefault(r["kernel"], []).append(r)
    print(f"{'kernel':28s} {'n':>2s} {'us':>8s} {'rd MB':>8s} {'wr MB':>8s} {'dram%':>6s} {'fp64%':>6s} {'dmma%':>6s} {'issue%':>6s} {'warps%':>6s} {'regs':>4s}")
    summary = {}
    for k, rs in by.items():
        f = lambda m: sum(r.get(m, float("nan")) for r in rs) / len(rs)
        print(f"{k:28s} {len(rs):2d} {f('gpu__time_duration.sum') * 1e6:8.1f} {f('dram__bytes_read.sum') / 1e6:8.2f} {f('dram__bytes_write.sum') / 1e6:8.2f} "
              f"{f('dram__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} {f('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'):6.1f} "
              f"{f('sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active'):6.1f} {f('sm__issue_active.avg.pct_of_peak_sustained_active'):6.1f} "
              f"{f('sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f} {f('launch__registers_per_thread'):4.0f}")
        if k in CLASS:
            summary[CLASS[k]] = {"kernel": k, "n": a.n, "n_gpus": a.gpus, "launches_averaged": len(rs), "dram_bytes_read": f("dram__bytes_read.sum"),
                                 "dram_bytes_write": f("dram__bytes_write.sum"), "duration_us_under_ncu": f("gpu__time_duration.sum") * 1e6,
                                 "source": os.path.relpath(os.path.abspath(a.csv), ROOT)}
    if a.json:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as fh:
            json.dump(summary, fh, indent=1)
        print("wrote profiles/ncu_traffic.json")


if __name__ == "__main__":
    main()
