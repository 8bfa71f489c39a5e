"""Summarise `ncu --page source --csv` output: top stalled instructions and opcode mix per kernel section."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        sections.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
for sec in sections:
    if want not in sec["name"]:
        continue
    hdr, data = sec["hdr"], sec["data"]
    ci = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[ci["# Samples"]]) for r in data)
    print("==", sec["name"][:60], "samples", tot, "instructions", len(data))
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = collections.Counter()
    for r in data:
        for h in stalls:
            agg[h[6:]] += int(r[ci[h]])
    print("   stall mix:", [(k, round(100 * v / tot, 1)) for k, v in agg.most_common(8)])
    for r in sorted(data, key=lambda r: -int(r[ci["# Samples"]]))[:18]:
        s = int(r[ci["# Samples"]])
        st = sorted(((int(r[ci[h]]), h) for h in stalls), reverse=True)[:3]
        print(f"{s:6d} {100*s/tot:5.1f}% {r[ci['Source']].strip()[:52]:52s} exec={r[ci['Instructions Executed']]:>8s} {[(h[6:], v) for v, h in st]}")
    op, sm = collections.Counter(), collections.Counter()
    for r in data:
        t = r[ci["Source"]].strip().split()
        o = t[1] if t[0].startswith("@") else t[0]
        op[o.split(".")[0]] += int(r[ci["Instructions Executed"]])
        sm[o.split(".")[0]] += int(r[ci["# Samples"]])
    te = sum(op.values())
    for o, c in op.most_common(12):
        print(f"   {o:10s} exec {c:10d} {100*c/te:5.1f}%  samples {100*sm[o]/tot:5.1f}%")
