#!/usr/bin/env python
"""top stall-sample SASS instructions of one kernel from an ncu report:
    ncu -i rep --page source --csv --print-source sass --launch-skip K --launch-count 1 > x.csv; python tools/ncu_hot.py x.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "Address"][0]
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != "Address"]
tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print(rows[0][1][:120], "total samples", tot)
agg = {s: sum(int(r[idx[s]] or 0) for r in data) for s in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
top = sorted(data, key=lambda r: -int(r[idx["# Samples"]] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for r in top:
    n = int(r[idx["# Samples"]] or 0)
    why = sorted(((int(r[idx[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print(f"{100*n/tot:5.1f}%  {r[idx['Source']][:90]:90s} {why}")
