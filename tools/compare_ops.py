#!/usr/bin/env python
"""compare two per-op timing tables written by bench.py --ops-out"""
import json, sys
a = json.load(open(sys.argv[1])); b = json.load(open(sys.argv[2]))
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.03
fa, fb = {}, {}
for o, n in zip(a["ops"], b["ops"]):
    fa[o["kind"]] = fa.get(o["kind"], 0) + o["ms"]; fb[n["kind"]] = fb.get(n["kind"], 0) + n["ms"]
    if abs(o["ms"] - n["ms"]) / max(o["ms"], 1e-9) > thr and o["ms"] > 0.02:
        tf = n["flops"] / n["ms"] / 1e9 if n["flops"] else 0
        print(f"  {o['name']:30s} {o['kind']:10s} {o['ms']:.4f} -> {n['ms']:.4f}  ({100*(n['ms']/o['ms']-1):+.0f}%)  {tf:7.0f} TF {n['bytes']/n['ms']/1e6:7.0f} GB/s")
print("families:", {k: (round(fa[k], 3), round(fb[k], 3)) for k in fa})
print("total", round(sum(fa.values()), 3), "->", round(sum(fb.values()), 3))
