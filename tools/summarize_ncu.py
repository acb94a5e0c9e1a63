#!/usr/bin/env python
"""Joins the ncu per-launch metrics of ONE forward (tools/gpu_profile.sh -> forward_metrics.csv) with the plan's op
table (op_table.json) and writes a markdown summary for profiles/: per launch duration, DRAM traffic vs algorithmic
bytes, tensor-pipe activity; per kernel-family shares."""
import csv
import io
import json
import sys
from collections import OrderedDict, defaultdict


def load_metrics(path):
    txt = open(path).read()
    txt = txt[txt.index('"ID"'):]
    by = OrderedDict()
    for row in csv.DictReader(io.StringIO(txt)):
        k = int(row["ID"])
        e = by.setdefault(k, {"kernel": row["Kernel Name"]})
        e[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
    return list(by.values())


def main(metrics_csv, table_json, out_md, title):
    launches = load_metrics(metrics_csv)
    table = json.load(open(table_json))
    ops = table["ops"]
    # expand ops to launches: the conditioning ops are 4 launches, everything else 1
    exp = []
    for o in ops:
        n = 4 if o["kind"] in ("cond", "dit_cond") else (2 if o["kind"] == "head" else 1)
        for i in range(n):
            exp.append(dict(o, part=i, parts=n))
    assert len(exp) == len(launches), (len(exp), len(launches))
    fam = defaultdict(lambda: dict(ns=0.0, dram=0.0, alg=0.0, flops=0.0, n=0, tens=0.0))
    lines = []
    total_ns = sum(l["gpu__time_duration.sum"] for l in launches)
    for o, l in zip(exp, launches):
        ns = l["gpu__time_duration.sum"]
        dram = l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"]
        tens = l.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0)
        kern = l["kernel"].split("(")[0].replace("void ", "")
        key = kern if o["kind"] in ("conv",) else o["kind"]
        f = fam[key]
        f["ns"] += ns; f["dram"] += dram; f["alg"] += o["bytes"] / o["parts"]; f["flops"] += o["flops"] / o["parts"]
        f["n"] += 1; f["tens"] += tens * ns
        tf = o["flops"] / ns / 1e3 if o["flops"] else 0.0
        lines.append(f"| {o['index']} | {o['name']} | {kern} | {ns/1e3:.1f} | {100*ns/total_ns:.2f} | {dram/1e6:.1f} | "
                     f"{o['bytes']/o['parts']/1e6:.1f} | {dram/max(o['bytes']/o['parts'],1):.2f} | {dram/ns:.0f} | {tf:.0f} | {tens:.1f} | "
                     f"{l['lts__t_bytes.sum']/ns:.0f} |")
    with open(out_md, "w") as fh:
        fh.write(f"# {title}\n\n")
        fh.write(f"Source: `ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
                 f"dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active...` over ONE forward of {table['images']} "
                 f"images (tools/gpu_profile.sh).  ncu serialises launches and flushes caches between them: compare SHARES, not "
                 f"absolute times, with bench.py's CUDA-event numbers.\n\n")
        fh.write(f"Total: {len(launches)} launches, {total_ns/1e6:.3f} ms under ncu.\n\n## Per kernel family\n\n")
        fh.write("| family | launches | ms | share % | DRAM GB | algorithmic GB | DRAM/alg | DRAM GB/s | TFLOP/s | tensor pipe active % (time-weighted) |\n|---|---|---|---|---|---|---|---|---|---|\n")
        for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["ns"]):
            fh.write(f"| {k} | {f['n']} | {f['ns']/1e6:.3f} | {100*f['ns']/total_ns:.1f} | {f['dram']/1e9:.2f} | {f['alg']/1e9:.2f} | "
                     f"{f['dram']/max(f['alg'],1):.2f} | {f['dram']/f['ns']:.0f} | {f['flops']/f['ns']/1e3:.0f} | {f['tens']/f['ns']:.1f} |\n")
        fh.write("\n## Per launch\n\n| op | layer | kernel | us | share % | DRAM MB | algorithmic MB | DRAM/alg | DRAM GB/s | TFLOP/s | tensor % | L2 GB/s |\n|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        fh.write("\n".join(lines) + "\n")
    # machine-readable DRAM traffic per launch of the dominant kernel family (bench.py's roofline.traffic)
    conv = [f for k, f in fam.items() if k.startswith("conv_umma_kernel")]
    if conv:
        n = sum(f["n"] for f in conv)
        js = dict(images_per_launch=table["images"], kernel="conv_umma_kernel", launches=n,
                  dram_bytes_per_launch=sum(f["dram"] for f in conv) / n,
                  algorithmic_bytes_per_launch=sum(f["alg"] for f in conv) / n,
                  tensor_pipe_active_pct=sum(f["tens"] for f in conv) / sum(f["ns"] for f in conv),
                  source=out_md.split("/")[-1])
        json.dump(js, open(out_md.replace(".md", ".json"), "w"), indent=1)
    print("wrote", out_md)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "ncu per-launch summary of one forward")
