#!/usr/bin/env python
"""Markdown table of an `ncu --set full` capture (read with `ncu -i X.ncu-rep --page raw --csv > raw.csv`): one row per launch with
duration, DRAM bytes, pipe activity and the warp-stall mix.  usage: summarize_ncu_full.py raw.csv out.md "title" [name1,name2,...]"""
import csv
import sys


def main(raw_csv, out_md, title, names):
    rows = list(csv.reader(open(raw_csv)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def g(r, k, scale=1.0, fmt="{:.1f}"):
        if k not in ix or r[ix[k]] in ("", "n/a"):
            return ""
        return fmt.format(float(r[ix[k]].replace(",", "")) * scale)

    tu = units[ix["gpu__time_duration.sum"]] if "gpu__time_duration.sum" in ix else "?"
    cols = [(f"time {tu}", "gpu__time_duration.sum", 1.0, "{:.3f}"), ("DRAM rd", "dram__bytes_read.sum", 1.0, "{:.3f}"),
            ("DRAM wr", "dram__bytes_write.sum", 1.0, "{:.3f}"),
            ("tensor pipe %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1.0, "{:.1f}"),
            ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0, "{:.1f}"),
            ("XU %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1.0, "{:.1f}"),
            ("FMA %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1.0, "{:.1f}"),
            ("ALU %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1.0, "{:.1f}"),
            ("smem LSU %", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", 1.0, "{:.1f}"),
            ("warp inst", "smsp__inst_executed.sum", 1e-6, "{:.1f} M"), ("regs", "launch__registers_per_thread", 1.0, "{:.0f}"),
            ("stall long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 1.0, "{:.2f}"),
            ("stall wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", 1.0, "{:.2f}"),
            ("stall barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", 1.0, "{:.2f}"),
            ("stall mio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", 1.0, "{:.2f}"),
            ("stall math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", 1.0, "{:.2f}")]
    out = [f"# {title}", "", "Read with `ncu -i ... --page raw --csv`; times are cold-cache and serialised (compare shares and pipe "
           "percentages, not absolute times).  DRAM columns in the unit ncu printed (" +
           (units[ix["dram__bytes_read.sum"]] if "dram__bytes_read.sum" in ix else "?") + ").", "",
           "| launch | " + " | ".join(c[0] for c in cols) + " |", "|---|" + "---|" * len(cols)]
    for i, r in enumerate(data):
        k = r[ix["Kernel Name"]]
        k = k[:k.index("(")] if "(" in k else k
        label = f"`{k.replace('void ', '')}`" + (f" {names[i]}" if i < len(names) else "")
        out.append("| " + label + " | " + " | ".join(g(r, c[1], c[2], c[3]) for c in cols) + " |")
    open(out_md, "w").write("\n".join(out) + "\n")
    print("\n".join(out[4:]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4].split(",") if len(sys.argv) > 4 else [])
