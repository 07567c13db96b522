#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into a markdown table: python tools/ncu_summary.py rep.ncu-rep > profiles/x.md"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H = rows[0]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU %"), ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("dram__bytes_read.sum", "dram rd MB"),
        ("dram__bytes_write.sum", "dram wr MB"), ("dram__bytes_read.sum.per_second", "rd GB/s"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("smsp__inst_executed.sum", "warp inst"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst")]
idx = [(H.index(a), b) for a, b in want if a in H]
print("| " + " | ".join(b for _, b in idx) + " | top stalls (warps per issue) |")
print("|" + "---|" * (len(idx) + 1))
for r in rows[2:]:
    out = []
    for i, b in idx:
        v = r[i]
        if b == "kernel":
            v = v.split("(")[0].replace("<unnamed>::", "")
        else:
            try:
                v = "%.1f" % float(v.replace(",", ""))
            except ValueError:
                pass
        out.append(v)
    st = sorted([(float(r[i]), h.split("issue_stalled_")[1].split("_per_")[0]) for i, h in enumerate(H)
                 if "average_warps_issue_stalled" in h and "per_issue_active" in h and r[i]], reverse=True)[:3]
    print("| " + " | ".join(out) + " | " + ", ".join("%s %.1f" % (n, v) for v, n in st) + " |")
