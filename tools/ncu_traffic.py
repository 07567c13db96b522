#!/usr/bin/env python3
"""Per-kernel DRAM traffic / pipe figures of an `ncu --set full` capture -> profiles/ncu_traffic.json (read by bench.py for
`roofline.traffic`, which therefore always names the capture it comes from):
   python tools/ncu_traffic.py gpurun_out/prof_r2x.ncu-rep profiles/r2x_ncu_full_summary.md 64 640 480"""
import csv
import json
import os
import subprocess
import sys

rep, source, frames, w, h = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H = rows[0]
U = rows[1]                                  # units row


def col(name):
    return H.index(name) if name in H else None


def num(r, i, scale_unit=True):
    if i is None or not r[i]:
        return None
    v = float(r[i].replace(",", ""))
    if scale_unit:
        u = U[i].lower()
        v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    return v


c_name, c_rd, c_wr = col("Kernel Name"), col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
c_alu, c_inst, c_us = col("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), col("smsp__inst_executed.sum"), col("gpu__time_duration.sum")
c_thr = col("smsp__thread_inst_executed_per_inst_executed.ratio")
acc = {}
for r in rows[2:]:
    k = r[c_name].split("(")[0].replace("<unnamed>::", "").replace("void ", "").split("<")[0]
    a = acc.setdefault(k, {"n": 0, "dram": 0.0, "alu": 0.0, "inst": 0.0, "thr": 0.0, "t": 0.0})
    a["n"] += 1
    a["dram"] += (num(r, c_rd) or 0) + (num(r, c_wr) or 0)
    a["alu"] += num(r, c_alu, False) or 0
    a["inst"] += num(r, c_inst, False) or 0
    a["thr"] += num(r, c_thr, False) or 0
    a["t"] += num(r, c_us, False) or 0
scale = [1.0]
for _ in range(1, 8):
    scale.append(scale[-1] / 1.2)
px = sum(round(w * s) * round(h * s) for s in scale)
out = {"source": source, "report": os.path.basename(rep), "frames_per_launch": frames, "width": w, "height": h, "kernels": {}}
for k, a in acc.items():
    n = a["n"]
    if k == "k_resize":                      # seven launches (one per level) make one pass
        n = max(n // 7, 1)
    e = {"launches_in_capture": a["n"], "dram_bytes_per_frame": a["dram"] / n / frames, "alu_pipe_pct": a["alu"] / a["n"],
         "warp_inst_per_frame": a["inst"] / n / frames}
    if k == "k_fast_nms":
        e["warp_inst_per_pixel"] = a["inst"] / n / frames / px
        e["thread_inst_per_pixel"] = e["warp_inst_per_pixel"] * (a["thr"] / a["n"])
    out["kernels"][k] = e
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
