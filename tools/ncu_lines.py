#!/usr/bin/env python3
"""Per-source-line instruction and stall-sample shares of one kernel from an .ncu-rep (needs --import-source on):
   python tools/ncu_lines.py rep.ncu-rep k_describe [top]"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "-k", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
H = rows[hi]
iL, iS, iI, iSamp = 0, 1, H.index("Instructions Executed"), H.index("# Samples")
inst = collections.Counter(); samp = collections.Counter(); src = {}
first = True
for r in rows[hi + 1:]:
    if len(r) < len(H):
        continue
    if r[0] == "Line No":          # a second launch of the same kernel: keep the first only
        break
    try:
        ln = int(r[iL])
    except ValueError:
        continue
    src[ln] = r[iS]
    num = lambda v: int(v) if v.strip().isdigit() else 0
    inst[ln] += num(r[iI]); samp[ln] += num(r[iSamp])
ti, ts = sum(inst.values()), sum(samp.values())
print("kernel %s: %d warp instructions, %d samples" % (kern, ti, ts))
for ln, n in inst.most_common(top):
    print("%5d  inst %5.1f %%  samples %5.1f %%  %s" % (ln, 100.0 * n / ti, 100.0 * samp[ln] / max(ts, 1), src[ln].strip()[:150]))
