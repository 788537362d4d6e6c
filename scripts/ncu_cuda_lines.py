#!/usr/bin/env python
"""Warp instructions and stall samples per CUDA source line of one kernel in an ncu report
(needs --import-source on and -lineinfo).  usage: ncu_cuda_lines.py report.ncu-rep kernel-regex [top-n]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep, kre = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 50
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
agg = defaultdict(lambda: [0.0, 0.0, ""])
fname, cur, h = "?", None, None
for r in rows:
    if len(r) >= 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 4 and r[0] == "Line No":
        h = r
        ie, sam = h.index("Instructions Executed"), h.index("# Samples")
        continue
    if h is None or len(r) <= max(ie, sam):
        continue
    if r[0] != "":
        cur = (fname, int(r[0]))
        agg[cur][2] = r[1]
        continue
    if cur is None:
        continue
    try:
        agg[cur][0] += float(r[ie] or 0)
        agg[cur][1] += float(r[sam] or 0)
    except ValueError:
        pass
ti = sum(a[0] for a in agg.values()) or 1
ts = sum(a[1] for a in agg.values()) or 1
print(f"total warp instructions {ti:.0f}, samples {ts:.0f}")
for (f, n), (i_, s_, t) in sorted(agg.items(), key=lambda x: -x[1][0])[:topn]:
    print(f"{i_ / ti * 100:5.1f}% inst {s_ / ts * 100:5.1f}% smp  {f}:{n:<4} {t.strip()[:110]}")
