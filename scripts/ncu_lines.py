#!/usr/bin/env python
"""Instruction / stall-sample share per CUDA source line of one kernel in an ncu report.
usage: ncu_lines.py report.ncu-rep kernel-regex [top-n]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep, kre = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
lines = src.splitlines()
start = next(k for k, ln in enumerate(lines) if '"Source"' in ln)
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
h = rows[0]
si, ie, sam = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
tot_i = tot_s = 0
ops = defaultdict(lambda: [0.0, 0.0])
for r in rows[1:]:
    try:
        i_, s_ = float(r[ie] or 0), float(r[sam] or 0)
    except Exception:
        continue
    op = r[si].split()[0] if r[si].split() else "?"
    if op.startswith("@"):
        op = r[si].split()[1]
    op = op.split(".")[0]
    ops[op][0] += i_
    ops[op][1] += s_
    tot_i += i_
    tot_s += s_
print(f"total warp instructions {tot_i:.0f}, samples {tot_s:.0f}")
for op, (i_, s_) in sorted(ops.items(), key=lambda x: -x[1][0])[:topn]:
    print(f"{i_ / tot_i * 100:5.1f}% inst  {s_ / max(tot_s, 1) * 100:5.1f}% smp  {op}")
