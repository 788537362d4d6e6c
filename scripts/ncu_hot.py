#!/usr/bin/env python
"""Summarise an ncu report: headline metrics + hottest source lines by stall samples.
usage: ncu_hot.py report.ncu-rep [kernel-regex] [top-n]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else "."
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum",
        "smsp__inst_executed_op_shared_atom.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__grid_size", "lts__t_bytes.sum"]
for r in rows[2:]:
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w} = {r[i]} {units[i]}")
    stalls = [(hdr[i], float(r[i] or 0)) for i in range(len(hdr)) if hdr[i].startswith("smsp__pcsamp_warps_issue_stalled") and "not_issued" not in hdr[i]]
    tot = sum(v for _, v in stalls) or 1
    print("stalls:", ", ".join(f"{n.replace('smsp__pcsamp_warps_issue_stalled_', '')}={v / tot * 100:.0f}%" for n, v in sorted(stalls, key=lambda x: -x[1])[:8]))
    break

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}",
                      "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
lines = src.splitlines()
# find header row
for k, ln in enumerate(lines):
    if ln.startswith('"Address"') or ln.startswith('"#"') or '"Source"' in ln:
        start = k
        break
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
h = rows[0]
si = h.index("Source")
sam = h.index("# Samples") if "# Samples" in h else h.index("Warp Stall Sampling (All Samples)")
ie = h.index("Instructions Executed")
agg = []
for r in rows[1:]:
    try:
        agg.append((float(r[sam] or 0), float(r[ie] or 0), r[si]))
    except Exception:
        pass
tot = sum(a[0] for a in agg) or 1
print(f"--- top {topn} by stall samples (total {tot:.0f}); columns: samples% instr source")
for s_, i_, t in sorted(agg, key=lambda x: -x[0])[:topn]:
    print(f"{s_ / tot * 100:5.1f}% {i_:12.0f}  {t.strip()[:150]}")
