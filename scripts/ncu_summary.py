#!/usr/bin/env python
"""profiles/rNN_ncu_c3_batch64.json from an `ncu --set full` report holding one launch of the setup and of
the raster kernel (scripts/kernel_times.py c3).  usage: ncu_summary.py report.ncu-rep out.json [frames_per_launch]"""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 64
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
col = {n: i for i, n in enumerate(hdr)}


def f(r, name):
    return float(r[col[name]].replace(",", "")) if name in col and r[col[name]] not in ("", "n/a") else None


def to_bytes(r, name):
    v, unit = f(r, name), rows[1][col[name]].lower()
    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
    return None if v is None else v * scale


kernels = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    key = ("setup" if ("setup_kernel" in name or "setup_list_kernel" in name) else "raster" if "raster_kernel" in name
           else "mirror" if "mirror_update_kernel" in name else "reject" if "reject_kernel" in name else None)
    if key is None or key in kernels:
        continue
    rd, wr = to_bytes(r, "dram__bytes_read.sum"), to_bytes(r, "dram__bytes_write.sum")
    dur, dunit = f(r, "gpu__time_duration.sum"), rows[1][col["gpu__time_duration.sum"]]
    dur_us = dur * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(dunit, 1)
    kernels[key] = {
        "kernel": name, "frames_per_launch": frames, "duration_us_under_ncu": dur_us,
        "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_total": rd + wr,
        "dram_throughput_pct": f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "sm_throughput_pct": f(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        "issue_active_pct": f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers": f(r, "launch__registers_per_thread"),
        "warp_instructions": f(r, "smsp__inst_executed.sum"),
        "threads_per_instruction": f(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
        "fp32_pipe_fma_pct": f(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "l1_hit_pct": f(r, "l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": f(r, "lts__t_sector_hit_rate.pct"),
    }
what = sys.argv[4] if len(sys.argv) > 4 else "scripts/kernel_times.py c3 (C3 scene)"
json.dump({"source": f"ncu --set full --clock-control none, {what}, one batched draw of {frames} frames, B200", "kernels": kernels},
          open(out, "w"), indent=1)
print(json.dumps(kernels, indent=1))
