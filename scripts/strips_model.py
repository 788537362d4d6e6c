#!/usr/bin/env python
"""Per-rank kernel times of the sort-first strips of the C4 frame, measured on ONE GPU: for world = 1, 2, 4, 8 and both
partitions (equal tile rows / balanced by busy tiles), every rank's strip is drawn alone and its setup / raster
kernels are timed with CUDA events.  max over ranks = what the exchange-free part of a strip frame costs; the
difference to bench.py --mode strips is hand-off + launch + host time."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import gorender_b200 as g
from gorender_b200 import parallel, workloads

W, H = 3840, 2160
FPC = int(sys.argv[1]) if len(sys.argv) > 1 else 1      # frames per draw call
objs, cam = workloads.config_c4(100)
dev = g.default_device(0)
fb = g.FrameBuffer(W, H, FPC, dev)
r = g.Renderer(fb)
packed = np.ascontiguousarray(np.concatenate([r.pack_objects(objs, [cam])] * FPC, axis=0))
r.draw_packed(packed, 0)
weights = fb.tile_flags(0).sum(axis=1)
out = {}
print("frames per call:", FPC, "(times per frame)")
for world in (1, 2, 4, 8):
    for mode in ("equal", "busy"):
        rows = ([parallel.strip_rows(H, world, k) for k in range(world)] if mode == "equal"
                else parallel.balanced_strip_rows(weights, world, H))
        per = []
        for (y0, y1) in rows:
            if y1 <= y0:
                per.append({"rows": [y0, y1], "setup": 0, "raster": 0, "wall_us": 0})
                continue
            rr = None if world == 1 else (y0, y1)
            for _ in range(3):
                r.draw_packed(packed, 0, rows=rr, sync=False)
            dev.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                r.draw_packed(packed, 0, rows=rr, sync=False)
            dev.synchronize()
            wall = (time.perf_counter() - t0) / 20 * 1e6
            dev.set_kernel_timing(True)
            dev.kernel_times()
            for _ in range(10):
                r.draw_packed(packed, 0, rows=rr, sync=False)
            dev.synchronize()
            kt, _ = dev.kernel_times()
            dev.set_kernel_timing(False)
            per.append({"rows": [y0, y1], "setup": round(kt["setup"] / 10 * 1e3 / FPC, 1), "raster": round(kt["raster"] / 10 * 1e3 / FPC, 1),
                        "wall_us": round(wall / FPC, 1)})
        worst = max(p["setup"] + p["raster"] for p in per)
        out[f"{world}_{mode}"] = {"max_kernel_us": worst, "max_wall_us": max(p["wall_us"] for p in per), "ranks": per}
        print(world, mode, "max kernels", worst, "us; max wall", max(p["wall_us"] for p in per), "us;",
              [(p["setup"], p["raster"]) for p in per], flush=True)
# host cost of one asynchronous draw call of this scene (10 objects)
t0 = time.perf_counter()
for _ in range(200):
    r.draw_packed(packed, 0, rows=(1024, 1280), sync=False)
host = (time.perf_counter() - t0) / 200 * 1e6
dev.synchronize()
print("host time per async draw call (submission only):", round(host, 1), "us")
out["host_us_per_draw_call"] = host
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"strips_model_f{FPC}.json"), "w"), indent=1)
