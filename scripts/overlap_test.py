#!/usr/bin/env python
"""Throughput with S concurrent contexts (streams), each issuing batched draws of the C3 scene."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import gorender_b200 as g
from gorender_b200 import geometry, workloads

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
objs, cam = workloads.config_c3(100)
for S in (1, 2, 3, 4):
    devs = [g.Device(0) for _ in range(S)]
    fbs = [g.FrameBuffer(1280, 720, B, d) for d in devs]
    rs = [g.Renderer(fb) for fb in fbs]
    packed = [r.pack_objects(objs, [cam] * B, geometry.spin_rotations(B, start=7 * i)) for i, r in enumerate(rs)]
    for r, p in zip(rs, packed):
        for _ in range(3):
            r.draw_packed(p, 0, sync=False)
    for d in devs:
        d.synchronize()
    N = 48
    t0 = time.perf_counter()
    for i in range(N):
        k = i % S
        rs[k].draw_packed(packed[k], 0, sync=False)
    for d in devs:
        d.synchronize()
    dt = time.perf_counter() - t0
    print(f"streams={S} batch={B}: {N * B / dt:.0f} fps  {N * B / dt * 0.2:.0f} Mtri/s  ({dt / N * 1e6:.0f} us per batch)")
    del rs, fbs, devs
