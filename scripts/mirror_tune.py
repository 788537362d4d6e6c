#!/usr/bin/env python
"""Host-mirror update alone and beside rendering, for one setting of GRB_MIRROR_BLOCKS_PER_SM (C3, 64-frame batches)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import gorender_b200 as g
from gorender_b200 import _cabi, geometry, workloads
from gorender_b200.renderer import Mirror

B = 64
objs, cam = workloads.config_c3(100)
devs = [g.Device(0) for _ in range(2)]
fbs = [g.FrameBuffer(1280, 720, B, d) for d in devs]
rs = [g.Renderer(fb) for fb in fbs]
packed = [np.ascontiguousarray(rs[k].pack_objects(objs, [cam] * B, geometry.spin_rotations(B, start=k * B))) for k in range(2)]
mc = [Mirror(devs[k], 1280, 720, B, _cabi.GRB_PLANE_COLOR) for k in range(2)]
mz = [Mirror(devs[k], 1280, 720, B, _cabi.GRB_PLANE_DEPTH) for k in range(2)]
for k in range(2):
    rs[k].draw_packed(packed[k], 0, sync=False)
    fbs[k].update_mirrors_async(0, B, mc[k], mz[k])
    devs[k].synchronize()
w0 = mc[0].stats()[0] + mz[0].stats()[0]
t0 = time.perf_counter()
N = 20
for _ in range(N):
    fbs[0].update_mirrors_async(0, B, mc[0], mz[0])
devs[0].synchronize()
sec = (time.perf_counter() - t0) / N
tiles = (mc[0].stats()[0] + mz[0].stats()[0] - w0) / N
print(f"blocks/SM={os.environ.get('GRB_MIRROR_BLOCKS_PER_SM', 'default')}: mirror alone {sec * 1e3:.3f} ms per {B} frames, "
      f"{tiles * 4096 / sec / 1e9:.1f} GB/s", end="; ")
t0 = time.perf_counter()
N = 60
for i in range(N):
    k = i & 1
    rs[k].draw_packed(packed[k], 0, sync=False)
    fbs[k].update_mirrors_async(0, B, mc[k], mz[k])
for d in devs:
    d.synchronize()
sec = (time.perf_counter() - t0) / N
print(f"draw + mirror, two contexts alternating: {sec * 1e3:.3f} ms per batch = {B / sec:.0f} FPS = {B / sec * 0.2:.0f} Mtri/s")
