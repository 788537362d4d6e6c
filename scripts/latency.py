#!/usr/bin/env python
"""One frame at a time through the literal drop-in call (grb_draw_present: Draw + host framebuffer + stats, one
synchronisation per frame) on the C3 scene — the reference's contract is one Draw per frame (main.go:201-208).

    python scripts/latency.py [--frames 2000] [--config c3|c1|c2b] [--no-depth]
"""
import argparse
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import gorender_b200 as g  # noqa: E402
from gorender_b200 import _cabi, geometry, workloads  # noqa: E402


def run(config="c3", frames=2000, depth=True, dev=None, width=1280, height=720):
    dev = dev or g.default_device(0)
    if config == "c1":
        objs, cam = workloads.config_c1()
    elif config == "c2b":
        objs, cam = workloads.config_c2("B")
    else:
        objs, cam = workloads.config_c3(100)
    fb = g.FrameBuffer(width, height, 1, dev)
    r = g.Renderer(fb)
    n = min(frames, 512)
    packed = np.ascontiguousarray(r.pack_objects(objs, [cam] * n, geometry.spin_rotations(n)))
    p = r.draw_params(None)
    stats = np.zeros(1, dtype=_cabi.STATS_DTYPE)
    color, zb = fb.mirror("Pixels"), (fb.mirror("ZBuffer") if depth else None)
    lib, h = dev.lib, dev.h
    stride = packed.strides[0]
    nobj = packed.shape[1]

    def one(i):
        rc = lib.grb_draw_present(h, fb.handle, 0, 1, C.c_void_p(packed.ctypes.data + (i % n) * stride), nobj, C.byref(p),
                                  color.h, 0, zb.h if zb is not None else None, 0, C.c_void_p(stats.ctypes.data))
        if rc:
            dev.check(rc)

    for i in range(20):
        one(i)
    t0 = time.perf_counter()
    for i in range(frames):
        one(i)
    sec = time.perf_counter() - t0
    nfaces = sum(len(o.Mesh.Faces) for o in objs)
    w, full = color.stats()
    return {"fps": frames / sec, "us_per_frame": sec / frames * 1e6, "mtri_s": frames / sec * nfaces / 1e6,
            "frames": frames, "graph_replays": dev.graph_replays(), "tiles_written_frac": w / max(full, 1),
            "checksum": int(fb.Pixels.sum()), "tpf": int(stats["tpf"][0]), "reads_back": "pixels+z" if depth else "pixels"}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=2000)
    ap.add_argument("--config", default="c3")
    ap.add_argument("--no-depth", action="store_true")
    a = ap.parse_args()
    print(run(a.config, a.frames, not a.no_depth))
