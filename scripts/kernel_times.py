#!/usr/bin/env python
"""Per-kernel CUDA-event times of one batched draw, for tuning (run on the GPU box).

    python scripts/kernel_times.py [c1|c2a|c2b|c3|c4] [--batch 64] [--reps 10]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import gorender_b200 as g  # noqa: E402
from gorender_b200 import geometry, workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", nargs="?", default="c3")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--n", type=int, default=100)
    a = ap.parse_args()
    w, h = 1280, 720
    if a.config == "c1":
        objs, cam = workloads.config_c1()
    elif a.config == "c2a":
        objs, cam = workloads.config_c2("A")
    elif a.config == "c2b":
        objs, cam = workloads.config_c2("B")
    elif a.config == "empty":
        objs, cam = [], geometry.default_camera()
    elif a.config == "c4":
        objs, cam = workloads.config_c4(a.n)
        w, h = 3840, 2160
    else:
        objs, cam = workloads.config_c3(a.n)
    dev = g.default_device(0)
    fb = g.FrameBuffer(w, h, a.batch, dev)
    r = g.Renderer(fb)
    packed = r.pack_objects(objs, [cam] * a.batch, geometry.spin_rotations(a.batch))
    for _ in range(3):
        r.draw_packed(packed, 0, sync=False)
    dev.synchronize()
    dev.set_kernel_timing(True)
    dev.kernel_times()
    for _ in range(a.reps):
        r.draw_packed(packed, 0, sync=False)
    dev.synchronize()
    t, _ = dev.kernel_times()
    tot = sum(t.values()) / a.reps
    nf = sum(len(o.Mesh.Faces) for o in objs)
    print(a.config, f"batch={a.batch}", " ".join(f"{k}={v / a.reps * 1e3:.1f}us" for k, v in t.items()),
          f"total={tot * 1e3:.1f}us  per-frame={tot / a.batch * 1e3:.2f}us  fps={a.batch / tot * 1e3:.0f} "
          f"Mtri/s={a.batch / tot * 1e3 * nf / 1e6:.0f}")


if __name__ == "__main__":
    main()
