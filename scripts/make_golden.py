#!/usr/bin/env python
"""Regenerates gorender_b200/assets/ and tests/golden/ from the reference's model files (run in the
build container, where /root/reference exists; the GPU box only sees the
committed outputs).

  suzanne.npz, cube.npz   what gorender_b200.LoadObjFile produces from
                          models/suzanne.obj and models/cube.obj (+ cube.mtl,
                          textures-16.png): flattened arrays + premultiplied texels
  oracle_outputs.json     counts and SHA-256 of the oracle's framebuffers for the
                          pinned scenes of tests/scene_defs.py (a regression pin of
                          the oracle itself; the Go reference cannot run here)
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import gorender_b200 as g  # noqa: E402
from gorender_b200 import workloads  # noqa: E402

REF = "/root/reference/models"


def main():
    os.makedirs(workloads.GOLDEN_DIR, exist_ok=True)
    os.makedirs(workloads.ASSETS_DIR, exist_ok=True)
    if os.path.isdir(REF):
        for name in ("suzanne", "cube"):
            mesh = g.LoadMeshFile(os.path.join(REF, name + ".obj"), False)[0]
            workloads.save_mesh_fixture(os.path.join(workloads.ASSETS_DIR, name + ".npz"), mesh)
            print("wrote", name + ".npz", mesh.Vertices.shape, len(mesh.Faces))
    else:
        print("no /root/reference: keeping the committed mesh fixtures")

    from oracle_binding import Oracle
    import scene_defs

    orc = Oracle()
    out = {}
    for name, build in scene_defs.PINNED.items():
        sc = build()
        res = orc.draw(sc.renderer(None), sc.objects, sc.camera)
        out[name] = dict(
            width=sc.width, height=sc.height, tpf=res["tpf"], writes=res["writes"],
            covered=int((res["zbuffer"] > -1).sum()),
            pixels_sha256=hashlib.sha256(res["pixels"].tobytes()).hexdigest(),
            zbuffer_sha256=hashlib.sha256(res["zbuffer"].tobytes()).hexdigest())
        print(name, out[name])
    with open(os.path.join(workloads.GOLDEN_DIR, "oracle_outputs.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
