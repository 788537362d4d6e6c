#!/usr/bin/env python
"""Writes the pinned parity scenes (tests/scene_defs.py) as raw arrays for go/parity/parity_dump.go, the harness that
renders them with the unmodified Go reference.  No OBJ / PNG text goes in between: the Go side receives exactly the
float32 / int32 / RGBA8 values the oracle and the CUDA path are fed, and builds its Mesh with the reference's own
NewMesh.

    python scripts/export_go_scenes.py OUTDIR [scene ...]
"""
import json
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def write_mesh(path, mesh):
    F = mesh.Faces
    texs = list(F.Textures)
    nf = len(F)
    with open(path, "wb") as f:
        f.write(struct.pack("<4i", len(mesh.Vertices), len(mesh.VertexNormals), nf, len(texs)))
        f.write(np.ascontiguousarray(mesh.Vertices, "<f4").tobytes())
        f.write(np.ascontiguousarray(mesh.VertexNormals, "<f4").tobytes())
        f.write(np.ascontiguousarray(F.VertexIndices, "<i4").tobytes())
        f.write(np.ascontiguousarray(F.NormalIndices, "<i4").tobytes())
        f.write(np.ascontiguousarray(F.UVs, "<f4").tobytes())
        f.write(np.ascontiguousarray(F.TextureIndex, "<i4").tobytes())
        for t in texs:
            f.write(struct.pack("<3if4B", int(t.typ), int(t.width), int(t.height), float(t.scale), *[int(c) for c in t.color]))
            if t.pixels is not None:
                f.write(np.ascontiguousarray(t.pixels, np.uint8).tobytes())


def main():
    import scene_defs

    out = sys.argv[1]
    names = sys.argv[2:] or list(scene_defs.PINNED)
    os.makedirs(out, exist_ok=True)
    scenes, mesh_files = [], {}
    for name in names:
        sc = scene_defs.PINNED[name]()
        r = sc.renderer(None)
        if r.AffineTextures:
            print("skipped", name, "(affine texture mapping has no code path in the reference)")
            continue
        meshes, objs = [], []
        for o in sc.objects:
            key = id(o.Mesh)
            if key not in mesh_files:
                mesh_files[key] = (f"mesh{len(mesh_files):03d}.bin", o.Mesh)
                write_mesh(os.path.join(out, mesh_files[key][0]), o.Mesh)
            fn = mesh_files[key][0]
            if fn not in meshes:
                meshes.append(fn)
            objs.append({"mesh": meshes.index(fn), "translation": [float(x) for x in o.Translation],
                         "rotation": [float(x) for x in o.Rotation], "scale": [float(x) for x in o.Scale]})
        opts = {k: bool(getattr(r, k)) for k in ("FrustumClipping", "ShowVertices", "ShowEdges", "ShowFaces", "BackfaceCulling",
                                                 "Lighting", "FlatShading", "ShowTextures", "CrossHair", "Fog")}
        scenes.append({
            "name": name, "width": sc.width, "height": sc.height, "num_tiles": int(r.numTiles), "options": opts,
            "fog_start": float(r.FogStart), "fog_end": float(r.FogEnd), "fog_color": [int(c) & 0xff for c in r.FogColor],
            "camera": {"position": [float(x) for x in sc.camera.Position], "direction": [float(x) for x in sc.camera.Direction],
                       "up": [float(x) for x in sc.camera.Up]},
            "meshes": meshes, "objects": objs})
        print("exported", name, f"{len(objs)} objects")
    with open(os.path.join(out, "scenes.json"), "w") as f:
        json.dump(scenes, f, indent=1)
    print(f"{len(scenes)} scenes, {len(mesh_files)} meshes -> {out}")


if __name__ == "__main__":
    main()
