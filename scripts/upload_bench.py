#!/usr/bin/env python
"""Load-time path: NewMesh + upload of a large mesh, host-derived vs device-derived.

  host    numpy face normals + bounding box (mesh.py, the mirror of mesh.go:28-69), then grb_mesh_upload
  device  grb_mesh_new: the source arrays go up as they are, face normals / bounding box / the
          face-corner expansion are computed by mesh.cu, face normals and bbox copied back
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import gorender_b200 as g  # noqa: E402
from gorender_b200 import geometry  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 316      # 20 n^2 faces: 316 -> 1 997 120
src = geometry.geodesic_sphere(n, True)
F = src.Faces
dev = g.Device(0)
print(f"mesh: {len(src.Vertices)} vertices, {len(F)} faces, {len(src.VertexNormals)} vertex normals")


def faces():
    return g.FaceArray(F.VertexIndices, F.NormalIndices, F.UVs, F.TextureIndex, F.Textures)


for rep in range(3):
    t0 = time.perf_counter()
    m = g.NewMesh(src.Vertices, src.VertexNormals, faces())
    t1 = time.perf_counter()
    dev.mesh_id(m)
    t2 = time.perf_counter()
    md = g.NewMesh(src.Vertices, src.VertexNormals, faces(), device=dev)
    t3 = time.perf_counter()
    same = np.array_equal(m.FaceNormals.view(np.uint32), md.FaceNormals.view(np.uint32)) and \
        np.array_equal(m.BoundingBox.view(np.uint32), md.BoundingBox.view(np.uint32))
    print(f"rep {rep}: host NewMesh {1e3 * (t1 - t0):.1f} ms + upload {1e3 * (t2 - t1):.1f} ms = {1e3 * (t2 - t0):.1f} ms;  "
          f"device NewMesh (upload + derive + read back) {1e3 * (t3 - t2):.1f} ms;  identical={same}")
