"""Synthetic workloads of BASELINE.json / SURVEY.md §8(d).

Deterministic, no RNG: the frequency-n geodesic sphere (20*n^2 triangles,
10*n^2+2 vertices; n=100 is the 200k-triangle C3 mesh), an OBJ writer so the
same meshes can go through the OBJ loader, and the camera / pose sequences of
the five configurations.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np

from .mesh import FaceArray, Mesh, NewMesh


def _icosahedron() -> Tuple[np.ndarray, np.ndarray]:
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = np.array([
        [-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0],
        [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
        [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1],
    ], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([
        [0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
        [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
        [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
        [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1],
    ], dtype=np.int64)
    return v, f


def geodesic_sphere_arrays(n: int) -> Tuple[np.ndarray, np.ndarray]:
    """Vertices (10n^2+2, 3) float64 on the unit sphere and faces (20n^2, 3),
    outward counter-clockwise, face-major order (good screen-space locality)."""
    iv, ifc = _icosahedron()
    index = {}
    verts: List[np.ndarray] = []

    def vid(a, b, c, i, j, k):
        # point (i*A + j*B + k*C)/n, i+j+k == n; canonical key so shared edge /
        # corner points are created once
        key = tuple(sorted((p, q) for p, q in ((a, i), (b, j), (c, k)) if q != 0))
        r = index.get(key)
        if r is None:
            p = (i * iv[a] + j * iv[b] + k * iv[c]) / n
            p = p / math.sqrt(float(p @ p))
            r = index[key] = len(verts)
            verts.append(p)
        return r

    faces = []
    for a, b, c in ifc:
        # grid[r][s]: r steps from A toward B, s steps toward C
        grid = [[vid(a, b, c, n - r - s, r, s) for s in range(n - r + 1)] for r in range(n + 1)]
        for r in range(n):
            for s in range(n - r):
                faces.append((grid[r][s], grid[r + 1][s], grid[r][s + 1]))
                if s < n - r - 1:
                    faces.append((grid[r + 1][s], grid[r + 1][s + 1], grid[r][s + 1]))
    v = np.array(verts, dtype=np.float64)
    f = np.array(faces, dtype=np.int64)
    # orient outward (CCW seen from outside)
    p0, p1, p2 = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    flip = np.einsum("ij,ij->i", np.cross(p1 - p0, p2 - p0), p0 + p1 + p2) < 0
    f[flip] = f[flip][:, [0, 2, 1]]
    assert len(v) == 10 * n * n + 2 and len(f) == 20 * n * n
    return v, f


def geodesic_sphere(n: int = 100, with_normals_uvs: bool = False, texture=None, radius: float = 1.0) -> Mesh:
    """The C3 mesh (`f a b c`, no vn => flat branch) or, with
    `with_normals_uvs`, the `v/vt/vn` variant (vn = position, vt =
    equirectangular) used by the Gouraud / textured configurations (C4)."""
    v, f = geodesic_sphere_arrays(n)
    verts = np.ones((len(v), 4), dtype=np.float32)
    verts[:, :3] = (v * radius).astype(np.float32)
    if not with_normals_uvs:
        return NewMesh(verts, None, FaceArray(f.astype(np.int32)))
    vn = np.ones((len(v), 4), dtype=np.float32)
    vn[:, :3] = v.astype(np.float32)
    u = 0.5 + np.arctan2(v[:, 2], v[:, 0]) / (2 * math.pi)
    w = 0.5 - np.arcsin(np.clip(v[:, 1], -1, 1)) / math.pi
    vt = np.stack([u, w], axis=1).astype(np.float32)
    faces = FaceArray(f.astype(np.int32), f.astype(np.int32), vt[f])
    mesh = NewMesh(verts, vn, faces)
    if texture is not None:
        mesh.Faces.SetTexture(texture)
    return mesh


def write_obj(mesh: Mesh, filename: str, name: str = "mesh") -> None:
    """Write `mesh` as OBJ: `f a b c` without normals, `f v/vt/vn` with (never
    `v//vn`, which the reference mis-parses — SURVEY.md H10)."""
    with open(filename, "w") as out:
        out.write(f"o {name}\n")
        for x, y, z, _ in mesh.Vertices:
            out.write(f"v {x:.9g} {y:.9g} {z:.9g}\n")
        has_vn = len(mesh.VertexNormals) != 0
        F = mesh.Faces
        if has_vn:
            for x, y, z, _ in mesh.VertexNormals:
                out.write(f"vn {x:.9g} {y:.9g} {z:.9g}\n")
            uv = F.UVs.reshape(-1, 2)
            for u, v in uv:
                out.write(f"vt {u:.9g} {v:.9g}\n")
            for i in range(len(F)):
                a, b, c = F.VertexIndices[i] + 1
                na, nb, nc = F.NormalIndices[i] + 1
                t = 3 * i + 1
                out.write(f"f {a}/{t}/{na} {b}/{t + 1}/{nb} {c}/{t + 2}/{nc}\n")
        else:
            for a, b, c in F.VertexIndices + 1:
                out.write(f"f {a} {b} {c}\n")


# ------------------------------------------------------------------ poses

def vec3_to_radians_f32(deg) -> np.ndarray:
    """Vec3.ToRadians (vector.go:82-85)."""
    from .vecmath import vec3_to_radians

    return vec3_to_radians(np.asarray(deg, dtype=np.float32))


def default_camera():
    """main.go:192-196."""
    from .renderer import Camera

    return Camera(Position=(0, 0, 5), Direction=(0, 0, -1), Up=(0, 1, 0))


def spin_rotations(num_frames: int, start: int = 0) -> np.ndarray:
    """Rotation.Y of the demo spin (main.go:229-233): += float32(0.01) per frame, in f32."""
    out = np.empty(num_frames, dtype=np.float32)
    r = np.float32(0.0)
    step = np.float32(0.01)
    for _ in range(start):
        r = np.float32(r + step)
    for i in range(num_frames):
        out[i] = r
        r = np.float32(r + step)
    return out


def orbit_cameras(num_poses: int = 4096, radius: float = 5.0, height: float = 0.3):
    """C5: pose k on the circle radius*(sin t, height, cos t), t = 2*pi*k/num_poses,
    looking at the origin (direction = -position, normalised by NewViewMatrix)."""
    from .renderer import Camera

    cams = []
    for k in range(num_poses):
        t = 2.0 * math.pi * k / num_poses
        pos = np.array([radius * math.sin(t), radius * height, radius * math.cos(t)], dtype=np.float32)
        cams.append(Camera(Position=pos, Direction=-pos, Up=(0, 1, 0)))
    return cams
