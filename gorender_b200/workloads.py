"""The five BASELINE.json configurations as (objects, cameras)
builders, plus the .npz mesh fixture format used to carry the reference's
model files to machines where /root/reference does not exist.
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import numpy as np

from . import geometry
from .mesh import FaceArray, Mesh, NewMesh, NewObject, Object
from .renderer import Camera
from .texture import NewImageTexture, Texture, TextureTypeSolidColor

# the reference's two model files (models/suzanne.obj, models/cube.obj + textures-16.png) as LoadObjFile left them,
# carried as package data so that the workloads do not depend on /root/reference or on the test tree
ASSETS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")
# the oracle's pinned outputs live with the tests
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


# ------------------------------------------------------------------ fixtures

def save_mesh_fixture(path: str, mesh: Mesh) -> None:
    """Store what LoadObjFile produced: arrays + the (non-premultiplied is gone:
    premultiplied) texture pixels, so no OBJ/PNG parsing is needed to reload."""
    F = mesh.Faces
    d = dict(vertices=mesh.Vertices, vnormals=mesh.VertexNormals, vidx=F.VertexIndices, nidx=F.NormalIndices,
             uvs=F.UVs, texidx=F.TextureIndex, ntex=np.int32(len(F.Textures)))
    for i, t in enumerate(F.Textures):
        d[f"tex{i}_typ"] = np.int32(t.typ)
        d[f"tex{i}_color"] = np.array(t.color, dtype=np.uint8)
        d[f"tex{i}_scale"] = np.float32(t.scale)
        if t.pixels is not None:
            d[f"tex{i}_pixels"] = t.pixels
    np.savez_compressed(path, **d)


def load_mesh_fixture(path: str) -> Mesh:
    z = np.load(path)
    texs: List[Texture] = []
    for i in range(int(z["ntex"])):
        px = z[f"tex{i}_pixels"] if f"tex{i}_pixels" in z.files else None
        texs.append(Texture(int(z[f"tex{i}_typ"]), tuple(z[f"tex{i}_color"].tolist()), px, float(z[f"tex{i}_scale"])))
    faces = FaceArray(z["vidx"], z["nidx"], z["uvs"], z["texidx"], texs)
    return NewMesh(z["vertices"], z["vnormals"], faces)


def suzanne() -> Mesh:
    return load_mesh_fixture(os.path.join(ASSETS_DIR, "suzanne.npz"))


def cube() -> Mesh:
    return load_mesh_fixture(os.path.join(ASSETS_DIR, "cube.npz"))


def checker_texture(size: int = 64, cells: int = 8) -> Texture:
    """Deterministic RGBA test texture with partial alpha (exercises premultiply)."""
    y, x = np.mgrid[0:size, 0:size]
    c = ((x * cells // size) + (y * cells // size)) & 1
    img = np.zeros((size, size, 4), dtype=np.uint8)
    img[..., 0] = np.where(c, 230, 40) + (x % 7)
    img[..., 1] = (x * 255 // max(size - 1, 1)).astype(np.uint8)
    img[..., 2] = (y * 255 // max(size - 1, 1)).astype(np.uint8)
    img[..., 3] = np.where((x + y) % 5 == 0, 128, 255)
    return NewImageTexture(img)


# ------------------------------------------------------------------ configs

def config_c1() -> Tuple[List[Object], Camera]:
    """C1: suzanne, default camera, identity TRS."""
    return [NewObject(suzanne())], geometry.default_camera()


def config_c2(pose: str = "A") -> Tuple[List[Object], Camera]:
    """C2: textured cube with the camera partly inside the view volume."""
    obj = NewObject(cube())
    if pose == "A":
        cam = Camera(Position=(1.2, 0, 0.5), Direction=(0, 0, -1), Up=(0, 1, 0))
    else:
        cam = Camera(Position=(0.6, 0.3, 1.7), Direction=(0, 0, -1), Up=(0, 1, 0))
        obj.Rotation = np.array([0, 0.6, 0], dtype=np.float32)
    return [obj], cam


_sphere_cache = {}


def sphere(n: int = 100, with_normals_uvs: bool = False, texture: Optional[Texture] = None) -> Mesh:
    key = (n, with_normals_uvs, id(texture))
    m = _sphere_cache.get(key)
    if m is None:
        m = _sphere_cache[key] = geometry.geodesic_sphere(n, with_normals_uvs, texture)
    return m


def config_c3(n: int = 100) -> Tuple[List[Object], Camera]:
    """C3: 20*n^2-triangle geodesic sphere (n=100: 200k), flat, untextured, default camera."""
    return [NewObject(sphere(n))], geometry.default_camera()


def config_c4(n: int = 100, texture: Optional[Texture] = None) -> Tuple[List[Object], Camera]:
    """C4: 10 instances of the textured Gouraud sphere on a 5x2 grid, spacing 2.2,
    per-instance Y rotation 36 deg * k, texture scale 4, camera (0,0,9)."""
    if texture is None:
        texture = cube().Faces.Textures[0]
    tex = Texture(texture.typ, texture.color, texture.pixels, 4.0)
    mesh = sphere(n, True, tex)
    objs = []
    for k in range(10):
        o = NewObject(mesh)
        col, row = k % 5, k // 5
        o.Translation = np.array([(col - 2) * 2.2, (row - 0.5) * 2.2, 0], dtype=np.float32)
        o.Rotation = geometry.vec3_to_radians_f32([0, 36.0 * k, 0])
        objs.append(o)
    return objs, Camera(Position=(0, 0, 9), Direction=(0, 0, -1), Up=(0, 1, 0))


def config_c5(n: int = 100, poses: int = 4096) -> Tuple[List[Object], List[Camera]]:
    """C5: `poses` orbit cameras around the C3 sphere."""
    return [NewObject(sphere(n))], geometry.orbit_cameras(poses)
