"""Mesh / Object containers — host-side mirror of the reference's `mesh.go`.

The reference keeps `Faces []Face` with a Go pointer per face (mesh.go:12-17);
that layout cannot cross a C ABI, so faces are held struct-of-arrays from the
start (`FaceArray`): exactly the flattened form the C-ABI mesh upload takes
(include/gorender_b200.h, `grb_mesh_upload`).
"""
from __future__ import annotations

import os
from typing import List, Optional

import numpy as np

from .texture import Texture


class FaceArray:
    """`[]Face` (mesh.go:12-17), struct-of-arrays.

    VertexIndices (F,3) int32, NormalIndices (F,3) int32, UVs (F,3,2) float32,
    TextureIndex (F,) int32 into `Textures` (-1 == nil Texture).
    """

    def __init__(self, VertexIndices, NormalIndices=None, UVs=None, TextureIndex=None,
                 Textures: Optional[List[Texture]] = None):
        self.VertexIndices = np.ascontiguousarray(VertexIndices, dtype=np.int32).reshape(-1, 3)
        n = len(self.VertexIndices)
        if NormalIndices is None:
            NormalIndices = np.zeros((n, 3), dtype=np.int32)
        if UVs is None:
            UVs = np.zeros((n, 3, 2), dtype=np.float32)
        if TextureIndex is None:
            TextureIndex = np.full(n, -1, dtype=np.int32)
        self.NormalIndices = np.ascontiguousarray(NormalIndices, dtype=np.int32).reshape(n, 3)
        self.UVs = np.ascontiguousarray(UVs, dtype=np.float32).reshape(n, 3, 2)
        self.TextureIndex = np.ascontiguousarray(TextureIndex, dtype=np.int32).reshape(n)
        self.Textures: List[Texture] = list(Textures or [])

    def __len__(self) -> int:
        return len(self.VertexIndices)

    def SetTexture(self, texture: Optional[Texture]) -> None:
        """`for i := range mesh.Faces { mesh.Faces[i].Texture = texture }` (scene.go:94-101)."""
        if texture is None:
            self.Textures = []
            self.TextureIndex[:] = -1
        else:
            self.Textures = [texture]
            self.TextureIndex[:] = 0


def boundingBox(vertices: np.ndarray) -> np.ndarray:
    """mesh.go:28-51: the 8 corners, (x,y,z,1), min/max order of the reference."""
    v = np.asarray(vertices, dtype=np.float32)
    mn = v[:, :3].min(axis=0)
    mx = v[:, :3].max(axis=0)
    c = np.empty((8, 4), dtype=np.float32)
    k = 0
    for x in (mn[0], mx[0]):
        for y in (mn[1], mx[1]):
            for z in (mn[2], mx[2]):
                c[k] = (x, y, z, 1.0)
                k += 1
    return c


def faceNormals(vertices: np.ndarray, vidx: np.ndarray) -> np.ndarray:
    """mesh.go:54-60: normalize((v1-v0) x (v2-v0)).ToVec4(), float32, reference op order."""
    v = np.asarray(vertices, dtype=np.float32)
    if len(vidx) == 0:
        return np.zeros((0, 4), dtype=np.float32)
    v0 = v[vidx[:, 0], :3]
    v1 = v[vidx[:, 1], :3]
    v2 = v[vidx[:, 2], :3]
    a = v1 - v0
    b = v2 - v0
    with np.errstate(all="ignore"):
        x = a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1]
        y = a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2]
        z = a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]
        ln = np.sqrt(x * x + y * y + z * z)
        out = np.empty((len(vidx), 4), dtype=np.float32)
        out[:, 0] = x / ln
        out[:, 1] = y / ln
        out[:, 2] = z / ln
    out[:, 3] = 1.0
    return out


class Mesh:
    """mesh.go:19-26."""

    def __init__(self, Vertices, VertexNormals, Faces: FaceArray, Name: str = "", device=None):
        self.Name = Name
        self.Vertices = np.ascontiguousarray(Vertices, dtype=np.float32).reshape(-1, 4)
        self.VertexNormals = np.ascontiguousarray(
            VertexNormals if VertexNormals is not None else np.zeros((0, 4)), dtype=np.float32).reshape(-1, 4)
        self.Faces = Faces
        if device is None:
            if len(Faces) and (Faces.VertexIndices.min() < 0 or Faces.VertexIndices.max() >= len(self.Vertices)):
                raise IndexError("face vertex index out of range")  # the reference panics (mesh.go:56-58)
            self.FaceNormals = faceNormals(self.Vertices, Faces.VertexIndices)
            self.BoundingBox = boundingBox(self.Vertices)
        else:
            # NewMesh on the device (grb_mesh_new): the upload computes both and fills these in
            self.FaceNormals = self.BoundingBox = None
            device.mesh_id(self)


def NewMesh(vertices, vertexNormals, faces: FaceArray, device=None) -> Mesh:
    """mesh.go:53-69.  With `device`, the face normals and the bounding box are computed by the GPU
    while the mesh is uploaded instead of by numpy on the host."""
    return Mesh(vertices, vertexNormals, faces, device=device)


class Object:
    """mesh.go:71-79.  The per-frame scratch slices of the reference
    (TransformedVertices, WorldVertexNormals, WorldFaceNormals) live in HBM, not here."""

    def __init__(self, mesh: Mesh):
        self.Mesh = mesh
        self.Rotation = np.zeros(3, dtype=np.float32)
        self.Translation = np.zeros(3, dtype=np.float32)
        self.Scale = np.ones(3, dtype=np.float32)

    # embedded *Mesh (mesh.go:72)
    def __getattr__(self, name):
        return getattr(self.__dict__["Mesh"], name)


def NewObject(mesh: Mesh) -> Object:
    """mesh.go:81-89."""
    return Object(mesh)


def LoadMeshFile(filename: str, singleMesh: bool) -> List[Mesh]:
    """mesh.go:91-103."""
    ext = os.path.splitext(filename)[1]
    if ext == ".obj":
        from .obj import LoadObjFile

        return LoadObjFile(filename, singleMesh)
    raise ValueError(f"unsupported mesh format: {ext}")
