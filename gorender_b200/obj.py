"""Wavefront OBJ / MTL loading — host-side mirror of the reference's `obj.go`.

Load-time host code (out of scope for CUDA, kept for drop-in).
Behaviour follows obj.go line by line, including the index-offset logic for
multi-object files (obj.go:31-40) and the `v//vn` quirk (obj.go:77-89, SURVEY
H10: the third normal index is scanned into vn1, so NormalIndices[2] is -1).
"""
from __future__ import annotations

import ctypes
import ctypes.util
import logging
import os
from typing import Dict, List, Optional

import numpy as np

from .mesh import FaceArray, Mesh, NewMesh
from .texture import LoadTextureFile, Texture, TextureTypeSolidColor

log = logging.getLogger("gorender_b200")

_libc = ctypes.CDLL(ctypes.util.find_library("c") or "libc.so.6")
_libc.strtof.restype = ctypes.c_float
_libc.strtof.argtypes = [ctypes.c_char_p, ctypes.c_void_p]


def parse_f32(tokens) -> np.ndarray:
    """Decimal -> float32, correctly rounded like Go's strconv (`%f` into float32).

    float(s) is correctly rounded to double; narrowing that to float32 differs
    from a direct decimal->float32 rounding only when the double lands exactly
    on a float32 midpoint — those (vanishingly rare) tokens go through strtof.
    """
    d = np.array([float(t) for t in tokens], dtype=np.float64)
    out = d.astype(np.float32)
    bits = d.view(np.uint64)
    mid = (bits & np.uint64((1 << 29) - 1)) == np.uint64(1 << 28)
    for i in np.nonzero(mid)[0]:
        out[i] = _libc.strtof(str(tokens[i]).encode(), None)
    return out


class ObjMaterial:
    """obj.go:14-17."""

    def __init__(self, Name: str, MapKd: str = ""):
        self.Name = Name
        self.MapKd = MapKd


def parseMtlLibFile(filename: str) -> List[ObjMaterial]:
    """obj.go:153-192."""
    materials: List[ObjMaterial] = []
    mat: Optional[ObjMaterial] = None
    with open(filename, "r") as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            if line.startswith("newmtl "):
                if mat is not None:
                    materials.append(mat)
                mat = ObjMaterial(line[len("newmtl "):])
            elif line.startswith("map_Kd "):
                mat.MapKd = line[len("map_Kd "):]
    if mat is not None:
        materials.append(mat)
    return materials


class _ObjContext:
    """obj.go:19-40."""

    def __init__(self):
        self.Vertices: List[str] = []        # raw tokens, converted in bulk
        self.TextureVertices: List[str] = []
        self.VertexNormals: List[str] = []
        self.face_v: List[int] = []
        self.face_vt: List[int] = []
        self.face_vn: List[int] = []
        self.face_tex: List[Optional[Texture]] = []
        self.Textures: Dict[str, Texture] = {}
        self.VertexIndexOffset = 0
        self.TextureVertexOffset = 0
        self.VertexNormalOffset = 0

    def num_vertices(self) -> int:
        return len(self.Vertices) // 3

    def Clear(self) -> None:
        self.VertexIndexOffset += len(self.Vertices) // 3
        self.TextureVertexOffset += len(self.TextureVertices) // 2
        self.VertexNormalOffset += len(self.VertexNormals) // 3
        self.Vertices, self.TextureVertices, self.VertexNormals = [], [], []
        self.face_v, self.face_vt, self.face_vn, self.face_tex = [], [], [], []

    def build(self) -> Mesh:
        nv = len(self.Vertices) // 3
        verts = np.ones((nv, 4), dtype=np.float32)
        verts[:, :3] = parse_f32(self.Vertices).reshape(nv, 3)
        nvn = len(self.VertexNormals) // 3
        vns = np.ones((nvn, 4), dtype=np.float32)   # w = 1 (obj.go:57)
        if nvn:
            vns[:, :3] = parse_f32(self.VertexNormals).reshape(nvn, 3)
        nvt = len(self.TextureVertices) // 2
        vts = parse_f32(self.TextureVertices).reshape(nvt, 2) if nvt else np.zeros((0, 2), np.float32)

        nf = len(self.face_v) // 3
        vidx = np.array(self.face_v, dtype=np.int64).reshape(nf, 3)
        vtidx = np.array(self.face_vt, dtype=np.int64).reshape(nf, 3)
        nidx = np.array(self.face_vn, dtype=np.int64).reshape(nf, 3)
        uvs = np.zeros((nf, 3, 2), dtype=np.float32)
        has_vt = vtidx[:, 0] != np.iinfo(np.int64).min if nf else np.zeros(0, bool)
        if has_vt.any():
            sel = vtidx[has_vt]
            if sel.min() < 0 or sel.max() >= nvt:
                raise IndexError("texture vertex index out of range")  # Go: index panic (obj.go:107-109)
            uvs[has_vt] = vts[sel]
        textures: List[Texture] = []
        tex_index = np.full(nf, -1, dtype=np.int32)
        slot: Dict[int, int] = {}
        for i, t in enumerate(self.face_tex):
            if t is None:
                continue
            k = slot.get(id(t))
            if k is None:
                k = slot[id(t)] = len(textures)
                textures.append(t)
            tex_index[i] = k
        faces = FaceArray(vidx.astype(np.int32), nidx.astype(np.int32), uvs, tex_index, textures)
        return NewMesh(verts, vns, faces)


_NO_VT = np.iinfo(np.int64).min


def _vt_indices(c: _ObjContext, idx):
    """obj.go:107-109, 131-133 index c.TextureVertices while parsing the face: only the texture vertices read so
    far exist, and an index outside them is a Go runtime panic."""
    n = len(c.TextureVertices) // 2
    if any(i < 0 or i >= n for i in idx):
        raise IndexError("texture vertex index out of range")
    return idx


def _parseFace(c: _ObjContext, line: str) -> None:
    """obj.go:60-151."""
    if line.count(" ") != 3:
        raise ValueError("mesh is not triangulated")
    toks = line.split(" ")[1:]
    vo, to, no = c.VertexIndexOffset, c.TextureVertexOffset, c.VertexNormalOffset
    if line.count("//") == 3:
        v = [int(t.split("//")[0]) for t in toks]
        n = [int(t.split("//")[1]) for t in toks]
        # Sscanf target list is (&vn0, &vn1, &vn1): vn1 takes the third, vn2 stays 0
        vn0, vn1, vn2 = n[0], n[2], 0
        c.face_v += [v[0] - vo - 1, v[1] - vo - 1, v[2] - vo - 1]
        c.face_vn += [vn0 - no - 1, vn1 - no - 1, vn2 - no - 1]
        c.face_vt += [_NO_VT] * 3
    elif line.count("/") == 3:
        p = [t.split("/") for t in toks]
        c.face_v += [int(q[0]) - vo - 1 for q in p]
        c.face_vt += _vt_indices(c, [int(q[1]) - to - 1 for q in p])
        c.face_vn += [0, 0, 0]
    elif line.count("/") == 6:
        p = [t.split("/") for t in toks]
        c.face_v += [int(q[0]) - vo - 1 for q in p]
        c.face_vt += _vt_indices(c, [int(q[1]) - to - 1 for q in p])
        c.face_vn += [int(q[2]) - no - 1 for q in p]
    else:
        c.face_v += [int(t) - vo - 1 for t in toks]
        c.face_vt += [_NO_VT] * 3
        c.face_vn += [0, 0, 0]


def LoadObjFile(filename: str, singleMesh: bool) -> List[Mesh]:
    """obj.go:196-309."""
    dirname = os.path.dirname(filename)
    defaultTexture = Texture(TextureTypeSolidColor, color=(255, 0, 255, 255))  # obj.go:208
    currentTexture: Optional[Texture] = None
    c = _ObjContext()
    textureFiles: Dict[str, Texture] = {}
    meshes: List[Mesh] = []

    with open(filename, "r") as f:
        for raw in f:
            line = raw.strip()
            if not line:
                continue
            if line.startswith("mtllib "):
                mtlLibFile = line[len("mtllib "):]
                log.info("found mtllib file: %s", mtlLibFile)
                try:
                    materials = parseMtlLibFile(os.path.join(dirname, mtlLibFile))
                except OSError as e:
                    raise RuntimeError(f"failed to parse material library: {e}")
                for m in materials:
                    if m.MapKd == "":
                        log.info("using default texture for material: %s", m.Name)
                        c.Textures[m.Name] = defaultTexture
                    elif m.MapKd in textureFiles:
                        c.Textures[m.Name] = textureFiles[m.MapKd]
                    else:
                        log.info("loading texture: %s", m.MapKd)
                        texturePath = m.MapKd
                        if texturePath[0] != "/":
                            texturePath = os.path.join(dirname, m.MapKd)
                        try:
                            texture = LoadTextureFile(texturePath)
                        except Exception as e:
                            raise RuntimeError(f"failed to load texture: {e}")
                        textureFiles[m.MapKd] = texture
                        c.Textures[m.Name] = texture
            elif line.startswith("o "):
                if c.num_vertices() != 0 and not singleMesh:
                    meshes.append(c.build())
                    c.Clear()
            elif line.startswith("v "):
                t = line.split()
                if len(t) < 4:
                    raise ValueError("unexpected EOF")  # Sscanf error (obj.go:44)
                c.Vertices += t[1:4]
            elif line.startswith("vt "):
                t = line.split()
                if len(t) < 3:
                    raise ValueError("unexpected EOF")
                c.TextureVertices += t[1:3]
            elif line.startswith("vn "):
                t = line.split()
                if len(t) < 4:
                    raise ValueError("unexpected EOF")
                c.VertexNormals += t[1:4]
            elif line.startswith("usemtl "):
                currentTexture = c.Textures.get(line[len("usemtl "):])  # unknown name -> nil
            elif line.startswith("f "):
                _parseFace(c, line)
                c.face_tex.append(currentTexture)

    if c.num_vertices() != 0:
        meshes.append(c.build())
    if not meshes:
        raise ValueError("obj file does not have any vertices data")
    return meshes


def LoadObjFileNative(filename: str, singleMesh: bool, device=None) -> List[Mesh]:
    """LoadObjFile (obj.go:196-309) with the parsing done by the native library (`grb_obj_parse`,
    csrc/objparse.cpp: one pass over the file in memory instead of a scan per line; SURVEY.md §8f
    n4) — same meshes, bit for bit, as `LoadObjFile`.  Textures are decoded here, like in the
    reference's host code.  With `device`, NewMesh's face normals and bounding box are computed on
    the GPU while each mesh is uploaded (`grb_mesh_new`)."""
    import ctypes as C

    from . import _cabi

    lib = _cabi.load()
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    rc = lib.grb_obj_parse(os.fsencode(filename), int(bool(singleMesh)), C.byref(h), err, len(err))
    if rc != 0:
        msg = err.value.decode("utf-8", "replace")
        if msg.startswith("open ") or msg.startswith("failed to parse material library"):
            raise RuntimeError(msg)
        if msg == "index out of range":
            raise IndexError("texture vertex index out of range")
        raise ValueError(msg)
    try:
        sources: List[Texture] = []
        for i in range(lib.grb_obj_num_textures(h)):
            path = lib.grb_obj_texture_path(h, i).decode()
            if path == "":
                sources.append(Texture(TextureTypeSolidColor, color=(255, 0, 255, 255)))  # obj.go:208
            else:
                log.info("loading texture: %s", path)
                try:
                    sources.append(LoadTextureFile(path))
                except Exception as e:
                    raise RuntimeError(f"failed to load texture: {e}")
        meshes: List[Mesh] = []
        for i in range(lib.grb_obj_num_meshes(h)):
            d = _cabi.grb_mesh_desc()
            assert lib.grb_obj_mesh(h, i, C.byref(d)) == 0
            nv, nvn, nf = d.nv, d.nvn, d.nf

            def arr(ptr, n, dtype):
                if n == 0 or not ptr:
                    return np.zeros(0, dtype)
                return np.ctypeslib.as_array(ptr, (n,)).astype(dtype, copy=True)

            verts = arr(d.vertices, nv * 4, np.float32).reshape(nv, 4)
            vns = arr(d.vnormals, nvn * 4, np.float32).reshape(nvn, 4)
            vidx = arr(d.vidx, nf * 3, np.int32).reshape(nf, 3)
            nidx = arr(d.nidx, nf * 3, np.int32).reshape(nf, 3)
            uvs = arr(d.uvs, nf * 6, np.float32).reshape(nf, 3, 2)
            src = arr(d.tex, nf, np.int32)
            # Face.Texture pointers -> the mesh's texture list in order of first use
            used, first = np.unique(src[src >= 0], return_index=True) if nf else (np.zeros(0, np.int32), np.zeros(0, np.int64))
            order = used[np.argsort(first)]
            lut = np.full(len(sources) + 1, -1, np.int32)
            lut[order] = np.arange(len(order), dtype=np.int32)
            tex_index = lut[src] if nf else np.zeros(0, np.int32)
            faces = FaceArray(vidx, nidx, uvs, tex_index, [sources[k] for k in order])
            meshes.append(NewMesh(verts, vns, faces, device=device))
        return meshes
    finally:
        lib.grb_obj_free(h)
