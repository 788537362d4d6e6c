"""The C++ oracle's pixel side against oracle/py_raster.py, a second restatement of the same Go functions
in plain Python (Triangle, Line, Rect, Pixel, CrossHair, Fog, Sample, the tile lists and drawProjection's
overlay branches), on scenes small enough for Python loops.  Two independent transcriptions of
rasterizer.go / renderer.go:166-244 must agree bit for bit."""
import os
import sys

import numpy as np
import pytest

import gorender_b200 as g
from gorender_b200 import geometry, workloads

import scene_defs

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import py_raster  # noqa: E402


def small(builder, w, h, **opt):
    sc = builder(w=w, h=h, **opt) if "w" in builder.__code__.co_varnames else builder(**opt)
    sc.width, sc.height = w, h
    return sc


def textured_npot_cube():
    """A cube with a non-power-of-two texture (the `%` branch of Sample) crossing the camera plane."""
    y, x = np.mgrid[0:12, 0:20]
    img = np.stack([(x * 13) % 256, (y * 21) % 256, (x * y) % 256, np.full_like(x, 255)], axis=-1).astype(np.uint8)
    tex = g.NewImageTexture(img)
    tex.SetScale(2.5)
    cube = workloads.cube()
    F = cube.Faces
    faces = g.FaceArray(F.VertexIndices, F.NormalIndices, F.UVs, np.zeros(len(F), np.int32), [tex])
    return g.NewMesh(cube.Vertices, cube.VertexNormals, faces)


CASES = {
    "cube_poseA": lambda: small(scene_defs.c2, 96, 54, pose="A"),
    "cube_poseB_wire": lambda: small(scene_defs.c2, 96, 54, pose="B", ShowEdges=True, ShowVertices=True),
    "cube_serial_wire": lambda: scene_defs.SceneDef(80, 60, *workloads.config_c2("B"), options=dict(ShowEdges=True), parallel=False),
    "suzanne_faces": lambda: small(scene_defs.c1, 64, 36),
    "suzanne_wire_only": lambda: small(scene_defs.c1, 64, 36, ShowEdges=True, ShowFaces=False),
    "suzanne_verts_crosshair": lambda: small(scene_defs.c1, 61, 37, ShowVertices=True, CrossHair=True),
    "offscreen_wire": lambda: small(scene_defs.offscreen_no_clip, 80, 45, ShowEdges=True, ShowVertices=True),
    "fog_far": lambda: small(scene_defs.tiny_far, 48, 36, Fog=True, CrossHair=True),
    "fog_custom_gouraud": lambda: small(scene_defs.gouraud_sphere, 72, 54, n=4, Fog=True, FogStart=np.float32(0.7),
                                        FogEnd=np.float32(0.4), FogColor=(10, 200, 90, 128), ShowEdges=True),
    "npot_texture_cube": lambda: scene_defs.SceneDef(90, 60, [scene_defs._obj(textured_npot_cube(), r=(0.3, 0.6, 0.1))],
                                                     g.Camera(Position=(0.9, 0.2, 1.4)), {}),
    "no_textures": lambda: small(scene_defs.c2, 64, 48, pose="B", ShowTextures=False, ShowVertices=True),
}


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_python_restatement(name, oracle):
    sc = CASES[name]()
    r = sc.renderer(None)
    ref = oracle.draw(r, sc.objects, sc.camera, record=True)
    textures = []
    seen = {}
    for o in sc.objects:           # same texture numbering as oracle_binding.marshal
        if id(o.Mesh) in seen:
            continue
        seen[id(o.Mesh)] = True
        for t in o.Mesh.Faces.Textures:
            if not any(t is u for u in textures):
                textures.append(t)
    px, z, tpf = py_raster.draw(sc.width, sc.height, r.numTiles, ref["triangles"], textures,
                                ShowFaces=r.ShowFaces, ShowEdges=r.ShowEdges, ShowVertices=r.ShowVertices,
                                ShowTextures=r.ShowTextures, CrossHair=r.CrossHair, Fog=r.Fog,
                                FogStart=r.FogStart, FogEnd=r.FogEnd, FogColor=tuple(r.FogColor))
    assert tpf == ref["tpf"]
    assert np.array_equal(z.view(np.uint32), ref["zbuffer"].view(np.uint32)), name
    same = (px == ref["pixels"]).all(axis=-1)
    assert same.all(), f"{name}: {(~same).sum()} pixels differ, first at {np.argwhere(~same)[0]}"
    assert (ref["zbuffer"] > -1).sum() > 0 or not r.ShowFaces
