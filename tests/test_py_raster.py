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
import py_project  # noqa: E402
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
    # the affine mode is this repository's own definition (no code path in the reference): both restatements must agree on it
    "affine_cube_poseB": lambda: small(scene_defs.c2, 96, 54, pose="B", AffineTextures=True),
    "affine_npot_cube": lambda: scene_defs.SceneDef(90, 60, [scene_defs._obj(textured_npot_cube(), r=(0.3, 0.6, 0.1))],
                                                    g.Camera(Position=(0.9, 0.2, 1.4)), {"AffineTextures": True}),
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
                                FogStart=r.FogStart, FogEnd=r.FogEnd, FogColor=tuple(r.FogColor),
                                AffineTextures=r.AffineTextures)
    assert tpf == ref["tpf"]
    assert np.array_equal(z.view(np.uint32), ref["zbuffer"].view(np.uint32)), name
    same = (px == ref["pixels"]).all(axis=-1)
    assert same.all(), f"{name}: {(~same).sum()} pixels differ, first at {np.argwhere(~same)[0]}"
    assert (ref["zbuffer"] > -1).sum() > 0 or not r.ShowFaces


FULL = ["cube_poseA", "cube_poseB_wire", "suzanne_faces", "offscreen_wire", "fog_custom_gouraud", "npot_texture_cube"]


@pytest.mark.parametrize("name", FULL)
@pytest.mark.parametrize("variant", ["default", "flat_nocull"])
def test_whole_draw_restated_in_python(name, variant, oracle):
    """Renderer.Draw end to end, twice: oracle (C++) against py_project + py_raster (Python).  The projected
    triangle list must agree bit for bit (clip-space transform, cull, lighting, Sutherland-Hodgman clipping,
    divide, viewport), and so must the frame."""
    import gorender_b200.vecmath as vm

    sc = CASES[name]()
    if variant == "flat_nocull":
        sc.options = dict(sc.options, FlatShading=True, BackfaceCulling=False)
    r = sc.renderer(None)
    ref = oracle.draw(r, sc.objects, sc.camera, record=True)
    textures = []
    for o in sc.objects:
        for t in o.Mesh.Faces.Textures:
            if not any(t is u for u in textures):
                textures.append(t)
    persp = r.perspective()
    view = vm.NewViewMatrix(sc.camera.Position, sc.camera.Direction, sc.camera.Up)
    screen = vm.NewScreenMatrix(sc.width, sc.height)
    tris, vis = [], []
    for o in sc.objects:
        world, mvp = r.object_matrices(o, sc.camera, persp, view)
        tex_ids = [next(k for k, u in enumerate(textures) if u is t) for t in o.Mesh.Faces.Textures]
        v, out = py_project.project_object(o.Mesh, world, mvp, screen, vm.light_direction(), tex_ids=tex_ids,
                                           z_near=float(r.zNear), z_far=float(r.zFar), BackfaceCulling=r.BackfaceCulling,
                                           Lighting=r.Lighting, FlatShading=r.FlatShading, FrustumClipping=r.FrustumClipping)
        vis.append(v)
        tris += out
    assert vis == ref["visibility"].tolist()
    want = ref["triangles"]
    assert len(tris) == len(want)
    for k, (a, b) in enumerate(zip(tris, want)):
        for field in ("points", "uvs", "intensity"):
            x, y = np.asarray(a[field], np.float32), np.asarray(b[field], np.float32)
            nan = np.isnan(y)
            assert np.array_equal(np.isnan(x), nan) and np.array_equal(x.view(np.uint32)[~nan], y.view(np.uint32)[~nan]), (k, field)
        assert (a["tex"] >= 0) == (int(b["tex"]) >= 0)
    px, z, tpf = py_raster.draw(sc.width, sc.height, r.numTiles, tris, textures,
                                ShowFaces=r.ShowFaces, ShowEdges=r.ShowEdges, ShowVertices=r.ShowVertices,
                                ShowTextures=r.ShowTextures, CrossHair=r.CrossHair, Fog=r.Fog,
                                FogStart=r.FogStart, FogEnd=r.FogEnd, FogColor=tuple(r.FogColor),
                                AffineTextures=r.AffineTextures)
    assert tpf == ref["tpf"]
    assert np.array_equal(z.view(np.uint32), ref["zbuffer"].view(np.uint32))
    assert np.array_equal(px, ref["pixels"])
