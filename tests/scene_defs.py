"""Scenes shared by the CPU (oracle pin) and GPU (parity) tests.

Every scene is deterministic.  PINNED scenes are small enough for the oracle
to finish in well under a second; their oracle framebuffer hashes are
committed in tests/golden/oracle_outputs.json (scripts/make_golden.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List

import numpy as np

import gorender_b200 as g
from gorender_b200 import geometry, workloads


class StubFrameBuffer:
    """Width/Height only — lets the oracle be driven without a CUDA device."""

    def __init__(self, width, height):
        self.Width, self.Height, self.Frames, self.dev = width, height, 1, None


@dataclass
class SceneDef:
    width: int
    height: int
    objects: List[g.Object]
    camera: g.Camera
    options: Dict[str, bool] = field(default_factory=dict)
    parallel: bool = True

    def renderer(self, fb):
        r = g.Renderer(fb if fb is not None else StubFrameBuffer(self.width, self.height), self.parallel)
        for k, v in self.options.items():
            assert hasattr(r, k)
            setattr(r, k, v)
        return r


def _obj(mesh, t=(0, 0, 0), r=(0, 0, 0), s=(1, 1, 1)):
    o = g.NewObject(mesh)
    o.Translation = np.array(t, np.float32)
    o.Rotation = np.array(r, np.float32)
    o.Scale = np.array(s, np.float32)
    return o


def c1(w=1280, h=720, **opt):
    objs, cam = workloads.config_c1()
    return SceneDef(w, h, objs, cam, opt)


def c2(pose="A", w=1280, h=720, **opt):
    objs, cam = workloads.config_c2(pose)
    return SceneDef(w, h, objs, cam, opt)


def c3(n=100, w=1280, h=720, cam_z=5.0, **opt):
    objs, _ = workloads.config_c3(n)
    return SceneDef(w, h, objs, g.Camera(Position=(0, 0, cam_z), Direction=(0, 0, -1), Up=(0, 1, 0)), opt)


def c4(n=20, w=960, h=540, **opt):
    objs, cam = workloads.config_c4(n)
    return SceneDef(w, h, objs, cam, opt)


def gouraud_sphere(n=16, w=640, h=480, textured=True, **opt):
    tex = workloads.checker_texture(64) if textured else None
    mesh = geometry.geodesic_sphere(n, True, tex)
    return SceneDef(w, h, [_obj(mesh, r=(0.3, 0.5, 0.1))], g.Camera(Position=(0.2, 0.1, 2.2)), opt)


def npot_texture(w=400, h=300, **opt):
    """Non-power-of-two texture => TextureTypeImage (`%` wrap + idx<0 clamp, texture.go:81-87)."""
    y, x = np.mgrid[0:48, 0:80]
    img = np.stack([(x * 3) % 256, (y * 5) % 256, (x + y) % 256, np.full_like(x, 255)], axis=-1).astype(np.uint8)
    tex = g.NewImageTexture(img)
    tex.SetScale(3.0)
    mesh = geometry.geodesic_sphere(6, True, tex)
    return SceneDef(w, h, [_obj(mesh, r=(0.0, 1.0, 0.0))], g.Camera(Position=(0, 0, 2.5)), opt)


def inside_sphere(n=12, w=512, h=384, **opt):
    """Camera inside a big sphere, culling off: every plane clips something."""
    outer = geometry.geodesic_sphere(n, True, workloads.checker_texture(32))
    # reversed winding: the rasteriser only fills one winding (rasterizer.go:147), so the
    # inside of a sphere is visible only with its faces flipped
    F = outer.Faces
    faces = g.FaceArray(F.VertexIndices[:, ::-1], F.NormalIndices[:, ::-1], F.UVs[:, ::-1], F.TextureIndex, F.Textures)
    mesh = g.NewMesh(outer.Vertices, outer.VertexNormals, faces)
    o = {"BackfaceCulling": False}
    o.update(opt)
    return SceneDef(w, h, [_obj(mesh, s=(3, 3, 3))], g.Camera(Position=(0.3, 0.2, 0.4), Direction=(0.2, -0.1, -1)), o)


def multi_object(w=800, h=600, **opt):
    """Overlapping objects incl. two coincident copies (exact depth ties => submission order),
    one object behind the camera (Outside) and one crossing the right plane (Intersect)."""
    cube = workloads.cube()
    sph = geometry.geodesic_sphere(8)
    suz = workloads.suzanne()
    solid = g.NewColorTexture((30, 160, 220, 255))
    sph_solid = geometry.geodesic_sphere(5, True, solid)
    objs = [
        _obj(suz, t=(-1.5, 0, 0), r=(0, 0.4, 0)),
        _obj(cube, t=(1.2, -0.3, -1.0), r=(0.5, 0.7, 0.2), s=(0.8, 0.8, 0.8)),
        _obj(sph, t=(0.2, 0.4, 0.5)),
        _obj(sph, t=(0.2, 0.4, 0.5)),               # coincident: z ties, later object wins
        _obj(sph_solid, t=(0.9, 0.9, 1.0), s=(0.5, 0.5, 0.5)),
        _obj(cube, t=(0, 0, 9)),                    # behind the camera
        _obj(suz, t=(3.6, 0.5, 0), r=(0, -0.6, 0)),  # crosses the right plane
    ]
    return SceneDef(w, h, objs, geometry.default_camera(), opt)


def tiny_far(w=320, h=240, **opt):
    """A far-away dense sphere: most triangles snap to a single pixel (the all-three-biased
    degenerate case of rasterizer.go:120-129 still writes it)."""
    return SceneDef(w, h, [_obj(geometry.geodesic_sphere(24), t=(0, 0, -20))], geometry.default_camera(), opt)


def tiny_far_dense(w=320, h=240, **opt):
    """72 000-triangle sphere squeezed into a handful of device tiles: a tile's list grows past its
    in-place descriptor capacity, so the overflow list is exercised."""
    return SceneDef(w, h, [_obj(geometry.geodesic_sphere(60), t=(0, 0, -12))], geometry.default_camera(), opt)


def stacked_quads(w=96, h=96, layers=700, **opt):
    """`layers` screen-filling quads at different depths plus coincident duplicates (depth ties): more
    large triangles per tile than the block-wide queue holds, so the warp-sweep fallback runs too."""
    verts, faces = [], []
    for k in range(layers):
        z = -1.0 - 0.01 * (k % 350)          # the second half repeats the depths of the first: exact ties
        s = 4.0 + 0.001 * k
        b = len(verts)
        verts += [(-s, -s, z, 1), (s, -s, z, 1), (s, s, z, 1), (-s, s, z, 1)]
        faces += [(b, b + 1, b + 2), (b, b + 2, b + 3)]
    mesh = g.NewMesh(np.array(verts, np.float32), None, g.FaceArray(np.array(faces, np.int32)))
    return SceneDef(w, h, [_obj(mesh)], g.Camera(Position=(0.3, 0.2, 3.0)), opt)


def odd_size(w=333, h=211, **opt):
    """Width not a multiple of 4 and neither a multiple of the tile: scalar write-back path,
    ragged reference tiles."""
    return SceneDef(w, h, [_obj(workloads.suzanne(), r=(0.2, 0.3, 0))], geometry.default_camera(), opt)


def empty_scene(w=256, h=128, **opt):
    return SceneDef(w, h, [], geometry.default_camera(), opt)


def big_triangles(w=1280, h=720, **opt):
    """Scaled cube filling the screen: exercises the big-triangle list."""
    return SceneDef(w, h, [_obj(workloads.cube(), r=(0.3, 0.4, 0), s=(1.6, 1.6, 1.6))],
                    g.Camera(Position=(0, 0, 4.0)), opt)


def offscreen_no_clip(w=640, h=360, **opt):
    """FrustumClipping off with objects hanging over every screen edge (all in front of the camera, so
    |snapped coordinate| stays far inside the 16383 domain): negative and > width coordinates reach the
    tile test, the raster bbox clamps and — with overlays — FrameBuffer.Pixel's linear-index wrap."""
    suz = workloads.suzanne()
    objs = [
        _obj(suz, t=(-3.4, 0.3, 0.0), r=(0, 0.5, 0)),     # over the left edge
        _obj(suz, t=(3.5, -0.2, 0.5), r=(0, -0.4, 0)),    # over the right edge
        _obj(suz, t=(0.3, 2.0, 0.0), r=(0.3, 0, 0)),      # over the top
        _obj(suz, t=(-0.4, -2.1, 1.0), r=(-0.2, 0.2, 0)),  # over the bottom
        _obj(workloads.cube(), t=(0, 0, -2.0), r=(0.4, 0.3, 0.1)),
    ]
    o = {"FrustumClipping": False}
    o.update(opt)
    return SceneDef(w, h, objs, geometry.default_camera(), o)


PINNED: Dict[str, Callable[[], SceneDef]] = {
    "c1_suzanne_720p": c1,
    "c1_suzanne_800x600": lambda: c1(800, 600),
    "c1_serial_tiles1": lambda: SceneDef(640, 360, *workloads.config_c1(), parallel=False),
    "c2_cube_poseA": lambda: c2("A"),
    "c2_cube_poseB": lambda: c2("B"),
    "c2_no_textures": lambda: c2("B", ShowTextures=False),
    "c2_flat": lambda: c2("B", FlatShading=True),
    "c2_no_light": lambda: c2("A", Lighting=False),
    "c3_sphere_n20": lambda: c3(20),
    "c3_sphere_n100": lambda: c3(100),
    "c3_sphere_n30_clip": lambda: c3(30, cam_z=3.0),
    "c3_no_cull": lambda: c3(10, BackfaceCulling=False),
    "c4_small": c4,
    "gouraud_textured": gouraud_sphere,
    "gouraud_untextured": lambda: gouraud_sphere(textured=False),
    "npot_texture": npot_texture,
    "inside_sphere": inside_sphere,
    "multi_object": multi_object,
    "tiny_far": tiny_far,
    "tiny_far_dense": tiny_far_dense,
    "stacked_quads": stacked_quads,
    "odd_size": odd_size,
    "empty": empty_scene,
    "big_triangles": big_triangles,
    "no_faces": lambda: c1(640, 360, ShowFaces=False),
    "no_clipping_inside": lambda: c1(640, 360, FrustumClipping=False),
    "offscreen_no_clip": offscreen_no_clip,
    "offscreen_no_clip_serial": lambda: SceneDef(333, 211, offscreen_no_clip().objects, geometry.default_camera(),
                                                 {"FrustumClipping": False}, parallel=False),
    "offscreen_no_clip_wire": lambda: offscreen_no_clip(ShowEdges=True, ShowVertices=True),
    # overlays and post passes (renderer.go:191-216, 476-480)
    "wire_c1": lambda: c1(640, 360, ShowEdges=True),
    "wire_only_c1": lambda: c1(640, 360, ShowEdges=True, ShowFaces=False),
    "verts_c1": lambda: c1(640, 360, ShowVertices=True),
    "verts_only_c1": lambda: c1(333, 211, ShowVertices=True, ShowFaces=False),
    "wire_verts_c2A": lambda: c2("A", ShowEdges=True, ShowVertices=True),
    "wire_c2B_serial": lambda: SceneDef(800, 600, *workloads.config_c2("B"), options=dict(ShowEdges=True), parallel=False),
    "wire_verts_multi_object": lambda: multi_object(ShowEdges=True, ShowVertices=True),
    "wire_verts_sphere_n20": lambda: c3(20, ShowEdges=True, ShowVertices=True),
    "wire_inside_sphere": lambda: inside_sphere(ShowEdges=True, ShowVertices=True),
    "wire_odd_size_crosshair": lambda: odd_size(ShowEdges=True, ShowVertices=True, CrossHair=True),
    "wire_tiny_far": lambda: tiny_far(ShowEdges=True),
    "crosshair_c1": lambda: c1(640, 360, CrossHair=True),
    "crosshair_empty_tiny": lambda: empty_scene(7, 9, CrossHair=True),
    "fog_tiny_far": lambda: tiny_far(Fog=True, CrossHair=True),
    "fog_c1_custom": lambda: c1(640, 360, Fog=True, FogStart=np.float32(0.25), FogEnd=np.float32(0.17),
                                FogColor=(10, 200, 90, 128)),
    "fog_wire_gouraud": lambda: gouraud_sphere(ShowEdges=True, Fog=True, FogStart=np.float32(0.7), FogEnd=np.float32(0.4)),
    # affine texture mapping: this repository's own definition (GRB_OPT_AFFINE_TEXTURES) — no code path in the reference
    "affine_c2_poseB": lambda: c2("B", AffineTextures=True),
    "affine_gouraud_textured": lambda: gouraud_sphere(AffineTextures=True),
    "affine_npot_wire": lambda: npot_texture(AffineTextures=True, ShowEdges=True),
}
