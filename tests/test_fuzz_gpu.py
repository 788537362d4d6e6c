"""Seeded random scenes: triangle soups with random normals / UVs / textures, random object
transforms, cameras (inside, outside, grazing the frustum planes) and option combinations,
GPU (through the C ABI) against the oracle, bit for bit."""
import os

import numpy as np
import pytest

import gorender_b200 as g
from gorender_b200 import workloads

pytestmark = pytest.mark.gpu


def random_mesh(rng, ntri, spread, with_normals, textures):
    nv = max(3, ntri // 2 + 3)
    verts = np.ones((nv, 4), np.float32)
    verts[:, :3] = rng.normal(0, spread, (nv, 3)).astype(np.float32)
    vidx = rng.integers(0, nv, (ntri, 3)).astype(np.int32)
    # some shared-edge strips and some degenerate faces (repeated vertices)
    vidx[::7, 2] = vidx[::7, 1]
    vn = None
    nidx = None
    if with_normals:
        nvn = nv + 5
        vn = np.ones((nvn, 4), np.float32)
        vn[:, :3] = rng.normal(0, 1, (nvn, 3)).astype(np.float32)
        nidx = rng.integers(0, nvn, (ntri, 3)).astype(np.int32)
    uvs = rng.uniform(-1.5, 2.5, (ntri, 3, 2)).astype(np.float32)
    tex = rng.integers(-1, len(textures), ntri).astype(np.int32) if textures else None
    return g.NewMesh(verts, vn, g.FaceArray(vidx, nidx, uvs, tex, textures))


def random_textures(rng):
    out = [workloads.checker_texture(int(rng.choice([8, 32, 64])), 4)]
    h, w = int(rng.integers(3, 40)), int(rng.integers(3, 40))
    img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    t = g.NewImageTexture(img)            # usually non-power-of-two
    t.SetScale(float(rng.choice([0.5, 1.0, 3.0])))
    out.append(t)
    out.append(g.NewColorTexture(tuple(int(x) for x in rng.integers(0, 256, 4))))
    return out


# GORENDER_FUZZ_SEEDS=N widens the campaign (300 seeds were run clean before the round-1 hand-in)
@pytest.mark.parametrize("seed", range(int(os.environ.get("GORENDER_FUZZ_SEEDS", "24"))))
def test_random_scene(seed, device, oracle):
    rng = np.random.default_rng(1000 + seed)
    w = int(rng.choice([64, 97, 160, 256, 333, 640]))
    h = int(rng.choice([48, 75, 120, 200, 211, 360]))
    textures = random_textures(rng) if rng.random() < 0.7 else []
    objs = []
    for _ in range(int(rng.integers(1, 5))):
        ntri = int(rng.choice([1, 5, 40, 300, 1500]))
        mesh = random_mesh(rng, ntri, float(rng.choice([0.3, 1.0, 3.0])), rng.random() < 0.5, textures)
        o = g.NewObject(mesh)
        o.Translation = rng.normal(0, 1.5, 3).astype(np.float32)
        o.Rotation = (rng.uniform(-3, 3, 3) * (rng.random(3) < 0.7)).astype(np.float32)
        o.Scale = rng.choice([0.2, 1.0, 2.5], 3).astype(np.float32)
        objs.append(o)
    cam = g.Camera(Position=rng.normal(0, 2.5, 3).astype(np.float32) + np.array([0, 0, 3], np.float32),
                   Direction=(rng.normal(0, 0.4, 3) + np.array([0, 0, -1])).astype(np.float32), Up=(0, 1, 0))
    fb = g.FrameBuffer(w, h, 1, device)
    r = g.Renderer(fb, parallel=bool(rng.random() < 0.8))
    r.BackfaceCulling = bool(rng.random() < 0.6)
    r.Lighting = bool(rng.random() < 0.8)
    r.FlatShading = bool(rng.random() < 0.3)
    r.ShowTextures = bool(rng.random() < 0.8)
    # overlays / post passes on a third of the scenes (drawn after the other choices so that the
    # scenes of the seeds without them stay what they were)
    orng = np.random.default_rng(5000 + seed)
    if seed % 3 == 2:
        r.ShowEdges = bool(orng.random() < 0.7)
        r.ShowVertices = bool(orng.random() < 0.6)
        r.ShowFaces = bool(orng.random() < 0.7)
        r.CrossHair = bool(orng.random() < 0.5)
        r.Fog = bool(orng.random() < 0.4)
        r.FogStart, r.FogEnd = np.float32(orng.uniform(0.3, 1.0)), np.float32(orng.uniform(0.05, 0.3))
    if seed % 4 == 1:
        r.AffineTextures = True      # this repository's own mode (no reference code path): CUDA and oracle must still agree
    r.Draw(objs, cam)
    ref = oracle.draw(r, objs, cam)
    assert int(r.last_stats["out_of_domain"][0]) == 0
    assert r.TPF == ref["tpf"]
    same = (fb.Pixels == ref["pixels"]).all(axis=-1) & (fb.ZBuffer.view(np.uint32) == ref["zbuffer"].view(np.uint32))
    if not same.all():
        y, x = np.argwhere(~same)[0]
        raise AssertionError(f"seed {seed}: {(~same).sum()} differing pixels of {same.size}; first (x={x}, y={y}): "
                             f"gpu {fb.Pixels[y, x].tolist()} z={fb.ZBuffer[y, x]!r} oracle {ref['pixels'][y, x].tolist()} "
                             f"z={ref['zbuffer'][y, x]!r}")


@pytest.mark.parametrize("seed", range(int(os.environ.get("GORENDER_FUZZ_DENSE_SEEDS", "4"))))
def test_random_dense_scene(seed, device, oracle):
    """Dense meshes of tiny triangles (jittered spheres, two of them coincident: exact depth ties decided by
    submission order), random resolution / camera / options: long tile lists, overflow descriptors, the 4x4
    coverage masks and the fine-stage ring under load."""
    from gorender_b200 import geometry

    rng = np.random.default_rng(9000 + seed)
    w = int(rng.choice([320, 640, 801, 1280]))
    h = int(rng.choice([240, 360, 455, 720]))
    base = geometry.geodesic_sphere(int(rng.choice([24, 36, 48])), bool(rng.random() < 0.5),
                                    workloads.checker_texture(32) if rng.random() < 0.5 else None)
    verts = base.Vertices.copy()
    verts[:, :3] += rng.normal(0, 0.004, (len(verts), 3)).astype(np.float32)
    F = base.Faces
    mesh = g.NewMesh(verts, base.VertexNormals if len(base.VertexNormals) else None,
                     g.FaceArray(F.VertexIndices, F.NormalIndices, F.UVs, F.TextureIndex, F.Textures))
    objs = []
    for k in range(3):
        o = g.NewObject(mesh)
        o.Translation = (rng.normal(0, 0.6, 3) + np.array([0, 0, -1.0 * k])).astype(np.float32)
        o.Rotation = rng.uniform(-3, 3, 3).astype(np.float32)
        s = float(rng.choice([0.3, 0.8, 1.5]))
        o.Scale = np.array([s, s, s], np.float32)
        objs.append(o)
    twin = g.NewObject(mesh)                       # coincident with the first object: z ties everywhere
    twin.Translation, twin.Rotation, twin.Scale = objs[0].Translation.copy(), objs[0].Rotation.copy(), objs[0].Scale.copy()
    objs.append(twin)
    cam = g.Camera(Position=(rng.normal(0, 0.5, 3) + np.array([0, 0, float(rng.choice([2.5, 5.0, 12.0]))])).astype(np.float32),
                   Direction=(rng.normal(0, 0.15, 3) + np.array([0, 0, -1])).astype(np.float32), Up=(0, 1, 0))
    fb = g.FrameBuffer(w, h, 1, device)
    r = g.Renderer(fb, parallel=bool(rng.random() < 0.8))
    r.BackfaceCulling = bool(rng.random() < 0.7)
    r.FlatShading = bool(rng.random() < 0.3)
    r.ShowEdges = bool(rng.random() < 0.25)
    r.Draw(objs, cam)
    ref = oracle.draw(r, objs, cam)
    assert int(r.last_stats["out_of_domain"][0]) == 0
    assert r.TPF == ref["tpf"]
    same = (fb.Pixels == ref["pixels"]).all(axis=-1) & (fb.ZBuffer.view(np.uint32) == ref["zbuffer"].view(np.uint32))
    assert same.all(), f"seed {seed}: {(~same).sum()} differing pixels of {same.size}; first {np.argwhere(~same)[0]}"
