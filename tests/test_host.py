"""CPU tests of the host layer (the Python mirror of the reference's load-time
and per-object host code): matrices, mesh precompute, OBJ/MTL/scene/texture
loading, multi-GPU partition arithmetic."""
import json
import os

import numpy as np
import pytest

import gorender_b200 as g
import gorender_b200.vecmath as vm
from gorender_b200 import geometry, workloads
from gorender_b200.obj import parse_f32
from gorender_b200.parallel import pose_block, strip_rows
from gorender_b200.texture import premultiply_nrgba

REF_MODELS = "/root/reference/models"
have_ref = pytest.mark.skipif(not os.path.isdir(REF_MODELS), reason="reference checkout not present on this machine")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_matrices_match_oracle_bitwise(oracle):
    rng = np.random.default_rng(2)
    for _ in range(50):
        s, r, t = (rng.uniform(-3, 3, 3).astype(np.float32) for _ in range(3))
        if rng.random() < 0.3:
            r[rng.integers(3)] = 0  # exact-zero rotation short-circuit (matrix.go:33-35)
        w = vm.NewWorldMatrix(s, r, t)
        assert np.array_equal(bits(w), bits(oracle.world_matrix(s, r, t)))
        eye, d = rng.uniform(-5, 5, 3).astype(np.float32), rng.uniform(-1, 1, 3).astype(np.float32)
        v = vm.NewViewMatrix(eye, d, [0, 1, 0])
        assert np.array_equal(bits(v), bits(oracle.view_matrix(eye, d, [0, 1, 0])))
        fov, asp = np.float32(rng.uniform(0.3, 2)), np.float32(rng.uniform(0.5, 2.5))
        p = vm.NewPerspectiveMatrix(fov, asp, 0.0, 50.0)
        assert np.array_equal(bits(p), bits(oracle.perspective_matrix(fov, asp, 0.0, 50.0)))
        assert np.array_equal(bits(vm.mvp_matrix(p, v, w)), bits(oracle.mvp_matrix(p, v, w)))
    assert np.array_equal(bits(vm.NewScreenMatrix(1280, 720)), bits(oracle.screen_matrix(1280, 720)))
    assert np.array_equal(bits(vm.light_direction()), bits(oracle.light_direction()))


def test_new_mesh_precompute_matches_oracle(oracle):
    for mesh in (workloads.suzanne(), geometry.geodesic_sphere(7)):
        assert np.array_equal(bits(mesh.FaceNormals), bits(oracle.face_normals(mesh.Vertices, mesh.Faces.VertexIndices)))
        assert np.array_equal(bits(mesh.BoundingBox), bits(oracle.bounding_box(mesh.Vertices)))
    # corner order of mesh.go:41-50
    bb = workloads.cube().BoundingBox
    assert bb[0].tolist() == [-1, -1, -1, 1] and bb[1].tolist() == [-1, -1, 1, 1] and bb[4].tolist() == [1, -1, -1, 1]


def test_geodesic_sphere_counts():
    m = geometry.geodesic_sphere(100)
    assert m.Vertices.shape == (100002, 4) and len(m.Faces) == 200000 and len(m.VertexNormals) == 0
    v = m.Vertices[:, :3].astype(np.float64)
    assert np.allclose(np.linalg.norm(v, axis=1), 1.0, atol=1e-6)
    f = m.Faces.VertexIndices
    n = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    assert (np.einsum("ij,ij->i", n, v[f[:, 0]]) > 0).all()  # outward CCW


@have_ref
def test_obj_loader_on_reference_models():
    suz = g.LoadMeshFile(os.path.join(REF_MODELS, "suzanne.obj"), False)
    assert len(suz) == 1 and suz[0].Vertices.shape == (507, 4) and len(suz[0].Faces) == 967
    assert len(suz[0].VertexNormals) == 0 and (suz[0].Faces.TextureIndex == -1).all()
    cube = g.LoadMeshFile(os.path.join(REF_MODELS, "cube.obj"), False)[0]
    assert cube.Vertices.shape == (8, 4) and cube.VertexNormals.shape == (6, 4) and len(cube.Faces) == 12
    assert len(cube.Faces.Textures) == 1  # Top and Side share textures-16.png (obj.go:236-237)
    t = cube.Faces.Textures[0]
    assert (t.width, t.height, t.typ) == (512, 512, g.TextureTypeImageFast)
    assert cube.Faces.VertexIndices[0].tolist() == [3, 5, 7] and cube.Faces.NormalIndices[0].tolist() == [5, 5, 5]
    assert cube.Faces.UVs[0, 0].tolist() == [np.float32(0.062641), np.float32(0.499954)]
    # the committed fixtures are exactly what the loader produces
    for name, mesh in (("suzanne", suz[0]), ("cube", cube)):
        fx = workloads.load_mesh_fixture(os.path.join(workloads.ASSETS_DIR, name + ".npz"))
        assert np.array_equal(bits(fx.Vertices), bits(mesh.Vertices))
        assert np.array_equal(fx.Faces.VertexIndices, mesh.Faces.VertexIndices)
        assert np.array_equal(bits(fx.Faces.UVs), bits(mesh.Faces.UVs))
        assert np.array_equal(bits(fx.FaceNormals), bits(mesh.FaceNormals))
    assert np.array_equal(workloads.cube().Faces.Textures[0].pixels, t.pixels)


@have_ref
def test_reference_texture_is_premultiplied():
    from PIL import Image

    raw = np.asarray(Image.open(os.path.join(REF_MODELS, "textures-16.png")).convert("RGBA"))
    t = g.LoadTextureFile(os.path.join(REF_MODELS, "textures-16.png"))
    assert (t.pixels[raw[..., 3] == 0][:, :3] == 0).all()
    assert np.array_equal(t.pixels[raw[..., 3] == 255], raw[raw[..., 3] == 255])


def test_premultiply_formula():
    """color.RGBAModel.Convert on NRGBA: ((c*0x101)*a/0xff)>>8, alpha kept (texture.go:57)."""
    px = np.array([[[255, 128, 1, 128], [10, 20, 30, 0], [200, 100, 50, 255], [255, 255, 255, 1]]], np.uint8)
    out = premultiply_nrgba(px)
    assert out[0, 0].tolist() == [128, 64, 0, 128]
    assert out[0, 1].tolist() == [0, 0, 0, 0]
    assert out[0, 2].tolist() == [200, 100, 50, 255]
    assert out[0, 3].tolist() == [1, 1, 1, 1]


def test_parse_f32_is_correctly_rounded():
    toks = ["0.437500", "-1.367188", "1e-45", "3.4028235e38", "0.1", "16777217"]
    got = parse_f32(toks)
    want = np.array([np.float32(t) for t in toks])
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # a decimal just above a float32 midpoint: double-then-narrow rounds down, strtof rounds up
    tie = "1.00000005960464477539062500000000000000000001"
    assert parse_f32([tie])[0] == np.float32(1.0000001)


def test_obj_roundtrip_multi_object_and_quirks(tmp_path):
    a = geometry.geodesic_sphere(2)
    b = geometry.geodesic_sphere(3, True)
    pa, pb = tmp_path / "a.obj", tmp_path / "b.obj"
    geometry.write_obj(a, str(pa), "A")
    geometry.write_obj(b, str(pb), "B")
    # second object with indices continuing after the first (obj.go:31-40 offsets)
    lines_b = []
    for ln in open(pb).read().splitlines():
        if ln.startswith("f "):
            parts = []
            for tok in ln.split(" ")[1:]:
                v, vt, vn = (int(x) for x in tok.split("/"))
                parts.append(f"{v + len(a.Vertices)}/{vt}/{vn}")
            ln = "f " + " ".join(parts)
        lines_b.append(ln)
    both = tmp_path / "both.obj"
    both.write_text(open(pa).read() + "\n".join(lines_b) + "\n")
    meshes = g.LoadObjFile(str(both), False)
    assert len(meshes) == 2
    for got, want in zip(meshes, (a, b)):
        assert np.array_equal(bits(got.Vertices), bits(want.Vertices))
        assert np.array_equal(got.Faces.VertexIndices, want.Faces.VertexIndices)
        assert np.array_equal(bits(got.FaceNormals), bits(want.FaceNormals))
    assert np.array_equal(meshes[1].Faces.NormalIndices, b.Faces.NormalIndices)
    assert np.array_equal(bits(meshes[1].Faces.UVs), bits(b.Faces.UVs))
    single = g.LoadObjFile(str(both), True)
    assert len(single) == 1 and len(single[0].Vertices) == len(a.Vertices) + len(b.Vertices)

    quirk = tmp_path / "q.obj"
    quirk.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nvn 0 1 0\nvn 1 0 0\nf 1//1 2//2 3//3\n")
    m = g.LoadObjFile(str(quirk), False)[0]
    assert m.Faces.NormalIndices[0].tolist() == [0, 2, -1]  # obj.go:77-89
    quad = tmp_path / "quad.obj"
    quad.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0\nf 1 2 3 4\n")
    with pytest.raises(ValueError, match="not triangulated"):
        g.LoadObjFile(str(quad), False)
    empty = tmp_path / "e.obj"
    empty.write_text("# nothing\n")
    with pytest.raises(ValueError, match="does not have any vertices"):
        g.LoadObjFile(str(empty), False)
    with pytest.raises(ValueError, match="unsupported mesh format"):
        g.LoadMeshFile(str(tmp_path / "x.stl"), False)


def test_mtl_materials(tmp_path):
    from PIL import Image

    Image.fromarray(np.full((4, 4, 4), 200, np.uint8), "RGBA").save(tmp_path / "t.png")
    (tmp_path / "m.mtl").write_text("newmtl A\nmap_Kd t.png\n\nnewmtl B\nKd 1 1 1\n")
    (tmp_path / "m.obj").write_text(
        "mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 0 1\n"
        "f 1/1 2/2 3/3\nusemtl A\nf 1/1 2/2 3/3\nusemtl B\nf 1/1 2/2 3/3\nusemtl nope\nf 1/1 2/2 3/3\n")
    m = g.LoadObjFile(str(tmp_path / "m.obj"), False)[0]
    F = m.Faces
    assert F.TextureIndex.tolist() == [-1, 0, 1, -1]            # no material / unknown usemtl => nil (H18)
    assert F.Textures[0].typ == g.TextureTypeImageFast
    assert F.Textures[1].typ == g.TextureTypeSolidColor and F.Textures[1].color == (255, 0, 255, 255)


def test_scene_file(tmp_path):
    geometry.write_obj(geometry.geodesic_sphere(2, True), str(tmp_path / "s.obj"))
    from PIL import Image

    Image.fromarray(np.full((8, 8, 4), 255, np.uint8), "RGBA").save(tmp_path / "t.png")
    (tmp_path / "scene.json").write_text(json.dumps({
        "name": "t",
        "meshes": [{"id": "a", "objFile": "s.obj", "texture": "t.png", "textureScale": 4}, {"id": "b", "objFile": "s.obj"}],
        "objects": [{"meshID": "a", "position": [1, 2, 3], "rotation": [0, 90, 0], "scale": [1, 1, 1]},
                    {"meshID": "b", "position": [0, 0, 0], "rotation": [0, 0, 0], "scale": [2, 2, 2]},
                    {"meshID": "a", "position": [0, 0, 0], "rotation": [0, 0, 0], "scale": [1, 1, 1]}]}))
    sc = g.LoadSceneFile(str(tmp_path / "scene.json"))
    assert sc.NumObjects() == 3 and sc.NumTriangles() == 3 * 80 and sc.NumVertices() == 3 * 42
    assert sc.Objects[0].Mesh is sc.Objects[2].Mesh
    assert sc.Objects[0].Mesh.Faces.Textures[0].scale == 4 and sc.Objects[0].Translation.tolist() == [1, 2, 3]
    assert sc.Objects[0].Rotation[1] == np.float32(90) * (vm.pi32 / np.float32(180))
    assert sc.Objects[1].Mesh.Faces.Textures[0].color == (200, 200, 200, 255)   # scene.go:73-74,102-106
    (tmp_path / "bad.json").write_text(json.dumps({"meshes": [], "objects": [{"meshID": "zz"}]}))
    with pytest.raises(RuntimeError, match="mesh id not found"):
        g.LoadSceneFile(str(tmp_path / "bad.json"))


def test_spin_rotations_accumulate_in_f32():
    r = geometry.spin_rotations(1000)
    acc = np.float32(0)
    for i in range(1000):
        assert r[i] == acc
        acc = np.float32(acc + np.float32(0.01))
    assert r[999] != np.float32(9.99)  # accumulated, not multiplied


def test_partition_arithmetic():
    for n, w in ((4096, 8), (10, 4), (3, 8), (0, 2)):
        blocks = [pose_block(n, w, r) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        assert max(e - b for b, e in blocks) - min(e - b for b, e in blocks) <= 1
    for h, w in ((2160, 8), (720, 4), (720, 8), (100, 8), (2160, 1)):
        strips = [strip_rows(h, w, r) for r in range(w)]
        assert strips[0][0] == 0 and strips[-1][1] == h
        assert all(strips[i][1] == strips[i + 1][0] for i in range(w - 1))
        assert all(b % 32 == 0 or b == h for b, _ in strips)


def test_balanced_strip_rows_tile_the_frame():
    """Sort-first strips balanced by per-row weights: contiguous, tile-aligned, cover the frame, and no rank gets
    much more than its share."""
    from gorender_b200.parallel import balanced_strip_rows, strip_rows

    rng = np.random.default_rng(3)
    for world in (1, 2, 3, 4, 8):
        for h in (2160, 720, 363, 64):
            n = (h + 31) // 32
            for trial in range(4):
                w = rng.integers(0, 50, n).astype(float)
                if trial == 0:
                    w[: n // 3] = 0
                    w[-(n // 4):] = 0
                rows = balanced_strip_rows(w, world, h)
                assert len(rows) == world and rows[0][0] == 0 and rows[-1][1] == h
                for (a0, a1), (b0, b1) in zip(rows, rows[1:]):
                    assert a1 == b0 and a0 <= a1
                assert all(y0 % 32 == 0 for y0, _ in rows)
                share = [w[y0 // 32:(y1 + 31) // 32].sum() for y0, y1 in rows]
                assert abs(sum(share) - w.sum()) < 1e-9
                if w.sum() > 0:
                    assert max(share) <= w.sum() / world + w.max() + 1e-9
            assert balanced_strip_rows(np.zeros(n), world, h) == [strip_rows(h, world, r) for r in range(world)]
