"""CPU tests of the oracle: known answers, internal twins, and the committed
framebuffer pins.  (The reference has no golden vectors of its own — SURVEY.md
§4 — so the KATs here are the self-consistency values of SURVEY.md §8c, which
came from an independent numpy restatement made during the survey.)
"""
import hashlib
import json
import os

import numpy as np
import pytest

import gorender_b200 as g
import gorender_b200.vecmath as vm
from gorender_b200 import geometry, workloads

import scene_defs

GOLDEN = json.load(open(os.path.join(workloads.GOLDEN_DIR, "oracle_outputs.json")))


def hexf(x):
    return format(int(np.float32(x).view(np.uint32)), "08x")


def test_matrix_known_answers(oracle):
    """SURVEY.md §8c KATs (float32 hex)."""
    fovy = np.float32(45 * (np.pi / 180))
    assert hexf(fovy) == "3f490fdb"
    assert hexf(vm.tan32(fovy / np.float32(2))) == "3ed413cd"
    p169 = oracle.perspective_matrix(fovy, np.float32(1280) / np.float32(720), 0, 50)
    p43 = oracle.perspective_matrix(fovy, np.float32(800) / np.float32(600), 0, 50)
    assert hexf(p169[1, 1]) == "401a8279" and hexf(p169[0, 0]) == "3fadd2c9" and hexf(p43[0, 0]) == "3fe7c3b6"
    assert p169[2].tolist() == [0, 0, 1, 0] and p169[3].tolist() == [0, 0, -1, 0]
    light = oracle.light_direction()
    assert [hexf(abs(c)) for c in light] == ["3f13cd3a"] * 3 and light[0] < 0 < light[1]
    view = oracle.view_matrix([0, 0, 5], [0, 0, -1], [0, 1, 0])
    assert view.tolist() == [[-1, 0, 0, 0], [0, 1, 0, 0], [0, 0, -1, 5], [0, 0, 0, 1]]


def test_asm_test_vector(oracle):
    """asm_test.go:12-14 matrix applied to (1,1,1,1)."""
    m = vm.NewIdentityMatrix()
    m = vm.Multiply(vm.NewRotationMatrix(0.1, 0.2, 0.3), m)
    m = vm.Multiply(vm.NewTranslationMatrix(1, 2, 3), m)
    out = oracle.matvec4_batch(m, np.array([[1, 1, 1, 1]], np.float32))
    assert out[0].tolist() == [np.float32(1.8453332), np.float32(3.159851), np.float32(3.9696174), 1.0]


def test_scalar_and_sse_twins_agree(oracle):
    """asm_amd64.s:33-40 and asm_purego.go:13-16 use the same association."""
    rng = np.random.default_rng(1)
    for _ in range(5):
        m = rng.standard_normal((4, 4)).astype(np.float32)
        v = (rng.standard_normal((4099, 4)) * 10 ** rng.uniform(-3, 3, (4099, 1))).astype(np.float32)
        a, b = oracle.matvec4_batch(m, v), oracle.matvec4_batch(m, v, sse=True)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # asm_test.go inputs
    m = vm.Multiply(vm.NewTranslationMatrix(1, 2, 3), vm.Multiply(vm.NewRotationMatrix(0.1, 0.2, 0.3), vm.NewIdentityMatrix()))
    v = np.repeat(np.arange(1000, dtype=np.float32)[:, None], 4, axis=1)
    assert np.array_equal(oracle.matvec4_batch(m, v).view(np.uint32), oracle.matvec4_batch(m, v, sse=True).view(np.uint32))
    assert oracle.matvec4_batch(m, np.zeros((0, 4), np.float32)).shape == (0, 4)


def test_flat_intensity_kat(oracle):
    """flat intensity of an untranslated +Z unit normal = 0.70412415
    (normalize4 with w = 1: length sqrt(2))."""
    tri = np.array([[-1, -1, 0, 1], [1, -1, 0, 1], [0, 1, 0, 1]], np.float32)
    mesh = g.NewMesh(tri, None, g.FaceArray([[0, 1, 2]]))
    assert mesh.FaceNormals[0].tolist() == [0, 0, 1, 1]
    r = scene_defs.SceneDef(64, 64, [g.NewObject(mesh)], geometry.default_camera()).renderer(None)
    out = oracle.draw(r, [g.NewObject(mesh)], geometry.default_camera(), record=True)
    assert len(out["triangles"]) == 1
    assert out["triangles"]["intensity"][0].tolist() == [np.float32(0.70412415)] * 3


def test_survey_counts(oracle):
    """C1 / C3 counts of SURVEY.md §8c (survivors, covered pixels, bbox class)."""
    sc = scene_defs.c1()
    out = oracle.draw(sc.renderer(None), sc.objects, sc.camera, record=True)
    assert out["visibility"].tolist() == [g._cabi.GRB_TILE and 2]
    assert len(out["triangles"]) == 654 and int((out["zbuffer"] > -1).sum()) == 82758
    sc = scene_defs.c3(100)
    out = oracle.draw(sc.renderer(None), sc.objects, sc.camera, record=True)
    assert out["visibility"].tolist() == [2]
    assert len(out["triangles"]) == 79502 and int((out["zbuffer"] > -1).sum()) == 98957
    sc = scene_defs.c3(100, cam_z=3.0)
    out = oracle.draw(sc.renderer(None), sc.objects, sc.camera, record=True)
    assert out["visibility"].tolist() == [1] and len(out["triangles"]) == 66414
    for pose, covered in (("A", 377280), ("B", 804239)):
        sc = scene_defs.c2(pose)
        out = oracle.draw(sc.renderer(None), sc.objects, sc.camera)
        assert int((out["zbuffer"] > -1).sum()) == covered


@pytest.mark.parametrize("name", list(scene_defs.PINNED))
def test_oracle_pins(name, oracle):
    sc = scene_defs.PINNED[name]()
    out = oracle.draw(sc.renderer(None), sc.objects, sc.camera)
    gold = GOLDEN[name]
    assert (out["tpf"], out["writes"], int((out["zbuffer"] > -1).sum())) == (gold["tpf"], gold["writes"], gold["covered"])
    assert hashlib.sha256(out["pixels"].tobytes()).hexdigest() == gold["pixels_sha256"]
    assert hashlib.sha256(out["zbuffer"].tobytes()).hexdigest() == gold["zbuffer_sha256"]


def test_threaded_mode_equals_serial_for_one_object(oracle):
    """The reference's parallel mode is only nondeterministic across objects."""
    sc = scene_defs.c1()
    r = sc.renderer(None)
    a = oracle.draw(r, sc.objects, sc.camera)
    b = oracle.draw(r, sc.objects, sc.camera, threads=4)
    assert a["tpf"] == b["tpf"] and np.array_equal(a["pixels"], b["pixels"])
    assert np.array_equal(a["zbuffer"].view(np.uint32), b["zbuffer"].view(np.uint32))


def test_background_and_dot_grid(oracle):
    """Clear + DotGrid (rasterizer.go:36-52): (100,100,100) at x,y multiples of 10, >= 10."""
    sc = scene_defs.empty_scene()
    out = oracle.draw(sc.renderer(None), [], sc.camera)
    px = out["pixels"]
    assert (out["zbuffer"] == -1).all() and out["tpf"] == 0
    assert px[0, 0].tolist() == [50, 50, 50, 255] and px[10, 10].tolist() == [100, 100, 100, 255]
    assert px[0, 10].tolist() == [50, 50, 50, 255] and px[20, 30].tolist() == [100, 100, 100, 255]
    assert int((px[..., 0] == 100).sum()) == (sc.height - 1) // 10 * ((sc.width - 1) // 10)


def test_box_visibility_quirk(oracle):
    """clipping.go:141-149 returns Intersect at the first plane with 1..7 corners out (H14)."""
    inside = np.array([[x, y, z, -5.0] for x in (-1, 1) for y in (-1, 1) for z in (1, 2)], np.float32)
    assert oracle.box_visibility(inside) == 2
    out_right = inside.copy()
    out_right[:, 0] += 100
    assert oracle.box_visibility(out_right) == 0
    # one corner beyond the left plane, everything beyond the far plane: still Intersect
    mixed = inside.copy()
    mixed[:, 2] = 60
    mixed[0, 0] = -100
    assert oracle.box_visibility(mixed) == 1


def test_clip_triangle_properties(oracle):
    # fully inside: six rotations of a triangle's vertex order are the identity (clipping.go:195-214)
    pts = np.array([[-0.5, -0.5, 1, -2], [0.5, -0.5, 1, -2], [0, 0.5, 1, -2]], np.float32)
    uvs = np.array([[0, 0], [1, 0], [0, 1]], np.float32)
    ins = np.array([0.2, 0.5, 0.9], np.float32)
    po, uo, io = oracle.clip_triangle(pts, uvs, ins)
    assert len(po) == 1 and np.array_equal(po[0], pts) and np.array_equal(uo[0], uvs) and np.array_equal(io[0], ins)
    # fully outside one plane: nothing
    far = pts.copy()
    far[:, 2] = 60
    assert len(oracle.clip_triangle(far, uvs, ins)[0]) == 0
    # one vertex beyond the left plane (x < w, w negative): a quad => two triangles sharing vertex 0
    cut = pts.copy()
    cut[0, 0] = -5
    po, uo, io = oracle.clip_triangle(cut, uvs, ins)
    assert len(po) == 2 and np.array_equal(po[0][0], po[1][0]) and np.array_equal(po[0][2], po[1][1])
    for tri in po:
        for p in tri:
            assert p[0] + p[3] <= 1e-6 * abs(p[3]) + 1e-6  # inside-or-on the left plane: (q-P).N = x + w <= 0


def test_texture_sample_wrap(oracle):
    """texture.go:73-88: int() truncates toward zero, & wraps negatives for pow-2, % + clamp otherwise."""
    img = np.zeros((4, 8, 4), np.uint8)
    img[..., 0] = np.arange(8)[None, :]
    img[..., 1] = np.arange(4)[:, None]
    img[..., 3] = 255
    fast = g.NewImageTexture(img)
    assert fast.typ == g.TextureTypeImageFast
    # u = 1 - (x + .5)/8 -> x ; v = (y + .5)/4 -> y
    for x, y in ((0, 0), (7, 3), (3, 2)):
        assert oracle.texture_sample(fast, 1 - (x + 0.5) / 8, (y + 0.5) / 4)[:2] == (x, y)
    # negative v: int(-0.375*4) = -1 -> & 3 = 3
    assert oracle.texture_sample(fast, 0.9, -0.375)[1] == 3
    # u > 1: (1-u) negative: int(-0.25*8) = -2 -> & 7 = 6
    assert oracle.texture_sample(fast, 1.25, 0.1)[0] == 6
    slow = g.NewImageTexture(np.ascontiguousarray(img[:3, :5]))
    assert slow.typ == g.TextureTypeImage
    assert oracle.texture_sample(slow, 1 - 6.5 / 5, 0.1)[:2] == (1, 0)      # 6 % 5
    assert oracle.texture_sample(slow, 0.9, -0.4)[:2] == (0, 0)            # negative idx clamps to texel 0
    solid = g.NewColorTexture((1, 2, 3, 4))
    assert oracle.texture_sample(solid, 0.3, 0.3) == (1, 2, 3, 4)


def test_oracle_against_reference_c_prototype(oracle):
    """oracle/_ref/libref_cmatrix.so is the reference's own c/matrix_amd64.c (SSE prototype of
    matrixMultiplyVec4Batch), compiled unmodified by oracle/build_ref.sh.  It sums
    (p1+p2)+(p3+p4) while the Go assembly sums ((p1+p2)+p3)+p4, so: exact
    agreement where the two orders coincide (z == 0), a few ulp elsewhere."""
    import ctypes as C

    path = os.path.join(os.path.dirname(workloads.GOLDEN_DIR), "..", "oracle", "_ref", "libref_cmatrix.so")
    path = os.path.normpath(path)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built (no reference checkout on this machine)")
    ref = C.CDLL(path)
    ref.matrix_multiply_vec4.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    rng = np.random.default_rng(5)
    m = vm.Multiply(vm.NewTranslationMatrix(1, 2, 3), vm.Multiply(vm.NewRotationMatrix(0.1, 0.2, 0.3), vm.NewIdentityMatrix()))
    cols = np.ascontiguousarray(m.T)                      # the prototype takes the matrix column-major
    buf = (C.c_char * 80)()
    base = (C.addressof(buf) + 15) & ~15                 # __m128 loads need 16-byte alignment
    C.memmove(base, cols.ctypes.data, 64)
    v = (rng.standard_normal((5000, 4)) * 10).astype(np.float32)
    v[:2500, 2] = 0.0                                     # z == 0: both summation orders give the same bits
    want = oracle.matvec4_batch(m, v)
    got = np.ascontiguousarray(v.copy())
    ref.matrix_multiply_vec4(C.c_void_p(base), C.c_void_p(got.ctypes.data), len(got))
    assert np.array_equal(got[:2500].view(np.uint32), want[:2500].view(np.uint32))
    scale = np.abs(m).max() * np.abs(v).max(axis=1, keepdims=True)
    assert (np.abs(got - want) <= 4 * np.finfo(np.float32).eps * scale).all()


def test_go_reference_pin_if_present(oracle):
    """tests/golden/go_reference_outputs.json is written by scripts/compare_go_dump.py --pin from a run of the UNMODIFIED Go
    reference (go/parity/parity_dump.go).  It does not exist until somebody with a Go toolchain has produced it: until
    then parity stays "unpinned"  and this test has nothing to check."""
    import hashlib
    import json
    import os

    import scene_defs
    from gorender_b200 import workloads

    path = os.path.join(workloads.GOLDEN_DIR, "go_reference_outputs.json")
    if not os.path.exists(path):
        pytest.skip("no output of the Go reference has been committed yet")
    ref = json.load(open(path))
    assert ref, "empty pin"
    for name, gref in ref.items():
        sc = scene_defs.PINNED[name]()
        res = oracle.draw(sc.renderer(None), sc.objects, sc.camera)
        assert res["tpf"] == gref["tpf"], name
        assert hashlib.sha256(res["pixels"].tobytes()).hexdigest() == gref["pixels_sha256"], name
        assert hashlib.sha256(res["zbuffer"].tobytes()).hexdigest() == gref["zbuffer_sha256"], name


def test_go_harness_scene_export_round_trips(oracle, tmp_path):
    """The scene files handed to the Go harness carry exactly what the oracle is fed: read back with the byte layout
    go/parity/parity_dump.go's loadMesh uses, rebuilt with NewMesh and rendered, they give the same frames."""
    import json
    import struct
    import subprocess
    import sys

    import gorender_b200 as g
    import scene_defs

    names = ["c2_cube_poseB", "multi_object", "fog_wire_gouraud", "c1_serial_tiles1"]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run([sys.executable, os.path.join(root, "scripts", "export_go_scenes.py"), str(tmp_path)] + names, check=True,
                   capture_output=True)
    scenes = json.load(open(tmp_path / "scenes.json"))
    assert [s["name"] for s in scenes] == names

    def load_mesh(fn):
        b = open(tmp_path / fn, "rb").read()
        nv, nvn, nf, ntex = struct.unpack_from("<4i", b, 0)
        off = 16

        def take(dtype, count, shape):
            nonlocal off
            a = np.frombuffer(b, dtype, count, off).reshape(shape).copy()
            off += a.nbytes
            return a

        verts, vns = take("<f4", nv * 4, (nv, 4)), take("<f4", nvn * 4, (nvn, 4))
        vidx, nidx = take("<i4", nf * 3, (nf, 3)), take("<i4", nf * 3, (nf, 3))
        uvs, tidx = take("<f4", nf * 6, (nf, 3, 2)), take("<i4", nf, (nf,))
        texs = []
        for _ in range(ntex):
            typ, w, h, scale, c0, c1, c2, c3 = struct.unpack_from("<3if4B", b, off)
            off += 20
            px = None
            if typ != g.TextureTypeSolidColor:
                px = take(np.uint8, w * h * 4, (h, w, 4))
            texs.append(g.Texture(typ, (c0, c1, c2, c3), px, scale))
        assert off == len(b)
        return g.NewMesh(verts, vns if nvn else None, g.FaceArray(vidx, nidx, uvs, tidx, texs))

    for s in scenes:
        want_sc = scene_defs.PINNED[s["name"]]()
        objs = []
        meshes = [load_mesh(fn) for fn in s["meshes"]]
        for so in s["objects"]:
            o = g.NewObject(meshes[so["mesh"]])
            o.Translation, o.Rotation, o.Scale = (np.array(so[k], np.float32) for k in ("translation", "rotation", "scale"))
            objs.append(o)
        cam = g.Camera(s["camera"]["position"], s["camera"]["direction"], s["camera"]["up"])
        r = scene_defs.SceneDef(s["width"], s["height"], objs, cam, {k: v for k, v in s["options"].items()}, parallel=s["num_tiles"] == 16).renderer(None)
        r.FogStart, r.FogEnd, r.FogColor = np.float32(s["fog_start"]), np.float32(s["fog_end"]), tuple(s["fog_color"])
        got = oracle.draw(r, objs, cam)
        want = oracle.draw(want_sc.renderer(None), want_sc.objects, want_sc.camera)
        assert got["tpf"] == want["tpf"], s["name"]
        assert np.array_equal(got["pixels"], want["pixels"]) and np.array_equal(got["zbuffer"].view(np.uint32), want["zbuffer"].view(np.uint32)), s["name"]


def test_affine_mode_is_the_perspective_one_where_w_is_constant(oracle):
    """GRB_OPT_AFFINE_TEXTURES is this repository's own definition (the reference lists the feature, README.md:46, but has no
    code path).  Sanity of that definition: on a quad parallel to the screen (constant clip w) it samples the same texels as the
    reference's perspective-correct interpolation up to float rounding at texel borders, depth is untouched, and on a tilted
    quad the two differ — which is all "affine" means."""
    tex = workloads.checker_texture(64)
    verts = np.array([[-1, -1, 0, 1], [1, -1, 0, 1], [1, 1, 0, 1], [-1, 1, 0, 1]], np.float32)
    uv = np.array([[[0, 0], [1, 0], [1, 1]], [[0, 0], [1, 1], [0, 1]]], np.float32)
    faces = g.FaceArray(np.array([[0, 1, 2], [0, 2, 3]], np.int32), None, uv, np.zeros(2, np.int32), [tex])
    mesh = g.NewMesh(verts, None, faces)

    def render(rot_x, affine):
        o = g.NewObject(mesh)
        o.Rotation = np.array([rot_x, 0, 0], np.float32)
        sc = scene_defs.SceneDef(320, 240, [o], g.Camera(Position=(0, 0, 3)), {"AffineTextures": affine, "Lighting": False})
        return oracle.draw(sc.renderer(None), sc.objects, sc.camera)

    flat_p, flat_a = render(0.0, False), render(0.0, True)
    assert np.array_equal(flat_p["zbuffer"].view(np.uint32), flat_a["zbuffer"].view(np.uint32))
    covered = flat_p["zbuffer"] > -1
    assert covered.sum() > 10000
    differ = (flat_p["pixels"] != flat_a["pixels"]).any(axis=-1)
    assert differ.sum() < 0.02 * covered.sum()          # only pixels whose (u, v) sits on a texel border
    tilt_p, tilt_a = render(1.0, False), render(1.0, True)
    assert np.array_equal(tilt_p["zbuffer"].view(np.uint32), tilt_a["zbuffer"].view(np.uint32))
    covered = tilt_p["zbuffer"] > -1
    differ = (tilt_p["pixels"] != tilt_a["pixels"]).any(axis=-1)
    assert differ.sum() > 0.2 * covered.sum()
