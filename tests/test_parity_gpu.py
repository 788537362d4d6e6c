"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the
same inputs.  Integer / byte results must be bit-exact; float32 depth is also
compared bit-exactly (the kernels do the reference's IEEE ops in the
reference's order), which is stricter than north_star's 1e-5 relative.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import gorender_b200 as g
from gorender_b200 import geometry, workloads

import scene_defs

pytestmark = pytest.mark.gpu

GOLDEN = json.load(open(os.path.join(workloads.GOLDEN_DIR, "oracle_outputs.json")))


def render(sc, dev, **kw):
    fb = g.FrameBuffer(sc.width, sc.height, 1, dev)
    r = sc.renderer(fb)
    r.Draw(sc.objects, sc.camera)
    return fb, r


def assert_frame_equal(px, z, ref, what=""):
    zr = ref["zbuffer"]
    same_z = z.view(np.uint32) == zr.view(np.uint32)
    same_p = (px == ref["pixels"]).all(axis=-1)
    if not (same_z.all() and same_p.all()):
        bad = np.argwhere(~(same_z & same_p))
        y, x = bad[0]
        raise AssertionError(
            f"{what}: {len(bad)} differing pixels of {z.size}; first at (x={x}, y={y}): "
            f"gpu rgba={px[y, x].tolist()} z={z[y, x]!r}  oracle rgba={ref['pixels'][y, x].tolist()} z={zr[y, x]!r}")


@pytest.mark.parametrize("name", list(scene_defs.PINNED))
def test_framebuffer_bit_exact(name, device, oracle):
    sc = scene_defs.PINNED[name]()
    fb, r = render(sc, device)
    ref = oracle.draw(r, sc.objects, sc.camera)
    assert r.TPF == ref["tpf"], f"TPF {r.TPF} != {ref['tpf']}"
    assert_frame_equal(fb.Pixels, fb.ZBuffer, ref, name)
    # and against the committed oracle pin
    gold = GOLDEN[name]
    assert hashlib.sha256(fb.Pixels.tobytes()).hexdigest() == gold["pixels_sha256"]
    assert hashlib.sha256(fb.ZBuffer.tobytes()).hexdigest() == gold["zbuffer_sha256"]
    assert int(r.last_stats["out_of_domain"][0]) == 0


def test_matrix_multiply_vec4_batch_seam(device, oracle):
    """asm_test.go's inputs as a known-answer test + random vectors, bit-exact vs both CPU twins."""
    import gorender_b200.vecmath as vm

    m = vm.NewIdentityMatrix()
    m = vm.Multiply(vm.NewRotationMatrix(0.1, 0.2, 0.3), m)
    m = vm.Multiply(vm.NewTranslationMatrix(1, 2, 3), m)
    rng = np.random.default_rng(0)
    cases = [
        np.repeat(np.arange(1000, dtype=np.float32)[:, None], 4, axis=1),   # asm_test.go:16-24
        np.repeat(np.arange(3, dtype=np.float32)[:, None], 4, axis=1),      # asm_test.go:38-41
        rng.standard_normal((100003, 4)).astype(np.float32) * np.float32(1e3),
        np.zeros((0, 4), np.float32),
        np.array([[np.inf, 1, -0.0, 1e-42], [np.nan, 0, 1, 1]], np.float32),
    ]
    for vecs in cases:
        want = oracle.matvec4_batch(m, vecs)
        assert np.array_equal(want.view(np.uint32), oracle.matvec4_batch(m, vecs, sse=True).view(np.uint32))
        got = np.ascontiguousarray(vecs.copy())
        device.matrixMultiplyVec4Batch(m, got)
        # NaN payloads differ by architecture (x86 default NaN 0xffc00000, sm_100 0x7fffffff) and
        # never reach an output (a NaN fails every z-test); everything else is bit-exact
        nan = np.isnan(want)
        assert np.array_equal(np.isnan(got), nan)
        assert np.array_equal(got.view(np.uint32)[~nan], want.view(np.uint32)[~nan])
    one = np.array([[1, 1, 1, 1]], np.float32)
    device.matrixMultiplyVec4Batch(m, one)
    assert one.tolist() == [[np.float32(1.8453332), np.float32(3.159851), np.float32(3.9696174), 1.0]]


@pytest.mark.parametrize("name", ["c1_suzanne_720p", "c3_sphere_n20", "c2_cube_poseB", "multi_object", "c4_small"])
def test_stage_parity(name, device, oracle):
    """Stage-wise: BoxVisibility, clip-space vertices, emitted triangles (snap, w, intensity, uv, texture)."""
    sc = scene_defs.PINNED[name]()
    device.set_stage_capture(True)   # also run the standalone transform kernel and keep its output
    try:
        fb, r = render(sc, device)
        tv = r.debug_transformed(0)
        recs, uvs = r.debug_triangles(0)
    finally:
        device.set_stage_capture(False)
    ref = oracle.draw(r, sc.objects, sc.camera, record=True)
    vis = r.debug_visibility(0, len(sc.objects))
    assert vis.tolist() == ref["visibility"].tolist()

    # Object.TransformedVertices of every visible object
    base = 0
    persp = r.perspective()
    for o, v in zip(sc.objects, vis):
        n = len(o.Mesh.Vertices)
        if v != 0:
            _, mvp = r.object_matrices(o, sc.camera, persp)
            want = oracle.matvec4_batch(mvp, o.Mesh.Vertices)
            assert np.array_equal(tv[base:base + n].view(np.uint32), want.view(np.uint32))
        base += n

    tris = ref["triangles"]
    # the device keeps only triangles with a non-empty raster bbox that sit in some
    # reference tile list; filter the oracle's list the same way
    pts = tris["points"]
    with np.errstate(invalid="ignore"):
        xi = np.trunc(pts[:, :, 0]).astype(np.int64)
        yi = np.trunc(pts[:, :, 1]).astype(np.int64)
    keep = (xi.max(1) >= 0) & (xi.min(1) <= sc.width - 1) & (yi.max(1) >= 0) & (yi.min(1) <= sc.height - 1)
    keep &= (pts[:, :, 0].max(1) >= 0) & (pts[:, :, 1].max(1) >= 0)
    tris = tris[keep]
    assert len(recs) == len(tris)
    for k, (cx, cy, cw, ci) in enumerate((("x0", "y0", "w0", "i0"), ("x1", "y1", "w1", "i1"), ("x2", "y2", "w2", "i2"))):
        assert np.array_equal(recs[cx], np.trunc(tris["points"][:, k, 0]).astype(np.int32))
        assert np.array_equal(recs[cy], np.trunc(tris["points"][:, k, 1]).astype(np.int32))
        assert np.array_equal(recs[cw].view(np.uint32), tris["points"][:, k, 3].view(np.uint32))
        assert np.array_equal(recs[ci].view(np.uint32), tris["intensity"][:, k].view(np.uint32))
    want_tex = tris["tex"] if r.ShowTextures else np.full(len(tris), -1)
    assert np.array_equal(recs["tex"] >= 0, want_tex >= 0)
    textured = recs["tex"] >= 0
    assert np.array_equal(uvs[textured].view(np.uint32), tris["uvs"].reshape(-1, 6)[textured].view(np.uint32))


def test_c3_counts_match_survey(device):
    """SURVEY.md §8c self-consistency numbers for the 200k sphere at the default camera."""
    sc = scene_defs.c3(100)
    fb, r = render(sc, device)
    assert int(r.last_stats["triangles"][0]) == 79502
    assert int((fb.ZBuffer > -1).sum()) == 98957
    assert r.TPF == 81256


def test_batch_equals_single_frames(device, oracle):
    """Frame-parallel batch (C5 style): every frame of a batched draw equals its single-frame oracle render."""
    objs, cams = workloads.config_c5(n=24, poses=12)
    fb = g.FrameBuffer(640, 360, len(cams), device)
    r = g.Renderer(fb)
    px, z, tpf = r.DrawBatch(objs, cams)
    for f in (0, 3, 7, 11):
        ref = oracle.draw(r, objs, cams[f])
        assert int(tpf[f]) == ref["tpf"]
        assert_frame_equal(px[f], z[f], ref, f"pose {f}")


def test_spin_batch(device, oracle):
    """The demo spin (main.go:229-233) as a batch with per-frame Rotation.Y."""
    objs, cam = workloads.config_c1()
    rot = geometry.spin_rotations(6, start=40)
    fb = g.FrameBuffer(800, 600, len(rot), device)
    r = g.Renderer(fb)
    px, z, tpf = r.DrawBatch(objs, [cam] * len(rot), rotations_y=rot)
    for f in (0, 5):
        ref = oracle.draw(r, objs, cam, rotation_y=rot[f])
        assert int(tpf[f]) == ref["tpf"]
        assert_frame_equal(px[f], z[f], ref, f"spin frame {f}")


@pytest.mark.parametrize("strips", [2, 3, 5])
def test_strips_compose_to_full_frame(strips, device, oracle):
    """Sort-first: rendering tile-aligned row strips separately and stacking them equals the full frame."""
    sc = scene_defs.multi_object()
    fb = g.FrameBuffer(sc.width, sc.height, 1, device)
    r = sc.renderer(fb)
    ref = oracle.draw(r, sc.objects, sc.camera)
    packed = r.pack_objects(sc.objects, [sc.camera])
    from gorender_b200.parallel import strip_rows

    px = np.zeros((sc.height, sc.width, 4), np.uint8)
    z = np.zeros((sc.height, sc.width), np.float32)
    tpf = 0
    for k in range(strips):
        y0, y1 = strip_rows(sc.height, strips, k)
        if y0 == y1:
            continue
        stats = r.draw_packed(packed, 0, rows=(y0, y1))
        tpf += int(stats["tpf"][0])                    # a triangle is counted by the strip that owns its top row
        p, zz = fb.read(0, 1)
        px[y0:y1] = p[0, y0:y1]
        z[y0:y1] = zz[0, y0:y1]
    assert tpf == ref["tpf"]                           # ... so the strips' TPFs add up to the frame's
    assert_frame_equal(px, z, ref, f"{strips} strips")


def test_strips_with_overlays(device, oracle):
    """Overlay pixels cross strip borders (lines are not clipped to tiles): every rank draws all of
    them and composes its own rows."""
    sc = scene_defs.multi_object(ShowEdges=True, ShowVertices=True, CrossHair=True)
    fb = g.FrameBuffer(sc.width, sc.height, 1, device)
    r = sc.renderer(fb)
    ref = oracle.draw(r, sc.objects, sc.camera)
    packed = r.pack_objects(sc.objects, [sc.camera])
    from gorender_b200.parallel import strip_rows

    px = np.zeros((sc.height, sc.width, 4), np.uint8)
    z = np.zeros((sc.height, sc.width), np.float32)
    for k in range(3):
        y0, y1 = strip_rows(sc.height, 3, k)
        r.draw_packed(packed, 0, rows=(y0, y1))
        p, zz = fb.read(0, 1)
        px[y0:y1] = p[0, y0:y1]
        z[y0:y1] = zz[0, y0:y1]
    assert_frame_equal(px, z, ref, "3 strips with overlays")


def test_overlay_batch(device, oracle):
    """Overlays in a frame-parallel batch: per-frame event-key planes."""
    objs, cam = workloads.config_c1()
    rot = geometry.spin_rotations(4, start=10)
    fb = g.FrameBuffer(640, 360, len(rot), device)
    r = g.Renderer(fb)
    r.ShowEdges = True
    r.ShowVertices = True
    px, z, tpf = r.DrawBatch(objs, [cam] * len(rot), rotations_y=rot)
    for f in range(len(rot)):
        ref = oracle.draw(r, objs, cam, rotation_y=rot[f])
        assert int(tpf[f]) == ref["tpf"]
        assert_frame_equal(px[f], z[f], ref, f"overlay batch frame {f}")
    # and the plain path right after, on the same context (the overlay plane must not leak)
    r.ShowEdges = r.ShowVertices = False
    px, z, tpf = r.DrawBatch(objs, [cam] * len(rot), rotations_y=rot)
    assert_frame_equal(px[1], z[1], oracle.draw(r, objs, cam, rotation_y=rot[1]), "plain after overlay")


def test_full_size_properties_4k(device):
    """C4 at its BASELINE size (2M faces, 3840x2160) through size-independent properties:
    determinism, strip/full agreement, batch/single agreement, and depth-vs-colour consistency."""
    objs, cam = workloads.config_c4(n=100)
    fb = g.FrameBuffer(3840, 2160, 1, device)
    r = g.Renderer(fb)
    r.Draw(objs, cam)
    px1, z1, tpf1 = fb.Pixels.copy(), fb.ZBuffer.copy(), r.TPF
    assert int(r.last_stats["out_of_domain"][0]) == 0
    r.Draw(objs, cam)
    assert r.TPF == tpf1 and np.array_equal(px1, fb.Pixels) and np.array_equal(z1.view(np.uint32), fb.ZBuffer.view(np.uint32))
    # uncovered pixels are exactly the cleared background / dot grid
    bg = z1 == -1.0
    yy, xx = np.nonzero(bg)
    dot = (xx % 10 == 0) & (yy % 10 == 0) & (xx >= 10) & (yy >= 10)
    want = np.where(dot[:, None], np.array([100, 100, 100, 255], np.uint8), np.array([50, 50, 50, 255], np.uint8))
    assert np.array_equal(px1[bg], want)
    assert (z1[~bg] > 0).all() and px1[~bg][:, 3].max() <= 255
    # lower half as a strip
    packed = r.pack_objects(objs, [cam])
    r.draw_packed(packed, 0, rows=(1088, 2160))
    p, zz = fb.read(0, 1)
    assert np.array_equal(p[0, 1088:], px1[1088:]) and np.array_equal(zz[0, 1088:].view(np.uint32), z1[1088:].view(np.uint32))


def test_full_size_c3_spin_matches_oracle(device, oracle):
    """C3 at full size on three frames of the demo spin, bit-exact."""
    objs, cam = workloads.config_c3(100)
    rot = geometry.spin_rotations(3, start=100)
    fb = g.FrameBuffer(1280, 720, 3, device)
    r = g.Renderer(fb)
    px, z, tpf = r.DrawBatch(objs, [cam] * 3, rotations_y=rot)
    for f in range(3):
        ref = oracle.draw(r, objs, cam, rotation_y=rot[f])
        assert int(tpf[f]) == ref["tpf"]
        assert_frame_equal(px[f], z[f], ref, f"C3 spin frame {f}")


def test_full_size_c4_matches_oracle(device, oracle):
    """C4 at its BASELINE size (BASELINE.json configs[3]: 10 textured Gouraud spheres, 2.0 M faces, 3840x2160)
    bit for bit against the oracle: the whole frame on one GPU, and the same frame composed from the 8
    tile-aligned row strips of the sort-first mode.  The oracle renders this frame in about 0.7 s."""
    from gorender_b200.parallel import strip_rows

    objs, cam = workloads.config_c4(n=100)
    W4, H4 = 3840, 2160
    fb = g.FrameBuffer(W4, H4, 1, device)
    r = g.Renderer(fb)
    ref = oracle.draw(r, objs, cam)
    r.Draw(objs, cam)
    assert r.TPF == ref["tpf"]
    assert int(r.last_stats["out_of_domain"][0]) == 0
    assert_frame_equal(fb.Pixels, fb.ZBuffer, ref, "C4 whole frame")
    packed = r.pack_objects(objs, [cam])
    px = np.zeros((H4, W4, 4), np.uint8)
    z = np.zeros((H4, W4), np.float32)
    tpf = 0
    for k in range(8):
        y0, y1 = strip_rows(H4, 8, k)
        stats = r.draw_packed(packed, 0, rows=(y0, y1))
        tpf += int(stats["tpf"][0])
        p, zz = fb.read(0, 1)
        px[y0:y1] = p[0, y0:y1]
        z[y0:y1] = zz[0, y0:y1]
    assert tpf == ref["tpf"]
    assert_frame_equal(px, z, ref, "C4 composed from 8 strips")
    # ... and from 8 strips balanced by the busy tiles of the frame (what parallel.StripGroup uses)
    from gorender_b200.parallel import balanced_strip_rows

    r.draw_packed(packed, 0)
    rows = balanced_strip_rows(fb.tile_flags(0).sum(axis=1), 8, H4)
    px[:] = 0
    z[:] = 0
    tpf = 0
    for y0, y1 in rows:
        if y1 == y0:
            continue
        stats = r.draw_packed(packed, 0, rows=(y0, y1))
        tpf += int(stats["tpf"][0])
        p, zz = fb.read(0, 1)
        px[y0:y1] = p[0, y0:y1]
        z[y0:y1] = zz[0, y0:y1]
    assert tpf == ref["tpf"]
    assert_frame_equal(px, z, ref, "C4 composed from 8 balanced strips")


def test_full_size_c5_pose_batch(device, oracle):
    """C5 at its BASELINE size: 4096 orbit poses of the 200k-triangle sphere at 1280x720, drawn in batches of
    64 into alternating framebuffers.  64 poses spread over the orbit (one per batch, at a different position
    inside each batch) are compared bit for bit with the oracle, and a second pass over one batch must
    reproduce the first (run-to-run determinism of the atomics)."""
    objs, cams = workloads.config_c5(n=100, poses=4096)
    B = 64
    fbs = [g.FrameBuffer(1280, 720, B, device) for _ in range(2)]
    rs = [g.Renderer(fb) for fb in fbs]
    sample = {k * B + (k * 37) % B: None for k in range(4096 // B)}
    sample[4095] = None
    assert len(sample) >= 64
    tpf_all = np.zeros(4096, np.int64)
    # view matrices differ per pose, the object does not move: pack once per batch
    for b0 in range(0, 4096, B):
        k = (b0 // B) & 1
        packed = rs[k].pack_objects(objs, cams[b0:b0 + B])
        stats = rs[k].draw_packed(packed, 0)
        tpf_all[b0:b0 + B] = stats["tpf"]
        assert int(stats["out_of_domain"].sum()) == 0
        for p in sample:
            if b0 <= p < b0 + B:
                px, z = fbs[k].read(p - b0, 1)
                sample[p] = (px[0].copy(), z[0].copy())
    for p, (px, z) in sample.items():
        ref = oracle.draw(rs[0], objs, cams[p])
        assert int(tpf_all[p]) == ref["tpf"]
        assert_frame_equal(px, z, ref, f"C5 pose {p}")
    assert tpf_all.min() > 70000          # every pose sees the sphere
    # determinism: the first batch again
    packed = rs[0].pack_objects(objs, cams[:B])
    rs[0].draw_packed(packed, 0)
    px, z = fbs[0].read(0, 1)
    assert np.array_equal(px[0], sample[0][0]) and np.array_equal(z[0].view(np.uint32), sample[0][1].view(np.uint32))


def _bits_equal_nan_aware(got, want):
    nan = np.isnan(want)
    return np.array_equal(np.isnan(got), nan) and np.array_equal(got.view(np.uint32)[~nan], want.view(np.uint32)[~nan])


@pytest.mark.parametrize("which", ["suzanne", "cube", "sphere30_vn", "soup"])
def test_new_mesh_on_device(which, device, oracle):
    """NewMesh (mesh.go:53-69) on the device: face normals and bounding box bit for bit against the
    oracle's restatement and the host (numpy) mirror; the mesh then renders exactly like a host-built one."""
    if which == "suzanne":
        host = workloads.suzanne()
    elif which == "cube":
        host = workloads.cube()
    elif which == "sphere30_vn":
        host = geometry.geodesic_sphere(30, True, workloads.checker_texture(32))
    else:
        rng = np.random.default_rng(7)
        nv = 5000
        verts = np.ones((nv, 4), np.float32)
        verts[:, :3] = (rng.normal(0, 2, (nv, 3)) * rng.choice([1e-3, 1.0, 1e3], (nv, 1))).astype(np.float32)
        verts[17, :3] = (-0.0, 0.0, -0.0)
        vidx = rng.integers(0, nv, (20000, 3)).astype(np.int32)
        vidx[::5, 2] = vidx[::5, 1]                      # degenerate faces: 0/0 -> NaN normals
        host = g.NewMesh(verts, None, g.FaceArray(vidx))
    F = host.Faces
    dev_mesh = g.NewMesh(host.Vertices, host.VertexNormals if len(host.VertexNormals) else None,
                         g.FaceArray(F.VertexIndices, F.NormalIndices, F.UVs, F.TextureIndex, F.Textures), device=device)
    want_fn = oracle.face_normals(host.Vertices, F.VertexIndices)
    want_bb = oracle.bounding_box(host.Vertices)
    assert dev_mesh.FaceNormals.shape == want_fn.shape
    assert _bits_equal_nan_aware(dev_mesh.FaceNormals, want_fn)
    assert _bits_equal_nan_aware(host.FaceNormals, want_fn)
    assert np.array_equal(dev_mesh.BoundingBox.view(np.uint32), want_bb.reshape(8, 4).view(np.uint32))
    if which != "soup":
        cam = geometry.default_camera()
        fb = g.FrameBuffer(640, 360, 2, device)
        r = g.Renderer(fb)
        px, z, tpf = r.DrawBatch([g.NewObject(host)], [cam])
        px2, z2, tpf2 = r.DrawBatch([g.NewObject(dev_mesh)], [cam])
        assert int(tpf[0]) == int(tpf2[0]) and np.array_equal(px[0], px2[0])
        assert np.array_equal(z[0].view(np.uint32), z2[0].view(np.uint32))


def test_new_mesh_on_device_rejects_bad_indices(device):
    verts = np.ones((4, 4), np.float32)
    with pytest.raises(g.renderer._cabi.GorenderError):
        g.NewMesh(verts, None, g.FaceArray(np.array([[0, 1, 4]], np.int32)), device=device)
    with pytest.raises(g.renderer._cabi.GorenderError):
        g.NewMesh(verts, None, g.FaceArray(np.array([[0, -1, 2]], np.int32)), device=device)


def test_errors_are_loud(device):
    fb = g.FrameBuffer(64, 64, 1, device)
    r = g.Renderer(fb)
    objs, cam = workloads.config_c1()
    packed = r.pack_objects(objs, [cam])
    with pytest.raises(g.renderer._cabi.GorenderError):
        r.draw_packed(packed, 0, rows=(3, 40))          # not tile aligned
    with pytest.raises(g.renderer._cabi.GorenderError):
        r.draw_packed(packed, 5)                        # frame outside the framebuffer
    bad = packed.copy()
    bad["mesh"] = 12345
    with pytest.raises(g.renderer._cabi.GorenderError):
        r.draw_packed(bad, 0)
    m = workloads.suzanne()
    F = m.Faces
    broken = g.Mesh(m.Vertices, np.ones((2, 4), np.float32),
                    g.FaceArray(F.VertexIndices, np.full_like(F.NormalIndices, 7)))
    with pytest.raises(g.renderer._cabi.GorenderError):
        device.mesh_id(broken)                          # normal index out of range: the reference panics
