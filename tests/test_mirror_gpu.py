"""Host mirrors (tile-sparse read-back), the one-call Draw with its CUDA-graph cache, batches split by the
workspace limit, and the bounded overflow pool with its big-list fallback.

The property under test everywhere: what lands in host memory is byte for byte what a full-frame copy
(`grb_read_frames`) delivers — and that, in turn, is the oracle's frame.
"""
import ctypes as C

import numpy as np
import pytest

import gorender_b200 as g
from gorender_b200 import _cabi, geometry, workloads
from gorender_b200.renderer import Mirror

import scene_defs

pytestmark = pytest.mark.gpu


def full_read(fb, frame=0):
    px, z = fb.read(frame, 1)
    return px[0], z[0]


def assert_host_equals_device(fb, what, frame=0):
    px, z = full_read(fb, frame)
    assert np.array_equal(fb.Pixels, px), f"{what}: host pixels differ from the device frame"
    assert np.array_equal(fb.ZBuffer.view(np.uint32), z.view(np.uint32)), f"{what}: host z-buffer differs from the device frame"


def test_draw_keeps_host_framebuffer_exact_while_object_moves(device, oracle):
    """Suzanne wanders over the screen: tiles go busy -> empty -> busy.  After every Draw (one-call path: CUDA graph
    + mirrors) fb.Pixels / fb.ZBuffer equal a full read-back; sampled frames also equal the oracle."""
    objs, cam = workloads.config_c1()
    fb = g.FrameBuffer(640, 360, 1, device)
    r = g.Renderer(fb)
    replays0 = device.graph_replays()
    path = [(-2.5, 0.0), (-1.2, 0.8), (0.0, 0.0), (0.0, 0.0), (1.4, -0.9), (2.6, 0.3), (9.0, 0.0), (0.3, 0.2)]
    for k, (tx, ty) in enumerate(path):
        objs[0].Translation = np.array([tx, ty, 0], np.float32)
        objs[0].Rotation = np.array([0, 0.3 * k, 0], np.float32)
        r.Draw(objs, cam)
        assert_host_equals_device(fb, f"step {k}")
        if k in (0, 2, 6, 7):
            ref = oracle.draw(r, objs, cam)
            assert r.TPF == ref["tpf"]
            assert np.array_equal(fb.Pixels, ref["pixels"]) and np.array_equal(fb.ZBuffer.view(np.uint32), ref["zbuffer"].view(np.uint32))
    # same scene, options and framebuffer: replays, except when the object's BoxVisibility class changes
    # (inside <-> intersecting the frustum selects other kernels)
    assert device.graph_replays() - replays0 >= 3
    written, full = fb.mirror("Pixels").stats()
    assert full == len(path) * 20 * 12
    assert written < full * 0.7      # first frame writes everything, later ones only what changed


def test_option_changes_recapture_and_stay_exact(device, oracle):
    """Overlays / post passes mark every tile busy; switching them off again must bring every tile back."""
    sc = scene_defs.multi_object()
    fb = g.FrameBuffer(sc.width, sc.height, 1, device)
    r = sc.renderer(fb)
    for opts in ({}, dict(ShowEdges=True, ShowVertices=True), dict(CrossHair=True), {}, dict(Fog=True), dict(ShowFaces=False),
                 dict(FlatShading=True), {}):
        for k in ("ShowEdges", "ShowVertices", "CrossHair", "Fog", "FlatShading"):
            setattr(r, k, False)
        r.ShowFaces = True
        for k, v in opts.items():
            setattr(r, k, v)
        r.Draw(sc.objects, sc.camera)
        ref = oracle.draw(r, sc.objects, sc.camera)
        assert np.array_equal(fb.Pixels, ref["pixels"]), opts
        assert np.array_equal(fb.ZBuffer.view(np.uint32), ref["zbuffer"].view(np.uint32)), opts


def test_swap_buffers_double_buffer(device):
    """Draw / SwapBuffers loop of main.go:198-227: Pixels and Pixels2 are separate mirrors with separate histories."""
    objs, cam = workloads.config_c1()
    fb = g.FrameBuffer(480, 320, 1, device)
    r = g.Renderer(fb)
    shown = None
    for k in range(6):
        objs[0].Translation = np.array([-2.0 + 0.8 * k, 0, 0], np.float32)
        r.Draw(objs, cam)
        px, _ = full_read(fb)
        assert np.array_equal(fb.Pixels, px)
        fb.SwapBuffers()
        assert np.array_equal(fb.Pixels2, px)      # the presenter's buffer is the frame just drawn
        if shown is not None:
            assert not np.array_equal(shown, px)
        shown = px.copy()


def test_batch_mirror_updates(device, oracle):
    """Multi-frame mirrors updated from batched draws (the streaming form of bench.py's e2e leg): two batches of
    different poses into the same mirror frames, against full copies."""
    objs, cams = workloads.config_c5(n=24, poses=16)
    B = 8
    fb = g.FrameBuffer(640, 360, B, device)
    r = g.Renderer(fb)
    mc = Mirror(device, 640, 360, B, _cabi.GRB_PLANE_COLOR)
    mz = Mirror(device, 640, 360, B, _cabi.GRB_PLANE_DEPTH)
    for b0 in (0, 8, 0):
        packed = r.pack_objects(objs, cams[b0:b0 + B])
        r.draw_packed(packed, 0, sync=False)
        fb.update_mirrors_async(0, B, mc, mz)
        mc.wait()
        mz.wait()
        px, z = fb.read(0, B)
        assert np.array_equal(mc.array, px) and np.array_equal(mz.array.view(np.uint32), z.view(np.uint32))
    ref = oracle.draw(r, objs, cams[3])
    assert np.array_equal(mc.array[3], ref["pixels"]) and np.array_equal(mz.array[3].view(np.uint32), ref["zbuffer"].view(np.uint32))
    # colour only, into other frames of the mirror than the device frames they come from
    packed = r.pack_objects(objs, cams[4:8])
    r.draw_packed(packed, 2, sync=False)
    fb.update_mirrors_async(2, 4, mc, None, color_frame0=0)
    mc.wait()
    px, _ = fb.read(2, 4)
    assert np.array_equal(mc.array[:4], px)
    w, full = mc.stats()
    assert 0 < w < full


def test_mirror_of_strips_and_odd_sizes(device, oracle):
    """Strip draws only touch their rows' tile flags; widths that are not multiples of 4 or 32 take the scalar path."""
    from gorender_b200.parallel import strip_rows

    for (w, h) in ((642, 363), (300, 200)):
        objs, cam = workloads.config_c1()
        fb = g.FrameBuffer(w, h, 1, device)
        r = g.Renderer(fb)
        r.Draw(objs, cam)
        ref = oracle.draw(r, objs, cam)
        assert np.array_equal(fb.Pixels, ref["pixels"]) and np.array_equal(fb.ZBuffer.view(np.uint32), ref["zbuffer"].view(np.uint32))
        objs[0].Translation = np.array([0.7, 0.2, 0], np.float32)
        packed = r.pack_objects(objs, [cam])
        for k in range(3):
            y0, y1 = strip_rows(h, 3, k)
            r.draw_packed(packed, 0, rows=(y0, y1), sync=False)
            fb.update_mirrors_async(0, 1, fb.mirror("Pixels"), fb.mirror("ZBuffer"))
        fb.mirror("Pixels").wait()
        fb.mirror("ZBuffer").wait()
        ref = oracle.draw(r, objs, cam)
        assert np.array_equal(fb.Pixels, ref["pixels"]) and np.array_equal(fb.ZBuffer.view(np.uint32), ref["zbuffer"].view(np.uint32))


def test_invalidate_after_host_writes(device):
    objs, cam = workloads.config_c1()
    fb = g.FrameBuffer(320, 256, 1, device)
    r = g.Renderer(fb)
    r.Draw(objs, cam)
    want = fb.Pixels.copy()
    fb.Pixels[:40, :40] = 7          # a HUD drawn by the host into a background corner
    r.Draw(objs, cam)
    assert (fb.Pixels[:40, :40] == 7).all()           # the mirror believes that corner is still background
    fb.mirror("Pixels").invalidate()
    r.Draw(objs, cam)
    assert np.array_equal(fb.Pixels, want)


def test_registered_host_memory_backs_a_mirror(device):
    """grb_host_register: caller-owned memory (a Go slice, a numpy array) pinned in place."""
    lib = device.lib
    objs, cam = workloads.config_c1()
    fb = g.FrameBuffer(320, 256, 1, device)
    r = g.Renderer(fb)
    raw = np.zeros(320 * 256 * 4 + 4096, np.uint8)
    off = (-raw.ctypes.data) % 4096
    plane = raw[off:off + 320 * 256 * 4].reshape(256, 320, 4)
    h = C.c_void_p()
    rc = lib.grb_mirror_create(device.h, 320, 256, 1, _cabi.GRB_PLANE_COLOR, C.c_void_p(plane.ctypes.data), C.byref(h))
    assert rc != 0                                     # pageable memory is refused, loudly
    assert lib.grb_host_register(C.c_void_p(plane.ctypes.data), plane.nbytes) == 0
    try:
        device.check(lib.grb_mirror_create(device.h, 320, 256, 1, _cabi.GRB_PLANE_COLOR, C.c_void_p(plane.ctypes.data), C.byref(h)))
        packed = r.pack_objects(objs, [cam])
        r.draw_packed(packed, 0, sync=False)
        device.check(lib.grb_mirror_update_async(device.h, fb.handle, 0, 1, h, 0, None, 0))
        device.check(lib.grb_mirror_wait(h))
        px, _ = full_read(fb)
        assert np.array_equal(plane, px)
        device.check(lib.grb_mirror_destroy(h))
    finally:
        assert lib.grb_host_unregister(C.c_void_p(plane.ctypes.data)) == 0


def test_workspace_limit_splits_batches(oracle):
    """A batch whose workspace exceeds the limit is rendered in several launches: same frames, same stats."""
    dev = g.Device(0)
    objs, cams = workloads.config_c5(n=24, poses=10)
    fb = g.FrameBuffer(640, 360, 10, dev)
    r = g.Renderer(fb)
    px, z, tpf = r.DrawBatch(objs, cams)
    dev.trim()
    dev.set_workspace_limit(3 * 1000 * 1000)       # about 3 frames' worth for this scene
    px2, z2, tpf2 = r.DrawBatch(objs, cams)
    assert np.array_equal(px, px2) and np.array_equal(z.view(np.uint32), z2.view(np.uint32)) and np.array_equal(tpf, tpf2)
    dev.set_workspace_limit(1)                     # one frame per launch
    px3, z3, tpf3 = r.DrawBatch(objs, cams)
    assert np.array_equal(px, px3) and np.array_equal(z.view(np.uint32), z3.view(np.uint32)) and np.array_equal(tpf, tpf3)
    ref = oracle.draw(r, objs, cams[9])
    assert np.array_equal(px3[9], ref["pixels"]) and int(tpf3[9]) == ref["tpf"]
    fb.close()
    dev.close()


@pytest.mark.parametrize("cap", [0, 64, 1])
def test_tile_list_overflow_and_big_list_fallback(cap, oracle):
    """A dense mesh far away: tens of thousands of triangles in a handful of tiles.  The tiles' in-place lists
    (512 descriptors) overflow into the frame's pool; with the pool forced tiny the rest falls back to the
    frame-wide list.  Pixels, depth and TPF never depend on any of it."""
    dev = g.Device(0)
    dev.check(dev.lib.grb_debug_set_overflow_cap(dev.h, cap))
    mesh = workloads.sphere(60)                       # 72 000 faces
    o = g.NewObject(mesh)
    o.Scale = np.array([0.09, 0.09, 0.09], np.float32)
    o.Translation = np.array([0.37, 0.21, 0], np.float32)
    o2 = g.NewObject(workloads.suzanne())
    o2.Translation = np.array([-1.5, 0, 0], np.float32)
    cam = geometry.default_camera()
    fb = g.FrameBuffer(640, 360, 1, dev)
    r = g.Renderer(fb)
    r.Draw([o, o2], cam)
    ref = oracle.draw(r, [o, o2], cam)
    assert r.TPF == ref["tpf"]
    assert np.array_equal(fb.Pixels, ref["pixels"]) and np.array_equal(fb.ZBuffer.view(np.uint32), ref["zbuffer"].view(np.uint32))
    fallbacks = int(r.last_stats["list_fallbacks"][0])
    if cap == 0:
        assert fallbacks == 0
    else:
        assert fallbacks > 0
    # and a clipped variant: the clip instantiation appends single-triangle descriptors
    o.Translation = np.array([0.37, 0.21, 4.93], np.float32)
    o.Scale = np.array([0.02, 0.02, 0.02], np.float32)
    r.Draw([o, o2], cam)
    ref = oracle.draw(r, [o, o2], cam)
    assert r.TPF == ref["tpf"]
    assert np.array_equal(fb.Pixels, ref["pixels"]) and np.array_equal(fb.ZBuffer.view(np.uint32), ref["zbuffer"].view(np.uint32))
    fb.close()
    dev.close()


def test_setup_without_raster_does_not_poison_the_next_draw(device, oracle):
    """descCount is re-zeroed by the raster kernel; a context whose setup ran but whose raster did not
    (here: forced through the internal flag by drawing a strip, then a different geometry) must clear it."""
    objs, cam = workloads.config_c1()
    fb = g.FrameBuffer(640, 360, 1, device)
    r = g.Renderer(fb)
    packed = r.pack_objects(objs, [cam])
    r.draw_packed(packed, 0, rows=(0, 160))
    fb2 = g.FrameBuffer(320, 200, 1, device)       # another tile geometry on the same context
    r2 = g.Renderer(fb2)
    r2.Draw(objs, cam)
    ref = oracle.draw(r2, objs, cam)
    assert np.array_equal(fb2.Pixels, ref["pixels"])
    r.Draw(objs, cam)
    ref = oracle.draw(r, objs, cam)
    assert np.array_equal(fb.Pixels, ref["pixels"]) and np.array_equal(fb.ZBuffer.view(np.uint32), ref["zbuffer"].view(np.uint32))


def test_mirror_errors_are_loud(device):
    """Wrong plane type, wrong size, frame ranges, misaligned rows, signals on a framebuffer that is not shared: error codes
    with messages, never a silent wrong copy."""
    Err = g.renderer._cabi.GorenderError
    fb = g.FrameBuffer(320, 256, 2, device)
    mc = Mirror(device, 320, 256, 2, _cabi.GRB_PLANE_COLOR)
    mz = Mirror(device, 320, 256, 2, _cabi.GRB_PLANE_DEPTH)
    small = Mirror(device, 160, 128, 1, _cabi.GRB_PLANE_COLOR)
    with pytest.raises(Err, match="plane"):
        fb.update_mirrors_async(0, 1, mz, None)                 # a depth mirror in the colour slot
    with pytest.raises(Err, match="sizes differ"):
        fb.update_mirrors_async(0, 1, small, None)
    with pytest.raises(Err, match="outside the mirror"):
        fb.update_mirrors_async(0, 2, mc, mz, color_frame0=1)
    with pytest.raises(Err, match="outside the framebuffer"):
        fb.update_mirrors_async(1, 2, mc, mz)
    with pytest.raises(Err, match="tile-aligned"):
        fb.update_mirrors_async(0, 1, mc, mz, rows=(5, 64))
    with pytest.raises(Err, match="not shared"):
        fb.signal(0, 1)
    with pytest.raises(Err, match="not shared"):
        fb.wait_signals(0, 1, 1)
    other = g.Device(0)
    try:
        with pytest.raises(Err, match="another context"):
            other.check(other.lib.grb_mirror_update_async(other.h, fb.handle, 0, 1, mc.h, 0, None, 0))
    finally:
        other.close()
    # and the valid call still works afterwards
    objs, cam = workloads.config_c1()
    r = g.Renderer(fb)
    r.draw_packed(r.pack_objects(objs, [cam, cam]), 0, sync=False)
    fb.update_mirrors_async(0, 2, mc, mz, rows=(0, 256))
    mc.wait()
    mz.wait()
    px, z = fb.read(0, 2)
    assert np.array_equal(mc.array, px) and np.array_equal(mz.array.view(np.uint32), z.view(np.uint32))


def test_mirror_of_wrapped_framebuffer_copies_everything(device):
    """A framebuffer over caller-owned device memory (torch tensors) may be written behind the library's back: its tile flags
    cannot be trusted, so a mirror update moves every tile."""
    import torch

    from gorender_b200.parallel import TorchFrameBuffer

    tfb = TorchFrameBuffer(320, 256, 1, device, torch.device("cuda", 0))
    tfb.color.zero_()
    tfb.depth.zero_()
    torch.cuda.synchronize()
    mc = Mirror(device, 320, 256, 1, _cabi.GRB_PLANE_COLOR)
    mz = Mirror(device, 320, 256, 1, _cabi.GRB_PLANE_DEPTH)
    objs, cam = workloads.config_c1()
    r = g.Renderer(tfb.fb)
    r.draw_packed(r.pack_objects(objs, [cam]), 0)
    tfb.fb.update_mirrors_async(0, 1, mc, mz)
    mc.wait()
    mz.wait()
    assert np.array_equal(mc.array[0], tfb.color[0].cpu().numpy())
    device.synchronize()
    tfb.color[0, :8, :8] = 9                      # the owner scribbles into a background corner of its tensor
    torch.cuda.synchronize()
    tfb.fb.update_mirrors_async(0, 1, mc, mz)
    mc.wait()
    assert (mc.array[0, :8, :8] == 9).all()
    w, full = mc.stats()
    assert w == full                              # every tile, both times


def test_signal_wait_times_out_instead_of_hanging(device):
    """Device-side hand-off flags of a shared framebuffer: a wait nobody answers gives up after its timeout and is counted
    (a dead peer must not hang the GPU); once the flag has been raised the same wait passes at once."""
    fb = g.FrameBuffer(64, 64, 1, device)
    fb.ipc_export()                                   # allocates the flag words (nobody opens the handle here)
    before = device.signal_timeouts()
    fb.wait_signals(5, 2, 1, timeout_ms=40)           # slots 5 and 6 are still 0
    device.synchronize()
    assert device.signal_timeouts() == before + 2     # one per unanswered slot
    fb.signal(5, 3)
    fb.signal(6, 1, on_copy_stream=True)
    device.synchronize()
    fb.wait_signals(5, 2, 1, timeout_ms=40)           # flags only grow: 3 >= 1 and 1 >= 1
    device.synchronize()
    assert device.signal_timeouts() == before + 2
    fb.close()
