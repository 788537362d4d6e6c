"""Frame pipeline — host-side mirror of the reference's `renderer.go` /
`rasterizer.go` public API, with the body of `Draw` running on the GPU through
the C ABI (include/gorender_b200.h).

Kept from the reference: `Camera` (renderer.go:22-26), `FrameBuffer` with
`Width/Height/ZBuffer/Pixels/Pixels2/SwapBuffers` (rasterizer.go:7-34),
`Renderer` with its option fields and `TPF` (renderer.go:83-101), `NewRenderer`
(renderer.go:114-164) and `Draw(objects, camera)` (renderer.go:443-483: no
return value; results are side effects on `fb.Pixels`, `fb.ZBuffer`, `r.TPF`;
failure is an exception, the analogue of the Go shim's panic).

Added for the B200: `DrawBatch` (frame-parallel pose batches)
and `rows=(begin, end)` (sort-first strips).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _cabi
from . import vecmath as vm
from .mesh import Mesh, Object
from .texture import Texture, TextureTypeSolidColor


class Camera:
    """renderer.go:22-26."""

    def __init__(self, Position=(0, 0, 0), Direction=(0, 0, -1), Up=(0, 1, 0)):
        self.Position = np.asarray(Position, dtype=np.float32)
        self.Direction = np.asarray(Direction, dtype=np.float32)
        self.Up = np.asarray(Up, dtype=np.float32)


class Device:
    """One grb_context (one GPU).  Owns uploaded meshes / textures."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self.lib = _cabi.load()
        h = C.c_void_p()
        rc = self.lib.grb_context_create(int(device), C.byref(h))
        if rc != 0:
            raise _cabi.GorenderError(rc, self.lib.grb_last_error(None).decode())
        self.h = h
        self.device = int(device)
        self._meshes = {}    # id(Mesh) -> (mesh id, Mesh kept alive, texture ids tuple)
        self._textures = {}  # id(Texture) -> (texture id, Texture kept alive, scale)
        if stream is not None:
            self.set_stream(stream)

    def check(self, rc: int) -> None:
        _cabi.check(self.h, rc)

    def set_stream(self, cuda_stream: Optional[int]) -> None:
        self.check(self.lib.grb_context_set_stream(self.h, C.c_void_p(cuda_stream or 0)))

    def synchronize(self) -> None:
        self.check(self.lib.grb_context_synchronize(self.h))

    def signal_timeouts(self) -> int:
        return int(self.lib.grb_context_signal_timeouts(self.h))

    def launch_count(self) -> int:
        return int(self.lib.grb_launch_count(self.h))

    def set_kernel_timing(self, enable: bool) -> None:
        self.check(self.lib.grb_context_set_kernel_timing(self.h, int(bool(enable))))

    def set_stage_capture(self, enable: bool) -> None:
        """Also run the standalone transform kernel so `Renderer.debug_transformed` has data."""
        self.check(self.lib.grb_context_set_stage_capture(self.h, int(bool(enable))))

    def set_workspace_limit(self, nbytes: int) -> None:
        """Batches whose per-frame workspace would exceed `nbytes` are rendered in several launches."""
        self.check(self.lib.grb_context_set_workspace_limit(self.h, int(nbytes)))

    def trim(self) -> None:
        self.check(self.lib.grb_context_trim(self.h))

    def graph_replays(self) -> int:
        return int(self.lib.grb_graph_replays(self.h))

    def kernel_times(self):
        ms = (C.c_double * 5)()
        n = C.c_int64()
        self.check(self.lib.grb_kernel_times(self.h, ms, C.byref(n)))
        names = ("transform", "setup", "bin_scan", "bin_fill", "raster")
        return dict(zip(names, [float(x) for x in ms])), int(n.value)

    # -- assets
    def texture_id(self, t: Optional[Texture]) -> int:
        if t is None:
            return -1
        ent = self._textures.get(id(t))
        if ent is not None:
            if ent[2] != float(t.scale):  # Texture.SetScale after upload
                self.check(self.lib.grb_texture_set_scale(self.h, ent[0], float(t.scale)))
                self._textures[id(t)] = (ent[0], t, float(t.scale))
            return ent[0]
        out = C.c_int32(-1)
        color = (C.c_uint8 * 4)(*t.color)
        if t.typ == TextureTypeSolidColor:
            rc = self.lib.grb_texture_upload(self.h, t.typ, 0, 0, float(t.scale), color, None, C.byref(out))
        else:
            px = np.ascontiguousarray(t.pixels, dtype=np.uint8)
            rc = self.lib.grb_texture_upload(self.h, t.typ, t.width, t.height, float(t.scale), color,
                                             C.c_void_p(px.ctypes.data), C.byref(out))
        self.check(rc)
        self._textures[id(t)] = (out.value, t, float(t.scale))
        return out.value

    def mesh_id(self, m: Mesh) -> int:
        F = m.Faces
        tex_ids = tuple(self.texture_id(t) for t in F.Textures)
        ent = self._meshes.get(id(m))
        if ent is not None and ent[2] == tex_ids and ent[3] is F.TextureIndex:
            return ent[0]
        # a mesh built with NewMesh(..., device=dev) has no derived arrays yet: NewMesh's face
        # normals and bounding box (mesh.go:53-69) are computed on the device and copied back
        derive = m.FaceNormals is None
        d = _cabi.grb_mesh_desc()
        nf = len(F)
        verts = np.ascontiguousarray(m.Vertices, dtype=np.float32)
        vns = np.ascontiguousarray(m.VertexNormals, dtype=np.float32)
        fns = None if derive else np.ascontiguousarray(m.FaceNormals, dtype=np.float32)
        vidx = np.ascontiguousarray(F.VertexIndices, dtype=np.int32)
        nidx = np.ascontiguousarray(F.NormalIndices, dtype=np.int32)
        uvs = np.ascontiguousarray(F.UVs, dtype=np.float32)
        lut = np.array(list(tex_ids) + [-1], dtype=np.int32)  # index -1 -> nil
        tex = np.ascontiguousarray(lut[F.TextureIndex], dtype=np.int32)
        d.nv, d.nvn, d.nf = len(verts), len(vns), nf
        d.vertices = _cabi.ptr(verts, C.c_float)
        d.vnormals = _cabi.ptr(vns, C.c_float) if len(vns) else None
        d.fnormals = _cabi.ptr(fns, C.c_float) if (nf and not derive) else None
        d.vidx = _cabi.ptr(vidx, C.c_int32) if nf else None
        d.nidx = _cabi.ptr(nidx, C.c_int32) if (nf and len(vns)) else None
        d.uvs = _cabi.ptr(uvs, C.c_float) if nf else None
        d.tex = _cabi.ptr(tex, C.c_int32) if nf else None
        out = C.c_int32(-1)
        if derive:
            self.check(self.lib.grb_mesh_new(self.h, C.byref(d), C.byref(out)))
            fn = np.empty((nf, 4), np.float32)
            bb = np.empty((8, 4), np.float32)
            self.check(self.lib.grb_mesh_read_derived(self.h, out.value, _cabi.ptr(fn) if nf else None, _cabi.ptr(bb)))
            m.FaceNormals, m.BoundingBox = fn, bb
        else:
            bbox = np.ascontiguousarray(m.BoundingBox, dtype=np.float32).reshape(32)
            d.bbox = (C.c_float * 32)(*bbox.tolist())
            self.check(self.lib.grb_mesh_upload(self.h, C.byref(d), C.byref(out)))
        if ent is not None:
            self.check(self.lib.grb_mesh_free(self.h, ent[0]))
        self._meshes[id(m)] = (out.value, m, tex_ids, F.TextureIndex)
        return out.value

    def pinned_array(self, shape, dtype) -> np.ndarray:
        """numpy array over cudaHostAlloc'ed memory (freed when the array is collected)."""
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        p = self.lib.grb_host_alloc(max(nbytes, 1))
        if not p:
            raise MemoryError(f"grb_host_alloc({nbytes}) failed")
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(p)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        import weakref

        weakref.finalize(buf, self.lib.grb_host_free, C.c_void_p(p))
        return arr

    def matrixMultiplyVec4Batch(self, m: np.ndarray, vecs: np.ndarray) -> None:
        """The reference's build-tag seam (asm_amd64.go:8 / asm_purego.go:9): in place."""
        m = np.ascontiguousarray(m, dtype=np.float32).reshape(16)
        assert vecs.dtype == np.float32 and vecs.flags["C_CONTIGUOUS"] and vecs.size % 4 == 0
        self.check(self.lib.grb_matrix_multiply_vec4_batch(
            self.h, _cabi.ptr(m, C.c_float), C.c_void_p(vecs.ctypes.data), vecs.size // 4))

    def close(self) -> None:
        if getattr(self, "h", None):
            self.lib.grb_context_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_devices = {}


def default_device(device: int = 0) -> Device:
    d = _default_devices.get(device)
    if d is None:
        d = _default_devices[device] = Device(device)
    return d


class Mirror:
    """A host plane (colour or depth) kept in sync with device frames tile by tile (`grb_mirror`):
    `array` is pinned host memory of shape (frames, H, W[, 4]); `update` writes only the tiles that are
    busy now or were busy in the host copy."""

    def __init__(self, dev: Device, width: int, height: int, frames: int, plane: int, target: "Optional[FrameBuffer]" = None):
        self.dev = dev
        self.plane = plane
        h = C.c_void_p()
        if target is not None:
            # the plane is a plane of another framebuffer (normally one opened from another process: a strip pushed
            # to the frame's owner over NVLink); there is no host array
            self.array = None
            self.target = target
            dev.check(dev.lib.grb_mirror_create_on_framebuffer(dev.h, target.handle, plane, C.byref(h)))
        else:
            shape = (frames, height, width, 4) if plane == _cabi.GRB_PLANE_COLOR else (frames, height, width)
            self.array = dev.pinned_array(shape, np.uint8 if plane == _cabi.GRB_PLANE_COLOR else np.float32)
            dev.check(dev.lib.grb_mirror_create(dev.h, width, height, frames, plane, C.c_void_p(self.array.ctypes.data), C.byref(h)))
        self.h = h

    def wait(self) -> None:
        self.dev.check(self.dev.lib.grb_mirror_wait(self.h))

    def invalidate(self) -> None:
        self.dev.check(self.dev.lib.grb_mirror_invalidate(self.h))

    def stats(self) -> Tuple[int, int]:
        """(tiles written into the host plane, tiles full copies would have moved) since creation."""
        w, f = C.c_int64(), C.c_int64()
        self.dev.check(self.dev.lib.grb_mirror_stats(self.h, C.byref(w), C.byref(f)))
        return int(w.value), int(f.value)

    def close(self) -> None:
        if getattr(self, "h", None) and getattr(self.dev, "h", None):
            self.dev.lib.grb_mirror_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FrameBuffer:
    """rasterizer.go:7-23.  `Pixels`, `Pixels2` are (H, W, 4) uint8 RGBA,
    `ZBuffer` is (H, W) float32 — host arrays, as in the reference; the device
    copy (`frames` of them for batches) lives behind `handle`.  The host arrays are pinned memory
    behind host mirrors (created on first use): `Renderer.Draw` moves only the tiles that changed."""

    def __init__(self, width: int, height: int, frames: int = 1, device: Optional[Device] = None,
                 device_color: Optional[int] = None, device_depth: Optional[int] = None, ipc_handle: Optional[bytes] = None):
        self.Width = int(width)
        self.Height = int(height)
        self.Frames = int(frames)
        self.dev = device or default_device()
        self._mirrors = {}   # "Pixels" / "Pixels2" / "ZBuffer" -> Mirror (one frame each)
        h = C.c_void_p()
        if ipc_handle is not None:
            buf = (C.c_uint8 * _cabi.GRB_IPC_HANDLE_BYTES).from_buffer_copy(ipc_handle)
            rc = self.dev.lib.grb_framebuffer_ipc_open(self.dev.h, buf, C.byref(h))
        elif device_color is not None:
            rc = self.dev.lib.grb_framebuffer_wrap(self.dev.h, width, height, frames, C.c_void_p(device_color),
                                                   C.c_void_p(device_depth), C.byref(h))
        else:
            rc = self.dev.lib.grb_framebuffer_create(self.dev.h, width, height, frames, C.byref(h))
        self.dev.check(rc)
        self.handle = h

    def mirror(self, name: str) -> Mirror:
        m = self._mirrors.get(name)
        if m is None:
            plane = _cabi.GRB_PLANE_DEPTH if name == "ZBuffer" else _cabi.GRB_PLANE_COLOR
            m = self._mirrors[name] = Mirror(self.dev, self.Width, self.Height, 1, plane)
            m.array[...] = 0     # the reference's NewFrameBuffer hands out zeroed slices
        return m

    @property
    def Pixels(self) -> np.ndarray:
        return self.mirror("Pixels").array[0]

    @property
    def Pixels2(self) -> np.ndarray:
        return self.mirror("Pixels2").array[0]

    @property
    def ZBuffer(self) -> np.ndarray:
        return self.mirror("ZBuffer").array[0]

    def SwapBuffers(self) -> None:
        """rasterizer.go:32-34."""
        a, b = self.mirror("Pixels"), self.mirror("Pixels2")
        self._mirrors["Pixels"], self._mirrors["Pixels2"] = b, a

    # -- a framebuffer shared by the processes of a node (sort-first strips, parallel.StripGroup)
    def ipc_export(self) -> bytes:
        buf = (C.c_uint8 * _cabi.GRB_IPC_HANDLE_BYTES)()
        self.dev.check(self.dev.lib.grb_framebuffer_ipc_export(self.handle, buf))
        return bytes(buf)

    def signal(self, slot: int, value: int, on_copy_stream: bool = False) -> None:
        self.dev.check(self.dev.lib.grb_framebuffer_signal(self.dev.h, self.handle, slot, value & 0xffffffff, int(on_copy_stream)))

    def wait_signals(self, slot0: int, nslots: int, value: int, timeout_ms: int = 5000, on_copy_stream: bool = False) -> None:
        self.dev.check(self.dev.lib.grb_framebuffer_wait_signals(self.dev.h, self.handle, slot0, nslots, value & 0xffffffff, timeout_ms,
                                                                 int(on_copy_stream)))

    def tile_flags(self, frame: int = 0) -> np.ndarray:
        """(tile rows, tile columns) uint8: 0 where the device tile holds only the cleared background."""
        nty, ntx = (self.Height + _cabi.GRB_TILE - 1) // _cabi.GRB_TILE, (self.Width + _cabi.GRB_TILE - 1) // _cabi.GRB_TILE
        out = np.zeros((nty, ntx), np.uint8)
        self.dev.check(self.dev.lib.grb_framebuffer_read_tile_flags(self.handle, frame, C.c_void_p(out.ctypes.data)))
        return out

    def device_ptrs(self) -> Tuple[int, int]:
        c, d = C.c_void_p(), C.c_void_p()
        self.dev.check(self.dev.lib.grb_framebuffer_device_ptrs(self.handle, C.byref(c), C.byref(d)))
        return int(c.value), int(d.value)

    def read(self, frame0: int = 0, nframes: int = 1, pixels: Optional[np.ndarray] = None,
             zbuffer: Optional[np.ndarray] = None, want_z: bool = True):
        """Copy device frames to host arrays (allocated if not given): whole frames, one DMA each."""
        if pixels is None:
            pixels = np.empty((nframes, self.Height, self.Width, 4), dtype=np.uint8)
        if zbuffer is None and want_z:
            zbuffer = np.empty((nframes, self.Height, self.Width), dtype=np.float32)
        self.dev.check(self.dev.lib.grb_read_frames(
            self.dev.h, self.handle, frame0, nframes, C.c_void_p(pixels.ctypes.data),
            C.c_void_p(zbuffer.ctypes.data) if zbuffer is not None else None))
        return pixels, zbuffer

    def read_async(self, frame0: int, nframes: int, pixels: Optional[np.ndarray], zbuffer: Optional[np.ndarray]) -> None:
        """Queue the copy on the context's copy stream (targets should be `Device.pinned_array`s);
        it overlaps draws into other framebuffers.  `Device.synchronize()` waits for it."""
        self.dev.check(self.dev.lib.grb_read_frames_async(
            self.dev.h, self.handle, frame0, nframes,
            C.c_void_p(pixels.ctypes.data) if pixels is not None else None,
            C.c_void_p(zbuffer.ctypes.data) if zbuffer is not None else None))

    def update_mirrors_async(self, frame0: int, nframes: int, color: Optional[Mirror], depth: Optional[Mirror],
                             color_frame0: int = 0, depth_frame0: int = 0, rows: Optional[Tuple[int, int]] = None) -> None:
        """Bring mirrors up to date with device frames [frame0, frame0 + nframes): only tiles that are busy
        now or were busy in the mirror's copy cross PCIe (host mirrors) or NVLink (mirrors on another GPU's
        framebuffer); `rows` restricts the update to the tile rows a strip draw has rendered."""
        y0, y1 = rows if rows is not None else (0, 0)
        self.dev.check(self.dev.lib.grb_mirror_update_rows_async(
            self.dev.h, self.handle, frame0, nframes, color.h if color is not None else None, color_frame0,
            depth.h if depth is not None else None, depth_frame0, y0, y1))

    def close(self) -> None:
        for m in getattr(self, "_mirrors", {}).values():
            m.close()
        self._mirrors = {}
        if getattr(self, "handle", None) and getattr(self.dev, "h", None):
            self.dev.lib.grb_framebuffer_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def NewFrameBuffer(width: int, height: int, frames: int = 1, device: Optional[Device] = None) -> FrameBuffer:
    """rasterizer.go:15-23."""
    return FrameBuffer(width, height, frames, device)


class Renderer:
    """renderer.go:83-112 (public fields) and :114-164 (NewRenderer)."""

    def __init__(self, fb: FrameBuffer, parallel: bool = True):
        self.fb = fb
        self.dev = fb.dev
        # renderer.go:115-122
        self.aspectX = np.float32(np.float32(fb.Width) / np.float32(fb.Height))
        self.aspectY = np.float32(np.float32(fb.Height) / np.float32(fb.Width))
        self.fovY = np.float32(45 * (math.pi / 180))
        self.fovX = np.float32(2 * math.atan(math.tan(float(self.fovY / np.float32(2))) * float(self.aspectX)))
        self.zNear, self.zFar = np.float32(0.0), np.float32(50.0)
        # renderer.go:130-137
        self.FrustumClipping = True
        self.ShowVertices = False
        self.ShowEdges = False
        self.ShowFaces = True
        self.BackfaceCulling = True
        self.Lighting = True
        self.FlatShading = False
        self.ShowTextures = True
        self.TPF = 0
        self.DebugEnabled = False
        self.DebugInfo: list = []
        # renderer.go:476-480: `if !demoMode { CrossHair; // Fog(0.100, 0.033, {100,100,100,255}) }` —
        # compile-time switches in the reference (main.go:22), run-time fields here
        self.CrossHair = False
        self.Fog = False
        self.FogStart, self.FogEnd = np.float32(0.100), np.float32(0.033)
        self.FogColor = (100, 100, 100, 255)
        # README.md:46 lists "Affine texture mapping"; the reference has no code path for it.  This repository's
        # own, unpinned definition (include/gorender_b200.h, GRB_OPT_AFFINE_TEXTURES); off by default
        self.AffineTextures = False
        # renderer.go:144,151: numTiles is 1 when !parallel, else max(NumCPU, 16) (16 on any
        # host the reference runs on: more than 16 CPUs panic)
        self.numTiles = 16 if parallel else 1
        self.last_stats = None

    # -- options -> ABI bitmask
    def options(self) -> int:
        o = 0
        if self.FrustumClipping: o |= _cabi.GRB_OPT_FRUSTUM_CLIPPING
        if self.ShowFaces: o |= _cabi.GRB_OPT_SHOW_FACES
        if self.BackfaceCulling: o |= _cabi.GRB_OPT_BACKFACE_CULLING
        if self.Lighting: o |= _cabi.GRB_OPT_LIGHTING
        if self.FlatShading: o |= _cabi.GRB_OPT_FLAT_SHADING
        if self.ShowTextures: o |= _cabi.GRB_OPT_SHOW_TEXTURES
        if self.ShowEdges: o |= _cabi.GRB_OPT_SHOW_EDGES
        if self.ShowVertices: o |= _cabi.GRB_OPT_SHOW_VERTICES
        if self.CrossHair: o |= _cabi.GRB_OPT_CROSSHAIR
        if self.Fog: o |= _cabi.GRB_OPT_FOG
        if self.AffineTextures: o |= _cabi.GRB_OPT_AFFINE_TEXTURES
        return o

    def draw_params(self, rows: Optional[Tuple[int, int]] = None) -> _cabi.grb_draw_params:
        key = (rows, self.options(), self.numTiles, float(self.zNear), float(self.zFar), self.fb.Width, self.fb.Height,
               float(self.FogStart), float(self.FogEnd), tuple(self.FogColor))
        cached = getattr(self, "_params_cache", None)
        if cached is not None and cached[0] == key:
            return cached[1]
        p = self._build_draw_params(rows)
        self._params_cache = (key, p)
        return p

    def _build_draw_params(self, rows: Optional[Tuple[int, int]] = None) -> _cabi.grb_draw_params:
        p = _cabi.grb_draw_params()
        screen = vm.NewScreenMatrix(self.fb.Width, self.fb.Height)          # renderer.go:264
        light = vm.light_direction()                                        # renderer.go:265
        p.screen = (C.c_float * 16)(*screen.reshape(16).tolist())
        p.light = (C.c_float * 3)(*light.tolist())
        p.options = self.options()
        p.z_near, p.z_far = float(self.zNear), float(self.zFar)
        p.ref_tiles = self.numTiles
        p.row_begin, p.row_end = (rows if rows is not None else (0, 0))
        p.fog_start, p.fog_end = float(self.FogStart), float(self.FogEnd)
        p.fog_color = (C.c_uint8 * 4)(*[int(c) & 0xff for c in self.FogColor])
        return p

    def perspective(self) -> np.ndarray:
        return vm.NewPerspectiveMatrix(self.fovY, self.aspectX, self.zNear, self.zFar)  # renderer.go:257

    def object_matrices(self, obj: Object, camera: Camera, perspective=None, view=None):
        """renderer.go:255-262."""
        world = vm.NewWorldMatrix(obj.Scale, obj.Rotation, obj.Translation)
        if view is None:
            view = vm.NewViewMatrix(camera.Position, camera.Direction, camera.Up)
        if perspective is None:
            perspective = self.perspective()
        return world, vm.mvp_matrix(perspective, view, world)

    def pack_objects(self, objects: Sequence[Object], cameras: Sequence[Camera],
                     rotations_y: Optional[np.ndarray] = None) -> np.ndarray:
        """grb_object array [len(cameras)][len(objects)].  `rotations_y[f]`, when
        given, overrides every object's Rotation.Y in frame f (the demo spin,
        main.go:229-233) without touching the objects."""
        nobj = len(objects)
        mesh_ids = [self.dev.mesh_id(o.Mesh) for o in objects]
        arr = np.zeros((len(cameras), max(nobj, 0)), dtype=_cabi.OBJECT_DTYPE)
        persp = self.perspective()
        for f, cam in enumerate(cameras):
            view = vm.NewViewMatrix(cam.Position, cam.Direction, cam.Up)
            for i, o in enumerate(objects):
                if rotations_y is not None:
                    rot = np.array([o.Rotation[0], rotations_y[f], o.Rotation[2]], dtype=np.float32)
                    world = vm.NewWorldMatrix(o.Scale, rot, o.Translation)
                    mvp = vm.mvp_matrix(persp, view, world)
                else:
                    world, mvp = self.object_matrices(o, cam, persp, view)
                arr[f, i]["mesh"] = mesh_ids[i]
                arr[f, i]["world"] = world.reshape(16)
                arr[f, i]["mvp"] = mvp.reshape(16)
        return arr

    # -- the hot path
    def draw_packed(self, packed: np.ndarray, frame0: int = 0, rows=None, sync: bool = True):
        """Issue one batched draw of packed[frames][nobj] into fb frames [frame0, ...)."""
        nframes, nobj = packed.shape
        p = self.draw_params(rows)
        lib = self.dev.lib
        if sync:
            stats = np.zeros(nframes, dtype=_cabi.STATS_DTYPE)
            self.dev.check(lib.grb_draw(self.dev.h, self.fb.handle, frame0, nframes,
                                        C.c_void_p(packed.ctypes.data), nobj, C.byref(p),
                                        C.c_void_p(stats.ctypes.data)))
            self.last_stats = stats
            return stats
        self.dev.check(lib.grb_draw_async(self.dev.h, self.fb.handle, frame0, nframes,
                                          C.c_void_p(packed.ctypes.data), nobj, C.byref(p)))
        return None

    def Draw(self, objects: Sequence[Object], camera: Camera, read_z: bool = True) -> None:
        """renderer.go:443-483.  Side effects: fb.Pixels, fb.ZBuffer, self.TPF.  One C-ABI call
        (grb_draw_present): draw, bring the host mirrors behind fb.Pixels / fb.ZBuffer up to date, stats."""
        packed = self.pack_objects(objects, [camera])
        fb = self.fb
        stats = np.zeros(1, dtype=_cabi.STATS_DTYPE)
        p = self.draw_params(None)
        color, depth = fb.mirror("Pixels"), (fb.mirror("ZBuffer") if read_z else None)
        self.dev.check(self.dev.lib.grb_draw_present(
            self.dev.h, fb.handle, 0, 1, C.c_void_p(packed.ctypes.data), packed.shape[1], C.byref(p),
            color.h, 0, depth.h if depth is not None else None, 0, C.c_void_p(stats.ctypes.data)))
        self.last_stats = stats
        self.TPF = int(stats["tpf"][0])

    def DrawBatch(self, objects: Sequence[Object], cameras: Sequence[Camera],
                  rotations_y: Optional[np.ndarray] = None, read_back: bool = True, read_z: bool = True):
        """Frame-parallel batch: frame f = Draw(objects, cameras[f]).  The
        framebuffer must have at least len(cameras) frames.  Returns (pixels,
        zbuffer, tpf) host arrays when read_back."""
        if len(cameras) > self.fb.Frames:
            raise ValueError("framebuffer has fewer frames than the batch")
        packed = self.pack_objects(objects, cameras, rotations_y)
        stats = self.draw_packed(packed, 0)
        self.TPF = int(stats["tpf"][-1])
        if not read_back:
            return None, None, stats["tpf"].copy()
        px, z = self.fb.read(0, len(cameras), want_z=read_z)
        return px, z, stats["tpf"].copy()

    # -- stage read-backs (parity tests)
    def debug_transformed(self, frame: int = 0) -> np.ndarray:
        n = C.c_int64()
        lib = self.dev.lib
        self.dev.check(lib.grb_debug_read_transformed(self.dev.h, frame, None, 0, C.byref(n)))
        out = np.empty((n.value, 4), dtype=np.float32)
        self.dev.check(lib.grb_debug_read_transformed(self.dev.h, frame, C.c_void_p(out.ctypes.data), n.value, C.byref(n)))
        return out

    def debug_triangles(self, frame: int = 0):
        n = C.c_int64()
        lib = self.dev.lib
        self.dev.check(lib.grb_debug_read_triangles(self.dev.h, frame, None, None, 0, C.byref(n)))
        recs = np.zeros(n.value, dtype=_cabi.TRIANGLE_DTYPE)
        uvs = np.zeros((n.value, 6), dtype=np.float32)
        if n.value:
            self.dev.check(lib.grb_debug_read_triangles(self.dev.h, frame, C.c_void_p(recs.ctypes.data),
                                                        C.c_void_p(uvs.ctypes.data), n.value, C.byref(n)))
        return recs, uvs

    def debug_visibility(self, frame: int, nobj: int) -> np.ndarray:
        out = np.zeros(nobj, dtype=np.int32)
        self.dev.check(self.dev.lib.grb_debug_read_visibility(self.dev.h, frame, C.c_void_p(out.ctypes.data), nobj))
        return out


def NewRenderer(fb: FrameBuffer, parallel: bool = True) -> Renderer:
    """renderer.go:114-164."""
    return Renderer(fb, parallel)
