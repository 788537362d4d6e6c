"""ctypes wrapper of the CPU oracle (oracle/gorender_oracle.h).

Test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU
legs import this.  It consumes the product's host-side containers (Mesh,
Object, Texture, Camera, Renderer options) so both sides of a parity test are
fed the very same arrays and matrices.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "_build", "liboracle_gorender.so")

c_float_p = C.POINTER(C.c_float)
c_i32_p = C.POINTER(C.c_int32)


class orc_texture(C.Structure):
    _fields_ = [("type", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("scale", C.c_float),
                ("color", C.c_uint8 * 4), ("pixels", C.c_void_p)]


class orc_mesh(C.Structure):
    _fields_ = [("nv", C.c_int32), ("nvn", C.c_int32), ("nf", C.c_int32),
                ("vertices", c_float_p), ("vnormals", c_float_p), ("fnormals", c_float_p),
                ("vidx", c_i32_p), ("nidx", c_i32_p), ("uvs", c_float_p), ("tex", c_i32_p),
                ("bbox", C.c_float * 32)]


class orc_object(C.Structure):
    _fields_ = [("mesh", C.c_int32), ("world", C.c_float * 16), ("mvp", C.c_float * 16)]


ORC_TRIANGLE_DTYPE = np.dtype([("points", np.float32, (3, 4)), ("uvs", np.float32, (3, 2)),
                               ("intensity", np.float32, (3,)), ("tex", np.int32),
                               ("object", np.int32), ("face", np.int32), ("fan", np.int32)])
assert ORC_TRIANGLE_DTYPE.itemsize == 100


def build_oracle() -> str:
    src = os.path.join(ORACLE_DIR, "gorender_oracle.cpp")
    if (not os.path.exists(LIB)) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
    return LIB


def _fp(a):
    return a.ctypes.data_as(c_float_p)


def _ip(a):
    return a.ctypes.data_as(c_i32_p)


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.orc_renderer_create.restype = C.c_void_p
        L.orc_renderer_create.argtypes = [C.c_int32] * 4
        L.orc_renderer_destroy.argtypes = [C.c_void_p]
        L.orc_renderer_record_triangles.argtypes = [C.c_void_p, C.c_int32]
        L.orc_renderer_set_fog.argtypes = [C.c_void_p, C.c_float, C.c_float, C.POINTER(C.c_uint8)]
        L.orc_renderer_draw.restype = C.c_int32
        L.orc_renderer_draw.argtypes = [C.c_void_p, C.POINTER(orc_mesh), C.c_int32, C.POINTER(orc_texture), C.c_int32,
                                        C.POINTER(orc_object), C.c_int32, c_float_p, c_float_p, C.c_uint32]
        L.orc_renderer_draw_sequence.restype = C.c_double
        L.orc_renderer_draw_sequence.argtypes = [C.c_void_p, C.POINTER(orc_mesh), C.c_int32, C.POINTER(orc_texture), C.c_int32,
                                                 C.POINTER(orc_object), C.c_int32, C.c_int32, c_float_p, c_float_p, C.c_uint32]
        L.orc_renderer_pixels.restype = C.c_void_p
        L.orc_renderer_pixels.argtypes = [C.c_void_p]
        L.orc_renderer_zbuffer.restype = C.c_void_p
        L.orc_renderer_zbuffer.argtypes = [C.c_void_p]
        for f in ("orc_renderer_tpf", "orc_renderer_pixel_writes", "orc_renderer_num_triangles"):
            getattr(L, f).restype = C.c_int64
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_renderer_triangles.restype = C.c_void_p
        L.orc_renderer_triangles.argtypes = [C.c_void_p]
        L.orc_renderer_visibility.restype = C.c_int32
        L.orc_renderer_visibility.argtypes = [C.c_void_p, c_i32_p, C.c_int32]
        L.orc_matvec4_batch_scalar.argtypes = [c_float_p, c_float_p, C.c_int64]
        L.orc_matvec4_batch_sse.argtypes = [c_float_p, c_float_p, C.c_int64]
        L.orc_box_visibility.restype = C.c_int32
        L.orc_box_visibility.argtypes = [c_float_p, C.c_float, C.c_float]
        L.orc_clip_triangle.restype = C.c_int32
        L.orc_clip_triangle.argtypes = [c_float_p, c_float_p, c_float_p, C.c_float, C.c_float,
                                        c_float_p, c_float_p, c_float_p]
        L.orc_world_matrix.argtypes = [c_float_p] * 4
        L.orc_view_matrix.argtypes = [c_float_p] * 4
        L.orc_perspective_matrix.argtypes = [C.c_float] * 4 + [c_float_p]
        L.orc_screen_matrix.argtypes = [C.c_int32, C.c_int32, c_float_p]
        L.orc_matrix_multiply.argtypes = [c_float_p] * 3
        L.orc_mvp_matrix.argtypes = [c_float_p] * 4
        L.orc_light_direction.argtypes = [c_float_p]
        L.orc_face_normals.argtypes = [c_float_p, c_i32_p, C.c_int32, c_float_p]
        L.orc_bounding_box.argtypes = [c_float_p, C.c_int32, c_float_p]
        L.orc_texture_sample.argtypes = [C.POINTER(orc_texture), C.c_float, C.c_float, C.POINTER(C.c_uint8)]

    # ---- host-side helpers
    def world_matrix(self, s, r, t):
        out = np.zeros(16, np.float32)
        a = [np.ascontiguousarray(x, np.float32) for x in (s, r, t)]
        self.lib.orc_world_matrix(_fp(a[0]), _fp(a[1]), _fp(a[2]), _fp(out))
        return out.reshape(4, 4)

    def view_matrix(self, eye, d, up):
        out = np.zeros(16, np.float32)
        a = [np.ascontiguousarray(x, np.float32) for x in (eye, d, up)]
        self.lib.orc_view_matrix(_fp(a[0]), _fp(a[1]), _fp(a[2]), _fp(out))
        return out.reshape(4, 4)

    def perspective_matrix(self, fov, aspect, zn, zf):
        out = np.zeros(16, np.float32)
        self.lib.orc_perspective_matrix(float(fov), float(aspect), float(zn), float(zf), _fp(out))
        return out.reshape(4, 4)

    def screen_matrix(self, w, h):
        out = np.zeros(16, np.float32)
        self.lib.orc_screen_matrix(w, h, _fp(out))
        return out.reshape(4, 4)

    def mvp_matrix(self, p, v, w):
        out = np.zeros(16, np.float32)
        a = [np.ascontiguousarray(x, np.float32).reshape(16) for x in (p, v, w)]
        self.lib.orc_mvp_matrix(_fp(a[0]), _fp(a[1]), _fp(a[2]), _fp(out))
        return out.reshape(4, 4)

    def light_direction(self):
        out = np.zeros(3, np.float32)
        self.lib.orc_light_direction(_fp(out))
        return out

    def face_normals(self, verts, vidx):
        verts = np.ascontiguousarray(verts, np.float32)
        vidx = np.ascontiguousarray(vidx, np.int32)
        out = np.zeros((len(vidx), 4), np.float32)
        self.lib.orc_face_normals(_fp(verts), _ip(vidx), len(vidx), _fp(out))
        return out

    def bounding_box(self, verts):
        verts = np.ascontiguousarray(verts, np.float32)
        out = np.zeros((8, 4), np.float32)
        self.lib.orc_bounding_box(_fp(verts), len(verts), _fp(out))
        return out

    def matvec4_batch(self, m, vecs, sse=False):
        m = np.ascontiguousarray(m, np.float32).reshape(16)
        out = np.ascontiguousarray(vecs, np.float32).copy()
        fn = self.lib.orc_matvec4_batch_sse if sse else self.lib.orc_matvec4_batch_scalar
        fn(_fp(m), _fp(out), out.size // 4)
        return out

    def box_visibility(self, bbox_clip, zn=0.0, zf=50.0):
        b = np.ascontiguousarray(bbox_clip, np.float32).reshape(32)
        return int(self.lib.orc_box_visibility(_fp(b), zn, zf))

    def clip_triangle(self, pts, uvs, intens, zn=0.0, zf=50.0):
        pts = np.ascontiguousarray(pts, np.float32).reshape(12)
        uvs = np.ascontiguousarray(uvs, np.float32).reshape(6)
        intens = np.ascontiguousarray(intens, np.float32).reshape(3)
        po = np.zeros((9, 3, 4), np.float32)
        uo = np.zeros((9, 3, 2), np.float32)
        io = np.zeros((9, 3), np.float32)
        n = self.lib.orc_clip_triangle(_fp(pts), _fp(uvs), _fp(intens), zn, zf, _fp(po), _fp(uo), _fp(io))
        return po[:n], uo[:n], io[:n]

    def _texture_struct(self, t, keep):
        s = orc_texture()
        s.type, s.width, s.height, s.scale = t.typ, t.width, t.height, float(t.scale)
        s.color = (C.c_uint8 * 4)(*t.color)
        if t.pixels is not None:
            px = np.ascontiguousarray(t.pixels, np.uint8)
            keep.append(px)
            s.pixels = px.ctypes.data
        return s

    def texture_sample(self, t, u, v):
        keep = []
        s = self._texture_struct(t, keep)
        out = (C.c_uint8 * 4)()
        self.lib.orc_texture_sample(C.byref(s), float(u), float(v), out)
        return tuple(out)

    # ---- Renderer.Draw
    def marshal(self, renderer, objects, cameras, rotations_y=None):
        """ctypes arrays for `frames = len(cameras)` consecutive draws of `objects`."""
        import gorender_b200.vecmath as vm

        fb = renderer.fb
        keep = []
        meshes, mesh_index, textures, tex_index = [], {}, [], {}
        nobj = len(objects)
        objs = (orc_object * max(nobj * len(cameras), 1))()
        persp = renderer.perspective()
        for o in objects:
            m = o.Mesh
            if id(m) in mesh_index:
                continue
            F = m.Faces
            ids = []
            for t in F.Textures:
                if id(t) not in tex_index:
                    tex_index[id(t)] = len(textures)
                    textures.append(self._texture_struct(t, keep))
                ids.append(tex_index[id(t)])
            lut = np.array(ids + [-1], dtype=np.int32)
            arrs = dict(
                v=np.ascontiguousarray(m.Vertices, np.float32), vn=np.ascontiguousarray(m.VertexNormals, np.float32),
                fn=np.ascontiguousarray(m.FaceNormals, np.float32), vi=np.ascontiguousarray(F.VertexIndices, np.int32),
                ni=np.ascontiguousarray(F.NormalIndices, np.int32), uv=np.ascontiguousarray(F.UVs, np.float32),
                tx=np.ascontiguousarray(lut[F.TextureIndex], np.int32))
            keep.append(arrs)
            s = orc_mesh()
            s.nv, s.nvn, s.nf = len(arrs["v"]), len(arrs["vn"]), len(arrs["vi"])
            s.vertices, s.vnormals, s.fnormals = _fp(arrs["v"]), _fp(arrs["vn"]), _fp(arrs["fn"])
            s.vidx, s.nidx, s.uvs, s.tex = _ip(arrs["vi"]), _ip(arrs["ni"]), _fp(arrs["uv"]), _ip(arrs["tx"])
            s.bbox = (C.c_float * 32)(*np.asarray(m.BoundingBox, np.float32).reshape(32).tolist())
            mesh_index[id(m)] = len(meshes)
            meshes.append(s)
        for f, camera in enumerate(cameras):
            view = vm.NewViewMatrix(camera.Position, camera.Direction, camera.Up)
            for i, o in enumerate(objects):
                if rotations_y is not None:
                    rot = np.array([o.Rotation[0], rotations_y[f], o.Rotation[2]], dtype=np.float32)
                    world = vm.NewWorldMatrix(o.Scale, rot, o.Translation)
                    mvp = vm.mvp_matrix(persp, view, world)
                else:
                    world, mvp = renderer.object_matrices(o, camera, persp, view)
                k = f * nobj + i
                objs[k].mesh = mesh_index[id(o.Mesh)]
                objs[k].world = (C.c_float * 16)(*world.reshape(16).tolist())
                objs[k].mvp = (C.c_float * 16)(*mvp.reshape(16).tolist())
        mesh_arr = (orc_mesh * max(len(meshes), 1))(*meshes)
        tex_arr = (orc_texture * max(len(textures), 1))(*textures)
        screen = np.ascontiguousarray(vm.NewScreenMatrix(fb.Width, fb.Height)).reshape(16)
        light = np.ascontiguousarray(vm.light_direction())
        return dict(keep=keep, meshes=mesh_arr, nmesh=len(meshes), textures=tex_arr, ntex=len(textures), objs=objs,
                    nobj=nobj, nframes=len(cameras), screen=screen, light=light, options=renderer.options())

    def draw(self, renderer, objects, camera, threads=0, record=False, rotation_y=None, handle=None):
        """Run the oracle on the same inputs `renderer.Draw(objects, camera)` would get.
        Returns dict(pixels, zbuffer, tpf, writes, triangles, visibility)."""
        fb = renderer.fb
        m = self.marshal(renderer, objects, [camera], None if rotation_y is None else [rotation_y])
        own = handle is None
        r = handle if handle is not None else self.lib.orc_renderer_create(fb.Width, fb.Height, renderer.numTiles, threads)
        assert r
        try:
            self.lib.orc_renderer_record_triangles(r, int(record))
            fog = (C.c_uint8 * 4)(*[int(c) & 0xff for c in getattr(renderer, "FogColor", (100, 100, 100, 255))])
            self.lib.orc_renderer_set_fog(r, float(getattr(renderer, "FogStart", 0.100)),
                                          float(getattr(renderer, "FogEnd", 0.033)), fog)
            rc = self.lib.orc_renderer_draw(r, m["meshes"], m["nmesh"], m["textures"], m["ntex"], m["objs"], m["nobj"],
                                            _fp(m["screen"]), _fp(m["light"]), m["options"])
            assert rc == 0
            n = fb.Width * fb.Height
            px = np.ctypeslib.as_array(C.cast(self.lib.orc_renderer_pixels(r), C.POINTER(C.c_uint8)), (n * 4,))
            zb = np.ctypeslib.as_array(C.cast(self.lib.orc_renderer_zbuffer(r), c_float_p), (n,))
            out = dict(pixels=px.reshape(fb.Height, fb.Width, 4).copy(), zbuffer=zb.reshape(fb.Height, fb.Width).copy(),
                       tpf=int(self.lib.orc_renderer_tpf(r)), writes=int(self.lib.orc_renderer_pixel_writes(r)))
            vis = np.zeros(max(len(objects), 1), np.int32)
            self.lib.orc_renderer_visibility(r, _ip(vis), len(objects))
            out["visibility"] = vis[:len(objects)].copy()
            if record:
                nt = int(self.lib.orc_renderer_num_triangles(r))
                if nt:
                    buf = (C.c_char * (nt * ORC_TRIANGLE_DTYPE.itemsize)).from_address(self.lib.orc_renderer_triangles(r))
                    out["triangles"] = np.frombuffer(buf, dtype=ORC_TRIANGLE_DTYPE).copy()
                else:
                    out["triangles"] = np.zeros(0, ORC_TRIANGLE_DTYPE)
        finally:
            if own:
                self.lib.orc_renderer_destroy(r)
        return out

    def sequence_timer(self, renderer, objects, cameras, rotations_y=None, threads=16):
        """Persistent renderer + marshalled frames; `run(first, count)` times `count` consecutive
        Draw calls in C and returns (seconds, tpf of the last frame)."""
        fb = renderer.fb
        m = self.marshal(renderer, objects, cameras, rotations_y)
        r = self.lib.orc_renderer_create(fb.Width, fb.Height, renderer.numTiles, threads)
        assert r
        lib = self.lib
        stride = C.sizeof(orc_object) * max(m["nobj"], 1)

        class Timer:
            def run(self_, first, count):
                assert 0 <= first and first + count <= m["nframes"]
                objs = C.cast(C.c_void_p(C.addressof(m["objs"]) + first * stride), C.POINTER(orc_object))
                sec = lib.orc_renderer_draw_sequence(r, m["meshes"], m["nmesh"], m["textures"], m["ntex"], objs,
                                                     m["nobj"], count, _fp(m["screen"]), _fp(m["light"]), m["options"])
                assert sec >= 0
                return float(sec), int(lib.orc_renderer_tpf(r))

            def close(self_):
                lib.orc_renderer_destroy(r)

        t = Timer()
        t._keep = m
        return t

    def time_sequence(self, renderer, objects, cameras, rotations_y=None, threads=16, warmup=2):
        """Seconds (measured in C) for len(cameras) consecutive Draw calls in the reference's
        threaded structure; the timed CPU baseline of bench.py.  Returns (seconds, frames, tpf)."""
        fb = renderer.fb
        m = self.marshal(renderer, objects, cameras, rotations_y)
        r = self.lib.orc_renderer_create(fb.Width, fb.Height, renderer.numTiles, threads)
        assert r
        try:
            args = (m["meshes"], m["nmesh"], m["textures"], m["ntex"])
            tail = (_fp(m["screen"]), _fp(m["light"]), m["options"])
            if warmup:
                self.lib.orc_renderer_draw_sequence(r, *args, m["objs"], m["nobj"], min(warmup, m["nframes"]), *tail)
            sec = self.lib.orc_renderer_draw_sequence(r, *args, m["objs"], m["nobj"], m["nframes"], *tail)
            assert sec >= 0
            return float(sec), m["nframes"], int(self.lib.orc_renderer_tpf(r))
        finally:
            self.lib.orc_renderer_destroy(r)
