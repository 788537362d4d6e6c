"""ctypes binding of the C ABI (include/gorender_b200.h).

The shared library is built in-tree (`gorender_b200/lib/libgorender_b200.so`,
see `__graft_entry__.build()`); if it is missing, importing this module fails
loudly — there is no Python or CPU fallback for the hot path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# GORENDER_B200_LIB: another build of the same library (kernel tuning variants, scripts/tune.sh)
LIB_PATH = os.environ.get("GORENDER_B200_LIB") or os.path.join(_HERE, "lib", "libgorender_b200.so")

GRB_OK = 0
GRB_OPT_FRUSTUM_CLIPPING = 1 << 0
GRB_OPT_SHOW_FACES = 1 << 1
GRB_OPT_BACKFACE_CULLING = 1 << 2
GRB_OPT_LIGHTING = 1 << 3
GRB_OPT_FLAT_SHADING = 1 << 4
GRB_OPT_SHOW_TEXTURES = 1 << 5
GRB_OPT_SHOW_EDGES = 1 << 6
GRB_OPT_SHOW_VERTICES = 1 << 7
GRB_OPT_CROSSHAIR = 1 << 8
GRB_OPT_FOG = 1 << 9
GRB_OPT_AFFINE_TEXTURES = 1 << 10
GRB_OPT_DEFAULT = (GRB_OPT_FRUSTUM_CLIPPING | GRB_OPT_SHOW_FACES | GRB_OPT_BACKFACE_CULLING |
                   GRB_OPT_LIGHTING | GRB_OPT_SHOW_TEXTURES)
GRB_TILE = 32
GRB_IPC_HANDLE_BYTES = 320
GRB_SIGNAL_SLOTS = 64
GRB_ABI_VERSION = 3
GRB_PLANE_COLOR, GRB_PLANE_DEPTH = 0, 1

c_float_p = C.POINTER(C.c_float)
c_i32_p = C.POINTER(C.c_int32)
c_u8_p = C.POINTER(C.c_uint8)


class grb_mesh_desc(C.Structure):
    _fields_ = [
        ("nv", C.c_int32), ("nvn", C.c_int32), ("nf", C.c_int32),
        ("vertices", c_float_p), ("vnormals", c_float_p), ("fnormals", c_float_p),
        ("vidx", c_i32_p), ("nidx", c_i32_p), ("uvs", c_float_p), ("tex", c_i32_p),
        ("bbox", C.c_float * 32),
    ]


class grb_object(C.Structure):
    _fields_ = [("mesh", C.c_int32), ("world", C.c_float * 16), ("mvp", C.c_float * 16)]


class grb_draw_params(C.Structure):
    _fields_ = [
        ("screen", C.c_float * 16), ("light", C.c_float * 3), ("options", C.c_uint32),
        ("z_near", C.c_float), ("z_far", C.c_float), ("ref_tiles", C.c_int32),
        ("row_begin", C.c_int32), ("row_end", C.c_int32),
        ("fog_start", C.c_float), ("fog_end", C.c_float), ("fog_color", C.c_uint8 * 4),
    ]


class grb_triangle_rec(C.Structure):
    _fields_ = [
        ("x0", C.c_int32), ("y0", C.c_int32), ("x1", C.c_int32), ("y1", C.c_int32),
        ("x2", C.c_int32), ("y2", C.c_int32),
        ("w0", C.c_float), ("w1", C.c_float), ("w2", C.c_float),
        ("i0", C.c_float), ("i1", C.c_float), ("i2", C.c_float),
        ("bx0", C.c_int16), ("by0", C.c_int16), ("bx1", C.c_int16), ("by1", C.c_int16),
        ("tex", C.c_int32), ("order", C.c_uint32),
    ]


class grb_frame_stats(C.Structure):
    _fields_ = [("tpf", C.c_int64), ("triangles", C.c_int32), ("big_triangles", C.c_int32),
                ("out_of_domain", C.c_int32), ("list_fallbacks", C.c_int32)]


# numpy mirrors of the two array-of-struct ABI types
OBJECT_DTYPE = np.dtype([("mesh", np.int32), ("world", np.float32, (16,)), ("mvp", np.float32, (16,))])
TRIANGLE_DTYPE = np.dtype([
    ("x0", np.int32), ("y0", np.int32), ("x1", np.int32), ("y1", np.int32), ("x2", np.int32), ("y2", np.int32),
    ("w0", np.float32), ("w1", np.float32), ("w2", np.float32),
    ("i0", np.float32), ("i1", np.float32), ("i2", np.float32),
    ("bx0", np.int16), ("by0", np.int16), ("bx1", np.int16), ("by1", np.int16),
    ("tex", np.int32), ("order", np.uint32)])
STATS_DTYPE = np.dtype([("tpf", np.int64), ("triangles", np.int32), ("big_triangles", np.int32),
                        ("out_of_domain", np.int32), ("list_fallbacks", np.int32)])
assert OBJECT_DTYPE.itemsize == C.sizeof(grb_object) == 132
assert TRIANGLE_DTYPE.itemsize == C.sizeof(grb_triangle_rec) == 64
assert STATS_DTYPE.itemsize == C.sizeof(grb_frame_stats) == 24

# name -> (restype, argtypes): every symbol include/gorender_b200.h declares
_VP = C.c_void_p
SIGNATURES = {
    "grb_abi_version": (C.c_int32, []),
    "grb_last_error": (C.c_char_p, [_VP]),
    "grb_context_create": (C.c_int32, [C.c_int32, C.POINTER(_VP)]),
    "grb_context_destroy": (C.c_int32, [_VP]),
    "grb_context_set_stream": (C.c_int32, [_VP, _VP]),
    "grb_context_synchronize": (C.c_int32, [_VP]),
    "grb_context_set_kernel_timing": (C.c_int32, [_VP, C.c_int32]),
    "grb_context_set_stage_capture": (C.c_int32, [_VP, C.c_int32]),
    "grb_kernel_times": (C.c_int32, [_VP, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "grb_launch_count": (C.c_int64, [_VP]),
    "grb_context_set_workspace_limit": (C.c_int32, [_VP, C.c_uint64]),
    "grb_context_trim": (C.c_int32, [_VP]),
    "grb_host_alloc": (_VP, [C.c_uint64]),
    "grb_host_free": (None, [_VP]),
    "grb_host_register": (C.c_int32, [_VP, C.c_uint64]),
    "grb_host_unregister": (C.c_int32, [_VP]),
    "grb_texture_upload": (C.c_int32, [_VP, C.c_int32, C.c_int32, C.c_int32, C.c_float, c_u8_p, _VP, c_i32_p]),
    "grb_texture_set_scale": (C.c_int32, [_VP, C.c_int32, C.c_float]),
    "grb_mesh_upload": (C.c_int32, [_VP, C.POINTER(grb_mesh_desc), c_i32_p]),
    "grb_mesh_new": (C.c_int32, [_VP, C.POINTER(grb_mesh_desc), c_i32_p]),
    "grb_mesh_read_derived": (C.c_int32, [_VP, C.c_int32, _VP, _VP]),
    "grb_mesh_free": (C.c_int32, [_VP, C.c_int32]),
    "grb_obj_parse": (C.c_int32, [C.c_char_p, C.c_int32, C.POINTER(_VP), C.c_char_p, C.c_int32]),
    "grb_obj_num_meshes": (C.c_int32, [_VP]),
    "grb_obj_num_textures": (C.c_int32, [_VP]),
    "grb_obj_texture_path": (C.c_char_p, [_VP, C.c_int32]),
    "grb_obj_mesh": (C.c_int32, [_VP, C.c_int32, C.POINTER(grb_mesh_desc)]),
    "grb_obj_free": (None, [_VP]),
    "grb_framebuffer_create": (C.c_int32, [_VP, C.c_int32, C.c_int32, C.c_int32, C.POINTER(_VP)]),
    "grb_framebuffer_wrap": (C.c_int32, [_VP, C.c_int32, C.c_int32, C.c_int32, _VP, _VP, C.POINTER(_VP)]),
    "grb_framebuffer_destroy": (C.c_int32, [_VP]),
    "grb_framebuffer_ipc_export": (C.c_int32, [_VP, _VP]),
    "grb_framebuffer_ipc_open": (C.c_int32, [_VP, _VP, C.POINTER(_VP)]),
    "grb_framebuffer_signal": (C.c_int32, [_VP, _VP, C.c_int32, C.c_uint32, C.c_int32]),
    "grb_framebuffer_wait_signals": (C.c_int32, [_VP, _VP, C.c_int32, C.c_int32, C.c_uint32, C.c_int32, C.c_int32]),
    "grb_context_signal_timeouts": (C.c_int64, [_VP]),
    "grb_framebuffer_read_tile_flags": (C.c_int32, [_VP, C.c_int32, _VP]),
    "grb_framebuffer_device_ptrs": (C.c_int32, [_VP, C.POINTER(_VP), C.POINTER(_VP)]),
    "grb_draw_async": (C.c_int32, [_VP, _VP, C.c_int32, C.c_int32, _VP, C.c_int32, C.POINTER(grb_draw_params)]),
    "grb_draw": (C.c_int32, [_VP, _VP, C.c_int32, C.c_int32, _VP, C.c_int32, C.POINTER(grb_draw_params), _VP]),
    "grb_frame_stats_read": (C.c_int32, [_VP, C.c_int32, _VP]),
    "grb_read_frames": (C.c_int32, [_VP, _VP, C.c_int32, C.c_int32, _VP, _VP]),
    "grb_read_frames_async": (C.c_int32, [_VP, _VP, C.c_int32, C.c_int32, _VP, _VP]),
    "grb_framebuffer_wait": (C.c_int32, [_VP]),
    "grb_mirror_create": (C.c_int32, [_VP, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _VP, C.POINTER(_VP)]),
    "grb_mirror_create_on_framebuffer": (C.c_int32, [_VP, _VP, C.c_int32, C.POINTER(_VP)]),
    "grb_mirror_update_rows_async": (C.c_int32, [_VP, _VP, C.c_int32, C.c_int32, _VP, C.c_int32, _VP, C.c_int32, C.c_int32, C.c_int32]),
    "grb_mirror_destroy": (C.c_int32, [_VP]),
    "grb_mirror_invalidate": (C.c_int32, [_VP]),
    "grb_mirror_update_async": (C.c_int32, [_VP, _VP, C.c_int32, C.c_int32, _VP, C.c_int32, _VP, C.c_int32]),
    "grb_mirror_wait": (C.c_int32, [_VP]),
    "grb_mirror_stats": (C.c_int32, [_VP, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "grb_draw_present": (C.c_int32, [_VP, _VP, C.c_int32, C.c_int32, _VP, C.c_int32, C.POINTER(grb_draw_params),
                                     _VP, C.c_int32, _VP, C.c_int32, _VP]),
    "grb_graph_replays": (C.c_int64, [_VP]),
    "grb_matrix_multiply_vec4_batch": (C.c_int32, [_VP, c_float_p, _VP, C.c_int64]),
    "grb_matrix_multiply_vec4_batch_device": (C.c_int32, [_VP, c_float_p, _VP, C.c_int64]),
    "grb_debug_read_transformed": (C.c_int32, [_VP, C.c_int32, _VP, C.c_int64, C.POINTER(C.c_int64)]),
    "grb_debug_read_triangles": (C.c_int32, [_VP, C.c_int32, _VP, _VP, C.c_int64, C.POINTER(C.c_int64)]),
    "grb_debug_set_overflow_cap": (C.c_int32, [_VP, C.c_uint32]),
    "grb_debug_read_visibility": (C.c_int32, [_VP, C.c_int32, _VP, C.c_int32]),
}


class GorenderError(RuntimeError):
    """A C-ABI call returned non-zero.  The reference's Draw cannot fail
    (renderer.go:443); its Go shim panics — this is the Python equivalent."""

    def __init__(self, code: int, message: str):
        super().__init__(f"gorender_b200 error {code}: {message}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """Load the native library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C gorender_b200/csrc`).  gorender_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.grb_abi_version() != GRB_ABI_VERSION:
        raise ImportError("libgorender_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(ctx, code: int) -> None:
    if code != GRB_OK:
        msg = load().grb_last_error(ctx)
        raise GorenderError(code, msg.decode("utf-8", "replace") if msg else "")


def ptr(a: np.ndarray, ctype=None):
    """Pointer to a C-contiguous numpy array (None -> NULL)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    if ctype is None:
        return C.c_void_p(a.ctypes.data)
    return a.ctypes.data_as(C.POINTER(ctype))
