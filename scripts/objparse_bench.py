#!/usr/bin/env python
"""Load-time path: parsing a 200 000-face `v/vt/vn` OBJ file — the Python mirror of obj.go
against the native parser (grb_obj_parse: one serial classification pass + threaded bulk parsing)."""
import ctypes as C
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gorender_b200 as g  # noqa: E402
from gorender_b200 import _cabi, geometry  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
path = os.path.join(tempfile.mkdtemp(), "sphere.obj")
geometry.write_obj(geometry.geodesic_sphere(n, True), path)
print(f"{path}: {os.path.getsize(path) / 1e6:.1f} MB, {20 * n * n} faces, {os.cpu_count()} CPUs")
lib = _cabi.load()
err = C.create_string_buffer(256)
for rep in range(3):
    h = C.c_void_p()
    t0 = time.perf_counter()
    rc = lib.grb_obj_parse(path.encode(), 0, C.byref(h), err, len(err))
    dt = time.perf_counter() - t0
    assert rc == 0, err.value
    lib.grb_obj_free(h)
    print(f"grb_obj_parse: {dt * 1e3:.1f} ms ({os.path.getsize(path) / dt / 1e6:.0f} MB/s)")
if "--python" in sys.argv:
    t0 = time.perf_counter()
    g.LoadObjFile(path, False)
    print(f"Python mirror of obj.go: {time.perf_counter() - t0:.2f} s")
