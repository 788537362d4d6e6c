"""The C++ host mirror (gorender_b200/host/): its float32 matrix library, OBJ/MTL loader and PNG
reader against the Python host layer (CPU), and the headless driver end to end against the
Python path (GPU)."""
import os
import subprocess

import numpy as np
import pytest

import gorender_b200 as g
import gorender_b200.vecmath as vm
from gorender_b200 import geometry, workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "gorender_b200", "lib", "gorender_headless")


def run(*args):
    p = subprocess.run([EXE, *map(str, args)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    return p.stdout


def hexwords(m):
    return " ".join(format(int(x), "08x") for x in np.ascontiguousarray(m, np.float32).reshape(-1).view(np.uint32))


def fnv1a(b: bytes) -> int:
    h = 1469598103934665603
    for x in b:
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def write_cube(tmp_path):
    """cube.obj + cube.mtl + texture PNG rebuilt from the committed fixture (same layout as the
    reference's models/cube.*: two materials sharing one map_Kd)."""
    from PIL import Image

    cube = workloads.cube()
    tex = cube.Faces.Textures[0]
    Image.fromarray(tex.pixels, "RGBA").save(tmp_path / "tex.png")  # alpha is 0 or 255: premultiply is idempotent
    (tmp_path / "cube.mtl").write_text("newmtl Side\nKd 1 1 1\nmap_Kd tex.png\n\nnewmtl Top\nmap_Kd tex.png\n")
    lines = ["mtllib cube.mtl", "o Cube"]
    lines += [f"v {x:.9g} {y:.9g} {z:.9g}" for x, y, z, _ in cube.Vertices]
    lines += [f"vn {x:.9g} {y:.9g} {z:.9g}" for x, y, z, _ in cube.VertexNormals]
    F = cube.Faces
    for u, v in F.UVs.reshape(-1, 2):
        lines.append(f"vt {u:.9g} {v:.9g}")
    for i in range(len(F)):
        if i == 0:
            lines.append("usemtl Top")
        if i == 2:
            lines.append("usemtl Side")
        a, b, c = F.VertexIndices[i] + 1
        na, nb, nc = F.NormalIndices[i] + 1
        t = 3 * i + 1
        lines.append(f"f {a}/{t}/{na} {b}/{t + 1}/{nb} {c}/{t + 2}/{nc}")
    (tmp_path / "cube.obj").write_text("\n".join(lines) + "\n")
    return str(tmp_path / "cube.obj")


def test_cpp_matrices_match_python(tmp_path):
    obj = tmp_path / "s.obj"
    geometry.write_obj(workloads.suzanne(), str(obj))
    out = run("-matrices", "-frames", 57, "-start", 11, "-w", 800, "-h", 600, obj).splitlines()
    rot = geometry.spin_rotations(57, start=11)[-1]
    world = vm.NewWorldMatrix([1, 1, 1], [0, rot, 0], [0, 0, 0])
    persp = vm.NewPerspectiveMatrix(np.float32(45 * (np.pi / 180)), np.float32(800) / np.float32(600), 0, 50)
    mvp = vm.mvp_matrix(persp, vm.NewViewMatrix([0, 0, 5], [0, 0, -1], [0, 1, 0]), world)
    assert out[0] == "world " + hexwords(world)
    assert out[1] == "mvp " + hexwords(mvp)
    assert out[2] == "vertices=507 triangles=967"


def test_cpp_png_reader_matches_pil(tmp_path):
    from PIL import Image

    rng = np.random.default_rng(3)
    rgba = rng.integers(0, 256, (16, 32, 4), dtype=np.uint8)
    cases = {
        "rgba.png": Image.fromarray(rgba, "RGBA"),
        "rgb.png": Image.fromarray(rgba[..., :3].copy(), "RGB"),
        "gray.png": Image.fromarray(rgba[..., 0].copy(), "L"),
        "pal.png": Image.fromarray(rgba[..., :3].copy(), "RGB").quantize(16),
        "npot.png": Image.fromarray(rgba[:5, :7].copy(), "RGBA"),
    }
    for name, im in cases.items():
        path = tmp_path / name
        im.save(path)
        t = g.LoadTextureFile(str(path))
        w, h, typ, hsh = run("-texdump", path).split()
        assert (int(w), int(h), int(typ)) == (t.width, t.height, t.typ), name
        assert int(hsh, 16) == fnv1a(t.pixels.tobytes()), name


def test_cpp_obj_loader_errors(tmp_path):
    quad = tmp_path / "quad.obj"
    quad.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0\nf 1 2 3 4\n")
    p = subprocess.run([EXE, "-matrices", str(quad)], capture_output=True, text=True)
    assert p.returncode == 1 and "mesh is not triangulated" in p.stderr
    p = subprocess.run([EXE, "-matrices", str(tmp_path / "missing.obj")], capture_output=True, text=True)
    assert p.returncode == 1 and "no such file" in p.stderr
    p = subprocess.run([EXE, "-matrices", str(tmp_path / "x.stl")], capture_output=True, text=True)
    assert p.returncode == 1 and "unsupported mesh format" in p.stderr


@pytest.mark.gpu
def test_headless_driver_matches_python_path(tmp_path, device):
    """The C++ drop-in (OBJ + MTL + PNG -> Renderer.Draw through the C ABI) renders the same bytes as the
    Python one on frame 2 of the demo spin of the textured cube."""
    obj = write_cube(tmp_path)
    raw = tmp_path / "out.bin"
    out = run("-w", 640, "-h", 360, "-frames", 3, "-raw", raw, obj)
    blob = np.fromfile(raw, dtype=np.uint8)
    px = blob[:640 * 360 * 4].reshape(360, 640, 4)
    z = blob[640 * 360 * 4:].view(np.float32).reshape(360, 640)

    mesh = g.LoadObjFile(obj, False)[0]
    o = g.NewObject(mesh)
    o.Rotation = np.array([0, geometry.spin_rotations(3)[-1], 0], np.float32)
    fb = g.FrameBuffer(640, 360, 1, device)
    r = g.Renderer(fb)
    r.Draw([o], geometry.default_camera())
    assert f"tpf={r.TPF} " in out
    assert np.array_equal(px, fb.Pixels) and np.array_equal(z.view(np.uint32), fb.ZBuffer.view(np.uint32))
    assert (fb.ZBuffer > -1).sum() > 10000
    # the streaming loop (DrawAsync / SwapBuffers / WaitFront) ends on the same frame
    raw2 = tmp_path / "out_async.bin"
    run("-w", 640, "-h", 360, "-frames", 3, "-async", "-raw", raw2, obj)
    assert np.array_equal(np.fromfile(raw2, dtype=np.uint8), blob)
    # the same file through the native OBJ parser and NewMesh on the device
    raw4 = tmp_path / "out_native.bin"
    out4 = run("-w", 640, "-h", 360, "-frames", 3, "-native", "-raw", raw4, obj)
    assert np.array_equal(np.fromfile(raw4, dtype=np.uint8), blob) and f"tpf={r.TPF} " in out4
    # the option hot-keys (main.go:255-272) as flags: wireframe + vertex marks + crosshair
    raw3 = tmp_path / "out_wire.bin"
    run("-w", 640, "-h", 360, "-frames", 3, "-edges", "-vertices", "-crosshair", "-raw", raw3, obj)
    r.ShowEdges = r.ShowVertices = r.CrossHair = True
    r.Draw([o], geometry.default_camera())
    wire = np.fromfile(raw3, dtype=np.uint8)[:640 * 360 * 4].reshape(360, 640, 4)
    assert np.array_equal(wire, fb.Pixels) and not np.array_equal(wire, px)
