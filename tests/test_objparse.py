"""The native OBJ parser (grb_obj_parse, csrc/objparse.cpp) against the Python mirror of
obj.go: same meshes bit for bit, same quirks, same errors.  Host code only — no GPU needed."""
import time

import numpy as np
import pytest

import gorender_b200 as g
from gorender_b200 import geometry, workloads

from test_cpp_host import write_cube


def assert_same_meshes(a, b):
    assert len(a) == len(b)
    for ma, mb in zip(a, b):
        for name in ("Vertices", "VertexNormals", "FaceNormals", "BoundingBox"):
            x, y = getattr(ma, name), getattr(mb, name)
            assert x.shape == y.shape, name
            nan = np.isnan(x)
            assert np.array_equal(np.isnan(y), nan) and np.array_equal(x.view(np.uint32)[~nan], y.view(np.uint32)[~nan]), name
        fa, fb = ma.Faces, mb.Faces
        assert np.array_equal(fa.VertexIndices, fb.VertexIndices)
        assert np.array_equal(fa.NormalIndices, fb.NormalIndices)
        assert np.array_equal(fa.UVs.view(np.uint32), fb.UVs.view(np.uint32))
        assert np.array_equal(fa.TextureIndex, fb.TextureIndex)
        assert len(fa.Textures) == len(fb.Textures)
        for ta, tb in zip(fa.Textures, fb.Textures):
            assert ta.typ == tb.typ and tuple(ta.color) == tuple(tb.color) and float(ta.scale) == float(tb.scale)
            assert (ta.pixels is None) == (tb.pixels is None)
            if ta.pixels is not None:
                assert np.array_equal(ta.pixels, tb.pixels)


def test_suzanne_and_sphere(tmp_path):
    p = tmp_path / "suzanne.obj"
    geometry.write_obj(workloads.suzanne(), str(p))
    assert_same_meshes(g.LoadObjFile(str(p), False), g.LoadObjFileNative(str(p), False))
    q = tmp_path / "sphere.obj"
    geometry.write_obj(geometry.geodesic_sphere(12, True), str(q))
    assert_same_meshes(g.LoadObjFile(str(q), False), g.LoadObjFileNative(str(q), False))


def test_cube_with_materials(tmp_path):
    obj = write_cube(tmp_path)
    a, b = g.LoadObjFile(obj, False), g.LoadObjFileNative(obj, False)
    assert_same_meshes(a, b)
    assert len(b[0].Faces.Textures) == 1 and b[0].Faces.TextureIndex.tolist() == [0] * 12   # two materials, one map_Kd


MULTI = """# every face syntax, two objects (index offsets), the v//vn quirk, an unknown material, awkward floats
mtllib m.mtl
o first
v 0 0 0
v 1.5e0 0 -0
v 0 1 0
  v 0.1 0.2 0.30000001192092896   
v 16777217 -1e-45 3.4028236e38
vt 0 0
vt 1 0.25
vt 0.5 1
vn 0 0 1
vn 0 1 0
usemtl plain
f 1 2 3
usemtl nosuchmaterial
f 1/1 2/2 3/3
usemtl textured
f 1/1/1 2/2/2 4/3/1
f 1//1 2//2 3//1
o second
v 2 0 0
v 3 0 0
v 2 1 0
vt 0.75 0.75
vn 1 0 0
f 6 7 8
f 6/4/3 7/4/3 8/4/3
usemtl plain
f 6//3 7//3 8//3
"""


def write_multi(tmp_path):
    from PIL import Image

    Image.fromarray(np.arange(8 * 8 * 4, dtype=np.uint8).reshape(8, 8, 4), "RGBA").save(tmp_path / "t.png")
    (tmp_path / "m.mtl").write_text("newmtl plain\nKd 1 0 0\nnewmtl textured\nmap_Kd t.png\n")
    p = tmp_path / "multi.obj"
    p.write_text(MULTI)
    return str(p)


@pytest.mark.parametrize("single", [False, True])
def test_multi_object_file(tmp_path, single):
    p = write_multi(tmp_path)
    a, b = g.LoadObjFile(p, single), g.LoadObjFileNative(p, single)
    assert len(a) == (1 if single else 2)
    assert_same_meshes(a, b)
    if not single:
        assert b[0].Faces.NormalIndices[3].tolist() == [0, 0, -1]      # obj.go:77-89: vn1 takes the third index
        assert b[1].Faces.VertexIndices.tolist() == [[0, 1, 2]] * 3     # second object's indices are rebased
        assert b[0].Faces.TextureIndex.tolist() == [0, -1, 1, 1]        # default texture, nil, image, image


def test_errors(tmp_path):
    quad = tmp_path / "quad.obj"
    quad.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0\nf 1 2 3 4\n")
    with pytest.raises(ValueError, match="mesh is not triangulated"):
        g.LoadObjFileNative(str(quad), False)
    with pytest.raises(RuntimeError, match="no such file"):
        g.LoadObjFileNative(str(tmp_path / "missing.obj"), False)
    empty = tmp_path / "empty.obj"
    empty.write_text("# nothing\n")
    with pytest.raises(ValueError, match="does not have any vertices"):
        g.LoadObjFileNative(str(empty), False)
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0\n")
    with pytest.raises(ValueError, match="unexpected EOF"):
        g.LoadObjFileNative(str(bad), False)
    nomtl = tmp_path / "nomtl.obj"
    nomtl.write_text("mtllib nothere.mtl\nv 0 0 0\n")
    with pytest.raises(RuntimeError, match="failed to parse material library"):
        g.LoadObjFileNative(str(nomtl), False)
    oob = tmp_path / "oob.obj"
    oob.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nf 1/1 2/2 3/1\n")
    with pytest.raises(IndexError):
        g.LoadObjFileNative(str(oob), False)
    with pytest.raises(IndexError):
        g.LoadObjFile(str(oob), False)


def test_native_parser_is_faster(tmp_path):
    """The point of the native parser: a 20 000-face file parses an order of magnitude faster."""
    p = tmp_path / "s.obj"
    geometry.write_obj(geometry.geodesic_sphere(32, True), str(p))   # 20 480 faces, v/vt/vn
    t0 = time.perf_counter()
    a = g.LoadObjFile(str(p), False)
    t1 = time.perf_counter()
    b = g.LoadObjFileNative(str(p), False)
    t2 = time.perf_counter()
    assert_same_meshes(a, b)
    assert (t2 - t1) < (t1 - t0) / 3, (t1 - t0, t2 - t1)


def test_float_parsing_is_correctly_rounded(tmp_path):
    """`%f` into a float32 must be ONE correctly rounded decimal -> binary32 conversion (Go's strconv): the native
    fast path against libc strtof on random decimals, float32 rounding midpoints, subnormals and overflow."""
    import ctypes as C

    rng = np.random.default_rng(1)
    toks = []
    for _ in range(30000):
        k = int(rng.integers(1, 17))
        m = "".join(str(x) for x in rng.integers(0, 10, k))
        dot = int(rng.integers(0, k + 1))
        s = m[:dot] + "." + m[dot:] if rng.random() < 0.8 else m
        if rng.random() < 0.3:
            s += "e%d" % rng.integers(-30, 30)
        toks.append(("-" if rng.random() < 0.5 else "") + s)
    toks += ["16777217", "0.30000001192092896", "1.00000005960464477539", "8388609.5", "1e-45", "3.4028236e38", "1e23",
             "4.5e-39", "0.1", "123456789012345678901234567890", "+7.", ".5", "-0", "0e99"]
    p = tmp_path / "floats.obj"
    p.write_text("".join(f"v {t} 0 0\n" for t in toks))
    got = g.LoadObjFileNative(str(p), False)[0].Vertices[:, 0]
    libc = C.CDLL("libc.so.6")
    libc.strtof.restype = C.c_float
    libc.strtof.argtypes = [C.c_char_p, C.c_void_p]
    want = np.array([libc.strtof(t.encode(), None) for t in toks], np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_forward_references_and_order(tmp_path):
    """A face may only use the texture vertices read so far (obj.go:107-109 indexes c.TextureVertices, a slice that
    grows line by line): a forward reference panics in the reference and must fail here too, also in the threaded
    parser, and of several bad lines the first in file order is the one reported."""
    p = tmp_path / "fwd.obj"
    p.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nf 1/1 2/1 3/2\nvt 1 1\n")
    with pytest.raises(IndexError):
        g.LoadObjFileNative(str(p), False)
    with pytest.raises(IndexError):
        g.LoadObjFile(str(p), False)
    q = tmp_path / "two_errors.obj"
    q.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3 4\nv 1\n")
    with pytest.raises(ValueError, match="mesh is not triangulated"):
        g.LoadObjFileNative(str(q), False)
    r = tmp_path / "two_errors_b.obj"
    r.write_text("v 0 0 0\nv 1\nv 0 1 0\nf 1 2 3 4\n")
    with pytest.raises(ValueError, match="unexpected EOF"):
        g.LoadObjFileNative(str(r), False)


def test_large_file_threaded_path(tmp_path):
    """Enough lines for the threaded passes (>= 20 000 per thread): two objects, identical to the Python mirror."""
    a, b = tmp_path / "a.obj", tmp_path / "b.obj"
    geometry.write_obj(geometry.geodesic_sphere(48, True), str(a))      # 46 080 faces
    geometry.write_obj(geometry.geodesic_sphere(40, False), str(b))     # 32 000 faces, `f a b c`
    # second object's indices continue after the first's (OBJ numbering is global; obj.go:31-40 rebases them)
    na = sum(1 for ln in open(a) if ln.startswith("v "))
    with open(tmp_path / "both.obj", "w") as out:
        out.write(open(a).read())
        for ln in open(b):
            if ln.startswith("f "):
                i, j, k = (int(t) + na for t in ln.split()[1:])
                out.write(f"f {i} {j} {k}\n")
            else:
                out.write(ln)
    path = str(tmp_path / "both.obj")
    for single in (False, True):
        assert_same_meshes(g.LoadObjFile(path, single), g.LoadObjFileNative(path, single))
