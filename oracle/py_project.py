"""py_project.py — a second, independent restatement of the reference's geometry side, in plain Python.

TEST INFRASTRUCTURE, NOT PRODUCT (see oracle/gorender_oracle.h).  Companion of py_raster.py: together they
restate the whole of `Renderer.Draw` a second time, straight from the Go source.  This file covers
`projectObject` (renderer.go:254-407) up to the projected `Triangle` list: `matrixMultiplyVec4Batch`
(asm_purego.go:9-19), `Frustum.BoxVisibility` (clipping.go:131-154), `facingCamera` (renderer.go:246-250),
the lighting branches (renderer.go:326-346) with `Vec4.Normalize` dividing by the 4-component length
(vector.go:119-125), `Frustum.ClipTriangle` with `Plane.IsVertexInside / Intersect`, `lerp32`, `lerpUV` and
`Polygon.Triangulate` (clipping.go:35-86, 156-236), and the perspective divide + viewport
(renderer.go:361-370, matrix.go:167-173).  The per-object matrices come from the product's host layer
(gorender_b200/vecmath.py, itself bit-checked against the oracle's restatement of matrix.go).

Every float operation is an np.float32 scalar operation (IEEE binary32, one rounding each, no fusion), in
the reference's order.  Pure-Python loops: small meshes only.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
ONE, ZERO, HALF = f32(1), f32(0), f32(0.5)      # ambientStrength = diffuseStrength = 0.5 (renderer.go:12-13)

BOX_OUTSIDE, BOX_INTERSECT, BOX_INSIDE = 0, 1, 2


def dot4(a, b):            # vector.go:111-113
    return f32(f32(f32(f32(a[0] * b[0]) + f32(a[1] * b[1])) + f32(a[2] * b[2])) + f32(a[3] * b[3]))


def dot3(a, b):            # vector.go:74-76
    return f32(f32(f32(a[0] * b[0]) + f32(a[1] * b[1])) + f32(a[2] * b[2]))


def sub(a, b):
    return [f32(x - y) for x, y in zip(a, b)]


def add(a, b):
    return [f32(x + y) for x, y in zip(a, b)]


def mat_vec(m, v):         # asm_purego.go:13-16 == matrix.go:168-171: ((m0*x + m1*y) + m2*z) + m3*w per row
    return [f32(f32(f32(f32(m[r][0] * v[0]) + f32(m[r][1] * v[1])) + f32(m[r][2] * v[2])) + f32(m[r][3] * v[3]))
            for r in range(4)]


def normalize4(v):         # vector.go:119-125: W takes part in the length
    with np.errstate(all="ignore"):
        ln = f32(np.sqrt(f32(f32(f32(f32(v[0] * v[0]) + f32(v[1] * v[1])) + f32(v[2] * v[2])) + f32(v[3] * v[3]))))
        return [f32(c / ln) for c in v]


def frustum_planes(z_near, z_far):   # clipping.go:100-126, order L, R, T, B, N, F
    P = [(-1, 0, 0, 1), (1, 0, 0, 1), (0, -1, 0, 1), (0, 1, 0, 1), (0, 0, z_near, 1), (0, 0, z_far, 1)]
    N = [(1, 0, 0, 1), (-1, 0, 0, 1), (0, 1, 0, 1), (0, -1, 0, 1), (0, 0, -1, 0), (0, 0, 1, 0)]
    return [([f32(c) for c in p], [f32(c) for c in n]) for p, n in zip(P, N)]


def box_visibility(planes, corners):   # clipping.go:131-154
    for point, normal in planes:
        outside = 0
        for c in corners:
            if f32(dot4(normal, c) - dot4(normal, point)) > 0:      # Plane.DistanceToVertex (clipping.go:69-71)
                outside += 1
        if outside == 8:
            return BOX_OUTSIDE
        if outside > 0:
            return BOX_INTERSECT
    return BOX_INSIDE


def inside(plane, q):      # clipping.go:74-76
    point, normal = plane
    return dot4(sub(q, point), normal) <= 0


def intersect(plane, q0, q1):   # clipping.go:79-86
    point, normal = plane
    u = sub(q1, q0)
    w = sub(q0, point)
    d = dot4(normal, u)
    n = f32(-dot4(normal, w))
    with np.errstate(all="ignore"):
        factor = f32(n / d)
    return add(q0, [f32(c * factor) for c in u]), factor


def lerp32(a, b, t):       # clipping.go:163-165
    return f32(a + f32(f32(b - a) * t))


def clip_triangle(planes, pts, uvs, intens):   # clipping.go:167-236
    poly = [(pts[k], uvs[k], intens[k]) for k in range(3)]
    for plane in planes:
        out = []
        n = len(poly)
        for b in range(n):
            a = (b + 1) % n
            (vA, uvA, iA), (vB, uvB, iB) = poly[a], poly[b]
            if inside(plane, vA):
                if not inside(plane, vB):
                    x, t = intersect(plane, vA, vB)
                    out.append((x, [lerp32(uvA[0], uvB[0], t), lerp32(uvA[1], uvB[1], t)], lerp32(iA, iB, t)))
                out.append((vA, uvA, iA))
            elif inside(plane, vB):
                x, t = intersect(plane, vA, vB)
                out.append((x, [lerp32(uvA[0], uvB[0], t), lerp32(uvA[1], uvB[1], t)], lerp32(iA, iB, t)))
        if not out:
            return []
        poly = out
    if len(poly) < 3:
        return []
    return [(poly[0], poly[i + 1], poly[i + 2]) for i in range(len(poly) - 2)]   # Triangulate (clipping.go:41-62)


def facing_camera(v):      # renderer.go:246-250, on clip-space xyz
    e1, e2 = sub(v[1][:3], v[0][:3]), sub(v[2][:3], v[0][:3])
    n = [f32(f32(e1[1] * e2[2]) - f32(e1[2] * e2[1])), f32(f32(e1[2] * e2[0]) - f32(e1[0] * e2[2])),
         f32(f32(e1[0] * e2[1]) - f32(e1[1] * e2[0]))]                          # vector.go:67-72
    return dot3(n, sub([ZERO, ZERO, ZERO], v[0][:3])) > 0


def project_object(mesh, world, mvp, screen, light, *, tex_ids=(), z_near=0.0, z_far=50.0, BackfaceCulling=True,
                   Lighting=True, FlatShading=False, FrustumClipping=True):
    """projectObject for one object.  `mesh` is a gorender_b200.Mesh, matrices are (4,4) float32 arrays,
    `tex_ids[k]` the scene-wide number of the mesh's k-th texture.
    Returns (visibility, [triangle dicts with points (3,4), uvs (3,2), intensity (3,), tex])."""
    world = [[f32(c) for c in row] for row in np.asarray(world, np.float32)]
    mvp = [[f32(c) for c in row] for row in np.asarray(mvp, np.float32)]
    screen = [[f32(c) for c in row] for row in np.asarray(screen, np.float32)]
    light = [f32(c) for c in light]
    planes = frustum_planes(z_near, z_far)
    bbox = [mat_vec(mvp, [f32(c) for c in corner]) for corner in mesh.BoundingBox]     # renderer.go:268-269
    vis = box_visibility(planes, bbox)
    if vis == BOX_OUTSIDE:
        return vis, []
    tv = [mat_vec(mvp, [f32(c) for c in v]) for v in mesh.Vertices]                    # :303-304
    wfn = [mat_vec(world, [f32(c) for c in v]) for v in mesh.FaceNormals]              # :307-310
    wvn = [mat_vec(world, [f32(c) for c in v]) for v in mesh.VertexNormals]
    has_vn = len(mesh.VertexNormals) != 0
    F = mesh.Faces
    out = []
    for fi in range(len(F)):
        verts = [tv[int(i)] for i in F.VertexIndices[fi]]
        if BackfaceCulling and not facing_camera(verts):
            continue
        if Lighting:
            if has_vn and not FlatShading:
                vint = [f32(HALF + f32(dot3(normalize4(wvn[int(n)])[:3], light) * HALF)) for n in F.NormalIndices[fi]]
            else:
                i = f32(HALF + f32(dot3(normalize4(wfn[fi])[:3], light) * HALF))
                vint = [i, i, i]
        else:
            vint = [HALF, HALF, HALF]
        uvs = [[f32(F.UVs[fi][k][0]), f32(F.UVs[fi][k][1])] for k in range(3)]
        if FrustumClipping and vis != BOX_INSIDE:
            tris = clip_triangle(planes, verts, uvs, vint)
        else:
            tris = [tuple((verts[k], uvs[k], vint[k]) for k in range(3))]
        for tri in tris:
            pts = []
            for (p, _, _) in tri:                                                     # :365-370
                with np.errstate(all="ignore"):
                    q = [f32(c / p[3]) for c in p]
                s = mat_vec(screen, q)
                s[3] = p[3]
                pts.append(s)
            ti = int(F.TextureIndex[fi])
            out.append(dict(points=np.array(pts, np.float32), uvs=np.array([t[1] for t in tri], np.float32),
                            intensity=np.array([t[2] for t in tri], np.float32), tex=tex_ids[ti] if ti >= 0 else -1))
    return vis, out
