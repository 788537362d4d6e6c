"""py_raster.py — a second, independent restatement of the reference's pixel side, in plain Python.

TEST INFRASTRUCTURE, NOT PRODUCT (see oracle/gorender_oracle.h).  Pure-Python loops with numpy float32
scalars: small scenes only.  It exists to cross-check the C++ oracle: the same functions written a
second time, straight from the Go source, in another language — `identifyTriangleTiles`
(renderer.go:226-244), `calculateTileBoundaries` (:50-76), `drawProjection` (:166-217) with its
ShowFaces / ShowEdges / ShowVertices branches, `renderTile` (:219-223), the tail of `Draw`
(:476-482), `FrameBuffer.Clear / DotGrid / Pixel / Rect / Line / Triangle / Fog / CrossHair`
(rasterizer.go:25-217), `colorIntensity` (:81-88), `blendRGBA` (:185-191) and `Texture.Sample`
(texture.go:69-89).  Its input is the list of projected `Triangle`s (renderer.go:29-34) the C++
oracle recorded, so the geometry side is shared and the pixel side is independent.

Every float operation is an np.float32 scalar operation (IEEE binary32, one rounding each, no
fusion), in the reference's order; Go `int` is a Python int; `int(f)` truncates toward zero.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32

FACE_COLOR = (200, 200, 200, 255)     # renderer.go:17
VERTEX_COLOR = (255, 161, 0, 255)     # renderer.go:18
EDGE_COLOR = (0, 0, 0, 255)           # renderer.go:19


def go_int(f) -> int:
    """Go `int(f)` on amd64 (CVTTSS2SQ): truncation toward zero, INT64_MIN for NaN / out of range."""
    f = float(f)
    if not math.isfinite(f) or abs(f) >= 9223372036854775808.0:
        return -(1 << 63)
    return int(math.trunc(f))


def go_mod(a: int, b: int) -> int:
    """Go's integer `%`: the result takes the sign of the dividend."""
    r = abs(a) % abs(b)
    return -r if a < 0 else r


def go_u8(f) -> int:
    """`uint8(f)` on amd64: CVTTSS2SL then the low byte."""
    f = float(f)
    if not math.isfinite(f) or abs(f) >= 2147483648.0:
        return 0
    return int(math.trunc(f)) & 0xFF


class FrameBuffer:
    def __init__(self, width: int, height: int):
        self.Width, self.Height = width, height
        self.Pixels = np.zeros((height * width, 4), np.uint8)
        self.ZBuffer = np.zeros(height * width, np.float32)

    def Clear(self, c):                                  # rasterizer.go:36-44
        self.ZBuffer[:] = f32(-1.0)
        self.Pixels[:] = c

    def Pixel(self, x: int, y: int, c):                  # rasterizer.go:25-30
        idx = y * self.Width + x
        if 0 < idx < len(self.Pixels):
            self.Pixels[idx] = c

    def DotGrid(self, c, step: int):                     # rasterizer.go:46-52
        for y in range(step, self.Height, step):
            for x in range(step, self.Width, step):
                self.Pixel(x, y, c)

    def Rect(self, x, y, width, height, c):              # rasterizer.go:54-63
        if x >= self.Width or y >= self.Height:
            return
        for py in range(y, y + height):
            for px in range(x, x + width):
                self.Pixel(px, py, c)

    def Line(self, x0, y0, x1, y1, c):                   # rasterizer.go:65-79
        dx, dy = x1 - x0, y1 - y0
        side = max(abs(dx), abs(dy))
        with np.errstate(all="ignore"):
            xs, ys = f32(dx) / f32(side), f32(dy) / f32(side)
            cx, cy = f32(x0), f32(y0)
            for _ in range(side + 1):
                self.Pixel(go_int(cx), go_int(cy), c)
                cx = f32(cx + xs)
                cy = f32(cy + ys)

    def CrossHair(self, c):                              # rasterizer.go:209-217
        size, offset = 5, 3
        x, y = self.Width // 2, self.Height // 2
        self.Line(x - size, y, x - offset, y, c)
        self.Line(x + offset, y, x + size, y, c)
        self.Line(x, y - size, x, y - offset, c)
        self.Line(x, y + offset, x, y + size, c)

    def Fog(self, fog_start, fog_end, c):                # rasterizer.go:193-207
        fog_start, fog_end = f32(fog_start), f32(fog_end)
        for i in range(len(self.Pixels)):
            depth = self.ZBuffer[i]
            if depth >= fog_start:
                continue
            if depth <= fog_end:
                self.Pixels[i] = c
            else:
                f = f32(f32(1) - f32(f32(fog_end - depth) / f32(fog_end - fog_start)))
                self.Pixels[i] = blend_rgba(tuple(int(v) for v in self.Pixels[i]), c, f)

    def Triangle(self, x0, y0, z0, u0, v0, x1, y1, z1, u1, v1, x2, y2, z2, u2, v2,
                 tsx, tsy, tex_, tey, ia, ib, ic, texture):   # rasterizer.go:90-183
        minX, maxX = min(x0, x1, x2), max(x0, x1, x2)
        minY, maxY = min(y0, y1, y2), max(y0, y1, y2)
        minX, maxX = max(minX, tsx, 0), min(maxX, tex_, self.Width - 1)
        minY, maxY = max(minY, tsy, 0), min(maxY, tey, self.Height - 1)
        f01 = (y0 - y1) * minX + (x1 - x0) * minY + (x0 * y1 - x1 * y0)
        f12 = (y1 - y2) * minX + (x2 - x1) * minY + (x1 * y2 - x2 * y1)
        f20 = (y2 - y0) * minX + (x0 - x2) * minY + (x2 * y0 - x0 * y2)
        f01dx, f01dy = y0 - y1, x1 - x0
        f12dx, f12dy = y1 - y2, x2 - x1
        f20dx, f20dy = y2 - y0, x0 - x2

        def adjust(f, dx, dy):
            return f if (dy > 0 or (dy == 0 and dx > 0)) else f - 1

        f01, f12, f20 = adjust(f01, f01dx, f01dy), adjust(f12, f12dx, f12dy), adjust(f20, f20dx, f20dy)
        with np.errstate(all="ignore"):
            v0z0, u0z0 = f32(v0 / z0), f32(u0 / z0)
            u1z1, v1z1 = f32(u1 / z1), f32(v1 / z1)
            u2z2, v2z2 = f32(u2 / z2), f32(v2 / z2)
            for y in range(minY, maxY + 1):
                fx01, fx12, fx20 = f01, f12, f20
                for x in range(minX, maxX + 1):
                    if fx01 < 0 and fx12 < 0 and fx20 < 0:
                        s = f32(fx12 + fx20 + fx01)
                        alpha = f32(f32(fx12) / s)
                        beta = f32(f32(fx20) / s)
                        gamma = f32(f32(f32(1) - alpha) - beta)
                        zrec = f32(-f32(f32(f32(alpha / z0) + f32(beta / z1)) + f32(gamma / z2)))
                        index = y * self.Width + x
                        if zrec >= self.ZBuffer[index]:
                            u = f32(f32(f32(f32(alpha * u0z0) + f32(beta * u1z1)) + f32(gamma * u2z2)) / zrec)
                            v = f32(f32(f32(f32(alpha * v0z0) + f32(beta * v1z1)) + f32(gamma * v2z2)) / zrec)
                            if self.affine:   # not in the reference (gorender_oracle.h, ORC_OPT_AFFINE_TEXTURES)
                                u = f32(-f32(f32(f32(alpha * u0) + f32(beta * u1)) + f32(gamma * u2)))
                                v = f32(-f32(f32(f32(alpha * v0) + f32(beta * v1)) + f32(gamma * v2)))
                            intensity = f32(f32(f32(alpha * ia) + f32(beta * ib)) + f32(gamma * ic))
                            c = FACE_COLOR
                            if texture is not None:
                                c = sample(texture, u, v)
                            self.ZBuffer[index] = zrec
                            self.Pixels[index] = color_intensity(c, intensity)
                    fx01 += f01dx
                    fx12 += f12dx
                    fx20 += f20dx
                f01 += f01dy
                f12 += f12dy
                f20 += f20dy


def color_intensity(c, i):                               # rasterizer.go:81-88
    return (go_u8(f32(f32(c[0]) * i)), go_u8(f32(f32(c[1]) * i)), go_u8(f32(f32(c[2]) * i)), c[3])


def blend_rgba(a, b, f):                                 # rasterizer.go:185-191
    g = f32(f32(1) - f)
    return tuple(go_u8(f32(f32(f32(a[k]) * g) + f32(f32(b[k]) * f))) for k in range(4))


def sample(t, u, v):                                     # texture.go:69-89 on gorender_b200.Texture
    if t.typ == 0:
        return tuple(int(x) for x in t.color)
    with np.errstate(all="ignore"):
        fx = go_int(f32(f32(f32(f32(1) - u) * f32(t.scale)) * f32(t.width)))
        fy = go_int(f32(f32(v * f32(t.scale)) * f32(t.height)))
    if t.typ == 2:
        x, y = fx & (t.width - 1), fy & (t.height - 1)
        return tuple(int(c) for c in t.pixels.reshape(-1, 4)[y * t.width + x])
    x, y = go_mod(fx, t.width), go_mod(fy, t.height)
    idx = max(y * t.width + x, 0)
    return tuple(int(c) for c in t.pixels.reshape(-1, 4)[idx])


def tile_boundaries(tile: int, num_tiles: int, width: int, height: int):   # renderer.go:50-76
    if num_tiles == 1:
        return (f32(0), f32(0)), (f32(width), f32(height))
    ntx = int(math.sqrt(num_tiles))
    nty = (num_tiles + ntx - 1) // ntx
    tw = (width + ntx - 1) // ntx
    th = (height + nty - 1) // nty
    sx, sy = f32((tile % ntx) * tw), f32((tile // ntx) * th)
    ex, ey = f32(sx + f32(tw)), f32(sy + f32(th))
    if ex > f32(width):
        ex = f32(width)
    if ey > f32(height):
        ey = f32(height)
    return (sx, sy), (ex, ey)


def draw(width: int, height: int, num_tiles: int, triangles, textures, *, ShowFaces=True, ShowEdges=False,
         ShowVertices=False, ShowTextures=True, CrossHair=False, Fog=False, FogStart=0.100, FogEnd=0.033,
         FogColor=(100, 100, 100, 255), AffineTextures=False):
    """Renderer.Draw from the barrier between the two phases on (renderer.go:448-449, 461-482), serial branch.
    `triangles`: records with fields points (3,4), uvs (3,2), intensity (3,), tex — the oracle's recording, in
    submission order.  Returns (pixels (H,W,4) uint8, zbuffer (H,W) float32, TPF)."""
    fb = FrameBuffer(width, height)
    fb.affine = bool(AffineTextures)
    fb.Clear((50, 50, 50, 255))
    fb.DotGrid((100, 100, 100, 255), 10)
    bounds = [tile_boundaries(i, num_tiles, width, height) for i in range(num_tiles)]
    lists = [[] for _ in range(num_tiles)]
    for t in triangles:                                  # identifyTriangleTiles, renderer.go:226-244
        px, py = t["points"][:, 0], t["points"][:, 1]
        minX, maxX, minY, maxY = min(px), max(px), min(py), max(py)
        if any(np.isnan(v) for v in (minX, maxX, minY, maxY)):
            continue
        for i, ((sx, sy), (ex, ey)) in enumerate(bounds):
            if maxX >= sx and minX <= ex and maxY >= sy and minY <= ey:
                lists[i].append(t)
    for i in range(num_tiles):                           # renderTile / drawProjection, renderer.go:166-223
        (sx, sy), (ex, ey) = bounds[i]
        for t in lists[i]:
            a, b, c = t["points"]
            uv = t["uvs"]
            li = t["intensity"]
            texture = textures[int(t["tex"])] if (ShowTextures and int(t["tex"]) >= 0) else None
            ax, ay, bx, by, cx, cy = (go_int(v) for v in (a[0], a[1], b[0], b[1], c[0], c[1]))
            if ShowFaces:
                fb.Triangle(ax, ay, a[3], uv[0][0], uv[0][1], bx, by, b[3], uv[1][0], uv[1][1],
                            cx, cy, c[3], uv[2][0], uv[2][1], go_int(sx), go_int(sy), go_int(ex), go_int(ey),
                            li[0], li[1], li[2], texture)
            if ShowEdges:
                colr = EDGE_COLOR if ShowFaces else (255, 255, 255, 255)
                fb.Line(ax, ay, bx, by, colr)
                fb.Line(bx, by, cx, cy, colr)
                fb.Line(cx, cy, ax, ay, colr)
                if ShowFaces:
                    mx = f32(f32(f32(a[0] + b[0]) + c[0]) / f32(3))
                    my = f32(f32(f32(a[1] + b[1]) + c[1]) / f32(3))
                    fb.Rect(go_int(mx) - 1, go_int(my) - 1, 3, 3, colr)
            if ShowVertices:
                fb.Rect(ax - 1, ay - 1, 3, 3, VERTEX_COLOR)
                fb.Rect(bx - 1, by - 1, 3, 3, VERTEX_COLOR)
                fb.Rect(cx - 1, cy - 1, 3, 3, VERTEX_COLOR)
    if CrossHair:
        fb.CrossHair((255, 255, 0, 255))
    if Fog:
        fb.Fog(FogStart, FogEnd, FogColor)
    tpf = sum(len(x) for x in lists)
    return fb.Pixels.reshape(height, width, 4), fb.ZBuffer.reshape(height, width), tpf
