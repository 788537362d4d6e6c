// setup.cu — K1 vertex transform and K2 clip-and-emit / triangle setup.
//
// K1 replaces matrixMultiplyVec4Batch on the vertex array
//    (renderer.go:303-304; asm_amd64.s:22-47): one thread per vertex,
//    128-bit load, 16 separately rounded FMUL + 12 FADD, 128-bit store.  It backs the
//    build-tag seam (grb_matrix_multiply_vec4_batch) and the stage read-back of
//    Object.TransformedVertices; the frame path fuses the same transform into K2.
// K2 replaces the per-face loop of projectObject (renderer.go:315-396):
//    gather, backface cull (:246-250), lighting (:326-346), Sutherland–Hodgman
//    frustum clip (clipping.go:167-236), perspective divide + viewport
//    (:361-370), integer snap (:182-184), the reference's tile-list membership
//    rule (:226-244) and TPF, then the emitted triangle's 48-byte record and
//    its entry in the device-tile lists.
//
// One thread per face; a block owns kFaceBlock consecutive faces of one
// object and its warps are independent (no block barrier, no shared memory).  The emitted
// triangles of a warp are compacted with a ballot and stored at statically
// assigned record slots: slot = object base + warp index * slots-per-warp +
// rank in warp.  Slot order is submission order (object, face, fan), so the
// slot doubles as the depth-tie order key of the raster kernel and no scan or
// atomic is needed to place records.  Binning happens here too: the warp
// groups its triangles by (first device tile, 32-slot record segment) with
// match.any and appends ONE 8-byte descriptor {segment base, bit mask} per
// group to the tile's list — one atomic per ~8 triangles, no separate count /
// scan / fill passes.  Triangles straddling tile borders add single-triangle
// descriptors to their other tiles.

#include "gr_types.cuh"
#include "kernels.h"

#ifndef GRB_SETUP_BLOCKS
#define GRB_SETUP_BLOCKS (5 * 256 / GRB_FACE_BLOCK)   // resident blocks per SM the register budget is held to (40 warps: 48 registers)
#endif

namespace gr {

// ---------------------------------------------------------------- K1

__global__ void __launch_bounds__(256) transform_kernel(const __grid_constant__ DrawArgs a) {
    const int frame = blockIdx.y;
    const int vb = blockIdx.x;
    const int o = a.vblkObj[vb];
    const DrawObj ob = a.objs[o];
    const FrameObj &fo = a.frameObjs[(size_t)frame * a.nobj + o];
    if (fo.visibility == GRB_BOX_OUTSIDE) return;  // renderer.go:273-275
    const MeshDev &m = a.meshes[ob.mesh];
    const int v = (vb - ob.vertBlockBase) * 256 + threadIdx.x;
    if (v >= m.nv) return;
    const float4 p = __ldg(&m.verts[v]);
    a.tv[(size_t)frame * a.totalVerts + ob.vertBase + v] = mat_vec(fo.mvp, p);
}

// The build-tag seam itself (asm_amd64.go:8): in place over a device array.
__global__ void __launch_bounds__(256) matvec_batch_kernel(Mat4 m, float4 *vecs, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        vecs[i] = mat_vec(m.m, vecs[i]);
}

// ---------------------------------------------------------------- K2 helpers

struct ClipVert {
    float4 p;
    float u, v, in;
};

// clipping.go:93-126
__device__ __forceinline__ void frustum_plane(int i, float zn, float zf, float4 &P, float4 &N) {
    switch (i) {
    case 0: P = make_float4(-1, 0, 0, 1); N = make_float4(1, 0, 0, 1); break;
    case 1: P = make_float4(1, 0, 0, 1); N = make_float4(-1, 0, 0, 1); break;
    case 2: P = make_float4(0, -1, 0, 1); N = make_float4(0, 1, 0, 1); break;
    case 3: P = make_float4(0, 1, 0, 1); N = make_float4(0, -1, 0, 1); break;
    case 4: P = make_float4(0, 0, zn, 1); N = make_float4(0, 0, -1, 0); break;
    default: P = make_float4(0, 0, zf, 1); N = make_float4(0, 0, 1, 0); break;
    }
}

// clipping.go:74-76
__device__ __forceinline__ bool plane_inside(float4 P, float4 N, float4 q) { return dot4(sub4(q, P), N) <= 0.0f; }

// clipping.go:79-86 + lerp32 / lerpUV (clipping.go:156-165)
__device__ __forceinline__ ClipVert plane_intersect(float4 P, float4 N, const ClipVert &A, const ClipVert &B) {
    const float4 u = sub4(B.p, A.p);
    const float4 w = sub4(A.p, P);
    const float d = dot4(N, u);
    const float n = -dot4(N, w);
    const float t = fdiv(n, d);
    ClipVert r;
    r.p = add4(A.p, mul4(u, t));
    r.in = fadd(A.in, fmul(fsub(B.in, A.in), t));
    r.u = fadd(A.u, fmul(fsub(B.u, A.u), t));
    r.v = fadd(A.v, fmul(fsub(B.v, A.v), t));
    return r;
}

// Frustum.ClipTriangle up to the polygon (clipping.go:167-232); returns the
// vertex count (0 or 3..9) with the polygon left in `poly`.
__device__ int clip_polygon(ClipVert *poly, ClipVert *tmp, float zn, float zf) {
    int count = 3;
    ClipVert *src = poly, *dst = tmp;
#pragma unroll 1
    for (int pi = 0; pi < 6; pi++) {
        float4 P, N;
        frustum_plane(pi, zn, zf, P, N);
        int out = 0;
#pragma unroll 1
        for (int b = 0; b < count; b++) {
            const int ai = (b + 1 == count) ? 0 : b + 1;
            const ClipVert A = src[ai], B = src[b];
            const bool inA = plane_inside(P, N, A.p);
            const bool inB = plane_inside(P, N, B.p);
            // the reference's arrays hold 9 vertices (clipping.go:10) and Go would
            // panic beyond that; a convex polygon never gets there
            if (inA) {
                if (!inB && out < 9) dst[out++] = plane_intersect(P, N, A, B);
                if (out < 9) dst[out++] = A;
            } else if (inB) {
                if (out < 9) dst[out++] = plane_intersect(P, N, A, B);
            }
        }
        if (out == 0) return 0;
        count = out;
        ClipVert *t = src; src = dst; dst = t;
    }
    if (src != poly)
        for (int i = 0; i < count; i++) poly[i] = src[i];
    return count;
}

struct ScreenVert {
    float sx, sy, w;
};

// renderer.go:365-370: p / p.w (4 true divides), full screen-matrix rows, W restored.
// Two of the four divides are skipped when their result provably cannot change sx, sy:
//   * w / w is exactly 1 for every finite non-zero w;
//   * z / w only enters as (0 * qz): the screen matrix has m[0][2] = m[1][2] = 0
//     (matrix.go:108-118), so for a finite quotient the term is +-0 and x + (+-0) == x in every
//     later comparison and in the integer snap.  |z| <= 2^60 and |w| >= 2^-60 keep the quotient
//     finite; anything else (and any other screen matrix) takes the literal path.
__device__ __forceinline__ ScreenVert to_screen(const Mat4 &S, bool noZ, float4 p) {
    ScreenVert s;
    s.w = p.w;
    const float aw = fabsf(p.w);
    const bool fast = noZ && aw >= 8.673617e-19f && aw <= 1.1529215e18f && fabsf(p.z) <= 1.1529215e18f;
    if (fast) {
        const float qx = fdiv(p.x, p.w), qy = fdiv(p.y, p.w);
        // ((m0*qx + m1*qy) + m2*qz) + m3*1 with the m2*qz = +-0 term dropped
        s.sx = fadd(fadd(fmul(S.m[0], qx), fmul(S.m[1], qy)), S.m[3]);
        s.sy = fadd(fadd(fmul(S.m[4], qx), fmul(S.m[5], qy)), S.m[7]);
        return s;
    }
    const float4 q = make_float4(fdiv(p.x, p.w), fdiv(p.y, p.w), fdiv(p.z, p.w), fdiv(p.w, p.w));
    s.sx = mat_row(S.m, q);
    s.sy = mat_row(S.m + 4, q);
    return s;
}

// `int(a.X)` (renderer.go:182): truncation; outside the int32 edge-function
// domain (or NaN) the triangle is reported, not rendered.
__device__ __forceinline__ int snap(float f, bool &bad) {
    if (!(fabsf(f) < (float)(kCoordLimit + 1))) { bad = true; return 0; }
    return __float2int_rz(f);
}

struct Emit {
    TriRec rec;
    int tpf;        // number of reference tile lists this triangle is in
    bool valid;     // gets a record (non-empty raster bbox, ShowFaces)
    bool bad;       // out of the integer domain
    // overlay instantiation only
    bool listed;    // in at least one reference tile list, with snapped coordinates in rec
    int tmax;       // the last reference tile pass that lists it
    int ccx, ccy;   // int(center.X), int(center.Y) of the face-centre mark (renderer.go:203-208)
};

// ---- overlays: FrameBuffer.Pixel / Rect / Line (rasterizer.go:25-30, 54-79) as per-pixel event keys

// Pixel: the reference bounds-checks the LINEAR index only (0 < idx < len), so x outside the row
// wraps into the neighbouring row and pixel 0 is never written.
__device__ __forceinline__ void ovl_pixel(unsigned long long *ovl, long long x, long long y, int width, long long npix,
                                          unsigned long long key) {
    const long long idx = y * width + x;
    if (idx > 0 && idx < npix) atomicMax(ovl + idx, key);
}
__device__ __forceinline__ void ovl_rect3(unsigned long long *ovl, int x, int y, int width, int height, long long npix,
                                          unsigned long long key) {
    if (x >= width || y >= height) return;
    for (int py = y; py < y + 3; py++)
        for (int px = x; px < x + 3; px++) ovl_pixel(ovl, px, py, width, npix, key);
}
// DDA with float32 steps accumulated by repeated addition: inherently serial per line.
__device__ __forceinline__ void ovl_line(unsigned long long *ovl, int x0, int y0, int x1, int y1, int width, long long npix,
                                         unsigned long long key) {
    const int dx = x1 - x0, dy = y1 - y0;
    const int side = max(abs(dx), abs(dy));
    const float xs = fdiv((float)dx, (float)side), ys = fdiv((float)dy, (float)side);
    float cx = (float)x0, cy = (float)y0;
    for (int i = 0; i <= side; i++) {
        // |cx|, |cy| stay within the snapped-coordinate domain: plain truncation == Go's int()
        ovl_pixel(ovl, __float2int_rz(cx), __float2int_rz(cy), width, npix, key);
        cx = fadd(cx, xs);
        cy = fadd(cy, ys);
    }
}
// drawProjection's overlay branches (renderer.go:191-216) for one listed triangle.
__device__ __forceinline__ void draw_overlays(const DrawArgs &a, int frame, const TriRec &t, int tmax, int ccx, int ccy,
                                              uint32_t slot) {
    const long long npix = (long long)a.width * a.height;
    unsigned long long *ovl = a.ovl + (size_t)frame * npix;
    const unsigned long long base =
        ((unsigned long long)(tmax + 1) << kOvlTileShift) | ((unsigned long long)(slot + 1u) << kOvlSlotShift);
    if (a.options & GRB_OPT_SHOW_EDGES) {
        ovl_line(ovl, t.x0, t.y0, t.x1, t.y1, a.width, npix, base | kOvlKindEdge);
        ovl_line(ovl, t.x1, t.y1, t.x2, t.y2, a.width, npix, base | kOvlKindEdge);
        ovl_line(ovl, t.x2, t.y2, t.x0, t.y0, a.width, npix, base | kOvlKindEdge);
        if (a.options & GRB_OPT_SHOW_FACES) ovl_rect3(ovl, ccx - 1, ccy - 1, a.width, a.height, npix, base | kOvlKindEdge);
    }
    if (a.options & GRB_OPT_SHOW_VERTICES) {
        ovl_rect3(ovl, t.x0 - 1, t.y0 - 1, a.width, a.height, npix, base | kOvlKindVertex);
        ovl_rect3(ovl, t.x1 - 1, t.y1 - 1, a.width, a.height, npix, base | kOvlKindVertex);
        ovl_rect3(ovl, t.x2 - 1, t.y2 - 1, a.width, a.height, npix, base | kOvlKindVertex);
    }
}

// Everything between the clipper and the rasteriser for one output triangle.
template <bool OVL>
__device__ __forceinline__ Emit setup_triangle(const DrawArgs &a, ScreenVert s0, ScreenVert s1, ScreenVert s2,
                                               float i0, float i1, float i2, int tex) {
    Emit e;
    e.valid = false;
    e.bad = false;
    e.tpf = 0;
    e.listed = false;

    // identifyTriangleTiles (renderer.go:226-244) on the float screen points
    // Go's min/max propagate NaN (then every tile comparison below is false): one check up front
    // instead of one per operation
    const float nanSum = fadd(fadd(fadd(s0.sx, s1.sx), fadd(s2.sx, s0.sy)), fadd(s1.sy, s2.sy));
    if (nanSum != nanSum && (s0.sx != s0.sx || s1.sx != s1.sx || s2.sx != s2.sx || s0.sy != s0.sy || s1.sy != s1.sy ||
                             s2.sy != s2.sy))
        return e;  // in no tile list: tpf 0, no record
    const float minX = fminf(fminf(s0.sx, s1.sx), s2.sx), maxX = fmaxf(fmaxf(s0.sx, s1.sx), s2.sx);
    const float minY = fminf(fminf(s0.sy, s1.sy), s2.sy), maxY = fmaxf(fmaxf(s0.sy, s1.sy), s2.sy);
    // calculateTileBoundaries (renderer.go:50-76) is a regular grid: column c lists the triangle iff
    // maxX >= sx[c] && minX <= ex[c].  sx grows with c and ex does not shrink, so the columns passing
    // the first test are 0..c1, the ones failing the second 0..c0-1, and the listed ones c0..c1
    // (bounds precomputed by the host, NaN — every comparison false — beyond the grid).
    const RefTiles &g = a.ref;
    const int c1 = (int)(maxX >= g.sx[0]) + (int)(maxX >= g.sx[1]) + (int)(maxX >= g.sx[2]) + (int)(maxX >= g.sx[3]) - 1;
    const int c0 = (int)(minX > g.ex[0]) + (int)(minX > g.ex[1]) + (int)(minX > g.ex[2]) + (int)(minX > g.ex[3]);
    const int r1 = (int)(maxY >= g.sy[0]) + (int)(maxY >= g.sy[1]) + (int)(maxY >= g.sy[2]) + (int)(maxY >= g.sy[3]) - 1;
    const int r0 = (int)(minY > g.ey[0]) + (int)(minY > g.ey[1]) + (int)(minY > g.ey[2]) + (int)(minY > g.ey[3]);
    const int nc = max(c1 - c0 + 1, 0), nr = max(r1 - r0 + 1, 0);
    e.tpf = nc * nr;
    if (e.tpf == 0) return e;
    if (a.tileRowBegin != 0 || a.tileRowEnd != a.nty) {
        // sort-first strips: every rank sees the triangles that touch its rows, so a triangle crossing a strip
        // border is seen twice; its TPF is counted by the rank that owns its top row (clamped into the frame),
        // and the ranks' TPFs add up to the frame's
        const int oy = min(__float2int_rz(fminf(fmaxf(minY, 0.0f), 16384.0f)), a.height - 1);
        if (oy < a.tileRowBegin * kTile || oy >= a.tileRowEnd * kTile) e.tpf = 0;
    }
    if (!OVL && !(a.options & GRB_OPT_SHOW_FACES)) return e;

    bool bad = false;
    TriRec &t = e.rec;
    t.x0 = snap(s0.sx, bad); t.y0 = snap(s0.sy, bad);
    t.x1 = snap(s1.sx, bad); t.y1 = snap(s1.sy, bad);
    t.x2 = snap(s2.sx, bad); t.y2 = snap(s2.sy, bad);
    if (bad) { e.bad = true; return e; }
    if constexpr (OVL) {
        e.listed = true;
        e.tmax = r1 * a.ref.ntx + c1;   // tile index = row * numTilesX + column (renderer.go:62-63)
        // renderer.go:203-208: centre of the float screen points, then int() - 1
        e.ccx = __float2int_rz(fdiv(fadd(fadd(s0.sx, s1.sx), s2.sx), 3.0f));
        e.ccy = __float2int_rz(fdiv(fadd(fadd(s0.sy, s1.sy), s2.sy), 3.0f));
        if (!(a.options & GRB_OPT_SHOW_FACES)) return e;
    }
    t.w0 = s0.w; t.w1 = s1.w; t.w2 = s2.w;
    t.i0 = i0; t.i1 = i1; t.i2 = i2;
    t.tex = tex;
    t.order = 0;

    // Pixel (x,y) is finally owned by the highest-index reference tile that
    // contains it: column min(x / tw, ntx-1).  A triangle is drawn there iff it
    // is in that tile's list, i.e. iff the column and row pass the float test
    // above.  Passing columns/rows are contiguous.
    const int xlo = c0 * a.ref.tw, xhi = (c1 == a.ref.ntx - 1) ? a.width - 1 : (c1 + 1) * a.ref.tw - 1;
    const int ylo = r0 * a.ref.th, yhi = (r1 == a.ref.nty - 1) ? a.height - 1 : (r1 + 1) * a.ref.th - 1;
    // rasterizer.go:99-104 (union over the tiles the triangle is listed in)
    int bx0 = max(max(min(min(t.x0, t.x1), t.x2), xlo), 0);
    int bx1 = min(min(max(max(t.x0, t.x1), t.x2), xhi), a.width - 1);
    int by0 = max(max(min(min(t.y0, t.y1), t.y2), ylo), 0);
    int by1 = min(min(max(max(t.y0, t.y1), t.y2), yhi), a.height - 1);
    // sort-first strip: keep only this rank's rows
    by0 = max(by0, a.tileRowBegin * kTile);
    by1 = min(by1, a.tileRowEnd * kTile - 1);
    if (bx0 > bx1 || by0 > by1) return e;
    t.bx0 = (int16_t)bx0; t.by0 = (int16_t)by0; t.bx1 = (int16_t)bx1; t.by1 = (int16_t)by1;
    e.valid = true;
    return e;
}

// Device tiles covered by a record's raster bbox.
struct TileSpan {
    int tx0, ty0, tx1, ty1;
    __device__ __forceinline__ int count() const { return (tx1 - tx0 + 1) * (ty1 - ty0 + 1); }
};
__device__ __forceinline__ TileSpan tile_span(const TriRec &r) {
    return {r.bx0 / kTile, r.by0 / kTile, r.bx1 / kTile, r.by1 / kTile};
}

// Writes the record (packed to 48 B, 3 x 128-bit stores) and its UVs.
__device__ __forceinline__ void store_record(const DrawArgs &a, int frame, const TriRec &rec, const TriUV &uv,
                                             uint32_t slot) {
    auto pair = [](int lo, int hi) { return (int)(((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16)); };
    int4 *d = reinterpret_cast<int4 *>(a.rec + (size_t)frame * a.recCap + slot);
    d[0] = make_int4(pair(rec.x0, rec.y0), pair(rec.x1, rec.y1), pair(rec.x2, rec.y2), __float_as_int(rec.w0));
    d[1] = make_int4(__float_as_int(rec.w1), __float_as_int(rec.w2), pair(rec.bx0, rec.by0), pair(rec.bx1, rec.by1));
    d[2] = make_int4(__float_as_int(rec.i0), __float_as_int(rec.i1), __float_as_int(rec.i2), rec.tex);
    if (rec.tex >= 0) a.uv[(size_t)frame * a.recCap + slot] = uv;
}

// Appends one descriptor to a tile's list, or to the frame's overflow pool when the tile's in-place
// segment is full.  The pool is bounded (DrawArgs::overflowCap): false = no room anywhere, the caller
// sends the descriptor's triangles to the frame's big list instead, which every tile scans and which
// has a slot for every record — correctness never depends on the pool's size.
__device__ __forceinline__ bool append_desc(const DrawArgs &a, int frame, int tile, uint32_t base, uint32_t mask) {
    const int nTiles = a.ntx * a.nty;
    const uint32_t pos = atomicAdd(&a.descCount[(size_t)frame * nTiles + tile], 1u);
    if (pos < a.descCap) {
        a.desc[((size_t)frame * nTiles + tile) * a.descCap + pos] = {base, mask};
        return true;
    }
    const uint32_t o = atomicAdd(&a.counters[frame].overflowCount, 1u);
    if (o < a.overflowCap) {
        a.overflow[(size_t)frame * a.overflowCap + o] = {(uint32_t)tile, base, mask, 0u};
        return true;
    }
    return false;
}

__device__ __forceinline__ void push_big(const DrawArgs &a, int frame, uint32_t slot) {
    const uint32_t pos = atomicAdd(&a.counters[frame].bigCount, 1u);
    a.bigList[(size_t)frame * a.recCap + pos] = slot;   // at most one entry per record: never full
}

// Tile lists of one emitted triangle beyond its first tile (which the caller has handled: `firstOk`
// false = that append found no room).  Triangles spanning more than kMaxBinsPerTri tiles, and the
// ones whose descriptors found no room, go to the frame's big list — once.
__device__ __forceinline__ void bin_other_tiles(const DrawArgs &a, int frame, const TileSpan &sp, uint32_t slot, bool firstOk) {
    bool ok = firstOk && sp.count() <= kMaxBinsPerTri;
    if (ok)
        for (int ty = sp.ty0; ty <= sp.ty1 && ok; ty++)
            for (int tx = sp.tx0; tx <= sp.tx1 && ok; tx++)
                if (ty != sp.ty0 || tx != sp.tx0) ok = append_desc(a, frame, ty * a.ntx + tx, slot & ~31u, 1u << (slot & 31u));
    if (!ok) {
        push_big(a, frame, slot);
        if (sp.count() <= kMaxBinsPerTri) atomicAdd(&a.counters[frame].listFallbacks, 1u);
    }
}

// renderer.go:328-337: xyz(normalize4(world * n)) . L, with w (=1, translated)
// taking part in the length
__device__ __forceinline__ float light_intensity(const Mat4P &world, float4 n, float lx, float ly, float lz) {
    const float4 wn = mat_vec(world, n);
    const float len = fsqrt(fadd(fadd(fadd(fmul(wn.x, wn.x), fmul(wn.y, wn.y)), fmul(wn.z, wn.z)), fmul(wn.w, wn.w)));
    const float nx = fdiv(wn.x, len), ny = fdiv(wn.y, len), nz = fdiv(wn.z, len);
    return fadd(0.5f, fmul(dot3(nx, ny, nz, lx, ly, lz), 0.5f));
}

// Sort-first strips and off-screen geometry: true when nothing inside the object-space box [lo, hi] can reach the
// rows [tileRowBegin, tileRowEnd) of this draw (or the screen at all).  The box is projected corner by corner; with
// every corner in front of the eye (clip w < 0) the projection of the box — and of everything inside it,
// clipped or not — lies within the corners' screen bounds.  The test is conservative (2 pixels of margin against a
// float32 error of ~0.002 pixels, an approximate reciprocal included; any doubt keeps the box), so results do not
// depend on it: skipped faces emit nothing and count no TPF on this rank.
__device__ __forceinline__ bool box_rejected(const DrawArgs &a, const float *mvp, float4 lo, float4 hi) {
    const float inf = __int_as_float(0x7f800000);
    float x0 = inf, x1 = -inf, y0 = inf, y1 = -inf;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const float4 p = make_float4((k & 1) ? hi.x : lo.x, (k & 2) ? hi.y : lo.y, (k & 4) ? hi.z : lo.z, 1.0f);
        const float4 c = mat_vec(mvp, p);
        const float rw = __frcp_rn(c.w);
        const float sx = fadd(fmul(a.screen.m[0], fmul(c.x, rw)), a.screen.m[3]);
        const float sy = fadd(fmul(a.screen.m[5], fmul(c.y, rw)), a.screen.m[7]);
        ok = ok && c.w < -1e-6f && fabsf(sx) < 1e9f && fabsf(sy) < 1e9f;   // false for NaN / Inf
        x0 = fminf(x0, sx); x1 = fmaxf(x1, sx);
        y0 = fminf(y0, sy); y1 = fmaxf(y1, sy);
    }
    if (!ok) return false;
    const float margin = 2.0f, W = (float)a.width, H = (float)a.height;
    // entirely off screen: in no reference tile list (renderer.go:236-238 needs maxX >= 0, minX <= W, ...)
    if (x1 < -margin || x0 > W + margin || y1 < -margin || y0 > H + margin) return true;
    // rows its triangles can own (TPF) or rasterise, clamped into the frame like setup_triangle does
    const int rlo = min(max(__float2int_rd(y0 - margin), 0), a.height - 1);
    const int rhi = min(max(__float2int_ru(y1 + margin), 0), a.height - 1);
    return rhi < a.tileRowBegin * kTile || rlo >= a.tileRowEnd * kTile;
}

// One thread per (face block, warp of that block): tests the bounds of the warp's 32 faces (mesh.cu) and appends the
// block with the mask of its surviving warps to the frame's block list — in any order: record slots are static, so
// the list's order has no effect on the result.  The setup kernel then touches only what is listed.  Without this
// pass every one of a frame's blocks pays the block prologue's chain of dependent loads before it can retire, which
// bounds the kernel at ~25 us for the 7820 blocks of the C4 frame however little of it a strip keeps.
__global__ void __launch_bounds__(256) reject_kernel(const __grid_constant__ DrawArgs a) {
    const int frame = blockIdx.y;
    const int t = blockIdx.x * 256 + threadIdx.x;
    const int fb = t / kWarpsPerFaceBlock, w = t % kWarpsPerFaceBlock;
    bool keep = false;
    if (fb < a.nFaceBlocks) {
        const int o = a.fblkObj[fb];
        const DrawObj ob = a.objs[o];
        const FrameObj &fo = a.frameObjs[(size_t)frame * a.nobj + o];
        if (fo.visibility != GRB_BOX_OUTSIDE) {
            const MeshDev &m = a.meshes[ob.mesh];
            const int wi = (fb - ob.faceBlockBase) * kWarpsPerFaceBlock + w;
            if (wi * 32 < m.nf) keep = !box_rejected(a, fo.mvp, __ldg(&m.warpLo[wi]), __ldg(&m.warpHi[wi]));
        }
    }
    const unsigned all = __ballot_sync(0xffffffffu, keep);
    const unsigned mask = (all >> ((threadIdx.x & 31u) & ~(unsigned)(kWarpsPerFaceBlock - 1))) & ((1u << kWarpsPerFaceBlock) - 1u);
    if (w == 0 && mask != 0u && fb < a.nFaceBlocks) {
        const uint32_t pos = atomicAdd(&a.blockCount[frame], 1u);
        a.blockList[(size_t)frame * a.nFaceBlocks + pos] = (uint32_t)fb | (mask << 24);
    }
}

// ---------------------------------------------------------------- K2

// OVL: the instantiation that also draws the ShowEdges / ShowVertices overlays (as event keys into
// a.ovl); the frame path proper is the OVL = false one.
template <bool CLIP, bool OVL>
__device__ __forceinline__ void setup_block(const DrawArgs &a, const int frame, const int fb) {
    const int o = a.fblkObj[fb];
    const DrawObj ob = a.objs[o];
    const FrameObj &fo = a.frameObjs[(size_t)frame * a.nobj + o];
    const int vis = fo.visibility;
    if (vis == GRB_BOX_OUTSIDE) return;
    // renderer.go:349: the clipper runs iff FrustumClipping && bbox not fully inside
    const bool objClips = (a.options & GRB_OPT_FRUSTUM_CLIPPING) && vis != GRB_BOX_INSIDE;
    if (objClips != CLIP) return;  // the other instantiation owns this block

    const MeshDev &m = a.meshes[ob.mesh];
    const unsigned lane = threadIdx.x & 31u, warpInBlock = threadIdx.x >> 5;
    const unsigned ltMask = (1u << lane) - 1u;

    // ------------------------------------------------------------ phase 1: transform + cull
    // The MVP transform of the face's three corners (matrixMultiplyVec4Batch, renderer.go:303-304,
    // asm_amd64.s:22-47) is fused here: corners stream in coalesced from the upload-time
    // face-corner expansion, so there is no clip-space vertex array to write and gather back.
    int f = (fb - ob.faceBlockBase) * kFaceBlock + threadIdx.x;
    bool alive = f < m.nf;
    float4 v0, v1, v2;
    if (alive) {
        const Mat4P mvp = pack_mat(fo.mvp, a.fma);
        v0 = mat_vec(mvp, __ldg(&m.cv[0][f]));
        v1 = mat_vec(mvp, __ldg(&m.cv[1][f]));
        v2 = mat_vec(mvp, __ldg(&m.cv[2][f]));
        if (a.options & GRB_OPT_BACKFACE_CULLING) {
            // facingCamera (renderer.go:246-250) on clip-space xyz
            const float e1x = fsub(v1.x, v0.x), e1y = fsub(v1.y, v0.y), e1z = fsub(v1.z, v0.z);
            const float e2x = fsub(v2.x, v0.x), e2y = fsub(v2.y, v0.y), e2z = fsub(v2.z, v0.z);
            const float nx = fsub(fmul(e1y, e2z), fmul(e1z, e2y));
            const float ny = fsub(fmul(e1z, e2x), fmul(e1x, e2z));
            const float nz = fsub(fmul(e1x, e2y), fmul(e1y, e2x));
            const float d = dot3(nx, ny, nz, fsub(0.0f, v0.x), fsub(0.0f, v0.y), fsub(0.0f, v0.z));
            alive = d > 0.0f;
        }
    }

    uint32_t slot;  // record slot(s) of this thread's face: static, in submission order
    if constexpr (!CLIP) {
        // Warps are independent (no block barrier, no shared memory): a warp whose faces are all culled
        // retires here, the others carry their survivors through phase 2 where they are.  (Compacting
        // the survivors of a block across its warps through shared memory was measured on C1, C3 and
        // C4: no gain — culling is spatially coherent, most warps are all-alive or all-dead already.)
        const unsigned aliveMask = __ballot_sync(0xffffffffu, alive);
        const uint32_t warpGlobal = (uint32_t)fb * kWarpsPerFaceBlock + warpInBlock;
        const uint32_t slot0 = fo.slotBase + ((uint32_t)(fb - ob.faceBlockBase) * kWarpsPerFaceBlock + warpInBlock) * kWarpSlots;
        // slots [slot0, slot0 + count) are in use; the ones whose triangle turns out to be
        // invisible are marked with an empty bbox in phase 2 (only the stage read-back looks)
        if (lane == 0 && a.warpCount) a.warpCount[(size_t)frame * a.nFaceBlocks * kWarpsPerFaceBlock + warpGlobal] = __popc(aliveMask);
        if (aliveMask == 0) return;  // whole warp culled
        slot = slot0 + __popc(aliveMask & ltMask);
    }

    // ------------------------------------------------------------ phase 2: light, project, emit
    float in0 = 0.5f, in1 = 0.5f, in2 = 0.5f;  // ambientStrength (renderer.go:342-346)
    int tex = -1;
    if (alive) {
        if (a.options & GRB_OPT_LIGHTING) {
            const Mat4P world = pack_mat(fo.world, a.fma);
            if (m.nvn != 0 && !(a.options & GRB_OPT_FLAT_SHADING)) {
                in0 = light_intensity(world, __ldg(&m.cn[0][f]), a.lx, a.ly, a.lz);
                in1 = light_intensity(world, __ldg(&m.cn[1][f]), a.lx, a.ly, a.lz);
                in2 = light_intensity(world, __ldg(&m.cn[2][f]), a.lx, a.ly, a.lz);
            } else {
                in0 = in1 = in2 = light_intensity(world, __ldg(&m.fnormals[f]), a.lx, a.ly, a.lz);
            }
        }
        // drawProjection (renderer.go:176-178): ShowTextures off => nil texture
        if ((a.options & GRB_OPT_SHOW_TEXTURES) && m.tex != nullptr) tex = __ldg(&m.tex[f]);
        if (tex >= a.ntex) tex = -1;
    }

    TriUV fuv = {0, 0, 0, 0, 0, 0};
    if (alive && m.uvs != nullptr) {
        const float2 a0 = __ldg(&m.uvs[3 * f]), a1 = __ldg(&m.uvs[3 * f + 1]), a2 = __ldg(&m.uvs[3 * f + 2]);
        fuv = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y};
    }

    int tpf = 0, nbad = 0;
    uint32_t emitted = 0;  // warp total of records written

    if constexpr (!CLIP) {
        Emit e;
        e.valid = false;
        if (alive) {
            e = setup_triangle<OVL>(a, to_screen(a.screen, a.screenNoZ, v0), to_screen(a.screen, a.screenNoZ, v1), to_screen(a.screen, a.screenNoZ, v2), in0,
                                    in1, in2, tex);
            tpf = e.tpf;
            nbad = e.bad ? 1 : 0;
            if constexpr (OVL)
                if (e.listed) draw_overlays(a, frame, e.rec, e.tmax, e.ccx, e.ccy, slot);
        }
        const unsigned validMask = __ballot_sync(0xffffffffu, e.valid);
        emitted = __popc(validMask);
        if (e.valid) {
            const TileSpan sp = tile_span(e.rec);
            const bool big = sp.count() > kMaxBinsPerTri;
            e.rec.order = slot;
            store_record(a, frame, e.rec, fuv, slot);
            // one descriptor (one atomic) per (first tile, 32-slot segment) group of the warp
            const unsigned binMask = __ballot_sync(validMask, !big);
            bool firstOk = true;
            if (!big) {
                const int t0 = sp.ty0 * a.ntx + sp.tx0;
                const unsigned long long groupKey = ((unsigned long long)(uint32_t)t0 << 32) | (slot & ~31u);
                const unsigned peers = __match_any_sync(binMask, groupKey);
                const uint32_t groupMask = __reduce_or_sync(peers, 1u << (slot & 31u));
                const int leader = __ffs(peers) - 1;
                int ok = 1;
                if ((int)lane == leader) ok = append_desc(a, frame, t0, slot & ~31u, groupMask) ? 1 : 0;
                firstOk = __shfl_sync(peers, ok, leader) != 0;
            }
            if (big || !firstOk || sp.count() > 1) bin_other_tiles(a, frame, sp, slot, firstOk);
        } else if (alive && a.warpCount) {
            // survived the cull but draws nothing (off screen / ShowFaces off / out of domain):
            // leave an empty bbox in its slot so that the stage read-back skips it
            int4 q = make_int4(0, 0, 1, 0xffff);  // bx0 = 1, by0 = 0, bx1 = -1, by1 = 0
            reinterpret_cast<int4 *>(a.rec + (size_t)frame * a.recCap + slot)[1] = q;
        }
    } else {
        const uint32_t warpGlobal = (uint32_t)fb * kWarpsPerFaceBlock + warpInBlock;
        const uint32_t slot0 = fo.slotBase + ((uint32_t)(fb - ob.faceBlockBase) * kWarpsPerFaceBlock + warpInBlock) * kWarpSlotsClip;
        ClipVert poly[9], tmp[9];
        ScreenVert sv[9];
        int count = 0;
        unsigned validBits = 0;
        int nRaster = 0;   // triangles of this lane that reach the rasteriser
        if (alive) {
            poly[0] = {v0, fuv.u0, fuv.v0, in0};
            poly[1] = {v1, fuv.u1, fuv.v1, in1};
            poly[2] = {v2, fuv.u2, fuv.v2, in2};
            count = clip_polygon(poly, tmp, a.zNear, a.zFar);
            if (count < 3) count = 0;  // Triangulate (clipping.go:47-49)
            for (int i = 0; i < count; i++) sv[i] = to_screen(a.screen, a.screenNoZ, poly[i].p);
            // fan (0, i+1, i+2)  (clipping.go:54-59)
            for (int i = 0; i + 2 < count; i++) {
                const Emit e = setup_triangle<OVL>(a, sv[0], sv[i + 1], sv[i + 2], poly[0].in, poly[i + 1].in,
                                                   poly[i + 2].in, tex);
                tpf += e.tpf;
                nbad += e.bad ? 1 : 0;
                // with overlays every listed triangle takes a slot (its place in the serial order)
                if (OVL ? e.listed : e.valid) validBits |= 1u << i;
                nRaster += e.valid ? 1 : 0;
            }
        }
        // exclusive scan of the per-lane triangle counts over the warp
        const int mine = __popc(validBits);
        int inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, d);
            if ((int)lane >= d) inc += t;
        }
        const uint32_t slotsUsed = (uint32_t)__shfl_sync(0xffffffffu, inc, 31);
        emitted = OVL ? __reduce_add_sync(0xffffffffu, (uint32_t)nRaster) : slotsUsed;
        slot = slot0 + (uint32_t)(inc - mine);
        for (int i = 0; i + 2 < count; i++) {
            if (!(validBits & (1u << i))) continue;
            Emit e = setup_triangle<OVL>(a, sv[0], sv[i + 1], sv[i + 2], poly[0].in, poly[i + 1].in, poly[i + 2].in, tex);
            if constexpr (OVL) {
                draw_overlays(a, frame, e.rec, e.tmax, e.ccx, e.ccy, slot);
                if (!e.valid) {
                    // keeps its slot for the order, draws no face: empty bbox for the stage read-back
                    if (a.warpCount) reinterpret_cast<int4 *>(a.rec + (size_t)frame * a.recCap + slot)[1] = make_int4(0, 0, 1, 0xffff);
                    slot++;
                    continue;
                }
            }
            const TriUV uv = {poly[0].u, poly[0].v, poly[i + 1].u, poly[i + 1].v, poly[i + 2].u, poly[i + 2].v};
            const TileSpan sp = tile_span(e.rec);
            const bool big = sp.count() > kMaxBinsPerTri;
            e.rec.order = slot;
            store_record(a, frame, e.rec, uv, slot);
            bool firstOk = true;
            if (!big) firstOk = append_desc(a, frame, sp.ty0 * a.ntx + sp.tx0, slot & ~31u, 1u << (slot & 31u));
            if (big || !firstOk || sp.count() > 1) bin_other_tiles(a, frame, sp, slot, firstOk);
            slot++;
        }
        if (lane == 0 && a.warpCount) a.warpCount[(size_t)frame * a.nFaceBlocks * kWarpsPerFaceBlock + warpGlobal] = slotsUsed;
    }

    // TPF (renderer.go:436-441) and diagnostics: one atomic per warp
    tpf = (int)__reduce_add_sync(0xffffffffu, (unsigned)tpf);     // REDUX: one instruction each
    nbad = (int)__reduce_add_sync(0xffffffffu, (unsigned)nbad);
    if (lane == 0) {
        if (tpf) atomicAdd(&a.counters[frame].tpf, (unsigned long long)tpf);
        if (nbad) atomicAdd(&a.counters[frame].outOfDomain, (uint32_t)nbad);
        if (emitted) atomicAdd(&a.counters[frame].triCount, emitted);
    }
}

// Whole-frame draws: one block per face block.
template <bool CLIP, bool OVL>
__global__ void __launch_bounds__(kFaceBlock, OVL ? 4 * 256 / GRB_FACE_BLOCK : GRB_SETUP_BLOCKS) setup_kernel(const __grid_constant__ DrawArgs a) {
    setup_block<CLIP, OVL>(a, (int)blockIdx.y, (int)blockIdx.x);
}

// With a block list (strip draws, objects reaching outside the frustum: reject_kernel) a fixed grid walks the frame's
// list — only the blocks that were kept, and of those only the warps whose 32 faces can reach this draw's rows — so
// that a strip keeping a tenth of the frame does not launch the other nine.  (A kernel of its own: the loop costs the
// whole-frame kernel registers it does not have.)
template <bool CLIP>
__global__ void __launch_bounds__(kFaceBlock, GRB_SETUP_BLOCKS) setup_list_kernel(const __grid_constant__ DrawArgs a) {
    const int frame = blockIdx.y;
    const uint32_t n = a.blockCount[frame];
    const uint32_t *list = a.blockList + (size_t)frame * a.nFaceBlocks;
    for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
        const uint32_t entry = list[i];
        if ((entry >> 24) >> (threadIdx.x >> 5) & 1u) setup_block<CLIP, false>(a, frame, (int)(entry & 0xffffffu));
    }
}

// ---------------------------------------------------------------- launchers

void launch_transform(const DrawArgs &a, int nframes, cudaStream_t s) {
    if (a.nVertBlocks == 0) return;
    transform_kernel<<<dim3(a.nVertBlocks, nframes), 256, 0, s>>>(a);
}

void launch_reject(const DrawArgs &a, int nframes, cudaStream_t s) {
    if (a.nFaceBlocks == 0 || a.blockList == nullptr) return;
    reject_kernel<<<dim3((a.nFaceBlocks * kWarpsPerFaceBlock + 255) / 256, nframes), 256, 0, s>>>(a);
}

void launch_setup(const DrawArgs &a, int nframes, bool anyPlain, bool anyClip, cudaStream_t s) {
    if (a.nFaceBlocks == 0) return;
    const dim3 grid(a.nFaceBlocks, nframes);
    if (a.ovl != nullptr) {
        if (anyPlain) setup_kernel<false, true><<<grid, kFaceBlock, 0, s>>>(a);
        if (anyClip) setup_kernel<true, true><<<grid, kFaceBlock, 0, s>>>(a);
    } else if (a.blockList != nullptr) {
        // a few blocks per SM walk the list (the setup warps are independent: no barrier to respect)
        // two resident sets of blocks per frame walk the list (measured on the C4 strips: 13 us per frame against 16 us
        // with one block per list entry at 8 frames per call; equal at one frame per call)
        const dim3 lgrid(min(a.nFaceBlocks, 148 * GRB_SETUP_BLOCKS * 2), nframes);
        if (anyPlain) setup_list_kernel<false><<<lgrid, kFaceBlock, 0, s>>>(a);
        if (anyClip) setup_list_kernel<true><<<lgrid, kFaceBlock, 0, s>>>(a);
    } else {
        if (anyPlain) setup_kernel<false, false><<<grid, kFaceBlock, 0, s>>>(a);
        if (anyClip) setup_kernel<true, false><<<grid, kFaceBlock, 0, s>>>(a);
    }
}

void launch_matvec_batch(const float m[16], float4 *vecs, long long n, cudaStream_t s) {
    if (n <= 0) return;
    Mat4 mm;
    for (int i = 0; i < 16; i++) mm.m[i] = m[i];
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    matvec_batch_kernel<<<(unsigned)blocks, 256, 0, s>>>(mm, vecs, n);
}

}  // namespace gr
