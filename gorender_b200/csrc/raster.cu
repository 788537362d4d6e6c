// raster.cu — K3: the per-tile rasteriser.  Replaces FrameBuffer.Clear /
// DotGrid / Triangle (rasterizer.go:36-52, 90-183), colorIntensity (:81-88),
// Texture.Sample (texture.go:69-89) and renderTile (renderer.go:219-223).
//
// One 256-thread block per 32x32-pixel device tile.  The tile's z-buffer lives
// in shared memory as one 64-bit key per pixel:
//
//     key = orderable(zRec) << 32 | (submission order + 1)
//
// The reference processes triangles serially with `zRec >= ZBuffer[i]`
// (rasterizer.go:156): later triangles win ties, so the surviving fragment of
// a pixel is the lexicographic maximum of (zRec, submission order) — which is
// exactly max(key), in any processing order.  Phase A resolves coverage and
// depth for all triangles of the tile in parallel: the tile's descriptors are
// scattered into a shared list of record slots and the warps take 32 of them
// at a time (tiny triangles: each lane tests its own 4x4 block, the covered
// pixels of the whole warp then run the divide-heavy fine stage densely;
// medium ones: bbox rows dealt to the lanes; large ones: a block-wide pass with
// thread-owned pixels; otherwise a shared-memory atomic max on the key).  Phase B shades only the
// winning fragment of each pixel (perspective-correct UV, Gouraud intensity,
// nearest texel through the read-only path), generates the cleared
// background and dot grid for uncovered pixels, and writes colour and depth
// back with 128-bit stores.  Integer edge functions and every float32
// operation follow the reference's order; results are bit-identical to the
// serial CPU path.

#include <algorithm>

#include "gr_types.cuh"
#include "kernels.h"

namespace gr {

constexpr int kRasterThreads = 256;
#ifndef GRB_SLOTWIN
#define GRB_SLOTWIN 2048
#endif
#ifndef GRB_RASTER_BLOCKS
#define GRB_RASTER_BLOCKS 5   // 48 registers: measured faster than 6 blocks at 40 registers with spills
#endif
constexpr int kRasterBlocksPerSM = GRB_RASTER_BLOCKS;  // resident blocks per SM the register budget is held to
constexpr int kSmallArea = 32;  // bbox∩tile pixels up to which a triangle takes the warp-level coarse/fine path
constexpr int kSmallWidth = 8;  // ... and the widest bbox row that path walks
constexpr int kTinyEdge = 4;    // bbox∩tile of at most kTinyEdge x kTinyEdge pixels: tested by its own lane, fully unrolled
constexpr unsigned long long kBackgroundKey = 0x407FFFFFull << 32;  // orderable(-1.0f) (rasterizer.go:37)

__device__ __forceinline__ uint32_t orderable(float z) {
    const uint32_t b = __float_as_uint(z);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float from_orderable(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Integer edge functions f(x,y) = a*x + b*y + c with the top-left bias folded
// into c (rasterizer.go:107-129).
struct Edges {
    int a01, b01, c01;
    int a12, b12, c12;
    int a20, b20, c20;
};

__device__ __forceinline__ int edge_bias(int dx, int dy) { return (dy > 0 || (dy == 0 && dx > 0)) ? 0 : 1; }

__device__ __forceinline__ Edges make_edges(int x0, int y0, int x1, int y1, int x2, int y2) {
    Edges e;
    e.a01 = y0 - y1; e.b01 = x1 - x0; e.c01 = x0 * y1 - x1 * y0 - edge_bias(e.a01, e.b01);
    e.a12 = y1 - y2; e.b12 = x2 - x1; e.c12 = x1 * y2 - x2 * y1 - edge_bias(e.a12, e.b12);
    e.a20 = y2 - y0; e.b20 = x0 - x2; e.c20 = x2 * y0 - x0 * y2 - edge_bias(e.a20, e.b20);
    return e;
}

// rasterizer.go:149-153
__device__ __forceinline__ void barycentric(int f01, int f12, int f20, float &alpha, float &beta, float &gamma) {
    const float sum = (float)(f12 + f20 + f01);
    alpha = fdiv((float)f12, sum);
    beta = fdiv((float)f20, sum);
    gamma = fsub(fsub(1.0f, alpha), beta);
}
__device__ __forceinline__ float z_reciprocal(float alpha, float beta, float gamma, float z0, float z1, float z2) {
    return -fadd(fadd(fdiv(alpha, z0), fdiv(beta, z1)), fdiv(gamma, z2));
}

// The record slot is the submission order (setup.cu): key = depth | slot + 1; 0 in the low
// word marks the cleared background.
__device__ __forceinline__ unsigned long long fragment_key(float zrec, uint32_t slot) {
    return ((unsigned long long)orderable(zrec) << 32) | (slot + 1u);
}

// Go `int(f)` on amd64 (CVTTSS2SQ): truncation, INT64_MIN when out of range / NaN.
__device__ __forceinline__ long long go_int(float f) {
    return (fabsf(f) < 9223372036854775808.0f) ? __float2ll_rz(f) : (long long)0x8000000000000000ull;
}
// Go `uint8(f)` on amd64: CVTTSS2SL, low byte.
__device__ __forceinline__ unsigned char go_u8(float f) {
    const int i = (fabsf(f) < 2147483648.0f) ? __float2int_rz(f) : (int)0x80000000;
    return (unsigned char)(i & 0xff);
}

// Texture.Sample (texture.go:69-89); texels through the read-only path.
__device__ __forceinline__ uchar4 sample_texture(const TexDev &t, float u, float v) {
    if (t.type == GRB_TEX_SOLID) return t.color;
    const long long fx = go_int(fmul(fmul(fsub(1.0f, u), t.scale), t.widthF));
    const long long fy = go_int(fmul(fmul(v, t.scale), t.heightF));
    long long idx;
    if (t.type == GRB_TEX_IMAGE_FAST) {
        const long long x = fx & (long long)(t.width - 1);
        const long long y = fy & (long long)(t.height - 1);
        idx = y * t.width + x;
    } else if (t.type == GRB_TEX_IMAGE) {
        const long long x = fx % (long long)t.width;
        const long long y = fy % (long long)t.height;
        idx = y * t.width + x;
        if (idx < 0) idx = 0;
    } else {
        return make_uchar4(255, 0, 255, 255);
    }
    return __ldg(&t.pixels[idx]);
}

// ---- small triangles: warp-cooperative coarse / fine rasterisation ---------------------------
//
// A warp takes 32 list entries at a time.  Each lane sets up its triangle (edge functions, the
// bbox clipped to the tile) into a per-warp shared-memory table, struct-of-arrays so that lanes
// reading different triangles hit different banks.  Coarse stage, integer adds only: a triangle
// whose clipped bbox fits 4 x 4 pixels (nearly all of a dense mesh) is tested by its own lane, all
// 16 pixels unrolled into a bit mask; for wider ones (up to 8 x 4) the bbox ROWS of the warp's
// triangles are flattened into one work list (warp prefix sum of the row counts) and a lane takes
// a row.  Covered (triangle, pixel) pairs are pushed into a per-warp ring; whenever 32 are
// available the fine stage runs with every lane busy: 5 IEEE divides for zRec
// (rasterizer.go:149-153) and a shared-memory atomic max on the pixel's key.  One lane per
// triangle would execute the divide sequence at the occupancy of the rare covered pixels; this
// way the expensive part runs dense.
//
// Shared memory is kept small on purpose: the record gathers of this phase live on L1, and every
// 32 KB of carveout the blocks do not need costs measurable time (a double-buffered table that
// avoided the sparse drain after each batch was slower for exactly that reason).

struct WarpTris {            // one per warp, 32 triangles
    int a01[32], b01[32], c01[32];
    int a12[32], b12[32], c12[32];
    int a20[32], b20[32], c20[32];
    float w0[32], w1[32], w2[32];
    uint32_t slot[32];
    uint32_t box[32];        // local x0 | local y0 << 5 | bw << 10
};
constexpr int kFragRing = 64;
#ifndef GRB_LARGEQ
#define GRB_LARGEQ 1024
#endif
constexpr int kLargeQueue = GRB_LARGEQ;  // large triangles a block queues per window before falling back to warp sweeps
constexpr int kDescRound = 256;    // descriptors expanded per round (one per thread)
constexpr int kSlotWin = GRB_SLOTWIN;      // record slots of a round held in shared memory at a time

// fine stage for `count` (<= 32) queued fragments starting at ring position `head`
__device__ __forceinline__ void fine_stage(const WarpTris &wt, const uint32_t *ring, uint32_t head, int count, int lane,
                                           int tileX, int tileY, unsigned long long *keys) {
    if (lane < count) {
        const uint32_t f = ring[(head + lane) & (kFragRing - 1)];
        const int t = f >> 10, lx = f & 31, ly = (f >> 5) & 31;
        const int x = tileX + lx, y = tileY + ly;
        const int f01 = wt.a01[t] * x + wt.b01[t] * y + wt.c01[t];
        const int f12 = wt.a12[t] * x + wt.b12[t] * y + wt.c12[t];
        const int f20 = wt.a20[t] * x + wt.b20[t] * y + wt.c20[t];
        float al, be, ga;
        barycentric(f01, f12, f20, al, be, ga);
        const float z = z_reciprocal(al, be, ga, wt.w0[t], wt.w1[t], wt.w2[t]);
        if (z >= -1.0f) {  // can ever pass `zRec >= ZBuffer` (cleared to -1); false for NaN
            const unsigned long long key = fragment_key(z, wt.slot[t]);
            unsigned long long *p = &keys[ly * kTile + lx];
            if (key > *(volatile unsigned long long *)p) atomicMax(p, key);
        }
    }
    // the ring entries just read are rewritten by the pushes that follow (compute-sanitizer racecheck flags the
    // read-then-write of different lanes without it)
    __syncwarp();
}

// Raster bbox from the second 16 bytes of a record.
struct Box { int x0, y0, x1, y1; };
__device__ __forceinline__ Box unpack_box(int4 q1) {
    return {(int)(int16_t)(q1.z & 0xffff), q1.z >> 16, (int)(int16_t)(q1.w & 0xffff), q1.w >> 16};
}
__device__ __forceinline__ Box load_box(const PackedRec *p) { return unpack_box(__ldg(reinterpret_cast<const int4 *>(p) + 1)); }

// First sector of a record: snapped vertices, w, bbox — all the coverage / depth phase needs.
__device__ __forceinline__ TriRec load_rec_geom(const PackedRec *p) {
    const int4 q0 = __ldg(reinterpret_cast<const int4 *>(p)), q1 = __ldg(reinterpret_cast<const int4 *>(p) + 1);
    TriRec r;
    r.x0 = (int16_t)(q0.x & 0xffff); r.y0 = q0.x >> 16;
    r.x1 = (int16_t)(q0.y & 0xffff); r.y1 = q0.y >> 16;
    r.x2 = (int16_t)(q0.z & 0xffff); r.y2 = q0.z >> 16;
    r.w0 = __int_as_float(q0.w); r.w1 = __int_as_float(q1.x); r.w2 = __int_as_float(q1.y);
    const Box b = unpack_box(q1);
    r.bx0 = (int16_t)b.x0; r.by0 = (int16_t)b.y0; r.bx1 = (int16_t)b.x1; r.by1 = (int16_t)b.y1;
    return r;
}
// ... plus the shading half (intensities, texture).
__device__ __forceinline__ TriRec load_rec(const PackedRec *p) {
    TriRec r = load_rec_geom(p);
    const int4 q2 = __ldg(reinterpret_cast<const int4 *>(p) + 2);
    r.i0 = __int_as_float(q2.x); r.i1 = __int_as_float(q2.y); r.i2 = __int_as_float(q2.z);
    r.tex = q2.w;
    return r;
}

// Clear (rasterizer.go:36-44) + DotGrid step 10 (:46-52) for the four pixels (gx..gx+3, gy) of a
// thread: bit k of the result is set where pixel gx+k is a grid dot.  One modulo per axis.
__device__ __forceinline__ unsigned dot_mask(int gx, int gy) {
    if (gy < 10 || gy % 10 != 0) return 0u;
    const int r = gx % 10;  // gx + k is a multiple of 10 iff r + k is 0 or 10
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
        if ((r + k == 0 || r + k == 10) && gx + k >= 10) m |= 1u << k;
    return m;
}
__device__ __forceinline__ uchar4 background(unsigned dots, int k) {
    return (dots >> k) & 1u ? make_uchar4(100, 100, 100, 255) : make_uchar4(50, 50, 50, 255);
}

// Four consecutive pixels of one row: one 128-bit store each for colour and depth.
// mirC / mirZ (MIRROR instantiation): also store the quad into the host mirrors' planes (frame 0 of the launch).
__device__ __forceinline__ void write_quad(const DrawArgs &a, int frame, int gx, int gy, const uchar4 col[4],
                                           const float zo[4], bool mirC = false, bool mirZ = false) {
    const size_t pix = ((size_t)frame * a.height + gy) * a.width + gx;
    if ((a.width & 3) == 0) {
        // gx is a multiple of 4 and so is width: 16-byte aligned, whole quad in range
        uint4 cq;
        cq.x = *reinterpret_cast<const uint32_t *>(&col[0]); cq.y = *reinterpret_cast<const uint32_t *>(&col[1]);
        cq.z = *reinterpret_cast<const uint32_t *>(&col[2]); cq.w = *reinterpret_cast<const uint32_t *>(&col[3]);
        *reinterpret_cast<uint4 *>(a.color + pix) = cq;
        *reinterpret_cast<float4 *>(a.depth + pix) = make_float4(zo[0], zo[1], zo[2], zo[3]);
        if (mirC) *reinterpret_cast<uint4 *>(a.mirColor + pix) = cq;
        if (mirZ) *reinterpret_cast<float4 *>(a.mirDepth + pix) = make_float4(zo[0], zo[1], zo[2], zo[3]);
    } else {
        for (int k = 0; k < 4 && gx + k < a.width; k++) {
            a.color[pix + k] = col[k];
            a.depth[pix + k] = zo[k];
            if (mirC) a.mirColor[pix + k] = col[k];
            if (mirZ) a.mirDepth[pix + k] = zo[k];
        }
    }
}

// ---- overlays and post passes (POST instantiation of the raster kernel only) -----------------
//
// drawProjection's ShowEdges / ShowVertices writes (renderer.go:191-216) arrive as one event key
// per pixel from the setup kernel (gr_types.cuh, OverlayKey); the pixel's last FACE write happens
// in the pass of the reference tile that owns it, at the winning triangle's list position
// (DESIGN.md §4.3), so the overlay shows iff its key is not older than (owning tile, winner slot):
// a triangle's own edges are drawn right after its face.  Then Draw's post passes in reference
// order: CrossHair, Fog (renderer.go:476-480).  slot1[k] = winner slot + 1, 0 = no face.
__device__ __forceinline__ void post_quad(const DrawArgs &a, int frame, int gx, int gy, uchar4 col[4], const float zo[4],
                                          const uint32_t slot1[4]) {
    const long long npix = (long long)a.width * a.height;
    const long long idx0 = (long long)gy * a.width + gx;
    if (a.ovl != nullptr) {
        const unsigned long long *ovl = a.ovl + (size_t)frame * npix + idx0;
        const int trow = min(gy / a.ref.th, a.ref.nty - 1) * a.ref.ntx;
        const uchar4 edgeCol = (a.options & GRB_OPT_SHOW_FACES) ? make_uchar4(0, 0, 0, 255)         // edgeColor, renderer.go:19
                                                                : make_uchar4(255, 255, 255, 255);  // renderer.go:193-196
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (gx + k >= a.width) break;
            const unsigned long long ok = __ldg(ovl + k);
            if (ok == 0) continue;
            const int tile = trow + min((gx + k) / a.ref.tw, a.ref.ntx - 1);
            const unsigned long long faceKey =
                slot1[k] ? ((unsigned long long)(tile + 1) << kOvlTileShift) | ((unsigned long long)slot1[k] << kOvlSlotShift) : 0ull;
            if (ok >= faceKey) col[k] = (ok & 1ull) ? make_uchar4(255, 161, 0, 255) : edgeCol;  // vertexColor, renderer.go:18
        }
    }
    if (a.options & GRB_OPT_CROSSHAIR) {
        // rasterizer.go:209-217: four 3-pixel axis-aligned lines around (W/2, H/2); Pixel() checks
        // the linear index only, so the test is on the index distance from the centre
        const long long c0 = (long long)(a.height / 2) * a.width + a.width / 2;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const long long idx = idx0 + k;
            if (gx + k >= a.width || idx <= 0) continue;
            const long long d = idx - c0, ad = d < 0 ? -d : d;
            const bool horizontal = ad >= 3 && ad <= 5;
            const bool vertical = ad % a.width == 0 && ad / a.width >= 3 && ad / a.width <= 5;
            if (horizontal || vertical) col[k] = make_uchar4(255, 255, 0, 255);
        }
    }
    if (a.options & GRB_OPT_FOG) {
        // rasterizer.go:185-207
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float depth = zo[k];
            if (depth >= a.fogStart) continue;
            if (depth <= a.fogEnd) { col[k] = a.fogColor; continue; }
            const float f = fsub(1.0f, fdiv(fsub(a.fogEnd, depth), fsub(a.fogEnd, a.fogStart)));
            const float g = fsub(1.0f, f);
            const uchar4 c = col[k], fc = a.fogColor;
            col[k] = make_uchar4(go_u8(fadd(fmul((float)c.x, g), fmul((float)fc.x, f))),
                                 go_u8(fadd(fmul((float)c.y, g), fmul((float)fc.y, f))),
                                 go_u8(fadd(fmul((float)c.z, g), fmul((float)fc.z, f))),
                                 go_u8(fadd(fmul((float)c.w, g), fmul((float)fc.w, f))));
        }
    }
}

// One batch of up to 32 list entries of a warp: lane `lane` holds record slot `slot` when
// `have`.  Small triangles go through the coarse / fine stages; a large one (more than kSmallArea
// pixels inside the tile) is swept by the whole warp, one tile row at a time, lane = column.
__device__ __forceinline__ void process_batch(bool have, uint32_t slot, const PackedRec *rec, WarpTris &wt, uint32_t *ring,
                                              uint32_t &qHead, uint32_t &qTail, int lane, int tileX, int tileY,
                                              int tileX1, int tileY1, unsigned long long *keys, uint32_t *largeQ,
                                              int *largeCount) {
    const unsigned ltMask = (1u << lane) - 1u;
    int rows = 0;   // bbox rows of this lane's triangle inside the tile (0: nothing for the small paths)
    bool large = false, tiny = false;
    Edges e = {};          // this lane's triangle: kept in registers for the tiny path
    int ox = 0, oy = 0, obw = 0;   // its clipped bbox origin (absolute) and width
    if (have) {
        const TriRec r = load_rec_geom(rec + slot);
        const int x0 = max((int)r.bx0, tileX), x1 = min((int)r.bx1, tileX1);
        const int y0 = max((int)r.by0, tileY), y1 = min((int)r.by1, tileY1);
        if (x0 <= x1 && y0 <= y1) {
            const int bw = x1 - x0 + 1, bh = y1 - y0 + 1;
            if (bw <= kSmallWidth && bw * bh <= kSmallArea) {
                e = make_edges(r.x0, r.y0, r.x1, r.y1, r.x2, r.y2);
                ox = x0; oy = y0; obw = bw;
                wt.a01[lane] = e.a01; wt.b01[lane] = e.b01; wt.c01[lane] = e.c01;
                wt.a12[lane] = e.a12; wt.b12[lane] = e.b12; wt.c12[lane] = e.c12;
                wt.a20[lane] = e.a20; wt.b20[lane] = e.b20; wt.c20[lane] = e.c20;
                wt.w0[lane] = r.w0; wt.w1[lane] = r.w1; wt.w2[lane] = r.w2;
                wt.slot[lane] = slot;
                wt.box[lane] = (uint32_t)(x0 - tileX) | ((uint32_t)(y0 - tileY) << 5) | ((uint32_t)bw << 10);
                rows = bh;
                tiny = bw <= kTinyEdge && bh <= kTinyEdge;
            } else {
                large = true;
            }
        }
    }
    // ---- tiny triangles (the bulk of a dense mesh): every lane tests the few pixels of its own
    //      triangle's bbox with integer adds and keeps the covered ones as a bit mask (bit = 8 * row
    //      + column); the fragments of the 32 masks then go into the ring
    uint32_t cover = 0u;
    if (tiny) {
        const int bw = obw, x = ox, y = oy;
        const int a01 = e.a01, a12 = e.a12, a20 = e.a20;
        const int b01 = e.b01, b12 = e.b12, b20 = e.b20;
        int r01 = a01 * x + b01 * y + e.c01;
        int r12 = a12 * x + b12 * y + e.c12;
        int r20 = a20 * x + b20 * y + e.c20;
        // all 16 pixels of the 4 x 4 block at the bbox origin, no loop control; the ones outside the
        // bbox (or the tile) are masked off afterwards
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int f01 = r01, f12 = r12, f20 = r20;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                cover |= ((uint32_t)(f01 & f12 & f20) >> (31 - (8 * j + k))) & (1u << (8 * j + k));
                f01 += a01; f12 += a12; f20 += a20;
            }
            r01 += b01; r12 += b12; r20 += b20;
        }
        const uint32_t rowMask = ((1u << bw) - 1u) * 0x01010101u;
        cover &= rowMask & (0xffffffffu >> (32 - 8 * rows));   // rows in 1..4
        rows = 0;  // not for the row path below
    }
    {
        const uint32_t origin = ((uint32_t)lane << 10) | ((uint32_t)(oy - tileY) << 5) | (uint32_t)(ox - tileX);  // t | ly0 << 5 | lx0
        unsigned pending = __ballot_sync(0xffffffffu, cover != 0);
        while (pending) {
            const bool emit = cover != 0;
            if (emit) {
                const int bit = __ffs((int)cover) - 1;
                cover &= cover - 1u;
                ring[(qTail + __popc(pending & ltMask)) & (kFragRing - 1)] =
                    origin + (uint32_t)(bit & 7) + ((uint32_t)(bit >> 3) << 5);
            }
            qTail += __popc(pending);
            __syncwarp();
            if (qTail - qHead >= 32) {
                fine_stage(wt, ring, qHead, 32, lane, tileX, tileY, keys);
                qHead += 32;
            }
            pending = __ballot_sync(0xffffffffu, cover != 0);
        }
    }
    // ---- medium triangles: flatten the bbox ROWS of the warp's triangles into one list; a lane takes a
    //      row, sets the three edge functions up once and walks its (at most kSmallWidth) pixels
    int incl = rows, total = 0;
    if (__any_sync(0xffffffffu, rows != 0)) {
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        total = __shfl_sync(0xffffffffu, incl, 31);
    }
    const int excl = incl - rows;
    __syncwarp();
    for (int item0 = 0; item0 < total; item0 += 32) {
        const bool active = item0 + lane < total;
        const int item = min(item0 + lane, total - 1);
        // triangle owning this row: number of lanes whose inclusive sum is <= item
        int t = 0;
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
            const int v = __shfl_sync(0xffffffffu, incl, t + step - 1);
            if (v <= item) t += step;
        }
        const int j = item - __shfl_sync(0xffffffffu, excl, t);
        const uint32_t box = wt.box[t];
        const int bw = active ? (int)(box >> 10) : 0;
        const int lx0 = (int)(box & 31u), ly = (int)((box >> 5) & 31u) + j;
        const int x = tileX + lx0, y = tileY + ly;
        const int a01 = wt.a01[t], a12 = wt.a12[t], a20 = wt.a20[t];
        int f01 = a01 * x + wt.b01[t] * y + wt.c01[t];
        int f12 = a12 * x + wt.b12[t] * y + wt.c12[t];
        int f20 = a20 * x + wt.b20[t] * y + wt.c20[t];
        const uint32_t tag = ((uint32_t)t << 10) | ((uint32_t)ly << 5);
        const int maxBw = __reduce_max_sync(0xffffffffu, bw);
        for (int k = 0; k < maxBw; k++) {
            const bool inside = k < bw && (f01 & f12 & f20) < 0;
            const unsigned m = __ballot_sync(0xffffffffu, inside);
            if (m) {
                if (inside) ring[(qTail + __popc(m & ltMask)) & (kFragRing - 1)] = tag | (uint32_t)(lx0 + k);
                qTail += __popc(m);
                __syncwarp();
                if (qTail - qHead >= 32) {
                    fine_stage(wt, ring, qHead, 32, lane, tileX, tileY, keys);
                    qHead += 32;
                }
            }
            f01 += a01; f12 += a12; f20 += a20;
        }
    }
    // the table is rewritten by the next batch: drain what is left
    if (qTail != qHead) {
        fine_stage(wt, ring, qHead, (int)(qTail - qHead), lane, tileX, tileY, keys);
        qHead = qTail;
    }
    __syncwarp();

    // ---- large triangles of the batch are queued for the block-wide pass (thread-owned pixels,
    // no atomics); if the queue is full the warp sweeps the tile rows itself, lane = column
    if (large) {
        const int pos = atomicAdd(largeCount, 1);
        if (pos < kLargeQueue) {
            largeQ[pos] = slot;
            large = false;
        }
    }
    unsigned largeMask = __ballot_sync(0xffffffffu, large);
    while (largeMask) {
        const int src = __ffs(largeMask) - 1;
        largeMask &= largeMask - 1;
        const uint32_t s = __shfl_sync(0xffffffffu, slot, src);
        const TriRec r = load_rec_geom(rec + s);  // same address for every lane: one broadcast transaction
        const Edges e = make_edges(r.x0, r.y0, r.x1, r.y1, r.x2, r.y2);
        const int x = tileX + lane;
        const int y0 = max((int)r.by0, tileY), y1 = min((int)r.by1, tileY1);
        const bool col = x >= r.bx0 && x <= r.bx1 && x <= tileX1;
        int f01 = e.a01 * x + e.b01 * y0 + e.c01;
        int f12 = e.a12 * x + e.b12 * y0 + e.c12;
        int f20 = e.a20 * x + e.b20 * y0 + e.c20;
        for (int y = y0; y <= y1; y++) {
            if (col && (f01 & f12 & f20) < 0) {
                float al, be, ga;
                barycentric(f01, f12, f20, al, be, ga);
                const float z = z_reciprocal(al, be, ga, r.w0, r.w1, r.w2);
                if (z >= -1.0f) {
                    const unsigned long long key = fragment_key(z, s);
                    unsigned long long *p = &keys[(y - tileY) * kTile + lane];
                    if (key > *(volatile unsigned long long *)p) atomicMax(p, key);
                }
            }
            f01 += e.b01; f12 += e.b12; f20 += e.b20;
        }
    }
}

// Block-wide pass over the queued large triangles: every thread owns 4 consecutive pixels of the
// tile and keeps their keys in registers while it walks the queue; no atomics.
__device__ __forceinline__ void coop_pass(const uint32_t *largeQ, int nq, const PackedRec *rec, int gx, int gy, int px, int py,
                                          unsigned long long *keys) {
    unsigned long long k0 = keys[py * kTile + px], k1 = keys[py * kTile + px + 1];
    unsigned long long k2 = keys[py * kTile + px + 2], k3 = keys[py * kTile + px + 3];
    // bbox first (one 16-byte broadcast load): most threads are outside a medium triangle; the
    // next entry's bbox is fetched one iteration ahead
    Box bNext = load_box(rec + largeQ[0]);
    for (int q = 0; q < nq; q++) {
        const uint32_t slot = largeQ[q];
        const Box b = bNext;
        if (q + 1 < nq) bNext = load_box(rec + largeQ[q + 1]);
        const int bx0 = b.x0, by0 = b.y0, bx1 = b.x1, by1 = b.y1;
        if (gy < by0 || gy > by1 || gx > bx1 || gx + 3 < bx0) continue;
        const TriRec r = load_rec_geom(rec + slot);
        const Edges e = make_edges(r.x0, r.y0, r.x1, r.y1, r.x2, r.y2);
        int f01 = e.a01 * gx + e.b01 * gy + e.c01;
        int f12 = e.a12 * gx + e.b12 * gy + e.c12;
        int f20 = e.a20 * gx + e.b20 * gy + e.c20;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int x = gx + k;
            if ((f01 & f12 & f20) < 0 && x >= bx0 && x <= bx1) {
                float al, be, ga;
                barycentric(f01, f12, f20, al, be, ga);
                const float z = z_reciprocal(al, be, ga, r.w0, r.w1, r.w2);
                if (z >= -1.0f) {
                    const unsigned long long key = fragment_key(z, slot);
                    if (k == 0) k0 = max(k0, key);
                    if (k == 1) k1 = max(k1, key);
                    if (k == 2) k2 = max(k2, key);
                    if (k == 3) k3 = max(k3, key);
                }
            }
            f01 += e.a01; f12 += e.a12; f20 += e.a20;
        }
    }
    keys[py * kTile + px] = k0; keys[py * kTile + px + 1] = k1;
    keys[py * kTile + px + 2] = k2; keys[py * kTile + px + 3] = k3;
}

// POST: the instantiation that also composes the overlays and runs the post passes (post_quad).
// MIRROR: one-frame draws with host mirrors attached — the write-back also goes to the mirrors' host planes, under the
// mirrors' rule (present.cu): a tile is written iff it is busy now or was busy in the host copy.
template <bool POST, bool MIRROR>
__global__ void __launch_bounds__(kRasterThreads, kRasterBlocksPerSM) raster_kernel(const __grid_constant__ DrawArgs a) {
    __shared__ unsigned long long keys[kTilePix];
    __shared__ WarpTris tris[kRasterThreads / 32];
    __shared__ uint32_t fragRing[kRasterThreads / 32][kFragRing];
    __shared__ uint32_t slotList[kSlotWin];
    __shared__ uint32_t warpSums[kRasterThreads / 32];
    __shared__ uint32_t largeQ[kLargeQueue];
    // Two counters used alternately (one per window of entries): the one a window does NOT use is
    // re-zeroed behind that window's first barrier, so a reset never races with the reads of the
    // window before.
    __shared__ int largeCount[2];

    const int tid = threadIdx.x;
    const int nTiles = a.ntx * a.nty;
    // pixels owned by this thread inside the tile: 4 consecutive in x
    const int px = (tid & 7) * 4, py = tid >> 3;

    const int frame = blockIdx.z;
    const int tx = blockIdx.x, ty = blockIdx.y + a.tileRowBegin;
    const int tile = ty * a.ntx + tx;
    const int tileX = tx * kTile, tileY = ty * kTile;

    const size_t frameTile = (size_t)frame * nTiles + tile;
    uint32_t *descCount = a.descCount + frameTile;
    const uint32_t nDescAll = *descCount;
    const uint32_t nBig = a.counters[frame].bigCount;

    const int gx = tileX + px, gy = tileY + py;
    const bool inImage = gy < a.height && gx < a.width;

    // ---- no list entries: unless one of the frame's big triangles reaches this tile, nothing
    // touches it (block-uniform decision)
    bool empty = nDescAll == 0;
    if (empty && nBig != 0) {
        const uint32_t *big = a.bigList + (size_t)frame * a.recCap;
        const PackedRec *rec = a.rec + (size_t)frame * a.recCap;
        bool hit = false;
        for (uint32_t i = tid; i < nBig; i += kRasterThreads) {
            const Box b = load_box(rec + big[i]);
            const int bx0 = b.x0, by0 = b.y0, bx1 = b.x1, by1 = b.y1;
            hit |= bx0 < tileX + kTile && bx1 >= tileX && by0 < tileY + kTile && by1 >= tileY;
        }
        empty = !__syncthreads_or(hit);
    }
    // host mirrors and strip pushes (present.cu) skip tiles that hold nothing but the cleared background; with
    // overlays or post passes any tile may differ from it
    if (tid == 0 && a.tileBusy != nullptr) a.tileBusy[frameTile] = (POST || !empty) ? 1 : 0;
    bool mirC = false, mirZ = false;
    if constexpr (MIRROR) {   // (frame == 0: one-frame launches only)
        const bool busy = POST || !empty;
        mirC = a.mirColor != nullptr && (busy || a.mirDirtyColor[tile] != 0);
        mirZ = a.mirDepth != nullptr && (busy || a.mirDirtyDepth[tile] != 0);
        __syncthreads();      // every warp has read the host copy's flags before they are rewritten
        if (tid == 0) {
            if (mirC) { a.mirDirtyColor[tile] = busy ? 1 : 0; atomicAdd(a.mirWrittenColor, 1ull); }
            if (mirZ) { a.mirDirtyDepth[tile] = busy ? 1 : 0; atomicAdd(a.mirWrittenDepth, 1ull); }
        }
    }
    // ---- cleared background straight to HBM
    if (empty) {
        if (inImage) {
            uchar4 col[4];
            float zo[4];
            const unsigned dots = dot_mask(gx, gy);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                col[k] = background(dots, k);
                zo[k] = -1.0f;
            }
            if constexpr (POST) {
                const uint32_t none[4] = {0u, 0u, 0u, 0u};
                post_quad(a, frame, gx, gy, col, zo, none);
            }
            write_quad(a, frame, gx, gy, col, zo, mirC, mirZ);
        }
        return;
    }

    const int tileX1 = min(tileX + kTile, a.width) - 1, tileY1 = min(tileY + kTile, a.height) - 1;
    const PackedRec *rec = a.rec + (size_t)frame * a.recCap;
    const TileDesc *desc = a.desc + ((size_t)frame * nTiles + tile) * a.descCap;
    const uint32_t nDesc = min(nDescAll, a.descCap);
    // (only busy tiles get here: the 86 % of C3's tile blocks that are empty have returned above)
    const uint32_t nOverflow = nDescAll > a.descCap ? min(a.counters[frame].overflowCount, a.overflowCap) : 0u;

    for (int i = tid; i < kTilePix; i += kRasterThreads) keys[i] = kBackgroundKey;
    if (tid < 2) largeCount[tid] = 0;
    __syncthreads();
    if (tid == 0) *descCount = 0;  // ready for the next draw (nobody else reads this tile's counter)

    // ------------------------------------------------------------ phase A: coverage + depth
    {
        const int lane = tid & 31, warp = tid >> 5;
        constexpr int kWarps = kRasterThreads / 32;
        WarpTris &wt = tris[warp];
        uint32_t *ring = fragRing[warp];
        uint32_t qHead = 0, qTail = 0;  // warp-uniform
        uint32_t win = 0;               // windows / big rounds so far (block-uniform): picks the large-triangle counter
        const OverflowDesc *ov = a.overflow + (size_t)frame * a.overflowCap;
        const uint32_t *big = a.bigList + (size_t)frame * a.recCap;
        // Rounds of up to 256 descriptors: first the tile's in-place list, then (rarely) the frame's
        // overflow list filtered by tile, finally the frame's big list as pseudo-descriptors.
        const uint32_t nRoundsOwn = (nDesc + kDescRound - 1) / kDescRound;
        const uint32_t nRoundsOv = (nOverflow + kDescRound - 1) / kDescRound;
        const uint32_t nRoundsBig = (nBig + kDescRound - 1) / kDescRound;
        for (uint32_t round = 0; round < nRoundsOwn + nRoundsOv + nRoundsBig; round++) {
            // -- one descriptor per thread
            TileDesc d = {0u, 0u};
            if (round < nRoundsOwn) {
                const uint32_t i = round * kDescRound + tid;
                if (i < nDesc) d = desc[i];
            } else if (round < nRoundsOwn + nRoundsOv) {
                const uint32_t i = (round - nRoundsOwn) * kDescRound + tid;
                if (i < nOverflow) {
                    const OverflowDesc o = ov[i];
                    if ((int)o.tile == tile) d = {o.base, o.mask};
                }
            } else {
                // big-list entries skip the expansion: whatever overlaps the tile goes straight to
                // the block-wide pass (256 entries per round never overflow the queue)
                int *counter = &largeCount[win & 1];
                const uint32_t i = (round - nRoundsOwn - nRoundsOv) * kDescRound + tid;
                if (i < nBig) {
                    const uint32_t slot = big[i];
                    const Box b = load_box(rec + slot);
                    const int bx0 = b.x0, by0 = b.y0, bx1 = b.x1, by1 = b.y1;
                    if (bx0 <= tileX1 && bx1 >= tileX && by0 <= tileY1 && by1 >= tileY) largeQ[atomicAdd(counter, 1)] = slot;
                }
                __syncthreads();
                if (tid == 0) largeCount[(win + 1) & 1] = 0;
                const int nqBig = *counter;
                if (nqBig) coop_pass(largeQ, nqBig, rec, gx, gy, px, py, keys);
                __syncthreads();
                win++;
                continue;
            }
            // -- block-wide exclusive scan of the triangle counts: the slots of this thread's
            //    descriptor are entries [first, first + cnt) of the round
            const uint32_t cnt = __popc(d.mask);
            uint32_t incl = cnt;
#pragma unroll
            for (int s_ = 1; s_ < 32; s_ <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, s_);
                if (lane >= s_) incl += v;
            }
            if (lane == 31) warpSums[warp] = incl;
            __syncthreads();
            uint32_t wbase = 0, total = 0;
#pragma unroll
            for (int w = 0; w < kWarps; w++) {
                const uint32_t v = warpSums[w];
                if (w < warp) wbase += v;
                total += v;
            }
            const uint32_t first = wbase + incl - cnt;
            if (total == 0) {  // (an overflow round without entries for this tile) warpSums is rewritten next round
                __syncthreads();
                continue;
            }
            // -- windows of kSlotWin entries: every thread scatters its descriptor's record slots
            //    into the shared list (one store per triangle, no searching), then the warps take
            //    32 consecutive entries at a time
            for (uint32_t win0 = 0; win0 < total; win0 += kSlotWin, win++) {
                int *counter = &largeCount[win & 1];
                if (first < win0 + kSlotWin && first + cnt > win0) {
                    uint32_t m = d.mask, pos = first - win0;  // wraps below the window: fails the range test
                    while (m) {
                        const uint32_t b = (uint32_t)__ffs(m) - 1u;
                        m &= m - 1u;
                        if (pos < (uint32_t)kSlotWin) {
                            slotList[pos] = d.base + b;
                        }
                        pos++;
                    }
                }
                __syncthreads();
                if (tid == 0) largeCount[(win + 1) & 1] = 0;
                const uint32_t nWin = min(total - win0, (uint32_t)kSlotWin);
                const uint32_t nBatches = (nWin + 31) / 32;
                for (uint32_t b = warp; b < nBatches; b += kWarps) {
                    const uint32_t e = b * 32 + lane;
                    const bool have = e < nWin;
                    const uint32_t slot = have ? slotList[e] : 0u;
                    {   // the geometry sector of this warp's next batch goes to L1 one batch ahead of its use (-2 %)
                        const uint32_t en = e + kWarps * 32;
                        if (en < nWin) asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + slotList[en]));
                    }
                    process_batch(have, slot, rec, wt, ring, qHead, qTail, lane, tileX, tileY, tileX1, tileY1, keys, largeQ,
                                  counter);
                }
                __syncthreads();
                // -- large triangles found in this window
                const int nq = min(*counter, kLargeQueue);
                if (nq) {
                    coop_pass(largeQ, nq, rec, gx, gy, px, py, keys);
                    __syncthreads();
                }
            }
        }
    }

    // ------------------------------------------------------------ phase B
    if (inImage) {
    uchar4 col[4];
    float zo[4];
    TriRec r;
    Edges e;
    uint32_t have = 0;  // slot + 1 of the record held in r / e
    uint32_t winner[4];
    const unsigned dots = dot_mask(gx, gy);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int x = gx + k;
        const unsigned long long key = keys[py * kTile + px + k];
        const uint32_t slot1 = (uint32_t)key;
        winner[k] = slot1;
        if (slot1 == 0) {
            col[k] = background(dots, k);
            zo[k] = -1.0f;
            continue;
        }
        const uint32_t slot = slot1 - 1u;
        if (slot1 != have) {  // neighbouring pixels of a large triangle share the winner
            r = load_rec(rec + slot);
            e = make_edges(r.x0, r.y0, r.x1, r.y1, r.x2, r.y2);
            have = slot1;
        }
        const int f01 = e.a01 * x + e.b01 * gy + e.c01;
        const int f12 = e.a12 * x + e.b12 * gy + e.c12;
        const int f20 = e.a20 * x + e.b20 * gy + e.c20;
        float al, be, ga;
        barycentric(f01, f12, f20, al, be, ga);
        // zRec of the winner is in the key, bit for bit (phase A computed it with the same
        // operations): no need to redo the three divides of rasterizer.go:153
        const float z = from_orderable((uint32_t)(key >> 32));
        // rasterizer.go:162
        const float intensity = fadd(fadd(fmul(al, r.i0), fmul(be, r.i1)), fmul(ga, r.i2));
        uchar4 c = make_uchar4(200, 200, 200, 255);  // faceColor (renderer.go:17)
        if (r.tex >= 0) {
            const TexDev &t = a.textures[r.tex];
            if (t.type == GRB_TEX_SOLID) {
                c = t.color;
            } else {
                const TriUV uv = a.uv[(size_t)frame * a.recCap + slot];
                // rasterizer.go:132-137, 158-159
                const float u0z0 = fdiv(uv.u0, r.w0), v0z0 = fdiv(uv.v0, r.w0);
                const float u1z1 = fdiv(uv.u1, r.w1), v1z1 = fdiv(uv.v1, r.w1);
                const float u2z2 = fdiv(uv.u2, r.w2), v2z2 = fdiv(uv.v2, r.w2);
                float u = fdiv(fadd(fadd(fmul(al, u0z0), fmul(be, u1z1)), fmul(ga, u2z2)), z);
                float v = fdiv(fadd(fadd(fmul(al, v0z0), fmul(be, v1z1)), fmul(ga, v2z2)), z);
                if (a.options & GRB_OPT_AFFINE_TEXTURES) {   // not a reference code path: include/gorender_b200.h
                    u = -fadd(fadd(fmul(al, uv.u0), fmul(be, uv.u1)), fmul(ga, uv.u2));
                    v = -fadd(fadd(fmul(al, uv.v0), fmul(be, uv.v1)), fmul(ga, uv.v2));
                }
                c = sample_texture(t, u, v);
            }
        }
        // colorIntensity (rasterizer.go:81-88)
        col[k] = make_uchar4(go_u8(fmul((float)c.x, intensity)), go_u8(fmul((float)c.y, intensity)),
                             go_u8(fmul((float)c.z, intensity)), c.w);
        zo[k] = z;
    }
    if constexpr (POST) post_quad(a, frame, gx, gy, col, zo, winner);
    write_quad(a, frame, gx, gy, col, zo, mirC, mirZ);
    }
}

void launch_raster(const DrawArgs &a, int nframes, cudaStream_t s) {
    const int rows = a.tileRowEnd - a.tileRowBegin;
    if (rows <= 0 || a.ntx <= 0) return;
    // the record gathers of phase A live on L1: a larger shared-memory carveout than the blocks need costs time
#ifndef GRB_CARVEOUT
#define GRB_CARVEOUT -1   // percent of the SM's L1/shared storage asked for as shared memory; -1: the driver's choice
#endif
    static const bool carveout = [] {
        if (GRB_CARVEOUT < 0) return false;
        cudaFuncSetAttribute(raster_kernel<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, GRB_CARVEOUT);
        cudaFuncSetAttribute(raster_kernel<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, GRB_CARVEOUT);
        return true;
    }();
    (void)carveout;
    const dim3 grid(a.ntx, rows, nframes);
    const bool post = (a.options & kOptPostPass) != 0;
    const bool mirror = nframes == 1 && (a.mirColor != nullptr || a.mirDepth != nullptr);
    if (mirror) {
        if (post) raster_kernel<true, true><<<grid, kRasterThreads, 0, s>>>(a);
        else raster_kernel<false, true><<<grid, kRasterThreads, 0, s>>>(a);
    } else {
        if (post) raster_kernel<true, false><<<grid, kRasterThreads, 0, s>>>(a);
        else raster_kernel<false, false><<<grid, kRasterThreads, 0, s>>>(a);
    }
}

}  // namespace gr
