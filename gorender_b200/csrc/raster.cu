// raster.cu — K5: the per-tile rasteriser.  Replaces FrameBuffer.Clear /
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
// depth for all triangles of the tile in parallel (small triangles: one
// thread each, shared-memory atomic max; large triangles: the whole block,
// each thread owning four pixels, no atomics).  Phase B shades only the
// winning fragment of each pixel (perspective-correct UV, Gouraud intensity,
// nearest texel through the read-only path), generates the cleared
// background and dot grid for uncovered pixels, and writes colour and depth
// back with 128-bit stores.  Integer edge functions and every float32
// operation follow the reference's order; results are bit-identical to the
// serial CPU path.

#include "gr_types.cuh"
#include "kernels.h"

namespace gr {

constexpr int kRasterThreads = 256;
constexpr int kSmallArea = 32;  // bbox∩tile pixels up to which one thread rasterises a triangle alone
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

__device__ __forceinline__ unsigned long long fragment_key(float zrec, uint32_t seq1) {
    return ((unsigned long long)orderable(zrec) << 32) | seq1;
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

// One thread rasterises one small triangle into the tile (rasterizer.go:140-182,
// visibility part only).
__device__ __forceinline__ void raster_small(const TriRec &r, int x0, int y0, int x1, int y1, int tileX, int tileY,
                                             unsigned long long *keys) {
    const Edges e = make_edges(r.x0, r.y0, r.x1, r.y1, r.x2, r.y2);
    int row01 = e.a01 * x0 + e.b01 * y0 + e.c01;
    int row12 = e.a12 * x0 + e.b12 * y0 + e.c12;
    int row20 = e.a20 * x0 + e.b20 * y0 + e.c20;
    for (int y = y0; y <= y1; y++) {
        int f01 = row01, f12 = row12, f20 = row20;
        for (int x = x0; x <= x1; x++) {
            if ((f01 & f12 & f20) < 0) {  // all three negative
                float al, be, ga;
                barycentric(f01, f12, f20, al, be, ga);
                const float z = z_reciprocal(al, be, ga, r.w0, r.w1, r.w2);
                if (z >= -1.0f) {  // can ever pass `zRec >= ZBuffer` (cleared to -1); false for NaN
                    const unsigned long long key = fragment_key(z, r.seq1);
                    unsigned long long *p = &keys[(y - tileY) * kTile + (x - tileX)];
                    if (key > *(volatile unsigned long long *)p) atomicMax(p, key);
                }
            }
            f01 += e.a01; f12 += e.a12; f20 += e.a20;
        }
        row01 += e.b01; row12 += e.b12; row20 += e.b20;
    }
}

__device__ __forceinline__ TriRec load_rec(const TriRec *p) {
    TriRec r;
    const int4 *s = reinterpret_cast<const int4 *>(p);
    int4 *d = reinterpret_cast<int4 *>(&r);
    d[0] = __ldg(s); d[1] = __ldg(s + 1); d[2] = __ldg(s + 2); d[3] = __ldg(s + 3);
    return r;
}

__global__ void __launch_bounds__(kRasterThreads) raster_kernel(const __grid_constant__ DrawArgs a) {
    __shared__ unsigned long long keys[kTilePix];
    __shared__ uint32_t queue[kRasterThreads];
    __shared__ int queueCount;

    const int frame = blockIdx.z;
    const int tx = blockIdx.x, ty = blockIdx.y + a.tileRowBegin;
    const int tile = ty * a.ntx + tx;
    const int nTiles = a.ntx * a.nty;
    const int tileX = tx * kTile, tileY = ty * kTile;
    const int tileX1 = min(tileX + kTile, a.width) - 1, tileY1 = min(tileY + kTile, a.height) - 1;
    const int tid = threadIdx.x;

    const TriRec *rec = a.rec + (size_t)frame * a.recCap;
    const uint32_t *off = a.tileOff + (size_t)frame * (nTiles + 1);
    const uint32_t listBegin = off[tile], listEnd = off[tile + 1];
    const uint32_t *list = a.binList + (size_t)frame * a.recCap * kMaxBinsPerTri;
    const uint32_t nBig = a.counters[frame].bigCount;
    const uint32_t *big = a.bigList + (size_t)frame * a.recCap;

    // pixels owned by this thread: 4 consecutive in x
    const int px = (tid & 7) * 4, py = tid >> 3;

    for (int i = tid; i < kTilePix; i += kRasterThreads) keys[i] = kBackgroundKey;
    if (tid == 0) queueCount = 0;
    __syncthreads();

    // ------------------------------------------------------------ phase A
    const uint32_t nList = listEnd - listBegin;
    const uint32_t nWork = nList + nBig;
    for (uint32_t base = 0; base < nWork; base += kRasterThreads) {
        const uint32_t i = base + tid;
        if (i < nWork) {
            const uint32_t slot = i < nList ? list[listBegin + i] : big[i - nList];
            const TriRec r = load_rec(rec + slot);
            const int x0 = max((int)r.bx0, tileX), x1 = min((int)r.bx1, tileX1);
            const int y0 = max((int)r.by0, tileY), y1 = min((int)r.by1, tileY1);
            if (x0 <= x1 && y0 <= y1) {
                if ((x1 - x0 + 1) * (y1 - y0 + 1) <= kSmallArea)
                    raster_small(r, x0, y0, x1, y1, tileX, tileY, keys);
                else
                    queue[atomicAdd(&queueCount, 1)] = slot;
            }
        }
        __syncthreads();
        const int nq = queueCount;
        if (nq) {
            // large triangles: the whole block, thread-owned pixels, no atomics
            unsigned long long k0 = keys[py * kTile + px], k1 = keys[py * kTile + px + 1];
            unsigned long long k2 = keys[py * kTile + px + 2], k3 = keys[py * kTile + px + 3];
            const int gx = tileX + px, gy = tileY + py;
            for (int q = 0; q < nq; q++) {
                const TriRec r = load_rec(rec + queue[q]);
                if (gy < r.by0 || gy > r.by1 || gx > r.bx1 || gx + 3 < r.bx0) continue;
                const Edges e = make_edges(r.x0, r.y0, r.x1, r.y1, r.x2, r.y2);
                int f01 = e.a01 * gx + e.b01 * gy + e.c01;
                int f12 = e.a12 * gx + e.b12 * gy + e.c12;
                int f20 = e.a20 * gx + e.b20 * gy + e.c20;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int x = gx + k;
                    if ((f01 & f12 & f20) < 0 && x >= r.bx0 && x <= r.bx1) {
                        float al, be, ga;
                        barycentric(f01, f12, f20, al, be, ga);
                        const float z = z_reciprocal(al, be, ga, r.w0, r.w1, r.w2);
                        if (z >= -1.0f) {
                            const unsigned long long key = fragment_key(z, r.seq1);
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
            __syncthreads();
            if (tid == 0) queueCount = 0;
            __syncthreads();
        }
    }

    // ------------------------------------------------------------ phase B
    const int gx = tileX + px, gy = tileY + py;
    if (gy >= a.height || gx >= a.width) return;

    uchar4 col[4];
    float zo[4];
    const uint32_t *blockBase = a.blockBase + (size_t)frame * a.nFaceBlocks;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int x = gx + k;
        const unsigned long long key = keys[py * kTile + px + k];
        const uint32_t seq1 = (uint32_t)key;
        if (seq1 == 0) {
            // Clear (rasterizer.go:36-44) + DotGrid step 10 (:46-52)
            const bool dot = (x >= 10) && (gy >= 10) && (x % 10 == 0) && (gy % 10 == 0);
            col[k] = dot ? make_uchar4(100, 100, 100, 255) : make_uchar4(50, 50, 50, 255);
            zo[k] = -1.0f;
            continue;
        }
        const uint32_t s = seq1 - 1u;
        const uint32_t slot = blockBase[s / kSeqStride] + (s % kSeqStride);
        const TriRec r = load_rec(rec + slot);
        const Edges e = make_edges(r.x0, r.y0, r.x1, r.y1, r.x2, r.y2);
        const int f01 = e.a01 * x + e.b01 * gy + e.c01;
        const int f12 = e.a12 * x + e.b12 * gy + e.c12;
        const int f20 = e.a20 * x + e.b20 * gy + e.c20;
        float al, be, ga;
        barycentric(f01, f12, f20, al, be, ga);
        const float z = z_reciprocal(al, be, ga, r.w0, r.w1, r.w2);
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
                const float u = fdiv(fadd(fadd(fmul(al, u0z0), fmul(be, u1z1)), fmul(ga, u2z2)), z);
                const float v = fdiv(fadd(fadd(fmul(al, v0z0), fmul(be, v1z1)), fmul(ga, v2z2)), z);
                c = sample_texture(t, u, v);
            }
        }
        // colorIntensity (rasterizer.go:81-88)
        col[k] = make_uchar4(go_u8(fmul((float)c.x, intensity)), go_u8(fmul((float)c.y, intensity)),
                             go_u8(fmul((float)c.z, intensity)), c.w);
        zo[k] = z;
    }

    const size_t pix = ((size_t)frame * a.height + gy) * a.width + gx;
    if ((a.width & 3) == 0) {
        // gx is a multiple of 4 and so is width: 16-byte aligned, whole quad in range
        uint4 cq;
        cq.x = *reinterpret_cast<uint32_t *>(&col[0]); cq.y = *reinterpret_cast<uint32_t *>(&col[1]);
        cq.z = *reinterpret_cast<uint32_t *>(&col[2]); cq.w = *reinterpret_cast<uint32_t *>(&col[3]);
        *reinterpret_cast<uint4 *>(a.color + pix) = cq;
        *reinterpret_cast<float4 *>(a.depth + pix) = make_float4(zo[0], zo[1], zo[2], zo[3]);
    } else {
        for (int k = 0; k < 4 && gx + k < a.width; k++) {
            a.color[pix + k] = col[k];
            a.depth[pix + k] = zo[k];
        }
    }
}

void launch_raster(const DrawArgs &a, int nframes, cudaStream_t s) {
    const int rows = a.tileRowEnd - a.tileRowBegin;
    if (rows <= 0 || a.ntx <= 0) return;
    raster_kernel<<<dim3(a.ntx, rows, nframes), kRasterThreads, 0, s>>>(a);
}

}  // namespace gr
