// gorender_oracle.cpp — CPU oracle for gorender's per-frame hot path.
//
// TEST INFRASTRUCTURE, NOT PRODUCT.  PARITY UNPINNED (see gorender_oracle.h):
// the Go reference has no golden vectors and cannot be built in this image.
//
// A restatement, in C++, of the reference's algorithm with the reference's
// float32 operation order.  Build: g++ -O2 -ffp-contract=off -fno-fast-math
// (no -march=native: baseline x86-64 has no FMA, which is the GOAMD64=v1
// behaviour the parity target is defined on — SURVEY.md H17).
//
// All `file:line` citations are into /root/reference.

#include "gorender_oracle.h"

#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__SSE__)
#include <xmmintrin.h>
#endif

namespace {

// ---------------------------------------------------------------- vector.go

struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };
struct UV { float u, v; };

// vector.go:51-53
inline V3 sub3(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
// vector.go:67-72
inline V3 cross3(V3 a, V3 b) {
    float x = a.y * b.z - a.z * b.y;
    float y = a.z * b.x - a.x * b.z;
    float z = a.x * b.y - a.y * b.x;
    return {x, y, z};
}
// vector.go:74-76 (left-to-right association)
inline float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// math32.go:11-13: float32(math.Sqrt(float64(x))) == correctly rounded sqrtf
inline float sqrt32(float x) { return (float)std::sqrt((double)x); }
// vector.go:63-65, 59-61, 78-80
inline float len3(V3 a) { return sqrt32(a.x * a.x + a.y * a.y + a.z * a.z); }
inline V3 div3(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 norm3(V3 a) { return div3(a, len3(a)); }

// vector.go:95-101
inline V4 add4(V4 a, V4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline V4 sub4(V4 a, V4 b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
// vector.go:103-109
inline V4 div4(V4 a, float s) { return {a.x / s, a.y / s, a.z / s, a.w / s}; }
inline V4 mul4(V4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
// vector.go:111-113
inline float dot4(V4 a, V4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
// vector.go:119-125 — W takes part in the length
inline float len4(V4 a) { return sqrt32(a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w); }
inline V4 norm4(V4 a) { return div4(a, len4(a)); }
inline V3 xyz(V4 a) { return {a.x, a.y, a.z}; }

// Go's builtin min/max on floats: NaN if any operand is NaN, -0 < +0.
inline float gomin(float a, float b) {
    if (a != a || b != b) return NAN;
    if (a == b) return std::signbit(a) ? a : b;
    return a < b ? a : b;
}
inline float gomax(float a, float b) {
    if (a != a || b != b) return NAN;
    if (a == b) return std::signbit(a) ? b : a;
    return a > b ? a : b;
}
inline float gomin3(float a, float b, float c) { return gomin(gomin(a, b), c); }
inline float gomax3(float a, float b, float c) { return gomax(gomax(a, b), c); }

// Go `int(f)` on amd64 is CVTTSS2SQ: truncation, "integer indefinite"
// (INT64_MIN) for NaN and out-of-range inputs.
inline int64_t go_int(float f) {
    if (!(f > -9223372036854775808.0f && f < 9223372036854775808.0f)) return INT64_MIN;
    return (int64_t)f;
}
// Go `uint8(f)` on amd64: CVTTSS2SL then the low byte (rasterizer.go:81-88).
inline uint8_t go_u8(float f) {
    int32_t i;
    if (!(f > -2147483648.0f && f < 2147483648.0f)) i = INT32_MIN;
    else i = (int32_t)f;
    return (uint8_t)(i & 0xff);
}

// ---------------------------------------------------------------- matrix.go

struct M4 { float m[4][4]; };

inline M4 load_m4(const float *p) { M4 r; std::memcpy(r.m, p, 64); return r; }
inline void store_m4(const M4 &a, float *p) { std::memcpy(p, a.m, 64); }

// matrix.go:5-12
M4 identity() {
    M4 r{};
    for (int i = 0; i < 4; i++) r.m[i][i] = 1.0f;
    return r;
}
// matrix.go:155-165: res starts at 0 and accumulates k = 0..3
M4 mat_mul(const M4 &a, const M4 &b) {
    M4 r{};
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            for (int k = 0; k < 4; k++) r.m[i][j] += a.m[i][k] * b.m[k][j];
    return r;
}
// math32.go:15-25
inline float sin32(float x) { return (float)std::sin((double)x); }
inline float cos32(float x) { return (float)std::cos((double)x); }
inline float tan32(float x) { return (float)std::tan((double)x); }

// matrix.go:32-72: rotation by exactly 0 short-circuits to identity
M4 rot_x(float a) {
    if (a == 0) return identity();
    float s = sin32(a), c = cos32(a);
    M4 r = identity();
    r.m[1][1] = c; r.m[1][2] = -s; r.m[2][1] = s; r.m[2][2] = c;
    return r;
}
M4 rot_y(float a) {
    if (a == 0) return identity();
    float s = sin32(a), c = cos32(a);
    M4 r = identity();
    r.m[0][0] = c; r.m[0][2] = s; r.m[2][0] = -s; r.m[2][2] = c;
    return r;
}
M4 rot_z(float a) {
    if (a == 0) return identity();
    float s = sin32(a), c = cos32(a);
    M4 r = identity();
    r.m[0][0] = c; r.m[0][1] = -s; r.m[1][0] = s; r.m[1][1] = c;
    return r;
}
// matrix.go:74-80
M4 rotation(float x, float y, float z) {
    M4 m = identity();
    m = mat_mul(m, rot_x(x));
    m = mat_mul(m, rot_y(y));
    m = mat_mul(m, rot_z(z));
    return m;
}
// matrix.go:82-88: T * (R * (S * I))
M4 world_matrix(V3 s, V3 r, V3 t) {
    M4 m = identity();
    M4 sc = identity();
    sc.m[0][0] = s.x; sc.m[1][1] = s.y; sc.m[2][2] = s.z;
    m = mat_mul(sc, m);
    m = mat_mul(rotation(r.x, r.y, r.z), m);
    M4 tr = identity();
    tr.m[0][3] = t.x; tr.m[1][3] = t.y; tr.m[2][3] = t.z;
    m = mat_mul(tr, m);
    return m;
}
// matrix.go:92-106
M4 perspective(float fov, float aspect, float zn, float zf) {
    float th = tan32(fov / 2.0f);
    float m00 = 1 / (aspect * th);
    float m11 = 1 / th;
    float m22 = (zf + zn) / (zn - zf);
    float m23 = (2 * zf * zn) / (zn - zf);
    M4 r{};
    r.m[0][0] = m00;
    r.m[1][1] = m11;
    r.m[2][2] = -m22; r.m[2][3] = -m23;
    r.m[3][2] = -1;
    return r;
}
// matrix.go:108-118
M4 screen_matrix(int w, int h) {
    float hw = (float)w / 2;
    float hh = (float)h / 2;
    M4 r{};
    r.m[0][0] = hw; r.m[0][3] = hw;
    r.m[1][1] = hh; r.m[1][3] = hh;
    r.m[2][2] = 0.5f; r.m[2][3] = 0.5f;
    r.m[3][3] = 1;
    return r;
}
// matrix.go:133-144
M4 view_matrix(V3 eye, V3 dir, V3 up) {
    V3 z = norm3(dir);
    V3 x = norm3(cross3(up, z));
    V3 y = norm3(cross3(z, x));
    M4 r{};
    r.m[0][0] = x.x; r.m[0][1] = x.y; r.m[0][2] = x.z; r.m[0][3] = -dot3(x, eye);
    r.m[1][0] = y.x; r.m[1][1] = y.y; r.m[1][2] = y.z; r.m[1][3] = -dot3(y, eye);
    r.m[2][0] = z.x; r.m[2][1] = z.y; r.m[2][2] = z.z; r.m[2][3] = -dot3(z, eye);
    r.m[3][3] = 1;
    return r;
}

// matrix.go:167-173 / asm_purego.go:13-16: ((m0*x + m1*y) + m2*z) + m3*w
inline V4 mat_vec(const M4 &m, V4 v) {
    V4 r;
    r.x = m.m[0][0] * v.x + m.m[0][1] * v.y + m.m[0][2] * v.z + m.m[0][3] * v.w;
    r.y = m.m[1][0] * v.x + m.m[1][1] * v.y + m.m[1][2] * v.z + m.m[1][3] * v.w;
    r.z = m.m[2][0] * v.x + m.m[2][1] * v.y + m.m[2][2] * v.z + m.m[2][3] * v.w;
    r.w = m.m[3][0] * v.x + m.m[3][1] * v.y + m.m[3][2] * v.z + m.m[3][3] * v.w;
    return r;
}

// asm_purego.go:9-19
void matvec_batch_scalar(const M4 &m, V4 *v, int64_t n) {
    for (int64_t i = 0; i < n; i++) v[i] = mat_vec(m, v[i]);
}

// asm_amd64.go:8-11 (transpose) + asm_amd64.s:22-47 (broadcast, 4x MULPS,
// ADDPS x1->x0, x2->x0, x3->x0): the same association as the scalar twin.
void matvec_batch_sse(const M4 &m, V4 *v, int64_t n) {
#if defined(__SSE__)
    alignas(16) float t[4][4];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) t[i][j] = m.m[j][i];
    const __m128 c0 = _mm_load_ps(t[0]), c1 = _mm_load_ps(t[1]);
    const __m128 c2 = _mm_load_ps(t[2]), c3 = _mm_load_ps(t[3]);
    float *p = reinterpret_cast<float *>(v);
    for (int64_t i = 0; i < n; i++, p += 4) {
        __m128 x = _mm_mul_ps(c0, _mm_set1_ps(p[0]));
        __m128 y = _mm_mul_ps(c1, _mm_set1_ps(p[1]));
        __m128 z = _mm_mul_ps(c2, _mm_set1_ps(p[2]));
        __m128 w = _mm_mul_ps(c3, _mm_set1_ps(p[3]));
        x = _mm_add_ps(x, y);
        x = _mm_add_ps(x, z);
        x = _mm_add_ps(x, w);
        _mm_storeu_ps(p, x);
    }
#else
    matvec_batch_scalar(m, v, n);
#endif
}

// -------------------------------------------------------------- clipping.go

constexpr int kMaxClipPoints = 9;  // clipping.go:10

struct Plane { V4 point, normal; };

struct Frustum {
    Plane planes[6];
    // clipping.go:93-126, order Left, Right, Top, Bottom, Near, Far
    Frustum(float zn, float zf) {
        planes[0] = {{-1, 0, 0, 1}, {1, 0, 0, 1}};
        planes[1] = {{1, 0, 0, 1}, {-1, 0, 0, 1}};
        planes[2] = {{0, -1, 0, 1}, {0, 1, 0, 1}};
        planes[3] = {{0, 1, 0, 1}, {0, -1, 0, 1}};
        planes[4] = {{0, 0, zn, 1}, {0, 0, -1, 0}};
        planes[5] = {{0, 0, zf, 1}, {0, 0, 1, 0}};
    }
};

// clipping.go:69-71
inline float plane_distance(const Plane &p, V4 v) { return dot4(p.normal, v) - dot4(p.normal, p.point); }
// clipping.go:74-76
inline bool plane_inside(const Plane &p, V4 q) { return dot4(sub4(q, p.point), p.normal) <= 0; }
// clipping.go:79-86
inline V4 plane_intersect(const Plane &p, V4 q0, V4 q1, float *factor) {
    V4 u = sub4(q1, q0);
    V4 w = sub4(q0, p.point);
    float d = dot4(p.normal, u);
    float n = -dot4(p.normal, w);
    float f = n / d;
    *factor = f;
    return add4(q0, mul4(u, f));
}

// clipping.go:131-154 — returns Intersect at the first plane with 1..7 corners
// outside, before later planes could prove Outside.
int box_visibility(const Frustum &f, const V4 bbox[8]) {
    for (int i = 0; i < 6; i++) {
        int outside = 0;
        for (int c = 0; c < 8; c++)
            if (plane_distance(f.planes[i], bbox[c]) > 0) outside++;
        if (outside == 8) return ORC_BOX_OUTSIDE;
        if (outside > 0) return ORC_BOX_INTERSECT;
    }
    return ORC_BOX_INSIDE;
}

struct Polygon {
    float intensity[kMaxClipPoints];
    V4 points[kMaxClipPoints];
    UV uvs[kMaxClipPoints];
    int count = 0;
    // clipping.go:35-40
    void add(V4 v, UV uv, float i) {
        intensity[count] = i;
        points[count] = v;
        uvs[count] = uv;
        count++;
    }
};

// clipping.go:156-165
inline float lerp32(float a, float b, float f) { return a + (b - a) * f; }
inline UV lerp_uv(UV a, UV b, float f) { return {a.u + (b.u - a.u) * f, a.v + (b.v - a.v) * f}; }

// clipping.go:167-236 (+ Triangulate, clipping.go:42-62)
int clip_triangle(const Frustum &f, const V4 pin[3], const UV uvin[3], const float iin[3],
                  V4 pout[][3], UV uvout[][3], float iout[][3]) {
    Polygon a, b;
    Polygon *poly = &a, *poly2 = &b;
    for (int k = 0; k < 3; k++) poly->add(pin[k], uvin[k], iin[k]);

    for (int pi = 0; pi < 6; pi++) {
        const Plane &plane = f.planes[pi];
        poly2->count = 0;
        for (int bi = 0; bi < poly->count; bi++) {
            int ai = (bi + 1) % poly->count;
            UV uvA = poly->uvs[ai], uvB = poly->uvs[bi];
            V4 vA = poly->points[ai], vB = poly->points[bi];
            float iA = poly->intensity[ai], iB = poly->intensity[bi];
            if (plane_inside(plane, vA)) {
                if (!plane_inside(plane, vB)) {
                    float t;
                    V4 x = plane_intersect(plane, vA, vB, &t);
                    float in = lerp32(iA, iB, t);
                    UV uv = lerp_uv(uvA, uvB, t);
                    poly2->add(x, uv, in);
                }
                poly2->add(vA, uvA, iA);
            } else if (plane_inside(plane, vB)) {
                float t;
                V4 x = plane_intersect(plane, vA, vB, &t);
                float in = lerp32(iA, iB, t);
                UV uv = lerp_uv(uvA, uvB, t);
                poly2->add(x, uv, in);
            }
        }
        if (poly2->count == 0) return 0;
        std::swap(poly, poly2);
    }

    if (poly->count < 3) return 0;
    int n = 0;
    for (int i = 0; i < poly->count - 2; i++) {
        iout[n][0] = poly->intensity[0]; iout[n][1] = poly->intensity[i + 1]; iout[n][2] = poly->intensity[i + 2];
        pout[n][0] = poly->points[0]; pout[n][1] = poly->points[i + 1]; pout[n][2] = poly->points[i + 2];
        uvout[n][0] = poly->uvs[0]; uvout[n][1] = poly->uvs[i + 1]; uvout[n][2] = poly->uvs[i + 2];
        n++;
    }
    return n;
}

// --------------------------------------------------------------- texture.go

struct RGBA { uint8_t r, g, b, a; };

// texture.go:69-89
RGBA texture_sample(const orc_texture &t, float u, float v) {
    switch (t.type) {
    case ORC_TEX_SOLID:
        return {t.color[0], t.color[1], t.color[2], t.color[3]};
    case ORC_TEX_IMAGE_FAST: {
        float wf = (float)t.width, hf = (float)t.height;
        int64_t x = go_int((1 - u) * t.scale * wf) & (int64_t)(t.width - 1);
        int64_t y = go_int(v * t.scale * hf) & (int64_t)(t.height - 1);
        const uint8_t *p = t.pixels + 4 * (y * t.width + x);
        return {p[0], p[1], p[2], p[3]};
    }
    case ORC_TEX_IMAGE: {
        float wf = (float)t.width, hf = (float)t.height;
        int64_t x = go_int((1 - u) * t.scale * wf) % (int64_t)t.width;
        int64_t y = go_int(v * t.scale * hf) % (int64_t)t.height;
        int64_t idx = y * t.width + x;
        if (idx < 0) idx = 0;
        const uint8_t *p = t.pixels + 4 * idx;
        return {p[0], p[1], p[2], p[3]};
    }
    default:
        return {255, 0, 255, 255};
    }
}

// ------------------------------------------------------------ rasterizer.go

constexpr RGBA kFaceColor{200, 200, 200, 255};    // renderer.go:17
constexpr RGBA kVertexColor{255, 161, 0, 255};    // renderer.go:18
constexpr RGBA kEdgeColor{0, 0, 0, 255};          // renderer.go:19

struct FrameBuffer {
    int width = 0, height = 0;
    std::vector<float> z;
    std::vector<RGBA> pix;
    int64_t writes = 0;  // diagnostic: number of z-test passes (serial mode only)
    bool affine = false; // ORC_OPT_AFFINE_TEXTURES (not in the reference, see gorender_oracle.h)

    // rasterizer.go:36-44
    void clear(RGBA c) {
        std::fill(z.begin(), z.end(), -1.0f);
        std::fill(pix.begin(), pix.end(), c);
    }
    // rasterizer.go:25-30 (refuses index 0) and :46-52
    void dot_grid(RGBA c, int step) {
        for (int y = step; y < height; y += step)
            for (int x = step; x < width; x += step) {
                int64_t idx = (int64_t)y * width + x;
                if (idx > 0 && idx < (int64_t)pix.size()) pix[idx] = c;
            }
    }
    // rasterizer.go:25-30: bounds are checked on the LINEAR index only, so x outside [0, width)
    // wraps into the neighbouring row, and index 0 is never written
    void pixel(int64_t x, int64_t y, RGBA c) {
        // y * width + x in wrapping int64 arithmetic, like Go
        const int64_t idx = (int64_t)((uint64_t)y * (uint64_t)width + (uint64_t)x);
        if (idx > 0 && idx < (int64_t)pix.size()) pix[idx] = c;
    }
    // rasterizer.go:54-63
    void rect(int64_t x, int64_t y, int64_t w, int64_t h, RGBA c) {
        if (x >= width || y >= height) return;
        for (int64_t py = y; py < y + h; py++)
            for (int64_t px = x; px < x + w; px++) pixel(px, py, c);
    }
    // rasterizer.go:65-79: DDA with float32 steps accumulated by repeated addition
    void line(int64_t x0, int64_t y0, int64_t x1, int64_t y1, RGBA c) {
        const int64_t dx = x1 - x0, dy = y1 - y0;
        const int64_t side = std::max(dx < 0 ? -dx : dx, dy < 0 ? -dy : dy);
        const float xs = (float)dx / (float)side, ys = (float)dy / (float)side;
        float cx = (float)x0, cy = (float)y0;
        for (int64_t i = 0; i <= side; i++) {
            pixel(go_int(cx), go_int(cy), c);
            cx += xs;
            cy += ys;
        }
    }
    // rasterizer.go:209-217
    void cross_hair(RGBA c) {
        const int64_t size = 5, offset = 3;
        const int64_t x = width / 2, y = height / 2;
        line(x - size, y, x - offset, y, c);
        line(x + offset, y, x + size, y, c);
        line(x, y - size, x, y - offset, c);
        line(x, y + offset, x, y + size, c);
    }
    // rasterizer.go:185-191
    static RGBA blend(RGBA a, RGBA b, float f) {
        return {go_u8((float)a.r * (1 - f) + (float)b.r * f), go_u8((float)a.g * (1 - f) + (float)b.g * f),
                go_u8((float)a.b * (1 - f) + (float)b.b * f), go_u8((float)a.a * (1 - f) + (float)b.a * f)};
    }
    // rasterizer.go:193-207
    void fog(float fog_start, float fog_end, RGBA c) {
        for (size_t i = 0; i < pix.size(); i++) {
            const float depth = z[i];
            if (depth >= fog_start) {
                // noop
            } else if (depth <= fog_end) {
                pix[i] = c;
            } else {
                const float f = 1 - ((fog_end - depth) / (fog_end - fog_start));
                pix[i] = blend(pix[i], c, f);
            }
        }
    }
};

// rasterizer.go:81-88
inline RGBA color_intensity(RGBA c, float i) {
    return {go_u8((float)c.r * i), go_u8((float)c.g * i), go_u8((float)c.b * i), c.a};
}

// rasterizer.go:120-125
inline int64_t edge_adjust(int64_t f, int64_t dx, int64_t dy) {
    if (dy > 0 || (dy == 0 && dx > 0)) return f;
    return f - 1;
}

template <typename T> inline T min3(T a, T b, T c) { return std::min(std::min(a, b), c); }
template <typename T> inline T max3(T a, T b, T c) { return std::max(std::max(a, b), c); }

// rasterizer.go:90-183.  Go `int` is 64-bit on amd64.
template <bool kCountWrites>
void fb_triangle(FrameBuffer &fb,
                 int64_t x0, int64_t y0, float z0, float u0, float v0,
                 int64_t x1, int64_t y1, float z1, float u1, float v1,
                 int64_t x2, int64_t y2, float z2, float u2, float v2,
                 int64_t tsx, int64_t tsy, int64_t tex_, int64_t tey,
                 float ia, float ib, float ic, const orc_texture *tex) {
    int64_t minX = min3(x0, x1, x2), maxX = max3(x0, x1, x2);
    int64_t minY = min3(y0, y1, y2), maxY = max3(y0, y1, y2);

    minX = max3<int64_t>(minX, tsx, 0); maxX = min3<int64_t>(maxX, tex_, fb.width - 1);
    minY = max3<int64_t>(minY, tsy, 0); maxY = min3<int64_t>(maxY, tey, fb.height - 1);

    int64_t f01 = (y0 - y1) * minX + (x1 - x0) * minY + (x0 * y1 - x1 * y0);
    int64_t f12 = (y1 - y2) * minX + (x2 - x1) * minY + (x1 * y2 - x2 * y1);
    int64_t f20 = (y2 - y0) * minX + (x0 - x2) * minY + (x2 * y0 - x0 * y2);

    const int64_t f01dx = y0 - y1, f01dy = x1 - x0;
    const int64_t f12dx = y1 - y2, f12dy = x2 - x1;
    const int64_t f20dx = y2 - y0, f20dy = x0 - x2;

    f01 = edge_adjust(f01, f01dx, f01dy);
    f12 = edge_adjust(f12, f12dx, f12dy);
    f20 = edge_adjust(f20, f20dx, f20dy);

    // rasterizer.go:132-137
    const float v0z0 = v0 / z0, u0z0 = u0 / z0;
    const float u1z1 = u1 / z1, v1z1 = v1 / z1;
    const float u2z2 = u2 / z2, v2z2 = v2 / z2;

    for (int64_t y = minY; y <= maxY; y++) {
        int64_t fx01 = f01, fx12 = f12, fx20 = f20;
        for (int64_t x = minX; x <= maxX; x++) {
            if (fx01 < 0 && fx12 < 0 && fx20 < 0) {
                float alpha = (float)fx12 / (float)(fx12 + fx20 + fx01);
                float beta = (float)fx20 / (float)(fx12 + fx20 + fx01);
                float gamma = 1 - alpha - beta;

                float zrec = -(alpha / z0 + beta / z1 + gamma / z2);
                int64_t index = y * fb.width + x;

                if (zrec >= fb.z[index]) {
                    float u = (alpha * u0z0 + beta * u1z1 + gamma * u2z2) / zrec;
                    float v = (alpha * v0z0 + beta * v1z1 + gamma * v2z2) / zrec;
                    if (fb.affine) {   // screen-space linear interpolation; the sign keeps Texture.Sample's convention (H6)
                        u = -(alpha * u0 + beta * u1 + gamma * u2);
                        v = -(alpha * v0 + beta * v1 + gamma * v2);
                    }
                    float intensity = alpha * ia + beta * ib + gamma * ic;
                    RGBA c = kFaceColor;
                    if (tex != nullptr) c = texture_sample(*tex, u, v);
                    fb.z[index] = zrec;
                    fb.pix[index] = color_intensity(c, intensity);
                    if (kCountWrites) fb.writes++;
                }
            }
            fx01 += f01dx; fx12 += f12dx; fx20 += f20dx;
        }
        f01 += f01dy; f12 += f12dy; f20 += f20dy;
    }
}

// -------------------------------------------------------------- renderer.go

constexpr int kMaxTiles = 16;            // renderer.go:11
constexpr int kLocalBuf = 128;           // renderer.go:79
constexpr float kDiffuse = 0.5f;         // renderer.go:12
constexpr float kAmbient = 0.5f;         // renderer.go:13

struct Triangle {                        // renderer.go:29-34
    V4 points[3];
    UV uvs[3];
    float intensity[3];
    int32_t tex;
    int32_t object, face, fan;           // provenance (oracle only)
};

struct TileBounds { float sx, sy, ex, ey; };

// renderer.go:50-76
TileBounds tile_boundaries(unsigned tile, unsigned n, int width, int height) {
    if (n == 1) return {0, 0, (float)width, (float)height};
    unsigned ntx = (unsigned)std::sqrt((double)n);
    unsigned nty = (n + ntx - 1) / ntx;
    unsigned tw = ((unsigned)width + ntx - 1) / ntx;
    unsigned th = ((unsigned)height + nty - 1) / nty;
    TileBounds b;
    b.sx = (float)((tile % ntx) * tw);
    b.sy = (float)((tile / ntx) * th);
    b.ex = b.sx + (float)tw;
    b.ey = b.sy + (float)th;
    if (b.ex > (float)width) b.ex = (float)width;
    if (b.ey > (float)height) b.ey = (float)height;
    return b;
}

// Minimal fixed worker pool standing in for the goroutine workers
// (renderer.go:145-156, 409-420).
class Pool {
public:
    explicit Pool(int n) {
        for (int i = 0; i < n; i++) threads_.emplace_back([this] { run(); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> g(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    // run fn(0..count-1) on the pool and wait (wg.Add / wg.Wait)
    void parallel(int count, const std::function<void(int)> &fn) {
        {
            std::lock_guard<std::mutex> g(mu_);
            fn_ = &fn;
            next_ = 0;
            total_ = count;
            pending_ = count;
        }
        cv_.notify_all();
        std::unique_lock<std::mutex> g(mu_);
        done_.wait(g, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

private:
    void run() {
        std::unique_lock<std::mutex> g(mu_);
        for (;;) {
            cv_.wait(g, [this] { return stop_ || (fn_ && next_ < total_); });
            if (stop_) return;
            int i = next_++;
            const std::function<void(int)> *fn = fn_;
            g.unlock();
            (*fn)(i);
            g.lock();
            if (--pending_ == 0) done_.notify_all();
        }
    }
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    const std::function<void(int)> *fn_ = nullptr;
    int next_ = 0, total_ = 0, pending_ = 0;
    bool stop_ = false;
};

}  // namespace

struct orc_renderer {
    FrameBuffer fb;
    Frustum frustum{0.0f, 50.0f};         // renderer.go:121-122
    unsigned num_tiles = 16;
    TileBounds bounds[kMaxTiles];
    std::vector<Triangle> tile_tris[kMaxTiles];
    std::mutex tile_locks[kMaxTiles];
    int64_t tpf = 0;
    bool record = false;
    std::vector<orc_triangle> recorded;
    std::vector<int32_t> visibility;
    std::unique_ptr<Pool> pool;
    // per-object scratch == Object.TransformedVertices etc. (mesh.go:76-78)
    struct Local { Triangle tris[kMaxTiles][kLocalBuf]; int count[kMaxTiles]; };
    struct Scratch { std::vector<V4> tv, wfn, wvn; std::unique_ptr<Local> local; };
    std::vector<Scratch> scratch;

    // frame inputs
    const orc_mesh *meshes = nullptr;
    const orc_texture *textures = nullptr;
    int ntex = 0;
    M4 screen;
    V3 light;
    uint32_t options = 0;
    // Fog arguments (the call is commented out at renderer.go:479 with 0.100, 0.033, {100,100,100,255})
    float fog_start = 0.100f, fog_end = 0.033f;
    RGBA fog_color{100, 100, 100, 255};
};

namespace {

// renderer.go:226-244
int identify_tiles(const orc_renderer &r, const V4 p[3], uint8_t out[kMaxTiles]) {
    float minX = gomin3(p[0].x, p[1].x, p[2].x), maxX = gomax3(p[0].x, p[1].x, p[2].x);
    float minY = gomin3(p[0].y, p[1].y, p[2].y), maxY = gomax3(p[0].y, p[1].y, p[2].y);
    int n = 0;
    for (unsigned i = 0; i < r.num_tiles; i++) {
        const TileBounds &b = r.bounds[i];
        if (maxX >= b.sx && minX <= b.ex && maxY >= b.sy && minY <= b.ey) out[n++] = (uint8_t)i;
    }
    return n;
}

// renderer.go:246-250
inline bool facing_camera(const V4 p[3]) {
    V3 v0 = xyz(p[0]), v1 = xyz(p[1]), v2 = xyz(p[2]);
    V3 n = cross3(sub3(v1, v0), sub3(v2, v0));
    return dot3(n, sub3(V3{0, 0, 0}, v0)) > 0;
}

// renderer.go:254-407
void project_object(orc_renderer &r, int oi, const orc_object &obj, bool locked) {
    const orc_mesh &mesh = r.meshes[obj.mesh];
    const M4 world = load_m4(obj.world);
    const M4 mvp = load_m4(obj.mvp);
    const M4 &screen = r.screen;
    const V3 light = r.light;

    // :268-275
    V4 bbox[8];
    std::memcpy(bbox, mesh.bbox, sizeof(bbox));
    matvec_batch_sse(mvp, bbox, 8);
    int vis = box_visibility(r.frustum, bbox);
    r.visibility[oi] = vis;
    if (vis == ORC_BOX_OUTSIDE) return;

    // local tile buffers (:291-300); kept per object like the pooled LocalBuffer
    orc_renderer::Scratch &s = r.scratch[oi];
    if (!s.local) s.local.reset(new orc_renderer::Local);
    orc_renderer::Local *local = s.local.get();
    for (int i = 0; i < kMaxTiles; i++) local->count[i] = 0;

    // :303-304
    s.tv.resize(mesh.nv);
    std::memcpy(s.tv.data(), mesh.vertices, (size_t)mesh.nv * 16);
    matvec_batch_sse(mvp, s.tv.data(), mesh.nv);
    // :307-310
    s.wfn.resize(mesh.nf);
    std::memcpy(s.wfn.data(), mesh.fnormals, (size_t)mesh.nf * 16);
    s.wvn.resize(mesh.nvn);
    if (mesh.nvn) std::memcpy(s.wvn.data(), mesh.vnormals, (size_t)mesh.nvn * 16);
    matvec_batch_sse(world, s.wfn.data(), mesh.nf);
    matvec_batch_sse(world, s.wvn.data(), mesh.nvn);

    const bool has_vn = mesh.nvn != 0;  // :313
    const bool opt_cull = r.options & ORC_OPT_BACKFACE_CULLING;
    const bool opt_light = r.options & ORC_OPT_LIGHTING;
    const bool opt_flat = r.options & ORC_OPT_FLAT_SHADING;
    const bool opt_clip = r.options & ORC_OPT_FRUSTUM_CLIPPING;

    uint8_t tile_nums[kMaxTiles];
    V4 verts[3];
    float vint[3];
    V4 cverts[kMaxClipPoints][3];
    float cint[kMaxClipPoints][3];
    UV cuv[kMaxClipPoints][3];

    auto flush = [&](int tile, int count) {
        if (locked) r.tile_locks[tile].lock();
        r.tile_tris[tile].insert(r.tile_tris[tile].end(), local->tris[tile], local->tris[tile] + count);
        if (locked) r.tile_locks[tile].unlock();
    };

    for (int fi = 0; fi < mesh.nf; fi++) {
        const int32_t *vi = mesh.vidx + 3 * fi;
        verts[0] = s.tv[vi[0]];
        verts[1] = s.tv[vi[1]];
        verts[2] = s.tv[vi[2]];

        if (opt_cull && !facing_camera(verts)) continue;  // :322

        if (opt_light) {  // :326-346
            if (has_vn && !opt_flat) {
                const int32_t *ni = mesh.nidx + 3 * fi;
                for (int k = 0; k < 3; k++) {
                    V3 vn = xyz(norm4(s.wvn[ni[k]]));
                    vint[k] = kAmbient + dot3(vn, light) * kDiffuse;
                }
            } else {
                V3 fn = xyz(norm4(s.wfn[fi]));
                float diffuse = dot3(fn, light) * kDiffuse;
                float in = kAmbient + diffuse;
                vint[0] = vint[1] = vint[2] = in;
            }
        } else {
            vint[0] = vint[1] = vint[2] = kAmbient;
        }

        UV fuv[3];
        std::memcpy(fuv, mesh.uvs + 6 * fi, sizeof(fuv));

        int clip_count;
        if (opt_clip && vis != ORC_BOX_INSIDE) {  // :349-359
            clip_count = clip_triangle(r.frustum, verts, fuv, vint, cverts, cuv, cint);
        } else {
            for (int k = 0; k < 3; k++) { cint[0][k] = vint[k]; cverts[0][k] = verts[k]; cuv[0][k] = fuv[k]; }
            clip_count = 1;
        }

        for (int i = 0; i < clip_count; i++) {
            Triangle t;
            for (int j = 0; j < 3; j++) {  // :365-370
                V4 p = cverts[i][j];
                float ow = p.w;
                p = div4(p, p.w);
                p = mat_vec(screen, p);
                p.w = ow;
                t.points[j] = p;
                t.uvs[j] = cuv[i][j];
                t.intensity[j] = cint[i][j];
            }
            t.tex = mesh.tex ? mesh.tex[fi] : -1;
            t.object = oi; t.face = fi; t.fan = i;

            if (r.record) {
                orc_triangle rec;
                std::memcpy(rec.points, t.points, sizeof(rec.points));
                std::memcpy(rec.uvs, t.uvs, sizeof(rec.uvs));
                std::memcpy(rec.intensity, t.intensity, sizeof(rec.intensity));
                rec.tex = t.tex; rec.object = oi; rec.face = fi; rec.fan = i;
                r.recorded.push_back(rec);
            }

            int n = identify_tiles(r, t.points, tile_nums);  // :380
            for (int k = 0; k < n; k++) {
                int tile = tile_nums[k];
                local->tris[tile][local->count[tile]++] = t;
                if (local->count[tile] == kLocalBuf) {  // :388-393
                    flush(tile, kLocalBuf);
                    local->count[tile] = 0;
                }
            }
        }
    }

    for (int tile = 0; tile < kMaxTiles; tile++)  // :399-406
        if (local->count[tile] != 0) {
            flush(tile, local->count[tile]);
            local->count[tile] = 0;
        }
}

// drawProjection (renderer.go:166-217) over the tile's list (renderTile, :219-223).  The overlays
// (ShowEdges / ShowVertices) are drawn by every tile pass that lists the triangle, unclipped by the
// tile and without a depth test, so the image depends on the serial order tile 0..15, list order
// inside a tile, face -> edges -> centre mark -> vertex marks inside a triangle.
template <bool kCount>
void render_tile(orc_renderer &r, unsigned tile) {
    const TileBounds &b = r.bounds[tile];
    const bool show_tex = r.options & ORC_OPT_SHOW_TEXTURES;
    const bool show_faces = r.options & ORC_OPT_SHOW_FACES;
    const bool show_edges = r.options & ORC_OPT_SHOW_EDGES;
    const bool show_verts = r.options & ORC_OPT_SHOW_VERTICES;
    if (!show_faces && !show_edges && !show_verts) return;
    for (const Triangle &t : r.tile_tris[tile]) {
        const V4 &a = t.points[0], &bb = t.points[1], &c = t.points[2];
        if (show_faces) {  // :180-189
            const orc_texture *tex = nullptr;
            if (show_tex && t.tex >= 0 && t.tex < r.ntex) tex = &r.textures[t.tex];
            fb_triangle<kCount>(r.fb,
                                go_int(a.x), go_int(a.y), a.w, t.uvs[0].u, t.uvs[0].v,
                                go_int(bb.x), go_int(bb.y), bb.w, t.uvs[1].u, t.uvs[1].v,
                                go_int(c.x), go_int(c.y), c.w, t.uvs[2].u, t.uvs[2].v,
                                go_int(b.sx), go_int(b.sy), go_int(b.ex), go_int(b.ey),
                                t.intensity[0], t.intensity[1], t.intensity[2], tex);
        }
        if (show_edges) {  // :191-210
            RGBA colr = kEdgeColor;
            if (!show_faces) colr = {255, 255, 255, 255};
            r.fb.line(go_int(a.x), go_int(a.y), go_int(bb.x), go_int(bb.y), colr);
            r.fb.line(go_int(bb.x), go_int(bb.y), go_int(c.x), go_int(c.y), colr);
            r.fb.line(go_int(c.x), go_int(c.y), go_int(a.x), go_int(a.y), colr);
            if (show_faces) {
                const float cx = (a.x + bb.x + c.x) / 3, cy = (a.y + bb.y + c.y) / 3;
                r.fb.rect(go_int(cx) - 1, go_int(cy) - 1, 3, 3, colr);
            }
        }
        if (show_verts) {  // :212-216
            r.fb.rect(go_int(a.x) - 1, go_int(a.y) - 1, 3, 3, kVertexColor);
            r.fb.rect(go_int(bb.x) - 1, go_int(bb.y) - 1, 3, 3, kVertexColor);
            r.fb.rect(go_int(c.x) - 1, go_int(c.y) - 1, 3, 3, kVertexColor);
        }
    }
}

}  // namespace

extern "C" {

orc_renderer *orc_renderer_create(int32_t width, int32_t height, int32_t num_tiles, int32_t threads) {
    if (width <= 0 || height <= 0 || num_tiles < 1 || num_tiles > kMaxTiles) return nullptr;
    orc_renderer *r = new orc_renderer;
    r->fb.width = width;
    r->fb.height = height;
    r->fb.z.assign((size_t)width * height, 0.0f);
    r->fb.pix.assign((size_t)width * height, RGBA{0, 0, 0, 0});
    r->num_tiles = (unsigned)num_tiles;
    for (unsigned i = 0; i < r->num_tiles; i++) r->bounds[i] = tile_boundaries(i, r->num_tiles, width, height);
    if (threads > 0) r->pool.reset(new Pool(threads));
    return r;
}

void orc_renderer_destroy(orc_renderer *r) { delete r; }

void orc_renderer_record_triangles(orc_renderer *r, int32_t enable) { r->record = enable != 0; }

void orc_renderer_set_fog(orc_renderer *r, float fog_start, float fog_end, const uint8_t color[4]) {
    r->fog_start = fog_start;
    r->fog_end = fog_end;
    r->fog_color = {color[0], color[1], color[2], color[3]};
}

// renderer.go:443-483
int32_t orc_renderer_draw(orc_renderer *r, const orc_mesh *meshes, int32_t nmesh,
                          const orc_texture *textures, int32_t ntex,
                          const orc_object *objects, int32_t nobj,
                          const float screen[16], const float light[3], uint32_t options) {
    if (!r || nobj < 0) return -1;
    for (int i = 0; i < nobj; i++)
        if (objects[i].mesh < 0 || objects[i].mesh >= nmesh) return -1;
    r->meshes = meshes;
    r->textures = textures;
    r->ntex = ntex;
    r->screen = load_m4(screen);
    r->light = {light[0], light[1], light[2]};
    r->options = options;
    r->fb.affine = (options & ORC_OPT_AFFINE_TEXTURES) != 0;
    r->recorded.clear();
    r->visibility.assign(nobj, ORC_BOX_OUTSIDE);
    if ((int)r->scratch.size() < nobj) r->scratch.resize(nobj);

    for (unsigned i = 0; i < r->num_tiles; i++) r->tile_tris[i].clear();  // :444-446 (capacity kept)

    r->fb.writes = 0;
    r->fb.clear({50, 50, 50, 255});         // :448
    r->fb.dot_grid({100, 100, 100, 255}, 10);  // :449

    if (r->pool) {
        // :452-465 — one projection task per object, barrier, one raster task per tile
        std::function<void(int)> proj = [&](int i) { project_object(*r, i, objects[i], true); };
        r->pool->parallel(nobj, proj);
        std::function<void(int)> rast = [&](int t) { render_tile<false>(*r, (unsigned)t); };
        r->pool->parallel((int)r->num_tiles, rast);
    } else {
        // :467-473 — the canonical (deterministic) order
        for (int i = 0; i < nobj; i++) project_object(*r, i, objects[i], false);
        for (unsigned t = 0; t < r->num_tiles; t++) render_tile<true>(*r, t);
    }

    // :476-480 — `if !demoMode { CrossHair; // Fog }`: demoMode is a constant true in main.go:22
    // and the Fog call is commented out, so both are option bits here
    if (options & ORC_OPT_CROSSHAIR) r->fb.cross_hair({255, 255, 0, 255});
    if (options & ORC_OPT_FOG) r->fb.fog(r->fog_start, r->fog_end, r->fog_color);

    // :436-441
    r->tpf = 0;
    for (unsigned i = 0; i < r->num_tiles; i++) r->tpf += (int64_t)r->tile_tris[i].size();
    return 0;
}

// Timed CPU baseline: nframes consecutive Draw calls, frame f using
// objects[f*nobj .. f*nobj+nobj).  Returns wall-clock seconds spent inside the
// Draw calls (steady clock), or a negative value on bad arguments.
double orc_renderer_draw_sequence(orc_renderer *r, const orc_mesh *meshes, int32_t nmesh,
                                  const orc_texture *textures, int32_t ntex,
                                  const orc_object *objects, int32_t nobj, int32_t nframes,
                                  const float screen[16], const float light[3], uint32_t options) {
    const auto t0 = std::chrono::steady_clock::now();
    for (int32_t f = 0; f < nframes; f++)
        if (orc_renderer_draw(r, meshes, nmesh, textures, ntex, objects + (size_t)f * nobj, nobj, screen, light, options))
            return -1.0;
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

const uint8_t *orc_renderer_pixels(const orc_renderer *r) { return reinterpret_cast<const uint8_t *>(r->fb.pix.data()); }
const float *orc_renderer_zbuffer(const orc_renderer *r) { return r->fb.z.data(); }
int64_t orc_renderer_tpf(const orc_renderer *r) { return r->tpf; }
int64_t orc_renderer_pixel_writes(const orc_renderer *r) { return r->fb.writes; }
int64_t orc_renderer_num_triangles(const orc_renderer *r) { return (int64_t)r->recorded.size(); }
const orc_triangle *orc_renderer_triangles(const orc_renderer *r) { return r->recorded.data(); }
int32_t orc_renderer_visibility(const orc_renderer *r, int32_t *out, int32_t cap) {
    int32_t n = std::min<int32_t>(cap, (int32_t)r->visibility.size());
    for (int32_t i = 0; i < n; i++) out[i] = r->visibility[i];
    return n;
}

void orc_matvec4_batch_scalar(const float m[16], float *vecs, int64_t n) {
    matvec_batch_scalar(load_m4(m), reinterpret_cast<V4 *>(vecs), n);
}
void orc_matvec4_batch_sse(const float m[16], float *vecs, int64_t n) {
    matvec_batch_sse(load_m4(m), reinterpret_cast<V4 *>(vecs), n);
}

int32_t orc_box_visibility(const float bbox_clip[32], float z_near, float z_far) {
    Frustum f(z_near, z_far);
    V4 b[8];
    std::memcpy(b, bbox_clip, sizeof(b));
    return box_visibility(f, b);
}

int32_t orc_clip_triangle(const float pts[12], const float uvs[6], const float intens[3],
                          float z_near, float z_far, float *pts_out, float *uvs_out, float *intens_out) {
    Frustum f(z_near, z_far);
    V4 pin[3];
    UV uvin[3];
    std::memcpy(pin, pts, sizeof(pin));
    std::memcpy(uvin, uvs, sizeof(uvin));
    V4 pout[kMaxClipPoints][3];
    UV uvout[kMaxClipPoints][3];
    float iout[kMaxClipPoints][3];
    int n = clip_triangle(f, pin, uvin, intens, pout, uvout, iout);
    std::memcpy(pts_out, pout, (size_t)n * 48);
    std::memcpy(uvs_out, uvout, (size_t)n * 24);
    std::memcpy(intens_out, iout, (size_t)n * 12);
    return n;
}

void orc_world_matrix(const float s[3], const float rot[3], const float t[3], float out[16]) {
    store_m4(world_matrix({s[0], s[1], s[2]}, {rot[0], rot[1], rot[2]}, {t[0], t[1], t[2]}), out);
}
void orc_view_matrix(const float eye[3], const float dir[3], const float up[3], float out[16]) {
    store_m4(view_matrix({eye[0], eye[1], eye[2]}, {dir[0], dir[1], dir[2]}, {up[0], up[1], up[2]}), out);
}
void orc_perspective_matrix(float fov, float aspect, float zn, float zf, float out[16]) {
    store_m4(perspective(fov, aspect, zn, zf), out);
}
void orc_screen_matrix(int32_t w, int32_t h, float out[16]) { store_m4(screen_matrix(w, h), out); }
void orc_matrix_multiply(const float a[16], const float b[16], float out[16]) {
    store_m4(mat_mul(load_m4(a), load_m4(b)), out);
}
void orc_mvp_matrix(const float p[16], const float v[16], const float w[16], float out[16]) {
    M4 m = identity();
    m = mat_mul(m, load_m4(p));
    m = mat_mul(m, load_m4(v));
    m = mat_mul(m, load_m4(w));
    store_m4(m, out);
}
void orc_light_direction(float out[3]) {
    V3 l = norm3({-1, 1, 1});
    out[0] = l.x; out[1] = l.y; out[2] = l.z;
}

// mesh.go:53-62
void orc_face_normals(const float *vertices, const int32_t *vidx, int32_t nf, float *out) {
    const V4 *v = reinterpret_cast<const V4 *>(vertices);
    for (int32_t i = 0; i < nf; i++) {
        V3 v0 = xyz(v[vidx[3 * i]]), v1 = xyz(v[vidx[3 * i + 1]]), v2 = xyz(v[vidx[3 * i + 2]]);
        V3 n = norm3(cross3(sub3(v1, v0), sub3(v2, v0)));
        out[4 * i] = n.x; out[4 * i + 1] = n.y; out[4 * i + 2] = n.z; out[4 * i + 3] = 1.0f;
    }
}

// mesh.go:28-51
void orc_bounding_box(const float *vertices, int32_t nv, float out[32]) {
    const V4 *v = reinterpret_cast<const V4 *>(vertices);
    float minX = v[0].x, minY = v[0].y, minZ = v[0].z;
    float maxX = minX, maxY = minY, maxZ = minZ;
    for (int32_t i = 0; i < nv; i++) {
        minX = gomin(minX, v[i].x); minY = gomin(minY, v[i].y); minZ = gomin(minZ, v[i].z);
        maxX = gomax(maxX, v[i].x); maxY = gomax(maxY, v[i].y); maxZ = gomax(maxZ, v[i].z);
    }
    const float c[8][4] = {
        {minX, minY, minZ, 1}, {minX, minY, maxZ, 1}, {minX, maxY, minZ, 1}, {minX, maxY, maxZ, 1},
        {maxX, minY, minZ, 1}, {maxX, minY, maxZ, 1}, {maxX, maxY, minZ, 1}, {maxX, maxY, maxZ, 1},
    };
    std::memcpy(out, c, sizeof(c));
}

void orc_texture_sample(const orc_texture *t, float u, float v, uint8_t out[4]) {
    RGBA c = texture_sample(*t, u, v);
    out[0] = c.r; out[1] = c.g; out[2] = c.b; out[3] = c.a;
}

}  // extern "C"
