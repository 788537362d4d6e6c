// mesh.cu — upload-time kernels: what NewMesh (mesh.go:53-69) and the flattening of `[]Face`
// (mesh.go:12-17) do once per mesh, done on the device so that a 2 M-face scene does not spend its
// first frame in host loops.
//
//   mesh_prepare_kernel  one thread per face: validates the face's indices (the reference panics
//                        on an out-of-range index, renderer.go:318-320, 328-330), builds the
//                        face-corner expansion cv[k][f] / cn[k][f] the frame kernels stream
//                        (gr_types.cuh, MeshDev), and — for grb_mesh_new — the face normal
//                        normalize((v1-v0) x (v2-v0)).ToVec4() of mesh.go:54-60, in the reference's
//                        float32 operation order (vector.go:51-80).
//   bbox_kernel          boundingBox (mesh.go:28-51): per-axis min / max over the vertices with
//                        Go's builtin min/max semantics (-0 < +0, NaN wins), as a block reduction
//                        followed by integer atomics on order-preserving keys.

#include "gr_types.cuh"
#include "kernels.h"

namespace gr {

namespace {

// float -> uint32 whose unsigned order is the float order with -0 < +0 (Go's min/max rule).
__device__ __forceinline__ uint32_t order_key(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

}  // namespace

__global__ void __launch_bounds__(256) mesh_prepare_kernel(MeshPrepArgs p) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= p.nf) return;
    const int i0 = p.vidx[3 * f], i1 = p.vidx[3 * f + 1], i2 = p.vidx[3 * f + 2];
    const bool vok = i0 >= 0 && i0 < p.nv && i1 >= 0 && i1 < p.nv && i2 >= 0 && i2 < p.nv;
    if (!vok) atomicOr(p.error, 1);
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 v0 = vok ? p.verts[i0] : zero, v1 = vok ? p.verts[i1] : zero, v2 = vok ? p.verts[i2] : zero;
    p.cv[0][f] = v0;
    p.cv[1][f] = v1;
    p.cv[2][f] = v2;
    if (p.nvn > 0) {
        const int n0 = p.nidx[3 * f], n1 = p.nidx[3 * f + 1], n2 = p.nidx[3 * f + 2];
        const bool nok = n0 >= 0 && n0 < p.nvn && n1 >= 0 && n1 < p.nvn && n2 >= 0 && n2 < p.nvn;
        if (!nok) atomicOr(p.error, 2);
        p.cn[0][f] = nok ? p.vnormals[n0] : zero;
        p.cn[1][f] = nok ? p.vnormals[n1] : zero;
        p.cn[2][f] = nok ? p.vnormals[n2] : zero;
    }
    if (p.fnormalsOut != nullptr) {
        // mesh.go:55-59; Sub (vector.go:51-53), CrossProduct (:67-72), Length (:63-65), Divide (:59-61)
        const float ax = fsub(v1.x, v0.x), ay = fsub(v1.y, v0.y), az = fsub(v1.z, v0.z);
        const float bx = fsub(v2.x, v0.x), by = fsub(v2.y, v0.y), bz = fsub(v2.z, v0.z);
        const float x = fsub(fmul(ay, bz), fmul(az, by));
        const float y = fsub(fmul(az, bx), fmul(ax, bz));
        const float z = fsub(fmul(ax, by), fmul(ay, bx));
        const float len = fsqrt(fadd(fadd(fmul(x, x), fmul(y, y)), fmul(z, z)));
        p.fnormalsOut[f] = make_float4(fdiv(x, len), fdiv(y, len), fdiv(z, len), 1.0f);
    }
}

// out[0..2] = keys of min x,y,z (initialised to 0xffffffff), out[3..5] = keys of max (0),
// out[6] = per-axis NaN bits.
__global__ void __launch_bounds__(256) bbox_kernel(const float4 *verts, int nv, uint32_t *out) {
    uint32_t mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u}, nan = 0u;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += gridDim.x * blockDim.x) {
        const float4 v = __ldg(&verts[i]);
        const float c[3] = {v.x, v.y, v.z};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (c[k] != c[k]) { nan |= 1u << k; continue; }
            const uint32_t key = order_key(c[k]);
            mn[k] = min(mn[k], key);
            mx[k] = max(mx[k], key);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        mn[k] = __reduce_min_sync(0xffffffffu, mn[k]);
        mx[k] = __reduce_max_sync(0xffffffffu, mx[k]);
    }
    nan = __reduce_or_sync(0xffffffffu, nan);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            atomicMin(&out[k], mn[k]);
            atomicMax(&out[3 + k], mx[k]);
        }
        if (nan) atomicOr(&out[6], nan);
    }
}

// Object-space bounds of every 32 consecutive faces (what one warp of a setup block works on): lets a strip draw
// skip the warps whose faces cannot reach its rows (setup.cu, reject_kernel).  A run of 32 faces of a mesh in any
// reasonable order is compact where 256 may straddle half the object.  NaN coordinates poison the bounds, which
// disables the skip for that warp.
__global__ void __launch_bounds__(256) warp_bounds_kernel(const float4 *cv0, const float4 *cv1, const float4 *cv2, int nf,
                                                          float4 *warpLo, float4 *warpHi) {
    const int f = blockIdx.x * 256 + threadIdx.x;
    const float inf = __int_as_float(0x7f800000);
    float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
    bool nan = false;
    if (f < nf) {
        const float4 c[3] = {__ldg(&cv0[f]), __ldg(&cv1[f]), __ldg(&cv2[f])};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float v[3] = {c[k].x, c[k].y, c[k].z};
#pragma unroll
            for (int d = 0; d < 3; d++) {
                nan |= v[d] != v[d];
                lo[d] = fminf(lo[d], v[d]);
                hi[d] = fmaxf(hi[d], v[d]);
            }
        }
    }
    nan = __any_sync(0xffffffffu, nan);
#pragma unroll
    for (int d = 0; d < 3; d++)
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
    if ((threadIdx.x & 31) == 0 && blockIdx.x * 256 + (int)(threadIdx.x & ~31u) < nf) {
        const float q = __int_as_float(0x7fc00000);
        const int w = f >> 5;
        warpLo[w] = nan ? make_float4(q, q, q, 1.f) : make_float4(lo[0], lo[1], lo[2], 1.f);
        warpHi[w] = nan ? make_float4(q, q, q, 1.f) : make_float4(hi[0], hi[1], hi[2], 1.f);
    }
}

void launch_warp_bounds(const float4 *const cv[3], int nf, float4 *warpLo, float4 *warpHi, cudaStream_t s) {
    if (nf <= 0) return;
    warp_bounds_kernel<<<(nf + 255) / 256, 256, 0, s>>>(cv[0], cv[1], cv[2], nf, warpLo, warpHi);
}

void launch_mesh_prepare(const MeshPrepArgs &p, cudaStream_t s) {
    if (p.nf <= 0) return;
    mesh_prepare_kernel<<<(p.nf + 255) / 256, 256, 0, s>>>(p);
}

void launch_bbox(const float4 *verts, int nv, uint32_t *out7, cudaStream_t s) {
    if (nv <= 0) return;
    int blocks = (nv + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    bbox_kernel<<<blocks, 256, 0, s>>>(verts, nv, out7);
}

}  // namespace gr
