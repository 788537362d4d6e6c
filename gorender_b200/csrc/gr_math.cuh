// gr_math.cuh — exact float32 arithmetic helpers (host + device).
//
// Parity with the reference depends on every float32 operation being a
// separately rounded IEEE-754 binary32 op in the reference's order (SURVEY.md
// Appendix A; Go on amd64 never fuses a*b+c).  On the device the _rn
// intrinsics are used — the compiler never contracts or reorders them, with
// or without -fmad=false; on the host (BoxVisibility in capi.cu) plain
// operators are used and the file is built with -ffp-contract=off.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define GR_HD __host__ __device__ __forceinline__
#define GR_D __device__ __forceinline__

namespace gr {

GR_HD float fmul(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
GR_HD float fadd(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
GR_HD float fsub(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
GR_HD float fdiv(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
GR_D float fsqrt(float a) { return __fsqrt_rn(a); }  // math32.go:11-13 (SQRTSS)

struct Mat4 {
    float m[16];  // row-major, matrix.go:3
};

// vector.go:111-113: ((x*x' + y*y') + z*z') + w*w'
GR_HD float dot4(float4 a, float4 b) {
    return fadd(fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fmul(a.z, b.z)), fmul(a.w, b.w));
}
// vector.go:74-76
GR_HD float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return fadd(fadd(fmul(ax, bx), fmul(ay, by)), fmul(az, bz));
}
GR_HD float4 sub4(float4 a, float4 b) {
    return make_float4(fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z), fsub(a.w, b.w));
}
GR_HD float4 add4(float4 a, float4 b) {
    return make_float4(fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z), fadd(a.w, b.w));
}
GR_HD float4 mul4(float4 a, float s) {
    return make_float4(fmul(a.x, s), fmul(a.y, s), fmul(a.z, s), fmul(a.w, s));
}

// asm_amd64.s:33-40 == asm_purego.go:13-16 == matrix.go:168-171:
// row r: ((m[r][0]*x + m[r][1]*y) + m[r][2]*z) + m[r][3]*w
GR_HD float mat_row(const float *r, float4 v) {
    return fadd(fadd(fadd(fmul(r[0], v.x), fmul(r[1], v.y)), fmul(r[2], v.z)), fmul(r[3], v.w));
}
GR_HD float4 mat_vec(const float *m, float4 v) {
    return make_float4(mat_row(m, v), mat_row(m + 4, v), mat_row(m + 8, v), mat_row(m + 12, v));
}

#ifdef __CUDACC__
// Packed matrix * vector on Blackwell's two-wide FP32 pipe.  There is no separately rounded packed
// multiply or add in SASS (ptxas contracts mul.f32x2 + add.f32x2 into FFMA2 even with explicit
// .rn and -fmad=false), so each is spelled as an FMA whose result is exactly the rounded product
// or sum:  a*b == fma(a, b, -0)  and  a+b == fma(a, 1, b)  (one rounding each, signs of zero
// included).  The -0 and 1 operands must be RUN-TIME values (FmaConsts, filled by the host): with
// literals the compiler simplifies the FMAs back to mul / add and ptxas then fuses those — a
// one-ulp difference in clip-space w that the parity tests catch.  14 FFMA2 instead of 16 FMUL +
// 12 FADD, bit-identical to mat_vec().
struct FmaConsts {
    float2 negZero, one;     // (-0, -0) and (1, 1)
};
struct Mat4P {
    float2 c01[4], c23[4];   // column k of rows (0,1) and of rows (2,3)
    FmaConsts k;
};
GR_D Mat4P pack_mat(const float *m, const FmaConsts &k) {
    Mat4P p;
    p.k = k;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        p.c01[k] = make_float2(m[k], m[4 + k]);
        p.c23[k] = make_float2(m[8 + k], m[12 + k]);
    }
    return p;
}
GR_D float4 mat_vec(const Mat4P &p, float4 v) {
    auto mul2 = [&](float2 a, float2 b) { return __ffma2_rn(a, b, p.k.negZero); };
    auto add2 = [&](float2 a, float2 b) { return __ffma2_rn(a, p.k.one, b); };
    const float2 x = make_float2(v.x, v.x), y = make_float2(v.y, v.y), z = make_float2(v.z, v.z), w = make_float2(v.w, v.w);
    // row r: ((m[r][0]*x + m[r][1]*y) + m[r][2]*z) + m[r][3]*w  (asm_amd64.s:33-40)
    const float2 r01 = add2(add2(add2(mul2(p.c01[0], x), mul2(p.c01[1], y)), mul2(p.c01[2], z)), mul2(p.c01[3], w));
    const float2 r23 = add2(add2(add2(mul2(p.c23[0], x), mul2(p.c23[1], y)), mul2(p.c23[2], z)), mul2(p.c23[3], w));
    return make_float4(r01.x, r01.y, r23.x, r23.y);
}
#endif

// Go builtin min/max on floats propagate NaN (renderer.go:229-232).
GR_HD float gomin(float a, float b) {
    if (a != a || b != b) return a + b;  // NaN
    return a < b ? a : b;
}
GR_HD float gomax(float a, float b) {
    if (a != a || b != b) return a + b;  // NaN
    return a > b ? a : b;
}

}  // namespace gr
