// svgf_device.cuh — device-side building blocks shared by the SVGF kernels: storage codecs, the compact
// guide texel, and the edge-stopping weight in its exp2/log2 form.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace svgf {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kBackgroundZ = 1e30f;  // GetDepth's sentinel, reference src/Filter.cuh:204

// ---- storage codecs: colour+variance plane (half4 | float4), moments plane (half2 | float2) ----------------
template <bool F32> struct ColourPlane;
template <> struct ColourPlane<false> {
    using texel = uint2;  // 4 x fp16
    static constexpr int kBytes = 8;
    static __device__ __forceinline__ float4 decode(uint2 t) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&t.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&t.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
    static __device__ __forceinline__ uint2 encode(float4 v) {
        const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
        uint2 t;
        t.x = *reinterpret_cast<const uint32_t *>(&a);
        t.y = *reinterpret_cast<const uint32_t *>(&b);
        return t;
    }
};
template <> struct ColourPlane<true> {
    using texel = float4;
    static constexpr int kBytes = 16;
    static __device__ __forceinline__ float4 decode(float4 t) { return t; }
    static __device__ __forceinline__ float4 encode(float4 v) { return v; }
};
template <bool F32> struct MomentsPlane;
template <> struct MomentsPlane<false> {
    using texel = uint32_t;  // 2 x fp16
    static __device__ __forceinline__ float2 decode(uint32_t t) { return __half22float2(*reinterpret_cast<const __half2 *>(&t)); }
    static __device__ __forceinline__ uint32_t encode(float2 v) {
        const __half2 a = __floats2half2_rn(v.x, v.y);
        return *reinterpret_cast<const uint32_t *>(&a);
    }
};
template <> struct MomentsPlane<true> {
    using texel = float2;
    static __device__ __forceinline__ float2 decode(float2 t) { return t; }
    static __device__ __forceinline__ float2 encode(float2 v) { return v; }
};

// The reference clamps values to [0,1] on imageLoad / imageStore (src/Filter.cuh:63-69,78-83).
__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }
__device__ __forceinline__ float4 clamp01(float4 v) { return make_float4(clamp01(v.x), clamp01(v.y), clamp01(v.z), clamp01(v.w)); }

// CalculateLuminance, src/Filter.cuh:260-263.  Deliberately NOT contracted into FMAs: with a centre
// variance near 0 the luminance edge-stopping scale is phi*1e-5, so the weights amplify luminance rounding
// differences by ~1e4; each product and sum rounds once, left to right, exactly like the scalar oracle.
// Luminance is computed once per pixel (never per tap), so the two extra instructions are free.
__device__ __forceinline__ float luminance(float r, float g, float b) {
    return __fadd_rn(__fadd_rn(__fmul_rn(0.2126f, r), __fmul_rn(0.7152f, g)), __fmul_rn(0.0722f, b));
}
// glm::mix(x, y, a) = x*(1-a) + y*a without contraction (temporal pass: bit-exact against the oracle)
__device__ __forceinline__ float mix_rn(float x, float y, float a) {
    return __fadd_rn(__fmul_rn(x, __fsub_rn(1.0f, a)), __fmul_rn(y, a));
}

// ---- compact guide texel (16 B): everything the consistency tests and edge-stopping functions need from
// the three G-buffer planes (motion.zw, normal.xyz, uv.w = 32 B of texels), emitted once per frame by the
// temporal pass.  Exact: depth stays fp32, normals are already fp16 in the G-buffer, the mesh id keeps its
// fp16 bits.
//   x: z  (GetDepth: 0 -> 1e30, src/Filter.cuh:199-207)   y: dz (0 for background)
//   z: half2(nx, ny) bits                                  w: lo16 = half(nz) bits, hi16 = uv.w bits
__device__ __forceinline__ float4 make_guide(float4 motion, ushort4 nrm, ushort4 uv) {
    float4 g;
    const bool bg = (motion.z == 0.0f);
    g.x = bg ? kBackgroundZ : motion.z;
    g.y = bg ? 0.0f : motion.w;
    g.z = __uint_as_float((uint32_t)nrm.x | ((uint32_t)nrm.y << 16));
    g.w = __uint_as_float((uint32_t)nrm.z | ((uint32_t)uv.w << 16));
    return g;
}
__device__ __forceinline__ float3 guide_normal(float4 g) {
    const uint32_t a = __float_as_uint(g.z), b = __float_as_uint(g.w);
    const float2 xy = __half22float2(*reinterpret_cast<const __half2 *>(&a));
    const float z = __half2float(__ushort_as_half((unsigned short)(b & 0xffffu)));
    return make_float3(xy.x, xy.y, z);
}
// int(SampleCuTexture(UV).w) as GBuffer.frag:77 intended it: fp16 -> float -> int (cvt.rzi)
__device__ __forceinline__ int guide_mesh_id(float4 g) {
    return __float2int_rz(__half2float(__ushort_as_half((unsigned short)(__float_as_uint(g.w) >> 16))));
}
__device__ __forceinline__ float dot3(float3 a, float3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }

// ---- edge-stopping weight, reference computeWeight (src/Filter.cuh:407-427):
//   w = exp(-|dl|/phiL - |dz|/phiZ) * pow(sat(n.n'), phiN)
// evaluated with one MUFU.LG2 + one MUFU.EX2:
//   w = 2^( (phiN/4) * lg2(max(sat(d)^4, tiny)) - |dl|*kL - |dz|*kZ )
// with kL = log2e/phiL, kZ = log2e/phiZ (0 when phiZ == 0, the reference's special case :420) folded per pixel.
// Two exact-ish squarings before the log keep lg2.approx's absolute error (2^-22.6, which matters because
// normals of one surface give d = 1 - O(1e-3)) from being multiplied by the full phiN: exponent error is
// phiN/4 * 2^-22.6 (5e-6 at phiN = 128).  max(.., tiny) keeps d == 0 finite: tiny^(phiN/4) <= 1e-9 for the
// supported phiN >= 1 (pow(0, phiN) = 0), and phiN == 0 gives 2^0 = 1 like pow(x, 0).
__device__ __forceinline__ float fast_log2(float x) {  // x is a normal number at every call site
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float edge_weight_log2(float dl_abs_kL, float dz_abs_kZ, float ndot, float phiN_over_4) {
    const float d = __saturatef(ndot);
    const float d2 = d * d;
    const float d4 = fmaxf(d2 * d2, 1e-36f);
    return phiN_over_4 * fast_log2(d4) - dl_abs_kL - dz_abs_kZ;
}

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// read-only, L1-allocating vector loads
template <typename T> __device__ __forceinline__ T ldg(const T *p) { return __ldg(p); }

}  // namespace svgf
