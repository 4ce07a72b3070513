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

// ---- compact guide planes: everything the consistency tests and edge-stopping functions need from the three
// G-buffer planes (motion.zw, normal.xyz, uv.w = 32 B of texels), emitted once per frame by the temporal pass
// and kept by the context for the next frame's previous-frame tests.  Exact: depth stays fp32, the fp16
// normals are widened to fp32 (so the a-trous tiles can be bulk-copied into shared memory with no per-texel
// conversion), the mesh id keeps its fp16 bits.
//   n   : float4 (z', nx, ny, nz)   z' = GetDepth: 0 -> 1e30 (src/Filter.cuh:199-207)      16 B/px
//   dz  : float  depth derivative (0 for background)                                         4 B/px
//   mid : fp16 bits of uv.w, the instance index (GBuffer.frag:77)                            2 B/px
//   seg : float4 per 32-pixel ROW SEGMENT (x / 32, y): (nx, ny, nz, state) - state 0: the segment holds only background
//         texels (z' = 1e30) or lies outside the image, 1: all its non-background texels carry the normal xyz, 2: mixed.
//         Lets an a-trous tile decide "one normal everywhere" from 96 entries instead of 1000+ texels
//         (svgf_kernels_lattice.cuh); background texels are wildcards because their weight is 0 through z'.
struct Guide {
    float4 *n;
    float *dz;
    unsigned short *mid;
    float4 *seg;
};
// Row groups of 64 x 3 outputs per CTA of the lattice kernel: 4 -> 256 threads, tiles of 12 rows, two CTAs per SM;
// 2 -> 128 threads, tiles of 6 rows, four CTAs per SM (tools/build_exp.sh 2).  A launch's row-block unit stays 12 * STEP rows.
#ifndef SVGF_EXP
#define SVGF_EXP 0      // build-time experiment selector (tools/build_exp.sh); 0 = the shipped form
#endif
#if SVGF_EXP == 2
constexpr int kLatRowGroups = 2;
#else
constexpr int kLatRowGroups = 4;
#endif
// Padding (pixels / rows) around the context-owned lattice planes of svgf_kernels_lattice.cuh: the widest halo, 2 * 2^4.
constexpr int kLatPadX = 32, kLatPadY = 32;

// One warp = one 32-pixel row segment (lane = pixel).  `gn` is the lane's guide texel (z', nx, ny, nz); lanes outside the
// image pass valid = false.  Returns the segment's map entry (identical in every lane).
__device__ __forceinline__ float4 segment_state(float4 gn, bool valid) {
    const bool solid = valid && gn.x != 1e30f;
    const unsigned m = __ballot_sync(0xffffffffu, solid);
    if (m == 0) return make_float4(0.f, 0.f, 0.f, 0.0f);
    const int leader = __ffs(m) - 1;
    const float rx = __shfl_sync(0xffffffffu, gn.y, leader), ry = __shfl_sync(0xffffffffu, gn.z, leader),
                rz = __shfl_sync(0xffffffffu, gn.w, leader);
    const bool same = !solid || (gn.y == rx && gn.z == ry && gn.w == rz);   // compared as floats: +0 == -0, NaN never matches
    return make_float4(rx, ry, rz, __all_sync(0xffffffffu, same) ? 1.0f : 2.0f);
}
struct GuideTexel {
    float4 n;
    float dz;
    unsigned short mid;
};
__device__ __forceinline__ GuideTexel make_guide(float4 motion, ushort4 nrm, ushort4 uv) {
    GuideTexel g;
    const bool bg = (motion.z == 0.0f);
    g.n = make_float4(bg ? kBackgroundZ : motion.z, __half2float(__ushort_as_half(nrm.x)), __half2float(__ushort_as_half(nrm.y)),
                      __half2float(__ushort_as_half(nrm.z)));
    g.dz = bg ? 0.0f : motion.w;
    g.mid = uv.w;
    return g;
}
__device__ __forceinline__ float3 guide_normal(float4 gn) { return make_float3(gn.y, gn.z, gn.w); }
// int(SampleCuTexture(UV).w) as GBuffer.frag:77 intended it: fp16 -> float -> int (cvt.rzi)
__device__ __forceinline__ int guide_mesh_id(unsigned short bits) { return __float2int_rz(__half2float(__ushort_as_half(bits))); }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }

// ---- edge-stopping weight, reference computeWeight (src/Filter.cuh:407-427):
//   w = exp(-|dl|/phiL - |dz|/phiZ) * pow(sat(n.n'), phiN)
// evaluated as w = 2^e with one MUFU.EX2.  The normal term needs care: normals of one surface give
// d = n.n' = 1 - O(1e-3), and pow(d, 128) amplifies any error in log(d) by 128, so
//   * d itself is formed exactly like the reference: (x*x' + y*y') + z*z' (products of fp16-origin values are
//     exact in fp32, so FMA contraction cannot change the two roundings);
//   * for phiN >= 32 ("series" mode) log2(d) = log2(1 - u), u = 1 - sat(d) (exact), is a 5-term Taylor
//     polynomial with phiN*log2e folded into the coefficients: relative error u^5/6, i.e. < 1e-7 absolute on
//     every weight that is not itself < 1e-9 (lg2.approx would give 2^-22 ABSOLUTE error on log2 -> 2e-5
//     relative on the weight);
//   * for phiN < 32 lg2.approx is accurate enough (error phiN * 2^-22 on the exponent) and exact for d -> 0
//     (max(d, tiny): tiny^phiN underflows like pow(0, phiN) = 0; phiN == 0 gives 2^0 = 1 like pow(x, 0)).
// kL = log2e/phiL and kZ = log2e/phiZ (0 when phiZ == 0, the reference's special case :420) are folded per
// pixel by the callers; `base` = |dl|*kL + |dz|*kZ (+ -log2 of the tap's kernel weight where there is one).
__device__ __forceinline__ float fast_log2(float x) {  // x is a normal number at every call site
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct NormalTerm {      // uniform per launch
    float k1, k2, k3, k4, k5;   // series mode: phiN*log2e / i
    float phiN;
    int series;
};
__host__ __device__ inline NormalTerm make_normal_term(float phiN) {
    NormalTerm t;
    const float c = phiN * 1.4426950408889634f;
    t.k1 = c; t.k2 = c * 0.5f; t.k3 = c * (1.0f / 3.0f); t.k4 = c * 0.25f; t.k5 = c * 0.2f;
    t.phiN = phiN;
    t.series = (phiN >= 32.0f) ? 1 : 0;
    return t;
}
// Three-term form for phiN >= 100: u*(k1 + k2 u + k3 u^2) with the truncated c*u^4/4 term folded back into the u^2 and
// u^3 coefficients so that the error of the WEIGHT 2^-E is equi-oscillating.  In the scaled variable s = c*ln2*u the
// weight is ~e^-s and the problem min max_s e^-s |s^4 - alpha s^3 - beta s^2| has the universal solution
// alpha = 5.71, beta = -6.52 (tools/fit_series.py), which gives k2 = c/2 + beta/(4 c ln^2 2), k3 = c/3 + alpha/(4 ln 2).
// Max absolute weight error 0.56/c^3: 8.9e-8 at phiN = 128, 1.9e-7 at phiN = 100 (checked by tests/test_series.py) -
// the same class as the 4-term Taylor form (1.5e-8 .. 3.9e-8), three orders below the 1e-4 parity bar.
__host__ __device__ inline NormalTerm economised_series3(float phiN) {
    NormalTerm t = make_normal_term(phiN);
    const float c = phiN * 1.4426950408889634f;
    t.k2 = c * 0.5f - 3.3930f / c;
    t.k3 = c * (1.0f / 3.0f) + 2.0594f;
    t.k4 = t.k5 = 0.0f;
    return t;
}
// exponent of the weight: e = phiN*log2(sat(ndot)) - base
template <bool SERIES>
__device__ __forceinline__ float edge_exponent(float base, float ndot, const NormalTerm &t) {
    const float d = __saturatef(ndot);
    if (SERIES) {
        const float u = 1.0f - d;   // exact (d in [0,1]: Sterbenz for d >= 0.5, and exact enough below: weight ~ 0)
        float p = fmaf(u, t.k5, t.k4);
        p = fmaf(u, p, t.k3);
        p = fmaf(u, p, t.k2);
        p = fmaf(u, p, t.k1);
        return fmaf(-u, p, -base);
    } else {
        return fmaf(t.phiN, fast_log2(fmaxf(d, 1e-37f)), -base);
    }
}

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 1 / sqrt(x) and 1 / x on the MUFU for arguments that are normal numbers at every call site (1e-10 + variance >= 1e-10;
// a weight sum >= 1): rsqrtf() / __frcp_rn() spend a dozen instructions on denormal scaling and an exactly rounded
// reciprocal that a 1-ulp result (1.2e-7 relative, against the 1e-4 parity bar) does not need.
__device__ __forceinline__ float fast_rsqrt(float x) {
#if SVGF_EXP == 6
    return rsqrtf(x);
#else
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#endif
}
__device__ __forceinline__ float fast_rcp(float x) {
#if SVGF_EXP == 6
    return __frcp_rn(x);
#else
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#endif
}

// read-only, L1-allocating vector loads
template <typename T> __device__ __forceinline__ T ldg(const T *p) { return __ldg(p); }

// ---- lattice planes (svgf_kernels_lattice.cuh): a level's input pre-transformed by its producer ---------------------
// destination planes of a level whose successor is a lattice level (origin of the PADDED allocation)
struct LatticeColour { float4 *c0, *c1, *lz; };
struct LatticeNormals { float4 *n0; float2 *n1; };

__device__ __forceinline__ size_t lattice_index(int gx, int gy, int pitch_pairs) {
    return (size_t)(gy + kLatPadY) * pitch_pairs + ((gx + kLatPadX) >> 1);
}

// What the next level's imageLoad would see of a stored result (src/Filter.cuh:78-83 after :618): round through the
// storage format, clamp, luminance.
template <bool F32>
__device__ __forceinline__ void lattice_requantise(float4 &o) {
    if (!F32) o = ColourPlane<false>::decode(ColourPlane<false>::encode(o));
    o = make_float4(__saturatef(o.x), __saturatef(o.y), __saturatef(o.z), __saturatef(o.w));
}
template <bool F32>
__device__ __forceinline__ void lattice_store_pair(const LatticeColour &dst, size_t li, float4 o0, float4 o1, float2 z) {
    lattice_requantise<F32>(o0);
    lattice_requantise<F32>(o1);
    dst.c0[li] = make_float4(o0.x, o1.x, o0.y, o1.y);
    dst.c1[li] = make_float4(o0.z, o1.z, o0.w, o1.w);
    dst.lz[li] = make_float4(luminance(o0.x, o0.y, o0.z), luminance(o1.x, o1.y, o1.z), z.x, z.y);
}

}  // namespace svgf
