// svgf_kernels_fused.cuh — a-trous levels 0 and 1 in ONE launch, the level-0 result staying in shared memory.
//
// BASELINE.json configs[2] asks for the single-level and the two-level-fused variant side by side.  Fusion removes
// one colour-plane round trip through HBM (16 B/px written + re-read, plus the 20 B/px guide re-read of level 1) at
// the price of an apron: a tile of 64 x 24 level-1 outputs needs the level-0 result on 72 x 32 pixels (level-1 taps
// reach +-4), which needs inputs on 76 x 36 (level-0 taps reach +-2), so level 0 is evaluated 1.5 x redundantly.
// The level is bound by the FP32 pipe / register-operand bandwidth, not by HBM, so this variant is
// expected to LOSE; it exists to measure exactly that (svgf_params.flags & SVGF_FLAG_FUSE_LEVELS_01, off by default).
//
// Results are bit-identical to two single-level launches: the taps are the same packed FP32x2 code in the same order
// (pk_all_taps), the level-0 result is rounded through the storage format (fp16 or fp32) and re-clamped exactly like a
// plane written by one launch and staged by the next, and level 0 still writes the colour history for the tile interior.
//
// Geometry (all indices in the STAGING frame: pixel pair ip = 0..37, row ir = 0..36, global pixel
// (x0 - 6 + 2 ip, y0 - 6 + ir)):
//   inputs     rows 0..36, pairs 0..37        (colour, variance, luminance, depth, normal)
//   level 0    rows 2..34, pairs 1..36        396 tasks of 3 rows x 1 pair; written to the "mid" planes, same frame
//   level 1    rows 6..29, pairs 3..34        256 tasks; row phase ph = (ir - 6) & 1: lattice pitch = 2 rows
#pragma once
#include "svgf_kernels_packed.cuh"

namespace svgf {

constexpr int kFzW = 64, kFzH = 24, kFzThreads = 512;
struct FusedGeom {
    static constexpr int pairs = (kFzW + 12) / 2;            // 38
    static constexpr int rows = 37;                           // 33 level-0 rows (24 + 8, padded to 11 x 3) + 4
    static constexpr int npairs = pairs * rows;               // 1406
    static constexpr int l0_pairs = (kFzW + 8) / 2;           // 36
    static constexpr int l0_groups = 11;
    static constexpr int l0_tasks = l0_pairs * l0_groups;     // 396
    static constexpr int l1_tasks = (kFzW / 2) * (kFzH / kPkRows);   // 256
    // C0 C1 G0 G1 (16 B each) + L (8 B) for the inputs, C0 C1 + L for the level-0 result
    static constexpr size_t smem_bytes = (size_t)npairs * (72 + 40);
};

// centre setup + 24 taps + normalisation for kPkRows outputs x 2 pixels; the code of atrous_packed_kernel's body
template <int STEP, int TERMS, int PITCH>
__device__ __forceinline__ void fused_task(const AtrousTiledArgs &a, float kZ_scale, const float *__restrict__ guide_dz,
                                           const float4 *sC0, const float4 *sC1, const float4 *sG0, const float4 *sG1, const float2 *sL,
                                           int row0, int pcol, int gx, int gy0, int gy_stride, bool uniform_n, const PkCoef &k, float un,
                                           float pn, float4 (&o0)[kPkRows], float4 (&o1)[kPkRows], bool (&live0)[kPkRows],
                                           bool (&live1)[kPkRows]) {
    PkCentre C[kPkRows];
    PkAcc A[kPkRows];
    bool any_live = false;
#pragma unroll
    for (int j = 0; j < kPkRows; j++) {
        const int si = (row0 + j) * PITCH + pcol;
        const float4 c0 = sC0[si], c1 = sC1[si], g0 = sG0[si], g1 = sG1[si];
        const int gy = gy0 + j * gy_stride;
        A[j].S = f2bc(1.0f);
        A[j].r = make_float2(c0.x, c0.y); A[j].g = make_float2(c0.z, c0.w);
        A[j].b = make_float2(c1.x, c1.y); A[j].v = make_float2(c1.z, c1.w);
        C[j].lc = sL[si];
        C[j].zc = make_float2(g0.x, g0.y); C[j].nx = make_float2(g0.z, g0.w);
        C[j].ny = make_float2(g1.x, g1.y); C[j].nz = make_float2(g1.z, g1.w);
        const bool inside = (gx >= 0) && (gx < a.W) && (gy >= 0) && (gy < a.H);
        live0[j] = inside && (g0.x != kBackgroundZ);
        live1[j] = inside && (g0.y != kBackgroundZ);
        any_live |= live0[j] | live1[j];
        C[j].kL = make_float2(a.kL_scale * fast_rsqrt(1e-10f + c1.z), a.kL_scale * fast_rsqrt(1e-10f + c1.w));
        float2 dz = make_float2(0.f, 0.f);
        if (inside) dz = __ldg(reinterpret_cast<const float2 *>(guide_dz + (size_t)gy * a.W + gx));
        C[j].kZ = make_float2(__fdividef(kZ_scale, fmaxf(dz.x, 1e-6f)), __fdividef(kZ_scale, fmaxf(dz.y, 1e-6f)));
    }
    if (__any_sync(__activemask(), any_live)) {
        if (uniform_n) pk_all_taps<STEP, TERMS, true, PITCH, kPkRows>(A, C, sC0, sC1, sG0, sG1, sL, row0, pcol, k, un, pn);
        else pk_all_taps<STEP, TERMS, false, PITCH, kPkRows>(A, C, sC0, sC1, sG0, sG1, sL, row0, pcol, k, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < kPkRows; j++) {
        const int si = (row0 + j) * PITCH + pcol;
        const float4 c0 = sC0[si], c1 = sC1[si];
        const float i0 = fast_rcp(A[j].S.x), i1 = fast_rcp(A[j].S.y);
        o0[j] = make_float4(A[j].r.x * i0, A[j].g.x * i0, A[j].b.x * i0, A[j].v.x * (i0 * i0));
        o1[j] = make_float4(A[j].r.y * i1, A[j].g.y * i1, A[j].b.y * i1, A[j].v.y * (i1 * i1));
        if (!live0[j]) o0[j] = make_float4(c0.x, c0.z, c1.x, c1.z);
        if (!live1[j]) o1[j] = make_float4(c0.y, c0.w, c1.y, c1.w);
    }
}

template <bool F32> __device__ __forceinline__ void store_pair(typename ColourPlane<F32>::texel *p, size_t gi,
                                                               const typename ColourPlane<F32>::texel &e0,
                                                               const typename ColourPlane<F32>::texel &e1, bool w0, bool w1) {
    if (F32) {
        if (w0) p[gi] = e0;
        if (w1) p[gi + 1] = e1;
    } else {
        const uint2 u0 = *reinterpret_cast<const uint2 *>(&e0), u1 = *reinterpret_cast<const uint2 *>(&e1);
        if (w0 && w1) *reinterpret_cast<uint4 *>(p + gi) = make_uint4(u0.x, u0.y, u1.x, u1.y);
        else {
            if (w0) p[gi] = e0;
            if (w1) p[gi + 1] = e1;
        }
    }
}

template <bool F32, int TERMS>
__global__ void __launch_bounds__(kFzThreads, 1)
atrous_fused01_kernel(AtrousTiledArgs a, const float4 *__restrict__ guide_n, const float *__restrict__ guide_dz,
                      const typename ColourPlane<F32>::texel *__restrict__ in, typename ColourPlane<F32>::texel *__restrict__ out,
                      typename ColourPlane<F32>::texel *__restrict__ hist_colour) {
    using G = FusedGeom;
    using CT = typename ColourPlane<F32>::texel;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *sC0 = reinterpret_cast<float4 *>(smem_raw);
    float4 *sC1 = sC0 + G::npairs;
    float4 *sG0 = sC1 + G::npairs;
    float4 *sG1 = sG0 + G::npairs;
    float4 *mC0 = sG1 + G::npairs;                              // level-0 result, same frame as the inputs
    float4 *mC1 = mC0 + G::npairs;
    float2 *sL = reinterpret_cast<float2 *>(mC1 + G::npairs);
    float2 *mL = sL + G::npairs;

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * kFzW, y0 = blockIdx.y * kFzH;

    // ---- stage inputs (same per-texel work as atrous_packed_kernel) ----
    const float4 nref = __ldg(guide_n + (size_t)min(y0, a.H - 1) * a.W + min(x0, a.W - 2));
    bool same_n = a.uniform_tiles && (nref.y != 0.0f || nref.z != 0.0f || nref.w != 0.0f);
    for (int idx = tid; idx < G::npairs; idx += kFzThreads) {
        const int r = idx / G::pairs, pc = idx - r * G::pairs;
        const int gx = x0 - 6 + 2 * pc, gy = y0 - 6 + r;
        float4 rg0 = make_float4(__int_as_float(0x7f800000), 0.f, 0.f, 0.f), rg1 = rg0;
        CT rc0 = CT(), rc1 = CT();
        if (gx >= 0 && gx < a.W && gy >= 0 && gy < a.H) {
            const size_t gi = (size_t)gy * a.W + gx;
            if (F32) {
                rc0 = __ldg(in + gi);
                rc1 = __ldg(in + gi + 1);
            } else {
                const uint4 t = __ldg(reinterpret_cast<const uint4 *>(in + gi));
                *reinterpret_cast<uint2 *>(&rc0) = make_uint2(t.x, t.y);
                *reinterpret_cast<uint2 *>(&rc1) = make_uint2(t.z, t.w);
            }
            rg0 = __ldg(guide_n + gi);
            rg1 = __ldg(guide_n + gi + 1);
            same_n &= (rg0.y == nref.y) & (rg0.z == nref.z) & (rg0.w == nref.w) & (rg1.y == nref.y) & (rg1.z == nref.z) &
                      (rg1.w == nref.w);
        }
        const float4 c0 = ColourPlane<F32>::decode(rc0), c1 = ColourPlane<F32>::decode(rc1);
        const float r0 = __saturatef(c0.x), g0 = __saturatef(c0.y), b0 = __saturatef(c0.z), v0 = __saturatef(c0.w);
        const float r1 = __saturatef(c1.x), g1 = __saturatef(c1.y), b1 = __saturatef(c1.z), v1 = __saturatef(c1.w);
        sC0[idx] = make_float4(r0, r1, g0, g1);
        sC1[idx] = make_float4(b0, b1, v0, v1);
        sG0[idx] = make_float4(rg0.x, rg1.x, rg0.y, rg1.y);
        sG1[idx] = make_float4(rg0.z, rg1.z, rg0.w, rg1.w);
        sL[idx] = make_float2(luminance(r0, g0, b0), luminance(r1, g1, b1));
        mC0[idx] = mC1[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
        mL[idx] = make_float2(0.f, 0.f);
    }
    const bool uniform_n = __syncthreads_and(same_n) != 0;

    PkCoef k;
    k.k1 = a.k1; k.k2 = a.k2; k.k3 = a.k3; k.k4 = a.k4; k.k5 = a.k5;
    float un, pn;
    pk_normal_term<TERMS>(nref.y, nref.z, nref.w, nref.y, nref.z, nref.w, k, un, pn);

    // ---- level 0 on the 72 x 33 apron region -> mid planes (+ colour history for the tile interior) ----
    if (tid < G::l0_tasks) {
        const int g = tid / G::l0_pairs, pc = tid - g * G::l0_pairs;
        const int row0 = g * kPkRows + 2, pcol = pc + 1;
        const int gx = x0 - 6 + 2 * pcol, gy0 = y0 - 6 + row0;
        float4 o0[kPkRows], o1[kPkRows];
        bool live0[kPkRows], live1[kPkRows];
        fused_task<1, TERMS, G::pairs>(a, a.kZ_scale, guide_dz, sC0, sC1, sG0, sG1, sL, row0, pcol, gx, gy0, 1, uniform_n, k, un, pn, o0, o1,
                                       live0, live1);
#pragma unroll
        for (int j = 0; j < kPkRows; j++) {
            const int ir = row0 + j, gy = gy0 + j;
            const CT e0 = ColourPlane<F32>::encode(o0[j]), e1 = ColourPlane<F32>::encode(o1[j]);
            // what the next launch would stage from the plane this level writes: storage rounding, then the [0,1] clamp
            const float4 d0 = ColourPlane<F32>::decode(e0), d1 = ColourPlane<F32>::decode(e1);
            const float r0 = __saturatef(d0.x), g0 = __saturatef(d0.y), b0 = __saturatef(d0.z), v0 = __saturatef(d0.w);
            const float r1 = __saturatef(d1.x), g1 = __saturatef(d1.y), b1 = __saturatef(d1.z), v1 = __saturatef(d1.w);
            const int si = ir * G::pairs + pcol;
            const bool inside = gx >= 0 && gx < a.W && gy >= 0 && gy < a.H;
            if (inside) {
                mC0[si] = make_float4(r0, r1, g0, g1);
                mC1[si] = make_float4(b0, b1, v0, v1);
                mL[si] = make_float2(luminance(r0, g0, b0), luminance(r1, g1, b1));
            }
            const bool interior = inside && ir >= 6 && ir < 6 + kFzH && pcol >= 3 && pcol < 3 + kFzW / 2;
            if (interior && hist_colour) store_pair<F32>(hist_colour, (size_t)gy * a.W + gx, e0, e1, live0[j], live1[j]);
        }
    }
    __syncthreads();

    // ---- level 1 on the tile interior, colour from the mid planes, guide from the input planes ----
    if (tid < G::l1_tasks) {
        const int pc = tid & (kFzW / 2 - 1), t = tid / (kFzW / 2);   // t = 0..7: phase = t & 1, row group = t >> 1
        const int ph = t & 1, tg = t >> 1;
        const int row0 = tg * kPkRows + 3, pcol = pc + 3;            // lattice row l <-> staging row 2 l + ph
        const int gx = x0 + 2 * pc, gy0 = y0 - 6 + 2 * row0 + ph;
        const int off = ph * G::pairs;
        float4 o0[kPkRows], o1[kPkRows];
        bool live0[kPkRows], live1[kPkRows];
        fused_task<2, TERMS, 2 * G::pairs>(a, 0.5f * a.kZ_scale, guide_dz, mC0 + off, mC1 + off, sG0 + off, sG1 + off, mL + off, row0, pcol,
                                           gx, gy0, 2, uniform_n, k, un, pn, o0, o1, live0, live1);
#pragma unroll
        for (int j = 0; j < kPkRows; j++) {
            const int gy = gy0 + 2 * j;
            if (gx >= a.W || gy >= a.H) continue;
            store_pair<F32>(out, (size_t)gy * a.W + gx, ColourPlane<F32>::encode(o0[j]), ColourPlane<F32>::encode(o1[j]), true, true);
        }
    }
}

}  // namespace svgf
