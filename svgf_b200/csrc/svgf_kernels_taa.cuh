// svgf_kernels_taa.cuh — temporal anti-aliasing + sRGB resolve, the step right after the a-trous levels:
// reference filter::TAAFilterKernel (src/Filter.cuh:288-357), launched by application::TAA() (src/App.cu:516-522).
//
// What the reference kernel does per pixel (x, y):
//   * every "texture sample" is textureSample (:116-131), which returns the FLOOR texel of uv * (size - 1) (the bilinear
//     part is dead code), with uv = (x, y) / size: for almost every pixel that is texel (x - 1, y - 1) - the resolve shifts
//     the image by one pixel.  Reproduced by evaluating the same float expressions, not by assuming the shift;
//   * history (previous output, sRGB-encoded, alpha 1) and the current texel are blended in squared space with rate
//     min(alpha, 0.5), the result is clamped in PAL-YUV to the box of the 3x3 neighbourhood (plus-shaped box, then averaged
//     with the full box), decoded, NaN -> 0, sRGB-encoded (:145-148) and stored with alpha 1.  The mix-rate update of
//     :341-346 never reaches memory (the stored alpha is the constant 1) and is not evaluated.
// D13: the reference reads the history from the very plane it writes (another thread may have overwritten the texel,
// and application::TAA passes FilterBuffer[1], which the a-trous ping-pong has just used as scratch).  Here the history is
// a separate, caller-owned plane: the previous call's output (snapshot semantics, like the history-length plane, D3).
//
// Streaming pass: 8 B in + 8 B history + 8 B out per pixel (fp16 storage).  pow(x, 2) = x * x, pow(x, 0.5) = IEEE sqrt; only the final, continuous
// pow(x, 1/2.4) is ex2(lg2(x) / 2.4) on the MUFU (<= 2^-21 relative).
#pragma once
#include "svgf_device.cuh"

namespace svgf {

// The resolve has a DISCONTINUITY: a decoded component that comes out negative makes pow(x, 0.5) NaN and the reference
// then blacks out the whole pixel (:351).  The YUV matrices are inverses of each other only to ~5e-6, so components that
// should be 0 land within rounding of it; to take the same side as the scalar restatement everything up to that decision
// is evaluated un-contracted, in the reference's order, with IEEE square roots.
__device__ __forceinline__ float taa_dot3(float x, float y, float z, float a, float b, float c) {   // glm::dot: (x*a + y*b) + z*c
    return __fadd_rn(__fadd_rn(__fmul_rn(x, a), __fmul_rn(y, b)), __fmul_rn(z, c));
}
__device__ __forceinline__ float3 taa_encode_pal_yuv(float3 c) {                     // :267-275
    const float r = __fmul_rn(c.x, c.x), g = __fmul_rn(c.y, c.y), b = __fmul_rn(c.z, c.z);
    return make_float3(taa_dot3(r, g, b, 0.299f, 0.587f, 0.114f), taa_dot3(r, g, b, -0.14713f, -0.28886f, 0.436f),
                       taa_dot3(r, g, b, 0.615f, -0.51499f, -0.10001f));
}
__device__ __forceinline__ float taa_sqrt(float x) {   // pow(x, 0.5): NaN for x < 0, +0 for -0
    return __fadd_rn(__fsqrt_rn(x), 0.0f);
}
__device__ __forceinline__ float taa_to_srgb(float c) {                             // :145-148
    const float hi = 1.055f * fast_exp2(fast_log2(fmaxf(c, 1e-30f)) * (1.0f / 2.4f)) - 0.055f;
    return (c <= 0.0031308f) ? 12.92f * c : hi;
}
// floor texel of textureSample (:116-131) along one axis: uv * (size - 1), floor, clamp
__device__ __forceinline__ int taa_texel(float uv, int size) {
    const int t = __float2int_rd(uv * (float)(size - 1));
    return min(max(t, 0), size - 1);
}

constexpr int kTaaBW = 32, kTaaBH = 16, kTaaTW = kTaaBW + 4, kTaaTH = kTaaBH + 4;

// One output pixel.  WINDOW: all nine taps come from the block's shared-memory window (row / column offsets computed once,
// one add per tap); otherwise from global memory (a block whose window test failed - never, up to float rounding).
template <bool F32, bool WINDOW>
__device__ __forceinline__ void taa_pixel(int W, int x, int y, const int (&xs)[3], const int (&ys)[3], int tx0, int ty0, const float4 *sYuv,
                                          const float4 *sSq, const typename ColourPlane<F32>::texel *__restrict__ filtered,
                                          const typename ColourPlane<F32>::texel *__restrict__ history,
                                          typename ColourPlane<F32>::texel *__restrict__ out) {
    const int col[3] = {xs[0] - tx0, xs[1] - tx0, xs[2] - tx0};
    const int row[3] = {(ys[0] - ty0) * kTaaTW, (ys[1] - ty0) * kTaaTW, (ys[2] - ty0) * kTaaTW};
    // encoded texel (ix, iy) of the neighbourhood, and optionally its squared rgb
    auto tap = [&](int ix, int iy, float4 *sq) -> float4 {
        if (WINDOW) {
            const int li = row[iy] + col[ix];
            if (sq) *sq = sSq[li];
            return sYuv[li];
        }
        const float4 c = clamp01(ColourPlane<F32>::decode(__ldg(filtered + (size_t)ys[iy] * W + xs[ix])));
        const float r = __fmul_rn(c.x, c.x), g = __fmul_rn(c.y, c.y), b = __fmul_rn(c.z, c.z);
        if (sq) *sq = make_float4(r, g, b, 0.f);
        return make_float4(taa_dot3(r, g, b, 0.299f, 0.587f, 0.114f), taa_dot3(r, g, b, -0.14713f, -0.28886f, 0.436f),
                           taa_dot3(r, g, b, 0.615f, -0.51499f, -0.10001f), 0.f);
    };
    const float4 last = clamp01(ColourPlane<F32>::decode(__ldg(history + (size_t)ys[1] * W + xs[1])));   // :299
    float4 c0sq;
    const float4 e0 = tap(1, 1, &c0sq);                                              // :306
    const float rate = fminf(last.w, 0.5f);                                          // :302
    float3 aa = make_float3(mix_rn(__fmul_rn(last.x, last.x), c0sq.x, rate), mix_rn(__fmul_rn(last.y, last.y), c0sq.y, rate),
                            mix_rn(__fmul_rn(last.z, last.z), c0sq.z, rate));       // :308
    aa = taa_encode_pal_yuv(make_float3(__fsqrt_rn(aa.x), __fsqrt_rn(aa.y), __fsqrt_rn(aa.z)));   // :309,:320
    // plus-shaped box (in0..in4), then the diagonal texels (in5..in8) folded in: :331-336
    float3 mn = make_float3(e0.x, e0.y, e0.z), mx = mn;
    auto fold = [&](int ix, int iy, float3 &lo, float3 &hi) {
        const float4 e = tap(ix, iy, nullptr);
        lo = make_float3(fminf(lo.x, e.x), fminf(lo.y, e.y), fminf(lo.z, e.z));
        hi = make_float3(fmaxf(hi.x, e.x), fmaxf(hi.y, e.y), fmaxf(hi.z, e.z));
    };
    fold(2, 1, mn, mx); fold(0, 1, mn, mx); fold(1, 2, mn, mx); fold(1, 0, mn, mx);
    float3 mn2 = mn, mx2 = mx;
    fold(2, 2, mn2, mx2); fold(0, 2, mn2, mx2); fold(2, 0, mn2, mx2); fold(0, 0, mn2, mx2);
    auto half_mix = [](float a, float b) { return __fadd_rn(0.5f * a, 0.5f * b); };   // exact halves, one rounding: == the reference's FP64 mix
    mn = make_float3(half_mix(mn.x, mn2.x), half_mix(mn.y, mn2.y), half_mix(mn.z, mn2.z));
    mx = make_float3(half_mix(mx.x, mx2.x), half_mix(mx.y, mx2.y), half_mix(mx.z, mx2.z));
    aa = make_float3(fminf(fmaxf(aa.x, mn.x), mx.x), fminf(fmaxf(aa.y, mn.y), mx.y), fminf(fmaxf(aa.z, mn.z), mx.z));   // :339
    // decodePalYuv :277-285
    float3 rgb = make_float3(taa_dot3(aa.x, aa.y, aa.z, 1.0f, 0.0f, 1.13983f), taa_dot3(aa.x, aa.y, aa.z, 1.0f, -0.39465f, -0.58060f),
                             taa_dot3(aa.x, aa.y, aa.z, 1.0f, 2.03211f, 0.0f));
    rgb = make_float3(taa_sqrt(rgb.x), taa_sqrt(rgb.y), taa_sqrt(rgb.z));
    if (rgb.x != rgb.x || rgb.y != rgb.y || rgb.z != rgb.z) rgb = make_float3(0.f, 0.f, 0.f);   // :351
    const float4 o = make_float4(taa_to_srgb(rgb.x), taa_to_srgb(rgb.y), taa_to_srgb(rgb.z), 1.0f);   // :353
    out[(size_t)y * W + x] = ColourPlane<F32>::encode(clamp01(o));                   // :355 imageStore
}

// Block = 32 x 16 outputs.  The block's source texels (its outputs' floor texels and their neighbours: a 36 x 20 window
// starting two texels up-left of the block, see taa_texel) are loaded, clamped and PAL-YUV-encoded ONCE into shared memory -
// each is the neighbour of nine outputs - together with their squared rgb (the blend operand of the centre tap).  A tap
// that falls outside the window (it cannot, up to float rounding of the uv arithmetic at the clamped image borders, which
// the window covers) is evaluated from global memory instead, so the window size is an optimisation, not an assumption.

template <bool F32>
__global__ void __launch_bounds__(kTaaBW *kTaaBH)
taa_kernel(int W, int H, const typename ColourPlane<F32>::texel *__restrict__ filtered,
           const typename ColourPlane<F32>::texel *__restrict__ history, typename ColourPlane<F32>::texel *__restrict__ out) {
    __shared__ float4 sYuv[kTaaTW * kTaaTH];    // y u v (w unused)
    __shared__ float4 sSq[kTaaTW * kTaaTH];     // r^2 g^2 b^2 (w unused)
    __shared__ int sXs[3][kTaaBW], sYs[3][kTaaBH];   // floor texel of uv - 1/size, uv, uv + 1/size per column / row of the block
    const int bx0 = blockIdx.x * kTaaBW, by0 = blockIdx.y * kTaaBH;
    const int tx0 = max(bx0 - 3, 0), ty0 = max(by0 - 3, 0);      // window origin: floor texels reach x - 2 (one more for rounding slack)
    const float inv_w = 1.0f / (float)W, inv_h = 1.0f / (float)H;
    // the sampling positions depend on the column (row) only: evaluated once per block, in the reference's float expressions
    bool covered = true;
    if (threadIdx.x < 3 * kTaaBW) {
        const int o = threadIdx.x / kTaaBW, cxx = threadIdx.x - o * kTaaBW;
        const float u = (float)(bx0 + cxx) * inv_w;                                  // :296
        const int t = taa_texel(o == 0 ? u - inv_w : o == 1 ? u : u + inv_w, W);
        sXs[o][cxx] = t;
        covered = t >= tx0 && t < tx0 + kTaaTW;
    } else if (threadIdx.x < 3 * kTaaBW + 3 * kTaaBH) {
        const int k = threadIdx.x - 3 * kTaaBW, o = k / kTaaBH, cyy = k - o * kTaaBH;
        const float v = (float)(by0 + cyy) * inv_h;
        const int t = taa_texel(o == 0 ? v - inv_h : o == 1 ? v : v + inv_h, H);
        sYs[o][cyy] = t;
        covered = t >= ty0 && t < ty0 + kTaaTH;
    }
    for (int i = threadIdx.x; i < kTaaTW * kTaaTH; i += kTaaBW * kTaaBH) {
        const int ty = i / kTaaTW, tx = i - ty * kTaaTW;
        const int gx = min(tx0 + tx, W - 1), gy = min(ty0 + ty, H - 1);
        const float4 c = clamp01(ColourPlane<F32>::decode(__ldg(filtered + (size_t)gy * W + gx)));   // imageLoad :78-83
        const float r = __fmul_rn(c.x, c.x), g = __fmul_rn(c.y, c.y), b = __fmul_rn(c.z, c.z);
        sSq[i] = make_float4(r, g, b, 0.f);
        sYuv[i] = make_float4(taa_dot3(r, g, b, 0.299f, 0.587f, 0.114f), taa_dot3(r, g, b, -0.14713f, -0.28886f, 0.436f),
                              taa_dot3(r, g, b, 0.615f, -0.51499f, -0.10001f), 0.f);
    }
    // every tap of the block inside the window?  (it is, by the arithmetic above; the test makes the window an optimisation
    // instead of an assumption: a block that fails it takes its taps from global memory)
    const bool window_ok = __syncthreads_and(covered) != 0;
    const int cxl = threadIdx.x & (kTaaBW - 1), cyl = threadIdx.x / kTaaBW;
    const int x = bx0 + cxl, y = by0 + cyl;
    if (x >= W || y >= H) return;
    const int xs[3] = {sXs[0][cxl], sXs[1][cxl], sXs[2][cxl]};
    const int ys[3] = {sYs[0][cyl], sYs[1][cyl], sYs[2][cyl]};
    if (window_ok)
        taa_pixel<F32, true>(W, x, y, xs, ys, tx0, ty0, sYuv, sSq, filtered, history, out);
    else
        taa_pixel<F32, false>(W, x, y, xs, ys, tx0, ty0, sYuv, sSq, filtered, history, out);
}

}  // namespace svgf
