// svgf_kernels_tiled.cuh — the a-trous level as a shared-memory-tiled, register-blocked stencil.
//
// Why it looks like this (B200, measured: the one-thread-per-pixel kernel spends 0.36 ms per 4K level):
//  * The level is bound by the FP32 pipe, not HBM: 24 taps x ~18 FP32 ops + 1 MUFU per tap and pixel is
//    ~430 lane-ops per pixel against 32-48 bytes of compulsory traffic.  So the design minimises issued
//    instructions per tap and keeps traffic merely coalesced.
//  * Everything that is per PIXEL rather than per (pixel, tap) is done once when the tile is staged:
//    fp16 -> fp32 decode, the reference's [0,1] clamp (src/Filter.cuh:82), luminance (un-contracted,
//    see svgf_device.cuh) and the normal unpack.  Shared memory holds fp32 structure-of-planes texels
//    (rgbv | lum,z,nx,ny | nz = 36 B) so a tap is 2 x LDS.128 + 1 x LDS.32 and zero conversions.
//  * Each thread owns a column of R = 4 outputs spaced `STEP` rows apart, so a staged tap is loaded once
//    and used by up to 5 outputs: 40 tap loads serve 96 tap evaluations (shared-memory bandwidth would
//    otherwise be the bound: 36 B x 24 taps per pixel).
//  * Dilation: a CTA processes ONE row phase of the level's lattice — rows y0 + STEP*j — over a contiguous
//    128-pixel-wide x range, so global loads stay fully coalesced at every STEP and the tile in shared
//    memory is (128 + 4*STEP) x 12 texels whatever the dilation (58-83 KB, two CTAs per SM).
//  * STEP is a template parameter: every shared-memory offset is an immediate.
//  * Taps outside the image are staged as "null" texels (z = +inf, colour 0): |zc - inf| * kZ = inf drives the
//    exponent to -inf and the weight to exactly 0 — the reference skips those taps (src/Filter.cuh:579).
#pragma once
#include "svgf_device.cuh"

namespace svgf {

constexpr int kTileW = 128;      // output pixels per CTA row (contiguous in x)
constexpr int kRowsPerThread = 4;
constexpr int kRowGroups = 2;
constexpr int kTileRows = kRowsPerThread * kRowGroups;   // lattice rows of outputs per CTA
constexpr int kTiledThreads = kTileW * kRowGroups;

template <int STEP> struct TileGeom {
    static constexpr int cols = kTileW + 4 * STEP;
    static constexpr int rows = kTileRows + 4;
    static constexpr int texels = cols * rows;
    static constexpr size_t smem_bytes = (size_t)texels * 36;
};

struct AtrousTiledArgs {
    int W, H;
    float kL_scale;      // log2e / phi_colour
    float kZ_scale;      // log2e / (STEP * phi_depth)
    float k1, k2, k3, k4, k5;   // normal term series coefficients (make_normal_term)
    int level;
};

// -log2 of the reference's tap kernel KW[|xx|] * KW[|yy|], KW = {1, 2/3, 1/6} as floats (src/Filter.cuh:540,582):
// the kernel weight is folded into the exponent of the edge-stopping weight (one FFMA instead of an FMUL).
__device__ __forceinline__ constexpr float tap_neg_log2_kernel(int ax, int ay) {
    //            ay=0          ay=1          ay=2
    return ax == 0 ? (ay == 0 ? 0.0f : ay == 1 ? 0.584962458f : 2.58496246f)
         : ax == 1 ? (ay == 0 ? 0.584962458f : ay == 1 ? 1.16992489f : 3.16992489f)
                   : (ay == 0 ? 2.58496246f : ay == 1 ? 3.16992489f : 5.16992489f);
}

template <bool F32, int STEP, int TERMS>
__global__ void __launch_bounds__(kTiledThreads, 2)
atrous_tiled_kernel(AtrousTiledArgs a, const float4 *__restrict__ guide, const typename ColourPlane<F32>::texel *__restrict__ in,
                    typename ColourPlane<F32>::texel *__restrict__ out, typename ColourPlane<F32>::texel *__restrict__ hist_colour) {
    using G = TileGeom<STEP>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *sA = reinterpret_cast<float4 *>(smem_raw);                   // r g b v      (clamped)
    float4 *sB = sA + G::texels;                                         // lum z nx ny
    float *sC = reinterpret_cast<float *>(sB + G::texels);               // nz

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * kTileW;
    // blockIdx.y enumerates (row block, phase): rows y0 + STEP*j, j = 0..kTileRows-1
    const int yblock = blockIdx.y / STEP, phase = blockIdx.y % STEP;
    const int y0 = yblock * (kTileRows * STEP) + phase;

    // ---- stage the tile: decode, clamp, luminance, unpack — once per texel.  All global loads of the thread are
    // issued before the first conversion so the CTA pays one memory round trip, not one per texel. ----
    constexpr int kStageIters = (G::texels + kTiledThreads - 1) / kTiledThreads;
    typename ColourPlane<F32>::texel raw_c[kStageIters];
    float4 raw_g[kStageIters];
#pragma unroll
    for (int i = 0; i < kStageIters; i++) {
        const int idx = tid + i * kTiledThreads;
        const int r = idx / G::cols, c = idx - r * G::cols;
        const int gx = x0 - 2 * STEP + c, gy = y0 + (r - 2) * STEP;
        raw_g[i] = make_float4(__int_as_float(0x7f800000), 0.f, 0.f, 0.f);   // null texel: z = +inf, zero normal
        raw_c[i] = typename ColourPlane<F32>::texel();
        if (idx < G::texels && gx >= 0 && gx < a.W && gy >= 0 && gy < a.H) {
            const size_t gi = (size_t)gy * a.W + gx;
            raw_c[i] = __ldg(in + gi);
            raw_g[i] = __ldg(guide + gi);
        }
    }
#pragma unroll
    for (int i = 0; i < kStageIters; i++) {
        const int idx = tid + i * kTiledThreads;
        if (idx < G::texels) {
            const float4 col = ColourPlane<F32>::decode(raw_c[i]);
            const float4 g = raw_g[i];
            const float4 ta = make_float4(__saturatef(col.x), __saturatef(col.y), __saturatef(col.z), __saturatef(col.w));   // :543,:586
            const float3 n = guide_normal(g);
            sA[idx] = ta;
            sB[idx] = make_float4(luminance(ta.x, ta.y, ta.z), g.x, n.x, n.y);
            sC[idx] = n.z;
        }
    }
    __syncthreads();

    const int tx = tid & (kTileW - 1), tg = tid / kTileW;
    const int gx = x0 + tx;
    const int col = tx + 2 * STEP;
    const int row0 = tg * kRowsPerThread + 2;            // smem row of this thread's first output

    // ---- centre data of the R outputs ----
    float S[kRowsPerThread], Ar[kRowsPerThread], Ag[kRowsPerThread], Ab[kRowsPerThread], Av[kRowsPerThread];
    float lc[kRowsPerThread], zc[kRowsPerThread], nx[kRowsPerThread], ny[kRowsPerThread], nz[kRowsPerThread];
    float kL[kRowsPerThread], kZ[kRowsPerThread][5];
    bool live[kRowsPerThread];
    bool any_live = false;
#pragma unroll
    for (int j = 0; j < kRowsPerThread; j++) {
        const int si = (row0 + j) * G::cols + col;
        const float4 ca = sA[si], cb = sB[si];
        const int gy = y0 + (tg * kRowsPerThread + j) * STEP;
        S[j] = 1.0f; Ar[j] = ca.x; Ag[j] = ca.y; Ab[j] = ca.z; Av[j] = ca.w;     // :567-568
        lc[j] = cb.x; zc[j] = cb.y; nx[j] = cb.z; ny[j] = cb.w; nz[j] = sC[si];
        live[j] = (gx < a.W) && (gy < a.H) && (cb.y != kBackgroundZ);            // :554: background passes through
        any_live |= live[j];
        kL[j] = a.kL_scale * rsqrtf(1e-10f + ca.w);                              // :562
        float dz = 0.f;
        if (gx < a.W && gy < a.H) dz = __ldg(guide + (size_t)gy * a.W + gx).y;
        const float k = __fdividef(a.kZ_scale, fmaxf(dz, 1e-6f));                // :563
        kZ[j][0] = k;                              // length 1
        kZ[j][1] = k * 0.70710678f;                // sqrt(2)
        kZ[j][2] = k * 0.5f;                       // 2
        kZ[j][3] = k * 0.44721360f;                // sqrt(5)
        kZ[j][4] = k * 0.35355339f;                // sqrt(8)
    }

    if (__any_sync(0xffffffffu, any_live)) {
#pragma unroll
        for (int dx = -2; dx <= 2; dx++) {
#pragma unroll
            for (int t = -2; t < kRowsPerThread + 2; t++) {
                const int si = (row0 + t) * G::cols + col + dx * STEP;
                const float4 qa = sA[si], qb = sB[si];
                const float qnz = sC[si];
#pragma unroll
                for (int j = 0; j < kRowsPerThread; j++) {
                    const int dy = t - j;
                    if (dy < -2 || dy > 2 || (dx == 0 && dy == 0)) continue;
                    const int ax = dx < 0 ? -dx : dx, ay = dy < 0 ? -dy : dy;
                    const int l2 = ax * ax + ay * ay;                     // 1 2 4 5 8
                    const int cls = l2 == 1 ? 0 : l2 == 2 ? 1 : l2 == 4 ? 2 : l2 == 5 ? 3 : 4;
                    const float ck = tap_neg_log2_kernel(ax, ay);
                    float base = fmaf(fabsf(qb.x - lc[j]), kL[j], ck);
                    base = fmaf(fabsf(qb.y - zc[j]), kZ[j][cls], base);
                    const float d = __saturatef(fmaf(nz[j], qnz, fmaf(ny[j], qb.w, nx[j] * qb.z)));
                    const float u = 1.0f - d;
                    float p;
                    if (TERMS == 5) { p = fmaf(u, a.k5, a.k4); p = fmaf(u, p, a.k3); }
                    else p = fmaf(u, a.k4, a.k3);
                    p = fmaf(u, p, a.k2);
                    p = fmaf(u, p, a.k1);
                    const float w = fast_exp2(fmaf(-u, p, -base));
                    S[j] += w;                                                  // :607-608
                    Ar[j] = fmaf(w, qa.x, Ar[j]);
                    Ag[j] = fmaf(w, qa.y, Ag[j]);
                    Ab[j] = fmaf(w, qa.z, Ab[j]);
                    Av[j] = fmaf(w * w, qa.w, Av[j]);
                }
            }
        }
    }

    // ---- normalise and store (:615-622) ----
#pragma unroll
    for (int j = 0; j < kRowsPerThread; j++) {
        const int gy = y0 + (tg * kRowsPerThread + j) * STEP;
        if (gx >= a.W || gy >= a.H) continue;
        const size_t gi = (size_t)gy * a.W + gx;
        if (!live[j]) {
            out[gi] = ColourPlane<F32>::encode(sA[(row0 + j) * G::cols + col]);              // :556 (clamped centre)
            continue;
        }
        const float inv = __frcp_rn(S[j]);
        const typename ColourPlane<F32>::texel o =
            ColourPlane<F32>::encode(make_float4(Ar[j] * inv, Ag[j] * inv, Ab[j] * inv, Av[j] * (inv * inv)));
        out[gi] = o;
        if (a.level == 0 && hist_colour) hist_colour[gi] = o;
    }
}

}  // namespace svgf
