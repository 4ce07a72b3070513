// svgf_kernels_tiled.cuh — the a-trous level as a persistent, shared-memory-tiled, register-blocked stencil
// with asynchronous bulk staging (cp.async.bulk -> UBLKCP, completion on an mbarrier).
//
// Why it looks like this (B200, measured):
//  * The level is bound by the FP32 pipe, not by HBM: 24 taps x ~19 FP32 ops + 1 MUFU.EX2 per tap and pixel
//    against 32-48 bytes of compulsory traffic per pixel (tools/microbench.cu: 120 lane-FMA/clk/SM, FFMA2 buys
//    issue slots but no FMA throughput, 16 EX2/clk/SM).  So the design minimises issued instructions per tap
//    and hides every byte of global latency behind that arithmetic.
//  * Everything per PIXEL rather than per (pixel, tap) is done once when a tile is staged: fp16 -> fp32 decode,
//    the reference's [0,1] clamp (src/Filter.cuh:82) and luminance (un-contracted, see svgf_device.cuh).
//    The guide plane already holds (z, nx, ny, nz) as fp32, so it is bulk-copied straight into its final
//    shared-memory plane with no per-texel instruction at all.  A tap is then 2 x LDS.128 + 1 x LDS.32.
//  * Each thread owns a column of R = 4 outputs spaced STEP rows apart: a staged tap is loaded once and used by
//    up to 5 outputs (40 tap loads serve 96 tap evaluations); otherwise shared-memory bandwidth would bind.
//  * Dilation: a tile is ONE row phase of the level's lattice — rows y0 + STEP*j — over a contiguous 128-pixel
//    x range, so every row of the tile is one contiguous run in global memory (one bulk copy per row and plane,
//    issued by one warp) and the tile in shared memory is (128 + 4*STEP) x (rows + 4) texels at every dilation.
//  * Persistent CTAs walk the tile list; while tile i is being filtered the bulk copies of tile i+1 are in
//    flight (raw colour buffer + the other half of the double-buffered guide plane).
//  * STEP is a template parameter: every shared-memory offset in the tap loop is an immediate.
//  * Taps outside the image are "null" texels (z = +inf, colour 0) written by the conversion pass:
//    |zc - inf| * kZ = inf drives the exponent to -inf and the weight to exactly 0 — the reference skips those
//    taps (src/Filter.cuh:579).
#pragma once
#include "svgf_device.cuh"

namespace svgf {

constexpr int kTileW = 128;      // output pixels per tile row (contiguous in x)
constexpr int kRowsPerThread = 4;

template <bool F32, int STEP, int RG> struct TileGeom {   // RG = row groups = warps-in-y of the CTA
    static constexpr int tile_rows = kRowsPerThread * RG;  // lattice rows of outputs per tile
    static constexpr int threads = kTileW * RG;
    static constexpr int cols = kTileW + 4 * STEP;
    static constexpr int rows = tile_rows + 4;
    static constexpr int texels = cols * rows;
    static constexpr int raw_texel = F32 ? 16 : 8;
    // sA (16) + sG x2 (32) + sL (4) + raw colour (8|16) per texel, + mbarrier
    static constexpr size_t smem_bytes = (size_t)texels * (16 + 32 + 4 + raw_texel) + 16;
};

struct AtrousTiledArgs {
    int W, H;
    float kL_scale;      // log2e / phi_colour
    float kZ_scale;      // log2e / (STEP * phi_depth)
    float k1, k2, k3, k4, k5;   // normal term series coefficients (make_normal_term)
    int level;
    int tiles_x, tiles_y;       // tiles_y counts (row block, phase) pairs
    int uniform_tiles;          // packed kernel: allow the uniform-normal tile shortcut
    const float *var_blur;      // packed kernel: blurred variance plane (SVGF_VARIANCE_PREFILTER_GAUSS3) or nullptr
    int yblock0, nyblocks;      // packed / lattice kernels: restrict the launch to row blocks [yblock0, yblock0 + nyblocks) of the
                                // level's tile grid (a row block = 12 * STEP image rows, all STEP phases); nyblocks == 0: all
    int yblock1, nyblocks1;     // lattice kernel: a second range in the same launch (band driver: both boundary strips at once)
};

// -log2 of the reference's tap kernel KW[|xx|] * KW[|yy|], KW = {1, 2/3, 1/6} as floats (src/Filter.cuh:540,582):
// the kernel weight is folded into the exponent of the edge-stopping weight (one FFMA instead of an FMUL).
__device__ __forceinline__ constexpr float tap_neg_log2_kernel(int ax, int ay) {
    //            ay=0          ay=1          ay=2
    return ax == 0 ? (ay == 0 ? 0.0f : ay == 1 ? 0.584962458f : 2.58496246f)
         : ax == 1 ? (ay == 0 ? 0.584962458f : ay == 1 ? 1.16992489f : 3.16992489f)
                   : (ay == 0 ? 2.58496246f : ay == 1 ? 3.16992489f : 5.16992489f);
}

// ---- mbarrier / bulk-copy primitives (PTX ISA: mbarrier, cp.async.bulk) ------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (UBLKCP); size and both addresses multiples of 16 bytes
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <bool F32, int STEP, int RG, int TERMS>
__global__ void __launch_bounds__(kTileW *RG, (RG == 2 ? 2 : 1))
atrous_tiled_kernel(AtrousTiledArgs a, const float4 *__restrict__ guide_n, const float *__restrict__ guide_dz,
                    const typename ColourPlane<F32>::texel *__restrict__ in, typename ColourPlane<F32>::texel *__restrict__ out,
                    typename ColourPlane<F32>::texel *__restrict__ hist_colour) {
    using G = TileGeom<F32, STEP, RG>;
    using CT = typename ColourPlane<F32>::texel;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *sA = reinterpret_cast<float4 *>(smem_raw);                   // r g b v (clamped)           conversion output
    float4 *sG0 = sA + G::texels;                                        // z nx ny nz, buffer 0        bulk-copy destination
    float4 *sG1 = sG0 + G::texels;                                       // z nx ny nz, buffer 1
    CT *sR = reinterpret_cast<CT *>(sG1 + G::texels);                    // raw colour texels           bulk-copy destination
    float *sL = reinterpret_cast<float *>(sR + G::texels);               // luminance                   conversion output
    uint64_t *bar = reinterpret_cast<uint64_t *>(sL + G::texels);

    const int tid = threadIdx.x;
    const int n_tiles = a.tiles_x * a.tiles_y;

    auto tile_origin = [&](int tile, int &x0, int &y0) {
        const int ty = tile / a.tiles_x, tx = tile - ty * a.tiles_x;
        const int yblock = ty / STEP, phase = ty - yblock * STEP;
        x0 = tx * kTileW;
        y0 = yblock * (G::tile_rows * STEP) + phase;
    };
    // One warp issues the tile's bulk copies: lane r copies row r of the colour plane and of the guide plane.
    auto issue_tile = [&](int tile, float4 *sG) {
        if (tid < 32) {
            int x0, y0;
            tile_origin(tile, x0, y0);
            const int gx0 = x0 - 2 * STEP;
            const int cx0 = max(gx0, 0), cx1 = min(gx0 + G::cols, a.W);
            const int n = max(cx1 - cx0, 0);
            uint32_t bytes = 0;
            int gy = 0;
            if (tid < G::rows) {
                gy = y0 + (tid - 2) * STEP;
                if (gy >= 0 && gy < a.H) bytes = (uint32_t)n * (16 + G::raw_texel);
            }
            const uint32_t total = __reduce_add_sync(0xffffffffu, bytes);
            if (tid == 0) mbar_expect_tx(bar, total);
            __syncwarp();
            if (bytes) {
                const size_t gi = (size_t)gy * a.W + cx0;
                const int si = tid * G::cols + (cx0 - gx0);
                bulk_g2s(sR + si, in + gi, (uint32_t)n * G::raw_texel, bar);
                bulk_g2s(sG + si, guide_n + gi, (uint32_t)n * 16, bar);
            }
        }
    };

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    int tile = blockIdx.x;
    if (tile < n_tiles) issue_tile(tile, sG0);
    uint32_t parity = 0;
    int buf = 0;

    const int tx = tid & (kTileW - 1), tg = tid / kTileW;
    const int col = tx + 2 * STEP;
    const int row0 = tg * kRowsPerThread + 2;            // smem row of this thread's first output

    for (; tile < n_tiles; tile += gridDim.x) {
        int x0, y0;
        tile_origin(tile, x0, y0);
        const int gx = x0 + tx;
        float4 *sG = buf ? sG1 : sG0;

        // depth derivatives of this thread's outputs: issued before the wait so their latency hides behind it
        float dz[kRowsPerThread];
#pragma unroll
        for (int j = 0; j < kRowsPerThread; j++) {
            const int gy = y0 + (tg * kRowsPerThread + j) * STEP;
            dz[j] = (gx < a.W && gy < a.H) ? __ldg(guide_dz + (size_t)gy * a.W + gx) : 0.f;
        }

        mbar_wait(bar, parity);
        parity ^= 1;

        // ---- conversion pass: raw colour -> clamped fp32 + luminance; null texels outside the image ----
        for (int idx = tid; idx < G::texels; idx += G::threads) {
            const int r = idx / G::cols, c = idx - r * G::cols;
            const int px = x0 - 2 * STEP + c, py = y0 + (r - 2) * STEP;
            if (px >= 0 && px < a.W && py >= 0 && py < a.H) {
                const float4 v = ColourPlane<F32>::decode(sR[idx]);
                const float4 t = make_float4(__saturatef(v.x), __saturatef(v.y), __saturatef(v.z), __saturatef(v.w));   // :543,:586
                sA[idx] = t;
                sL[idx] = luminance(t.x, t.y, t.z);
            } else {
                sA[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
                sL[idx] = 0.f;
                sG[idx] = make_float4(__int_as_float(0x7f800000), 0.f, 0.f, 0.f);
            }
        }
        fence_proxy_async();      // generic-proxy reads of sR / the other guide buffer precede the async-proxy refill
        __syncthreads();
        const int next = tile + gridDim.x;
        if (next < n_tiles) issue_tile(next, buf ? sG0 : sG1);

        // ---- centre data of the R outputs ----
        float S[kRowsPerThread], Ar[kRowsPerThread], Ag[kRowsPerThread], Ab[kRowsPerThread], Av[kRowsPerThread];
        float lc[kRowsPerThread], zc[kRowsPerThread], nx[kRowsPerThread], ny[kRowsPerThread], nz[kRowsPerThread];
        float kL[kRowsPerThread], kZ[kRowsPerThread][5];
        bool live[kRowsPerThread];
        bool any_live = false;
#pragma unroll
        for (int j = 0; j < kRowsPerThread; j++) {
            const int si = (row0 + j) * G::cols + col;
            const float4 ca = sA[si], cg = sG[si];
            const int gy = y0 + (tg * kRowsPerThread + j) * STEP;
            S[j] = 1.0f; Ar[j] = ca.x; Ag[j] = ca.y; Ab[j] = ca.z; Av[j] = ca.w;     // :567-568
            lc[j] = sL[si]; zc[j] = cg.x; nx[j] = cg.y; ny[j] = cg.z; nz[j] = cg.w;
            live[j] = (gx < a.W) && (gy < a.H) && (cg.x != kBackgroundZ);            // :554: background passes through
            any_live |= live[j];
            kL[j] = a.kL_scale * rsqrtf(1e-10f + ca.w);                              // :562
            const float k = __fdividef(a.kZ_scale, fmaxf(dz[j], 1e-6f));             // :563
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
                    const float4 qa = sA[si], qg = sG[si];
                    const float ql = sL[si];
#pragma unroll
                    for (int j = 0; j < kRowsPerThread; j++) {
                        const int dy = t - j;
                        if (dy < -2 || dy > 2 || (dx == 0 && dy == 0)) continue;
                        const int ax = dx < 0 ? -dx : dx, ay = dy < 0 ? -dy : dy;
                        const int l2 = ax * ax + ay * ay;                     // 1 2 4 5 8
                        const int cls = l2 == 1 ? 0 : l2 == 2 ? 1 : l2 == 4 ? 2 : l2 == 5 ? 3 : 4;
                        const float ck = tap_neg_log2_kernel(ax, ay);
                        float base = fmaf(fabsf(ql - lc[j]), kL[j], ck);
                        base = fmaf(fabsf(qg.x - zc[j]), kZ[j][cls], base);
                        const float d = __saturatef(fmaf(nz[j], qg.w, fmaf(ny[j], qg.z, nx[j] * qg.y)));
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
            const CT o = ColourPlane<F32>::encode(make_float4(Ar[j] * inv, Ag[j] * inv, Ab[j] * inv, Av[j] * (inv * inv)));
            out[gi] = o;
            if (a.level == 0 && hist_colour) hist_colour[gi] = o;
        }
        __syncthreads();          // sA / sL / sG[buf] are rewritten by the next iteration
        buf ^= 1;
    }
}

}  // namespace svgf
