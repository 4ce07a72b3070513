// svgf_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see svgf_oracle.h).
//
// Scalar restatement of the reference's filter math.  Every function cites the reference lines it follows
// (paths relative to /root/reference).  Arithmetic notes that matter for parity:
//   * compiled with -O2 -ffp-contract=off: every float op rounds once, in source order;
//   * sub-expressions the reference evaluates in FP64 (un-suffixed literals) are kept in FP64:
//     src/Filter.cuh:381 (alpha), :424 (exp), :461 (phiDepth of FilterMoments), :514 (variance boost),
//     :562 (phiIllumination);
//   * glm semantics: mix = x*(1-a) + y*a (submodules/glm/glm/detail/func_common.inl:87),
//     clamp = min(max(x,lo),hi) with min(x,y) = (y<x)?y:x, max(x,y) = (x<y)?y:x (:17-30,240-246),
//     dot(vec3) = (x*x' + y*y') + z*z', length(vec2) = sqrt(x*x + y*y)
//     (submodules/glm/glm/detail/func_geometric.inl);
//   * un-templated max(float,float) in device code is CUDA's fmaxf, max(float,double) is fmax.
// Decisions D1-D12 of SURVEY.md §8 are encoded as documented in include/svgf.h.
#include "svgf_oracle.h"

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

int g_threads = 0;

// ---- half <-> float (src/Filter.cuh:18-52: __float2half = round-to-nearest-even, __half2float exact) ----
inline uint16_t f2h(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const uint32_t absx = x & 0x7fffffffu;
    if (absx > 0x7f800000u) return (uint16_t)(sign | 0x7fffu);  // NaN -> canonical NaN (CUDA returns 0x7fff)
    if (absx >= 0x47800000u) {                                  // |f| >= 65536 -> inf (65520..65536 handled below)
        return (uint16_t)(sign | 0x7c00u);
    }
    if (absx < 0x33000000u) return (uint16_t)sign;              // |f| < 2^-25 -> +-0
    int32_t exp = (int32_t)(absx >> 23) - 127;
    uint32_t man = (absx & 0x7fffffu) | 0x800000u;              // 24-bit significand
    uint32_t half;
    if (exp < -14) {
        // subnormal half: value = man * 2^(exp-23); target unit 2^-24
        const int shift = (-14 - exp) + 13;                     // bits to drop from the 24-bit significand
        const uint32_t q = man >> shift;
        const uint32_t rem = man & ((1u << shift) - 1u);
        const uint32_t halfway = 1u << (shift - 1);
        half = q;
        if (rem > halfway || (rem == halfway && (q & 1u))) half++;
    } else {
        const uint32_t q = ((uint32_t)(exp + 15) << 10) | ((man >> 13) & 0x3ffu);
        const uint32_t rem = man & 0x1fffu;
        half = q;
        if (rem > 0x1000u || (rem == 0x1000u && (q & 1u))) half++;  // carry may roll into exponent / inf: correct
    }
    return (uint16_t)(sign | half);
}

inline float h2f(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    const uint32_t exp = (h >> 10) & 0x1fu;
    const uint32_t man = h & 0x3ffu;
    uint32_t x;
    if (exp == 0) {
        if (man == 0) {
            x = sign;
        } else {
            int e = -1;
            uint32_t m = man;
            do { e++; m <<= 1; } while ((m & 0x400u) == 0);
            x = sign | ((uint32_t)(127 - 15 - e) << 23) | ((m & 0x3ffu) << 13);
        }
    } else if (exp == 31) {
        x = sign | 0x7f800000u | (man << 13);
    } else {
        x = sign | ((exp + 112u) << 23) | (man << 13);
    }
    float f;
    std::memcpy(&f, &x, 4);
    return f;
}

struct V4 { float x, y, z, w; };
struct V2 { float x, y; };
struct V3 { float x, y, z; };

inline float glm_max(float x, float y) { return (x < y) ? y : x; }
inline float glm_min(float x, float y) { return (y < x) ? y : x; }
inline float clamp01(float v) { return glm_min(glm_max(v, 0.0f), 1.0f); }
inline float sat(float v) { return (v != v) ? 0.0f : (v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v)); }  // __saturatef
inline float mixf(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline float dot3(const V3 &a, const V3 &b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }

// CUDA float->int conversion semantics (cvt.rzi.s32.f32): NaN -> 0, saturating.
inline int f2i_rz(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (-2147483647 - 1);
    return (int)f;
}

// ---- planes -------------------------------------------------------------------------------------------
template <bool F32> struct Colour {  // half4 / float4 plane, dense y*W+x (src/App.cu:763-771)
    static V4 raw(const void *p, size_t i) {
        if (F32) { const float *f = (const float *)p + 4 * i; return {f[0], f[1], f[2], f[3]}; }
        const uint16_t *h = (const uint16_t *)p + 4 * i;
        return {h2f(h[0]), h2f(h[1]), h2f(h[2]), h2f(h[3])};
    }
    static void store_raw(void *p, size_t i, V4 v) {
        if (F32) { float *f = (float *)p + 4 * i; f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; return; }
        uint16_t *h = (uint16_t *)p + 4 * i;
        h[0] = f2h(v.x); h[1] = f2h(v.y); h[2] = f2h(v.z); h[3] = f2h(v.w);
    }
    // imageLoad (src/Filter.cuh:78-83): coordinates are in range at every call site; value clamp [0,1].
    static V4 ld01(const void *p, size_t i) {
        V4 v = raw(p, i);
        return {clamp01(v.x), clamp01(v.y), clamp01(v.z), clamp01(v.w)};
    }
    // imageStore (src/Filter.cuh:63-69)
    static void st01(void *p, size_t i, V4 v) { store_raw(p, i, {clamp01(v.x), clamp01(v.y), clamp01(v.z), clamp01(v.w)}); }
    static void copy_bits(const void *src, void *dst, size_t i) {
        const size_t b = F32 ? 16 : 8;
        std::memcpy((char *)dst + b * i, (const char *)src + b * i, b);
    }
};
template <bool F32> struct Moments {  // half2 / float2
    static V2 raw(const void *p, size_t i) {
        if (F32) { const float *f = (const float *)p + 2 * i; return {f[0], f[1]}; }
        const uint16_t *h = (const uint16_t *)p + 2 * i;
        return {h2f(h[0]), h2f(h[1])};
    }
    static void store_raw(void *p, size_t i, V2 v) {
        if (F32) { float *f = (float *)p + 2 * i; f[0] = v.x; f[1] = v.y; return; }
        uint16_t *h = (uint16_t *)p + 2 * i;
        h[0] = f2h(v.x); h[1] = f2h(v.y);
    }
};

struct GBuf {
    const char *normal, *uv, *motion;
    size_t np, up, mp;
    GBuf(const svgf_gbuffer *g, int W)
        : normal((const char *)g->normal_mat), uv((const char *)g->uv_inst), motion((const char *)g->motion_depth),
          np(g->normal_pitch ? g->normal_pitch : (size_t)W * 8), up(g->uv_pitch ? g->uv_pitch : (size_t)W * 8),
          mp(g->motion_pitch ? g->motion_pitch : (size_t)W * 16) {}
    const float *mot(int x, int y) const { return (const float *)(motion + (size_t)y * mp) + 4 * x; }
    // GetDepth, src/Filter.cuh:199-207
    V2 depth(int x, int y) const {
        const float *m = mot(x, y);
        if (m[2] == 0.0f) return {1e30f, 0.0f};
        return {m[2], m[3]};
    }
    // SampleCuTextureHalf4(...).xyz, src/Filter.cuh:188-197
    V3 nrm(int x, int y) const {
        const uint16_t *n = (const uint16_t *)(normal + (size_t)y * np) + 4 * x;
        return {h2f(n[0]), h2f(n[1]), h2f(n[2])};
    }
    // src/Filter.cuh:245-246 with D2
    int mesh_id(int x, int y, int mode) const {
        if (mode == SVGF_MESH_ID_REFERENCE_VACUOUS) return 0;
        const uint16_t *u = (const uint16_t *)(uv + (size_t)y * up) + 4 * x;
        return f2i_rz(h2f(u[3]));
    }
};

// CalculateLuminance, src/Filter.cuh:260-263
inline float lum(float r, float g, float b) { return 0.2126f * r + 0.7152f * g + 0.0722f * b; }

// computeWeight, src/Filter.cuh:407-427
inline float compute_weight(float zc, float zq, float phiZ, const V3 &nc, const V3 &nq, float phiN, float lc, float lq,
                            float phiL) {
    const float wN = powf(sat(dot3(nc, nq)), phiN);                                 // :419
    const float wZ = (phiZ == 0.0f) ? 0.0f : fabsf(zc - zq) / phiZ;                 // :420
    const float wL = fabsf(lc - lq) / phiL;                                         // :422
    const double e = exp(0.0 - fmax((double)wL, 0.0) - fmax((double)wZ, 0.0));      // :424 (FP64)
    return (float)(e * (double)wN);
}

int check_common(const svgf_params *p, int W, int H, int storage) {
    if (!p || W <= 0 || H <= 0) return SVGF_INVALID_ARG;
    if (storage != SVGF_STORE_F16 && storage != SVGF_STORE_F32) return SVGF_INVALID_ARG;
    if (p->history_cap < 1 || p->history_cap > 255) return SVGF_INVALID_ARG;  // D9
    if (p->atrous_iterations < 0 || p->atrous_iterations > 10) return SVGF_INVALID_ARG;
    if (p->reproj_mode != SVGF_REPROJ_NEAREST_TRUNC && p->reproj_mode != SVGF_REPROJ_BILINEAR) return SVGF_UNSUPPORTED;
    if (p->depth_test_mode != SVGF_DEPTH_TEST_ABSOLUTE && p->depth_test_mode != SVGF_DEPTH_TEST_RELATIVE) return SVGF_INVALID_ARG;
    if (p->variance_prefilter != SVGF_VARIANCE_PREFILTER_NONE && p->variance_prefilter != SVGF_VARIANCE_PREFILTER_GAUSS3)
        return SVGF_UNSUPPORTED;
    return SVGF_OK;
}

// ---- A.1 temporal: src/Filter.cuh:359-404, LoadPreviousData :225-258 -------------------------------------
template <bool F32>
void temporal(const svgf_params &P, int W, int H, const GBuf &cur, const GBuf &prev, const void *prev_colour,
              void *cur_colour, const uint8_t *hprev, uint8_t *hout, void *cur_mom, const void *prev_mom) {
#pragma omp parallel for schedule(static) num_threads(g_threads > 0 ? g_threads : omp_get_max_threads())
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            const size_t i = (size_t)y * W + x;
            const V4 c4 = Colour<F32>::ld01(cur_colour, i);                          // :370
            const V3 c = {c4.x, c4.y, c4.z};
            V3 pc = {0, 0, 0};
            V2 pm = {0, 0};
            int h = 1;
            bool ok = false;
            // the reference's consistency tests for one previous-frame texel (src/Filter.cuh:235-252)
            auto consistent = [&](int qx, int qy) {
                if (qx < 0 || qx >= W || qy < 0 || qy >= H) return false;            // :235
                const V2 dc = cur.depth(x, y);                                       // :239
                const V2 dp = prev.depth(qx, qy);                                    // :240
                if (P.depth_test_mode == SVGF_DEPTH_TEST_RELATIVE) {
                    if (fabsf(dp.x - dc.x) / (dc.y + 1e-2f) > P.depth_threshold) return false;   // :241, the commented-out form
                } else if (fabsf(dp.x - dc.x) > P.depth_threshold) return false;     // :242
                if (cur.mesh_id(x, y, P.mesh_id_mode) != prev.mesh_id(qx, qy, P.mesh_id_mode)) return false;  // :245-247
                const V3 n0 = cur.nrm(x, y), n1 = prev.nrm(qx, qy);                  // :250-251
                return !(dot3(n0, n1) < P.normal_threshold);                         // :252
            };
            const float *mv = cur.mot(x, y);                                         // :230
            if (P.reproj_mode == SVGF_REPROJ_BILINEAR) {
                // SVGF_REPROJ_BILINEAR (include/svgf.h; the paper's 2x2 fetch, not in the reference)
                const float fxp = (float)x + mv[0], fyp = (float)y + mv[1];
                if (fabsf(fxp) < 1e8f && fabsf(fyp) < 1e8f) {                        // false for NaN
                    const float flx = floorf(fxp), fly = floorf(fyp);
                    const float tx = fxp - flx, ty = fyp - fly;
                    const int x0 = (int)flx, y0 = (int)fly;
                    const float wx[2] = {1.0f - tx, tx}, wy[2] = {1.0f - ty, ty};
                    float ws = 0.0f, hs = 0.0f;
                    V3 cs = {0, 0, 0};
                    V2 ms = {0, 0};
                    for (int j = 0; j < 2; j++)
                        for (int k = 0; k < 2; k++) {
                            const int qx = x0 + k, qy = y0 + j;
                            if (!consistent(qx, qy)) continue;
                            const float w = wx[k] * wy[j];
                            const size_t qi = (size_t)qy * W + qx;
                            const V4 p4 = Colour<F32>::ld01(prev_colour, qi);
                            const V2 m2 = Moments<F32>::raw(prev_mom, qi);
                            ws += w;
                            cs.x += w * p4.x; cs.y += w * p4.y; cs.z += w * p4.z;
                            ms.x += w * m2.x; ms.y += w * m2.y;
                            hs += w * (float)hprev[qi];
                        }
                    if (ws >= 0.01f) {
                        pc = {cs.x / ws, cs.y / ws, cs.z / ws};
                        pm = {ms.x / ws, ms.y / ws};
                        h = (int)(hs / ws + 0.5f);
                        ok = true;
                    }
                }
            } else {
                const int qx = x + f2i_rz(mv[0]);                                    // :232 ivec2(vec2) truncation
                const int qy = y + f2i_rz(mv[1]);
                if (consistent(qx, qy)) {
                    const size_t qi = (size_t)qy * W + qx;
                    const V4 p4 = Colour<F32>::ld01(prev_colour, qi);                // :254
                    pc = {p4.x, p4.y, p4.z};
                    h = (int)hprev[qi];                                              // :255
                    pm = Moments<F32>::raw(prev_mom, qi);                            // :256
                    ok = true;
                }
            }
            float alpha, alpha_m;
            if (ok) {
                h = (P.history_cap < h + 1) ? P.history_cap : h + 1;                 // :380
                alpha = (float)(1.0 / (double)h);                                    // :381 (FP64 divide)
                alpha_m = alpha;
                alpha = fmaxf(alpha, P.alpha_min);                                   // D11 (0 = reference)
                alpha_m = fmaxf(alpha_m, P.moments_alpha_min);
            } else {
                alpha = 1.0f; alpha_m = 1.0f; h = 1;                                 // :385-386
            }
            const float L = lum(c.x, c.y, c.z);                                      // :391
            V2 m = {L, L * L};                                                       // :392
            m = {mixf(pm.x, m.x, alpha_m), mixf(pm.y, m.y, alpha_m)};                // :393
            const float var = fmaxf(0.0f, m.y - m.x * m.x);                          // :396
            const V3 nc = {mixf(pc.x, c.x, alpha), mixf(pc.y, c.y, alpha), mixf(pc.z, c.z, alpha)};  // :398
            hout[i] = (uint8_t)h;                                                    // :400
            Colour<F32>::st01(cur_colour, i, {nc.x, nc.y, nc.z, var});               // :401
            Moments<F32>::store_raw(cur_mom, i, m);                                  // :402
        }
    }
}

// ---- A.2 variance: src/Filter.cuh:430-525 ----------------------------------------------------------------
template <bool F32>
void variance(const svgf_params &P, int W, int H, const GBuf &G, const void *in, const void *mom, const uint8_t *hist,
              void *out) {
#pragma omp parallel for schedule(dynamic, 4) num_threads(g_threads > 0 ? g_threads : omp_get_max_threads())
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            const size_t i = (size_t)y * W + x;
            const float h = (float)hist[i];                                          // :442
            if (!(h < 4.0f)) {                                                       // :444
                Colour<F32>::copy_bits(in, out, i);                                  // :521 (half->float->half is the identity)
                continue;
            }
            float sumW = 0.0f;
            V3 sumC = {0, 0, 0};
            V2 sumM = {0, 0};
            const V4 cc = Colour<F32>::raw(in, i);                                   // :450
            const float lc = lum(cc.x, cc.y, cc.z);
            const V2 zc = G.depth(x, y);                                             // :453  (:454-458 is dead: z is never < 0)
            const V3 nc = G.nrm(x, y);                                               // :459
            const float phiL = P.phi_colour;                                         // :460
            const float phiZ0 = (float)(fmax((double)zc.y, 1e-8) * 3.0) * P.phi_depth;  // :461 (FP64), D11
            for (int yy = -3; yy <= 3; yy++) {                                       // :465-469
                for (int xx = -3; xx <= 3; xx++) {
                    const int px = x + xx, py = y + yy;
                    if (!(px < W && py < H && px >= 0 && py >= 0)) continue;         // :473,:477
                    const size_t qi = (size_t)py * W + px;
                    const V4 cq = Colour<F32>::raw(in, qi);                          // :479
                    const V2 mq = Moments<F32>::raw(mom, qi);                        // :480
                    const float lq = lum(cq.x, cq.y, cq.z);
                    const float zq = G.depth(px, py).x;
                    const V3 nq = G.nrm(px, py);
                    const float len = sqrtf((float)(xx * xx + yy * yy));             // length(vec2(xx,yy))
                    const float w = compute_weight(zc.x, zq, phiZ0 * len, nc, nq, P.phi_normal, lc, lq, phiL);  // :485-495
                    sumW += w;                                                       // :497
                    sumC.x += cq.x * w; sumC.y += cq.y * w; sumC.z += cq.z * w;      // :498
                    sumM.x += mq.x * w; sumM.y += mq.y * w;                          // :499
                }
            }
            sumW = fmaxf(sumW, 1e-6f);                                               // :505
            sumC = {sumC.x / sumW, sumC.y / sumW, sumC.z / sumW};                    // :507
            sumM = {sumM.x / sumW, sumM.y / sumW};                                   // :508
            float var = sumM.y - sumM.x * sumM.x;                                    // :511
            var = (float)((double)var * (4.0 / (double)h));                          // :514 (FP64)
            Colour<F32>::store_raw(out, i, {sumC.x, sumC.y, sumC.z, var});           // :516
        }
    }
}

// SVGF_VARIANCE_PREFILTER_GAUSS3 (include/svgf.h; not in the reference): the clamped variance of the level's input
// blurred with (1 2 1; 2 4 2; 1 2 1) / 16, rows first, coordinates clamped to the image.
template <bool F32>
float variance_gauss3(const void *in, int W, int H, int x, int y) {
    float col[3];
    for (int dx = -1; dx <= 1; dx++) {
        const int px = x + dx < 0 ? 0 : (x + dx >= W ? W - 1 : x + dx);
        const int ym = y > 0 ? y - 1 : 0, yp = y < H - 1 ? y + 1 : H - 1;
        const float a = Colour<F32>::ld01(in, (size_t)ym * W + px).w, b = Colour<F32>::ld01(in, (size_t)y * W + px).w,
                    c = Colour<F32>::ld01(in, (size_t)yp * W + px).w;
        col[dx + 1] = (0.25f * a + 0.5f * b) + 0.25f * c;
    }
    return (0.25f * col[0] + 0.5f * col[1]) + 0.25f * col[2];
}

// ---- A.4 a-trous level: src/Filter.cuh:527-624 -----------------------------------------------------------
template <bool F32>
void atrous(const svgf_params &P, int W, int H, const GBuf &G, const void *in, void *out, void *hist_colour, int level) {
    const int step = 1 << level;
    const float KW[3] = {1.0f, (float)(2.0 / 3.0), (float)(1.0 / 6.0)};              // :540
#pragma omp parallel for schedule(dynamic, 4) num_threads(g_threads > 0 ? g_threads : omp_get_max_threads())
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            const size_t i = (size_t)y * W + x;
            const V4 c = Colour<F32>::ld01(in, i);                                   // :543
            const float lc = lum(c.x, c.y, c.z);                                     // :544
            const float var = (P.variance_prefilter == SVGF_VARIANCE_PREFILTER_GAUSS3) ? variance_gauss3<F32>(in, W, H, x, y)
                                                                                       : c.w;  // :547
            const V2 zc = G.depth(x, y);                                             // :552
            if (zc.x == 1e30f) {                                                     // :554-558
                Colour<F32>::store_raw(out, i, c);
                continue;
            }
            const V3 nc = G.nrm(x, y);                                               // :560
            const float eps = 1e-10f;                                                // :539
            const float phiL = (float)((double)P.phi_colour * sqrt(fmax(0.0, (double)(eps + var))));  // :562 (FP64)
            const float phiZ = fmaxf(zc.y, 1e-6f) * (float)step * P.phi_depth;       // :563, D11
            float sumW = 1.0f;                                                       // :567
            V4 sum = c;                                                              // :568
            for (int yy = -2; yy <= 2; yy++) {
                for (int xx = -2; xx <= 2; xx++) {
                    const int px = x + xx * step, py = y + yy * step;                // :576
                    const bool inside = (px < W && py < H && px >= 0 && py >= 0);    // :579
                    const float kernel = KW[xx < 0 ? -xx : xx] * KW[yy < 0 ? -yy : yy];  // :582
                    if (!(inside && (xx != 0 || yy != 0))) continue;                 // :584
                    const size_t qi = (size_t)py * W + px;
                    const V4 cq = Colour<F32>::ld01(in, qi);                         // :586
                    const float lq = lum(cq.x, cq.y, cq.z);                          // :587
                    const float zq = G.depth(px, py).x;                              // :588
                    const V3 nq = G.nrm(px, py);                                     // :589
                    const float len = sqrtf((float)(xx * xx + yy * yy));
                    const float w = compute_weight(zc.x, zq, phiZ * len, nc, nq, P.phi_normal, lc, lq, phiL);  // :592-602
                    const float iw = w * kernel;                                     // :604
                    sumW += iw;                                                      // :607
                    sum.x += iw * cq.x; sum.y += iw * cq.y; sum.z += iw * cq.z;      // :608
                    sum.w += (iw * iw) * cq.w;
                }
            }
            const float s2 = sumW * sumW;
            const V4 o = {sum.x / sumW, sum.y / sumW, sum.z / sumW, sum.w / s2};     // :615
            Colour<F32>::store_raw(out, i, o);                                       // :618
            if (level == 0 && hist_colour) Colour<F32>::store_raw(hist_colour, i, o);  // :619-622
        }
    }
}

// ---- A.6 TAA + sRGB resolve: TAAFilterKernel, src/Filter.cuh:288-357 (encodePalYuv / decodePalYuv :267-285, ToSRGB
// :145-148, textureSample :116-141 which returns the floor texel - the bilinear part is dead code behind `return c00`).
// D13: the reference reads its history from the plane it is writing (`Output`, :299 vs :355; another thread may already
// have overwritten the texel) - here the history is a separate plane (snapshot semantics, like D3).
inline V3 encode_pal_yuv(V3 rgb) {                                                   // :267-275
    rgb = {powf(rgb.x, 2.0f), powf(rgb.y, 2.0f), powf(rgb.z, 2.0f)};                 // glm::pow(vec3, vec3(2.0))
    return {dot3(rgb, {0.299f, 0.587f, 0.114f}), dot3(rgb, {-0.14713f, -0.28886f, 0.436f}), dot3(rgb, {0.615f, -0.51499f, -0.10001f})};
}
inline V3 decode_pal_yuv(V3 yuv) {                                                   // :277-285
    const V3 rgb = {dot3(yuv, {1.0f, 0.0f, 1.13983f}), dot3(yuv, {1.0f, -0.39465f, -0.58060f}), dot3(yuv, {1.0f, 2.03211f, 0.0f})};
    return {powf(rgb.x, 0.5f), powf(rgb.y, 0.5f), powf(rgb.z, 0.5f)};                // glm::pow(vec3, vec3(1.0 / 2.0))
}
inline float to_srgb(float c) {                                                      // :145-148
    return (c <= 0.0031308f) ? 12.92f * c : (1 + 0.055f) * powf(c, 1 / 2.4f) - 0.055f;
}
inline V3 min3(V3 a, V3 b) { return {glm_min(a.x, b.x), glm_min(a.y, b.y), glm_min(a.z, b.z)}; }
inline V3 max3(V3 a, V3 b) { return {glm_max(a.x, b.x), glm_max(a.y, b.y), glm_max(a.z, b.z)}; }
// glm::mix(vec3, vec3, double 0.5): evaluated in double, rounded to float once (submodules/glm/glm/detail/func_common.inl)
inline float mix_half_d(float x, float y) { return (float)((double)x * (1.0 - 0.5) + (double)y * 0.5); }

template <bool F32>
void taa(int W, int H, const void *filtered, const void *history, void *out) {
    const float inv_w = 1.0f / float(W), inv_h = 1.0f / float(H);                    // :294 (and `off`, :305)
    // textureSample :116-131: x = uv.x * (Width - 1); x0 = floor(x); clamp; imageLoad (value clamp [0,1])
    auto sample = [&](const void *plane, float u, float v) -> V4 {
        const float x = u * (float)(W - 1), y = v * (float)(H - 1);
        int x0 = (int)floorf(x), y0 = (int)floorf(y);
        x0 = x0 < 0 ? 0 : (x0 > W - 1 ? W - 1 : x0);
        y0 = y0 < 0 ? 0 : (y0 > H - 1 ? H - 1 : y0);
        return Colour<F32>::ld01(plane, (size_t)y0 * W + x0);
    };
#pragma omp parallel for schedule(static) num_threads(g_threads > 0 ? g_threads : omp_get_max_threads())
    for (int py = 0; py < H; py++) {
        for (int px = 0; px < W; px++) {
            const float u = (float)px * inv_w, v = (float)py * inv_h;                // :296
            const V4 last = sample(history, u, v);                                   // :299
            V3 aa = {last.x, last.y, last.z};
            const float mix_rate = (float)fmin((double)last.w, 0.5);                 // :302
            const V4 s0 = sample(filtered, u, v);                                    // :306
            V3 in[9];
            in[0] = {s0.x, s0.y, s0.z};
            aa = {mixf(aa.x * aa.x, in[0].x * in[0].x, mix_rate), mixf(aa.y * aa.y, in[0].y * in[0].y, mix_rate),
                  mixf(aa.z * aa.z, in[0].z * in[0].z, mix_rate)};                   // :308
            aa = {sqrtf(aa.x), sqrtf(aa.y), sqrtf(aa.z)};                            // :309
            const float du[8] = {+inv_w, -inv_w, 0.0f, 0.0f, +inv_w, -inv_w, +inv_w, -inv_w};   // :311-318
            const float dv[8] = {0.0f, 0.0f, +inv_h, -inv_h, +inv_h, +inv_h, -inv_h, -inv_h};
            for (int k = 0; k < 8; k++) {
                const V4 t = sample(filtered, u + du[k], v + dv[k]);
                in[k + 1] = {t.x, t.y, t.z};
            }
            aa = encode_pal_yuv(aa);                                                 // :320
            for (int k = 0; k < 9; k++) in[k] = encode_pal_yuv(in[k]);               // :321-329
            V3 mn = min3(min3(min3(in[0], in[1]), min3(in[2], in[3])), in[4]);       // :331
            V3 mx = max3(max3(max3(in[0], in[1]), max3(in[2], in[3])), in[4]);       // :332
            const V3 mn2 = min3(min3(min3(in[5], in[6]), min3(in[7], in[8])), mn);
            const V3 mx2 = max3(max3(max3(in[5], in[6]), max3(in[7], in[8])), mx);
            mn = {mix_half_d(mn.x, mn2.x), mix_half_d(mn.y, mn2.y), mix_half_d(mn.z, mn2.z)};   // :333-334
            mx = {mix_half_d(mx.x, mx2.x), mix_half_d(mx.y, mx2.y), mix_half_d(mx.z, mx2.z)};   // :335-336
            aa = {glm_min(glm_max(aa.x, mn.x), mx.x), glm_min(glm_max(aa.y, mn.y), mx.y), glm_min(glm_max(aa.z, mn.z), mx.z)};   // :339
            // :341-346 update mixRate, which is never stored (the output alpha is the constant 1, :350,:353)
            aa = decode_pal_yuv(aa);                                                 // :348
            V4 frag = {aa.x, aa.y, aa.z, 1.0f};
            if (aa.x != aa.x || aa.y != aa.y || aa.z != aa.z) frag = {0.0f, 0.0f, 0.0f, 0.0f};   // :351 IsFinite == !isnan (src/Common.cuh:85-93)
            frag = {to_srgb(frag.x), to_srgb(frag.y), to_srgb(frag.z), 1.0f};        // :353
            Colour<F32>::st01(out, (size_t)py * W + px, frag);                       // :355 imageStore
        }
    }
}

}  // namespace

extern "C" {

uint16_t svgf_oracle_f2h(float f) { return f2h(f); }
float svgf_oracle_h2f(uint16_t h) { return h2f(h); }
void svgf_oracle_set_threads(int n) { g_threads = n; }
int svgf_oracle_get_threads(void) {
#ifdef _OPENMP
    return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
    return 1;
#endif
}

int svgf_oracle_temporal(const svgf_params *p, int W, int H, int storage, const svgf_gbuffer *cur,
                         const svgf_gbuffer *prev, const void *prev_colour, void *cur_colour,
                         const uint8_t *history_prev, uint8_t *history_out, void *cur_moments,
                         const void *prev_moments) {
    int st = check_common(p, W, H, storage);
    if (st) return st;
    if (!cur || !prev || !prev_colour || !cur_colour || !history_prev || !history_out || !cur_moments || !prev_moments ||
        history_prev == history_out)
        return SVGF_INVALID_ARG;
    GBuf c(cur, W), q(prev, W);
    if (storage == SVGF_STORE_F32)
        temporal<true>(*p, W, H, c, q, prev_colour, cur_colour, history_prev, history_out, cur_moments, prev_moments);
    else
        temporal<false>(*p, W, H, c, q, prev_colour, cur_colour, history_prev, history_out, cur_moments, prev_moments);
    return SVGF_OK;
}

int svgf_oracle_variance(const svgf_params *p, int W, int H, int storage, const svgf_gbuffer *cur,
                         const void *colour_in, const void *moments, const uint8_t *history, void *colour_out) {
    int st = check_common(p, W, H, storage);
    if (st) return st;
    if (!cur || !colour_in || !moments || !history || !colour_out || colour_in == colour_out) return SVGF_INVALID_ARG;
    GBuf g(cur, W);
    if (storage == SVGF_STORE_F32) variance<true>(*p, W, H, g, colour_in, moments, history, colour_out);
    else variance<false>(*p, W, H, g, colour_in, moments, history, colour_out);
    return SVGF_OK;
}

int svgf_oracle_atrous_level(const svgf_params *p, int W, int H, int storage, const svgf_gbuffer *cur, const void *in,
                             void *out, void *history_colour_out, int level) {
    int st = check_common(p, W, H, storage);
    if (st) return st;
    if (!cur || !in || !out || in == out || level < 0 || level > 10) return SVGF_INVALID_ARG;
    GBuf g(cur, W);
    if (storage == SVGF_STORE_F32) atrous<true>(*p, W, H, g, in, out, history_colour_out, level);
    else atrous<false>(*p, W, H, g, in, out, history_colour_out, level);
    return SVGF_OK;
}

// src/App.cu:552-556 (stage order), :491-514 (level ping-pong, odd-N copy), D3 (history snapshot), D4 (current moments)
int svgf_oracle_frame(const svgf_params *p, int W, int H, int storage, const svgf_gbuffer gbuf[2],
                      const svgf_frame_buffers *b) {
    int st = check_common(p, W, H, storage);
    if (st) return st;
    if (!gbuf || !b || (b->ping_pong != 0 && b->ping_pong != 1)) return SVGF_INVALID_ARG;
    const int P = b->ping_pong, Q = 1 - P;
    const size_t n = (size_t)W * H;
    std::vector<uint8_t> hprev(b->history, b->history + n);
    st = svgf_oracle_temporal(p, W, H, storage, &gbuf[P], &gbuf[Q], b->render[Q], b->render[P], hprev.data(), b->history,
                              b->moments[P], b->moments[Q]);
    if (st) return st;
    st = svgf_oracle_variance(p, W, H, storage, &gbuf[P], b->render[P], b->moments[P], b->history, b->filter[0]);
    if (st) return st;
    int pp = 0;
    for (int i = 0; i < p->atrous_iterations; i++) {
        st = svgf_oracle_atrous_level(p, W, H, storage, &gbuf[P], b->filter[pp], b->filter[1 - pp], b->render[P], i);
        if (st) return st;
        pp = 1 - pp;
    }
    if (p->atrous_iterations % 2 != 0)  // src/App.cu:510-513
        std::memcpy(b->filter[0], b->filter[1], n * (storage == SVGF_STORE_F32 ? 16 : 8));
    return SVGF_OK;
}

// Albedo demodulation / remodulation (include/svgf.h; README.md:172-174: the paper's step the reference leaves out).
int svgf_oracle_demodulate(int W, int H, int storage, const void *albedo, void *colour) {
    if (W <= 0 || H <= 0 || !albedo || !colour) return SVGF_INVALID_ARG;
    for (size_t i = 0; i < (size_t)W * H; i++) {
        if (storage == SVGF_STORE_F32) {
            const V4 a = Colour<true>::raw(albedo, i), c = Colour<true>::raw(colour, i);
            Colour<true>::store_raw(colour, i, {c.x / fmaxf(a.x, 1e-3f), c.y / fmaxf(a.y, 1e-3f), c.z / fmaxf(a.z, 1e-3f), c.w});
        } else {
            const V4 a = Colour<false>::raw(albedo, i), c = Colour<false>::raw(colour, i);
            Colour<false>::store_raw(colour, i, {c.x / fmaxf(a.x, 1e-3f), c.y / fmaxf(a.y, 1e-3f), c.z / fmaxf(a.z, 1e-3f), c.w});
        }
    }
    return SVGF_OK;
}
int svgf_oracle_remodulate(int W, int H, int storage, const void *albedo, const void *in, void *out) {
    if (W <= 0 || H <= 0 || !albedo || !in || !out) return SVGF_INVALID_ARG;
    for (size_t i = 0; i < (size_t)W * H; i++) {
        if (storage == SVGF_STORE_F32) {
            const V4 a = Colour<true>::raw(albedo, i), c = Colour<true>::raw(in, i);
            Colour<true>::store_raw(out, i, {c.x * a.x, c.y * a.y, c.z * a.z, c.w});
        } else {
            const V4 a = Colour<false>::raw(albedo, i), c = Colour<false>::raw(in, i);
            Colour<false>::store_raw(out, i, {c.x * a.x, c.y * a.y, c.z * a.z, c.w});
        }
    }
    return SVGF_OK;
}

// src/Filter.cuh:288-357, launched at src/App.cu:516-522.  `history` is the previous call's output (D13); must not alias `out`.
int svgf_oracle_taa(int W, int H, int storage, const void *filtered, const void *history, void *out) {
    if (W <= 0 || H <= 0 || !filtered || !history || !out || history == out || filtered == out) return SVGF_INVALID_ARG;
    if (storage == SVGF_STORE_F32) taa<true>(W, H, filtered, history, out);
    else if (storage == SVGF_STORE_F16) taa<false>(W, H, filtered, history, out);
    else return SVGF_INVALID_ARG;
    return SVGF_OK;
}

}  // extern "C"
