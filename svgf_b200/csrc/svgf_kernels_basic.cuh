// svgf_kernels_basic.cuh — one-thread-per-pixel sm_100a kernels for every stage of the path.  These are the
// simple, always-correct variants: used for resolutions / parameter combinations the tiled kernels do not
// cover and as the in-library A/B baseline.  The optimised kernels live in svgf_kernels_tiled.cuh.
#pragma once
#include "svgf_device.cuh"

namespace svgf {

struct GBufView {  // G-buffer planes (include/svgf.h svgf_gbuffer): pitch-linear memory (pitches in bytes) or texture objects
    const char *normal, *uv, *motion;
    size_t normal_pitch, uv_pitch, motion_pitch;
    int tex;   // bit 0 normal, 1 uv, 2 motion: the "pointer" is a cudaTextureObject_t (SVGF_PITCH_TEXTURE) over a cudaArray in
               // the reference's channel format, fetched like src/Filter.cuh:182-207: tex2D at integer texel coordinates of a
               // point-sampled, un-normalised texture (src/CudaUtil.h:88-95)
    __device__ __forceinline__ float4 mot(int x, int y) const {
        if (tex & 4) return tex2D<float4>((cudaTextureObject_t)motion, (float)x, (float)y);
        return __ldg(reinterpret_cast<const float4 *>(motion + (size_t)y * motion_pitch) + x);
    }
    __device__ __forceinline__ ushort4 nrm(int x, int y) const {
        if (tex & 1) return tex2D<ushort4>((cudaTextureObject_t)normal, (float)x, (float)y);
        return __ldg(reinterpret_cast<const ushort4 *>(normal + (size_t)y * normal_pitch) + x);
    }
    __device__ __forceinline__ ushort4 uvw(int x, int y) const {
        if (tex & 2) return tex2D<ushort4>((cudaTextureObject_t)uv, (float)x, (float)y);
        return __ldg(reinterpret_cast<const ushort4 *>(uv + (size_t)y * uv_pitch) + x);
    }
};

// Fused-frame outputs of the temporal pass (svgf_frame): besides RenderBuffer[P] it also produces the variance
// pass's output for every pixel that pass would merely copy (history >= 4, reference src/Filter.cuh:518-523) or
// zero (zero centre normal: all 49 weights are pow(0, phiN) = 0), and queues the remaining short-history pixels
// for the sparse 7x7 kernel.  var_out == nullptr disables all of it (stand-alone svgf_temporal).
template <bool F32> struct TemporalFused {
    typename ColourPlane<F32>::texel *var_out;
    unsigned int *worklist;
    unsigned int *counter;
    int zero_normal_shortcut;   // phi_normal > 0
};

struct TemporalArgs {
    int W, H;
    float depth_threshold, normal_threshold;
    int history_cap;
    float alpha_min, moments_alpha_min;
    int vacuous_mesh_id;   // svgf_mesh_id_mode == REFERENCE_VACUOUS
    int relative_depth;    // svgf_depth_test_mode == RELATIVE (src/Filter.cuh:241)
    int force_fail;        // first frame after svgf_reset: every reprojection fails (D12)
};

// ---- build the compact guide plane from a G-buffer (used when a stage is called on a G-buffer the temporal
// pass has not seen) ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) build_guide_kernel(GBufView g, Guide guide, int W, int H) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (y >= H) return;                                   // uniform per warp
    const bool inb = x < W;
    float4 gn = make_float4(kBackgroundZ, 0.f, 0.f, 0.f);
    if (inb) {
        const GuideTexel t = make_guide(g.mot(x, y), g.nrm(x, y), g.uvw(x, y));
        const size_t i = (size_t)y * W + x;
        guide.n[i] = t.n; guide.dz[i] = t.dz; guide.mid[i] = t.mid;
        gn = t.n;
    }
    const float4 sg = segment_state(gn, inb);
    if ((threadIdx.x & 31) == 0) guide.seg[(size_t)y * gridDim.x + blockIdx.x] = sg;
}

// ---- temporal reprojection + accumulation: reference filter::TemporalFilter (src/Filter.cuh:359-404) with
// LoadPreviousData (:225-258).  PREV_GUIDE: previous-frame consistency data comes from the compact guide
// plane cached by the previous frame instead of the three previous G-buffer planes.
// BILINEAR: SVGF_REPROJ_BILINEAR (include/svgf.h) - the 2x2 texels around coord + motion, each under the same consistency
// tests, weighted means of colour / moments / history length over the texels that pass; un-contracted FP32 in the
// oracle's order so that the fetched history length rounds identically.
template <bool F32, bool PREV_GUIDE, bool BILINEAR = false>
__global__ void __launch_bounds__(256)
temporal_kernel(TemporalArgs a, GBufView cur, GBufView prev, Guide prev_guide, Guide cur_guide, const typename ColourPlane<F32>::texel *__restrict__ prev_colour,
                typename ColourPlane<F32>::texel *colour, const uint8_t *__restrict__ hist_prev, uint8_t *__restrict__ hist_out,
                typename MomentsPlane<F32>::texel *__restrict__ cur_mom,
                const typename MomentsPlane<F32>::texel *__restrict__ prev_mom, TemporalFused<F32> fused) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const bool inb = (x < a.W && y < a.H);
    bool queue = false;
    const size_t i = (size_t)y * a.W + x;
    float4 seg_n = make_float4(kBackgroundZ, 0.f, 0.f, 0.f);
    if (inb) {

    const float4 mv = cur.mot(x, y);
    const GuideTexel g = make_guide(mv, cur.nrm(x, y), cur.uvw(x, y));
    cur_guide.n[i] = g.n; cur_guide.dz[i] = g.dz; cur_guide.mid[i] = g.mid;
    seg_n = g.n;
    const float4 c = clamp01(ColourPlane<F32>::decode(colour[i]));            // :370

    float3 pc = make_float3(0.f, 0.f, 0.f);
    float2 pm = make_float2(0.f, 0.f);
    int h = 1;
    bool ok = false;
    // the reference's consistency tests for one previous-frame texel (:235-252)
    auto consistent = [&](int qx, int qy) -> bool {
        if (qx < 0 || qx >= a.W || qy < 0 || qy >= a.H) return false;          // :235
        const size_t qi = (size_t)qy * a.W + qx;
        float4 pn;
        unsigned short pmid;
        if (PREV_GUIDE) {
            pn = __ldg(prev_guide.n + qi);
            pmid = __ldg(prev_guide.mid + qi);
        } else {
            const GuideTexel pg = make_guide(prev.mot(qx, qy), prev.nrm(qx, qy), prev.uvw(qx, qy));
            pn = pg.n; pmid = pg.mid;
        }
        const float dzv = fabsf(pn.x - g.n.x);
        bool good = a.relative_depth ? !(__fdiv_rn(dzv, __fadd_rn(g.dz, 1e-2f)) > a.depth_threshold)      // :241 (commented-out form)
                                     : !(dzv > a.depth_threshold);                                        // :242
        good = good && (a.vacuous_mesh_id || guide_mesh_id(g.mid) == guide_mesh_id(pmid));  // :245-247
        return good && !(dot3(guide_normal(g.n), guide_normal(pn)) < a.normal_threshold);  // :250-252
    };
    if (BILINEAR) {
        const float fxp = __fadd_rn((float)x, mv.x), fyp = __fadd_rn((float)y, mv.y);
        if (!a.force_fail && fabsf(fxp) < 1e8f && fabsf(fyp) < 1e8f) {         // false for NaN
            const float flx = floorf(fxp), fly = floorf(fyp);
            const float tx = __fsub_rn(fxp, flx), ty = __fsub_rn(fyp, fly);
            const int x0 = (int)flx, y0 = (int)fly;
            const float wx[2] = {__fsub_rn(1.0f, tx), tx}, wy[2] = {__fsub_rn(1.0f, ty), ty};
            float ws = 0.0f, hs = 0.0f;
            float3 cs = make_float3(0.f, 0.f, 0.f);
            float2 ms = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 2; j++)
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    const int qx = x0 + k, qy = y0 + j;
                    if (!consistent(qx, qy)) continue;
                    const float w = __fmul_rn(wx[k], wy[j]);
                    const size_t qi = (size_t)qy * a.W + qx;
                    const float4 p4 = clamp01(ColourPlane<F32>::decode(__ldg(prev_colour + qi)));
                    const float2 m2 = MomentsPlane<F32>::decode(__ldg(prev_mom + qi));
                    ws = __fadd_rn(ws, w);
                    cs.x = __fadd_rn(cs.x, __fmul_rn(w, p4.x)); cs.y = __fadd_rn(cs.y, __fmul_rn(w, p4.y));
                    cs.z = __fadd_rn(cs.z, __fmul_rn(w, p4.z));
                    ms.x = __fadd_rn(ms.x, __fmul_rn(w, m2.x)); ms.y = __fadd_rn(ms.y, __fmul_rn(w, m2.y));
                    hs = __fadd_rn(hs, __fmul_rn(w, (float)hist_prev[qi]));
                }
            if (ws >= 0.01f) {
                pc = make_float3(__fdiv_rn(cs.x, ws), __fdiv_rn(cs.y, ws), __fdiv_rn(cs.z, ws));
                pm = make_float2(__fdiv_rn(ms.x, ws), __fdiv_rn(ms.y, ws));
                h = (int)__fadd_rn(__fdiv_rn(hs, ws), 0.5f);
                ok = true;
            }
        }
    } else {
        const int qx = x + __float2int_rz(mv.x), qy = y + __float2int_rz(mv.y);    // :232
        if (!a.force_fail && consistent(qx, qy)) {
            const size_t qi = (size_t)qy * a.W + qx;
            const float4 p4 = clamp01(ColourPlane<F32>::decode(__ldg(prev_colour + qi)));  // :254
            pc = make_float3(p4.x, p4.y, p4.z);
            h = hist_prev[qi];                                                 // :255 (snapshot plane)
            pm = MomentsPlane<F32>::decode(__ldg(prev_mom + qi));              // :256
            ok = true;
        }
    }
    float alpha = 1.0f, alpha_m = 1.0f;
    if (ok) {
        h = min(a.history_cap, h + 1);                                         // :380
        const float inv = __fdiv_rn(1.0f, (float)h);                          // :381 (== float(1.0/h) for h <= 255, tests/test_oracle_kat.py)
        alpha = fmaxf(inv, a.alpha_min);
        alpha_m = fmaxf(inv, a.moments_alpha_min);
    } else {
        h = 1;
    }
    const float L = luminance(c.x, c.y, c.z);                                  // :391
    float2 m;
    m.x = mix_rn(pm.x, L, alpha_m);                                            // :393 glm::mix
    m.y = mix_rn(pm.y, __fmul_rn(L, L), alpha_m);
    const float var = fmaxf(0.0f, __fsub_rn(m.y, __fmul_rn(m.x, m.x)));        // :396
    float4 o;
    o.x = mix_rn(pc.x, c.x, alpha);                                            // :398
    o.y = mix_rn(pc.y, c.y, alpha);
    o.z = mix_rn(pc.z, c.z, alpha);
    o.w = var;
    hist_out[i] = (uint8_t)h;                                                  // :400
    const typename ColourPlane<F32>::texel enc = ColourPlane<F32>::encode(clamp01(o));
    colour[i] = enc;                                                           // :401
    cur_mom[i] = MomentsPlane<F32>::encode(m);                                 // :402
    if (fused.var_out) {
        const bool zero_n = fused.zero_normal_shortcut && g.n.y == 0.0f && g.n.z == 0.0f && g.n.w == 0.0f;
        queue = (h < 4) && !zero_n;
        fused.var_out[i] = (h < 4 && zero_n) ? ColourPlane<F32>::encode(make_float4(0.f, 0.f, 0.f, 0.f)) : enc;
    }
    }
    if (y < a.H) {         // this warp's 32-pixel row segment of the guide's segment map (uniform per warp)
        const float4 sg = segment_state(seg_n, inb);
        if ((threadIdx.x & 31) == 0) cur_guide.seg[(size_t)y * gridDim.x + blockIdx.x] = sg;
    }
    if (fused.var_out) {   // warp-aggregated append of the short-history pixels
        const unsigned int m = __ballot_sync(0xffffffffu, queue);
        if (m) {
            const int lane = threadIdx.x & 31;
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(fused.counter, (unsigned int)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (queue) fused.worklist[base + __popc(m & ((1u << lane) - 1u))] = (unsigned int)i;
        }
    }
}

struct SpatialArgs {
    int W, H;
    float phi_colour, phi_depth;
    NormalTerm nt;
    int step, level;
    const float *var_blur;   // a-trous: 3x3-blurred variance of the level's input (SVGF_VARIANCE_PREFILTER_GAUSS3) or nullptr
};

// ---- SVGF_VARIANCE_PREFILTER_GAUSS3 (include/svgf.h): the [0,1]-clamped variance channel of one plane blurred with
// (1 2 1; 2 4 2; 1 2 1) / 16, coordinates clamped to the image.  One warp produces 30 pixels of one row: every lane
// combines rows y-1, y, y+1 of its own column (lanes 0 and 31 carry the neighbour columns), the horizontal pass is two
// warp shuffles.  The weights are powers of two, so the only roundings are the four additions, in the oracle's order.
constexpr int kGauss3Rows = 4;   // output rows per thread: 6 loaded rows serve 4 outputs
template <bool F32>
__global__ void __launch_bounds__(256)
variance_gauss3_kernel(int W, int H, const typename ColourPlane<F32>::texel *__restrict__ in, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int y0 = (blockIdx.y * 8 + (threadIdx.x >> 5)) * kGauss3Rows;
    if (y0 >= H) return;                                  // uniform per warp
    const int x = blockIdx.x * 30 + lane - 1;
    const int px = min(max(x, 0), W - 1);
    // only the variance channel of a texel is read: 2 bytes at offset 6 (fp16 storage) / 4 bytes at offset 12 (fp32)
    float v[kGauss3Rows + 2];
#pragma unroll
    for (int r = 0; r < kGauss3Rows + 2; r++) {
        const int py = min(max(y0 + r - 1, 0), H - 1);
        const size_t qi = (size_t)py * W + px;
        float raw;
        if (F32) raw = __ldg(reinterpret_cast<const float *>(in) + 4 * qi + 3);
        else raw = __half2float(__ushort_as_half(__ldg(reinterpret_cast<const unsigned short *>(in) + 4 * qi + 3)));
        v[r] = clamp01(raw);
    }
#pragma unroll
    for (int r = 0; r < kGauss3Rows; r++) {
        const int y = y0 + r;
        const float col = __fadd_rn(__fadd_rn(0.25f * v[r], 0.5f * v[r + 1]), 0.25f * v[r + 2]);
        const float l = __shfl_up_sync(0xffffffffu, col, 1), rr = __shfl_down_sync(0xffffffffu, col, 1);
        if (lane >= 1 && lane <= 30 && x < W && y < H) out[(size_t)y * W + x] = __fadd_rn(__fadd_rn(0.25f * l, 0.5f * col), 0.25f * rr);
    }
}

// ---- albedo demodulation / remodulation (include/svgf.h; the paper's step the reference leaves out, README.md:172-174) ----
template <bool F32>
__global__ void __launch_bounds__(256)
demodulate_kernel(size_t n, const typename ColourPlane<F32>::texel *__restrict__ albedo, typename ColourPlane<F32>::texel *colour) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 a = ColourPlane<F32>::decode(__ldg(albedo + i)), c = ColourPlane<F32>::decode(colour[i]);
        colour[i] = ColourPlane<F32>::encode(make_float4(__fdiv_rn(c.x, fmaxf(a.x, 1e-3f)), __fdiv_rn(c.y, fmaxf(a.y, 1e-3f)),
                                                         __fdiv_rn(c.z, fmaxf(a.z, 1e-3f)), c.w));
    }
}
template <bool F32>
__global__ void __launch_bounds__(256)
remodulate_kernel(size_t n, const typename ColourPlane<F32>::texel *__restrict__ albedo, const typename ColourPlane<F32>::texel *__restrict__ in,
                  typename ColourPlane<F32>::texel *__restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 a = ColourPlane<F32>::decode(__ldg(albedo + i)), c = ColourPlane<F32>::decode(__ldg(in + i));
        out[i] = ColourPlane<F32>::encode(make_float4(__fmul_rn(c.x, a.x), __fmul_rn(c.y, a.y), __fmul_rn(c.z, a.z), c.w));
    }
}

// ---- variance estimation: reference filter::FilterMoments (src/Filter.cuh:430-525) -------------------------
// 7x7 cross-bilateral estimate for one short-history pixel (:446-516)
template <bool F32, bool SERIES>
__device__ __forceinline__ typename ColourPlane<F32>::texel
variance_7x7(const SpatialArgs &a, const Guide &guide, const typename ColourPlane<F32>::texel *__restrict__ in,
             const typename MomentsPlane<F32>::texel *__restrict__ mom, int x, int y, float h) {
    const size_t i = (size_t)y * a.W + x;
    const float4 cc = ColourPlane<F32>::decode(__ldg(in + i));                 // :450 (no clamp)
    const float lc = luminance(cc.x, cc.y, cc.z);
    const float4 gc = __ldg(guide.n + i);
    const float3 nc = guide_normal(gc);
    if (nc.x == 0.0f && nc.y == 0.0f && nc.z == 0.0f && a.nt.phiN > 0.0f) {
        // zero centre normal (background, D7): every weight is exp(..) * pow(0, phiN) = 0, so the sums are 0 and
        // the output is exactly (0, 0, 0, 0) — same result as running the 49 taps.
        return ColourPlane<F32>::encode(make_float4(0.f, 0.f, 0.f, 0.f));
    }
    const float kL = kLog2e / a.phi_colour;                                    // :460
    const float phiZ0 = fmaxf(__ldg(guide.dz + i), 1e-8f) * 3.0f * a.phi_depth;  // :461
    float sw = 0.f, sr = 0.f, sg = 0.f, sb = 0.f, sm1 = 0.f, sm2 = 0.f;
    // One window row at a time: the 21 loads of a row are issued together (clamped addresses, so they are unconditional)
    // and only then consumed.  In steady state this pass is a single wave of a few ten thousand threads and nothing but
    // load latency: 49 dependent round trips cost 37 us per frame at any resolution, 7 cost a fifth of that.  The taps are
    // still accumulated in the reference's order (rows, then columns), so the result is bit-identical to the serial loop.
    for (int yy = -3; yy <= 3; yy++) {
        const int py = y + yy;
        if (py < 0 || py >= a.H) continue;
        typename ColourPlane<F32>::texel rc[7];
        typename MomentsPlane<F32>::texel rm[7];
        float4 rg[7];
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const int px = min(max(x + k - 3, 0), a.W - 1);
            const size_t qi = (size_t)py * a.W + px;
            rc[k] = __ldg(in + qi); rm[k] = __ldg(mom + qi); rg[k] = __ldg(guide.n + qi);
        }
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const int xx = k - 3, px = x + xx;
            if (px < 0 || px >= a.W) continue;                                 // :473
            const float4 cq = ColourPlane<F32>::decode(rc[k]);
            const float2 mq = MomentsPlane<F32>::decode(rm[k]);
            const float4 gq = rg[k];
            const float lq = luminance(cq.x, cq.y, cq.z);
            const float phiZ = phiZ0 * sqrtf((float)(xx * xx + yy * yy));      // :488
            const float kZ = (phiZ == 0.0f) ? 0.0f : kLog2e / phiZ;            // :420
            const float e = edge_exponent<SERIES>(fmaf(fabsf(gc.x - gq.x), kZ, fabsf(lc - lq) * kL), dot3(nc, guide_normal(gq)), a.nt);
            const float w = fast_exp2(e);
            sw += w;                                                           // :497-499
            sr += cq.x * w; sg += cq.y * w; sb += cq.z * w;
            sm1 += mq.x * w; sm2 += mq.y * w;
        }
    }
    sw = fmaxf(sw, 1e-6f);                                                     // :505
    const float inv = 1.0f / sw;
    const float m1 = sm1 * inv, m2 = sm2 * inv;
    float var = m2 - m1 * m1;                                                  // :511
    var = var * (4.0f / h);                                                    // :514  (h in {0,1,2,3}; h == 0 only if the caller supplies it)
    return ColourPlane<F32>::encode(make_float4(sr * inv, sg * inv, sb * inv, var));  // :516
}

// Dense form (stand-alone svgf_variance).  Also publishes the history plane (hist_publish, may be null).
template <bool F32, bool SERIES>
__global__ void __launch_bounds__(256)
variance_kernel(SpatialArgs a, Guide guide, const typename ColourPlane<F32>::texel *__restrict__ in,
                const typename MomentsPlane<F32>::texel *__restrict__ mom, const uint8_t *__restrict__ hist,
                uint8_t *__restrict__ hist_publish, typename ColourPlane<F32>::texel *__restrict__ out) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= a.W || y >= a.H) return;
    const size_t i = (size_t)y * a.W + x;
    const uint8_t hb = hist[i];
    if (hist_publish) hist_publish[i] = hb;
    if (hb >= 4) {                                                             // :444,:518-523
        out[i] = in[i];
        return;
    }
    out[i] = variance_7x7<F32, SERIES>(a, guide, in, mom, x, y, (float)hb);
}

// Sparse form (svgf_frame): only the pixels the temporal pass queued (history < 4, non-zero normal); every other
// pixel of `out` was already written by the temporal pass.  Grid-stride over the worklist.
// Two chores ride along so that a frame needs no memset and no copy node between its kernels: the grid publishes this
// frame's history lengths (hist -> hist_publish, 16 bytes per thread and iteration; the caller's plane was last read by
// the temporal pass that precedes this launch in the stream) and zeroes the OTHER worklist counter, the one the next
// frame's temporal pass will append to.
template <bool F32, bool SERIES>
__global__ void __launch_bounds__(256)
variance_sparse_kernel(SpatialArgs a, Guide guide, const typename ColourPlane<F32>::texel *__restrict__ in,
                       const typename MomentsPlane<F32>::texel *__restrict__ mom, const uint8_t *__restrict__ hist,
                       const unsigned int *__restrict__ worklist, const unsigned int *__restrict__ counter,
                       typename ColourPlane<F32>::texel *__restrict__ out, uint8_t *__restrict__ hist_publish,
                       unsigned int *__restrict__ next_counter) {
    const unsigned int n = *counter;
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsize = (size_t)gridDim.x * blockDim.x;
    if (gtid == 0 && next_counter) *next_counter = 0;
    if (hist_publish) {    // both planes 16-byte aligned (checked by the host)
        const size_t bytes = (size_t)a.W * a.H, nv = bytes >> 4;
        const uint4 *src = reinterpret_cast<const uint4 *>(hist);
        uint4 *dst = reinterpret_cast<uint4 *>(hist_publish);
        for (size_t k = gtid; k < nv; k += gsize) dst[k] = src[k];
        if (gtid < (bytes & 15)) hist_publish[(nv << 4) + gtid] = hist[(nv << 4) + gtid];
    }
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const unsigned int i = worklist[k];
        const int y = (int)(i / (unsigned int)a.W), x = (int)(i - (unsigned int)y * (unsigned int)a.W);
        out[i] = variance_7x7<F32, SERIES>(a, guide, in, mom, x, y, (float)hist[i]);
    }
}

// ---- one a-trous level: reference filter::FilterKernel (src/Filter.cuh:527-624) ---------------------------
template <bool F32, bool SERIES>
__global__ void __launch_bounds__(256)
atrous_kernel(SpatialArgs a, Guide guide, const typename ColourPlane<F32>::texel *__restrict__ in,
              typename ColourPlane<F32>::texel *__restrict__ out, typename ColourPlane<F32>::texel *__restrict__ hist_colour) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= a.W || y >= a.H) return;
    const size_t i = (size_t)y * a.W + x;
    const float4 c = clamp01(ColourPlane<F32>::decode(__ldg(in + i)));         // :543
    const float4 gc = __ldg(guide.n + i);
    if (gc.x == kBackgroundZ) {                                                // :554-558
        out[i] = ColourPlane<F32>::encode(c);
        return;
    }
    const float lc = luminance(c.x, c.y, c.z);
    const float3 nc = guide_normal(gc);
    const float var = a.var_blur ? __ldg(a.var_blur + i) : c.w;
    const float phiL = a.phi_colour * sqrtf(fmaxf(0.0f, 1e-10f + var));        // :562
    const float kL = kLog2e / phiL;
    const float phiZ = fmaxf(__ldg(guide.dz + i), 1e-6f) * (float)a.step * a.phi_depth;  // :563
    const float kZ1 = (phiZ == 0.0f) ? 0.0f : kLog2e / phiZ;
    const float KW[3] = {1.0f, (float)(2.0 / 3.0), (float)(1.0 / 6.0)};        // :540
    float sw = 1.0f, sr = c.x, sg = c.y, sb = c.z, sv = c.w;                   // :567-568
#pragma unroll
    for (int yy = -2; yy <= 2; yy++) {
        const int py = y + yy * a.step;
        if (py < 0 || py >= a.H) continue;
#pragma unroll
        for (int xx = -2; xx <= 2; xx++) {
            const int px = x + xx * a.step;
            if (px < 0 || px >= a.W || (xx == 0 && yy == 0)) continue;         // :579,:584
            const size_t qi = (size_t)py * a.W + px;
            const float4 cq = clamp01(ColourPlane<F32>::decode(__ldg(in + qi)));  // :586
            const float4 gq = __ldg(guide.n + qi);
            const float lq = luminance(cq.x, cq.y, cq.z);
            const float kZ = kZ1 / sqrtf((float)(xx * xx + yy * yy));          // phiZ * length(xx,yy), :595
            const float e = edge_exponent<SERIES>(fmaf(fabsf(gc.x - gq.x), kZ, fabsf(lc - lq) * kL), dot3(nc, guide_normal(gq)), a.nt);
            const float w = fast_exp2(e) * (KW[xx < 0 ? -xx : xx] * KW[yy < 0 ? -yy : yy]);  // :604
            sw += w;                                                           // :607-608
            sr += w * cq.x; sg += w * cq.y; sb += w * cq.z;
            sv += (w * w) * cq.w;
        }
    }
    const float inv = 1.0f / sw;
    const typename ColourPlane<F32>::texel o =
        ColourPlane<F32>::encode(make_float4(sr * inv, sg * inv, sb * inv, sv * (inv * inv)));  // :615
    out[i] = o;                                                                // :618
    if (a.level == 0 && hist_colour) hist_colour[i] = o;                       // :619-622
}

}  // namespace svgf
