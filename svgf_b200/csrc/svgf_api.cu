// svgf_api.cu — the C ABI of include/svgf.h over the sm_100a kernels.  Host side only: argument validation,
// context-owned scratch (history shadow plane, compact guide planes), stage sequencing and buffer rotation.
// There is no CPU implementation of any stage in this library.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>

#include "../../include/svgf.h"
#include "svgf_ctx.h"
#include "svgf_kernels_basic.cuh"
#include "svgf_kernels_taa.cuh"
#include <cstdlib>

using namespace svgf;

namespace {

inline size_t colour_bytes(const svgf_ctx *c) { return (size_t)c->W * c->H * (c->storage == SVGF_STORE_F32 ? 16 : 8); }
inline size_t moments_bytes(const svgf_ctx *c) { return (size_t)c->W * c->H * (c->storage == SVGF_STORE_F32 ? 8 : 4); }

svgf_status check_params(const svgf_params *p) {
    if (!p) return SVGF_INVALID_ARG;
    if (p->history_cap < 1 || p->history_cap > 255) return SVGF_INVALID_ARG;  // D9: stored into uint8
    if (p->atrous_iterations < 0 || p->atrous_iterations > 10) return SVGF_INVALID_ARG;
    if (!(p->phi_colour > 0.0f) || !(p->phi_normal >= 0.0f) || !(p->phi_depth >= 0.0f)) return SVGF_INVALID_ARG;
    if (!(p->depth_threshold >= 0.0f)) return SVGF_INVALID_ARG;
    if (!(p->alpha_min >= 0.0f && p->alpha_min <= 1.0f) || !(p->moments_alpha_min >= 0.0f && p->moments_alpha_min <= 1.0f))
        return SVGF_INVALID_ARG;
    if (p->mesh_id_mode != SVGF_MESH_ID_INTENDED && p->mesh_id_mode != SVGF_MESH_ID_REFERENCE_VACUOUS) return SVGF_INVALID_ARG;
    if (p->depth_test_mode != SVGF_DEPTH_TEST_ABSOLUTE && p->depth_test_mode != SVGF_DEPTH_TEST_RELATIVE) return SVGF_INVALID_ARG;
    if (p->reproj_mode != SVGF_REPROJ_NEAREST_TRUNC && p->reproj_mode != SVGF_REPROJ_BILINEAR) return SVGF_UNSUPPORTED;
    if (p->variance_prefilter != SVGF_VARIANCE_PREFILTER_NONE && p->variance_prefilter != SVGF_VARIANCE_PREFILTER_GAUSS3)
        return SVGF_UNSUPPORTED;
    return SVGF_OK;
}

svgf_status check_gbuf(const svgf_ctx *c, const svgf_gbuffer *g) {
    if (!g || !g->normal_mat || !g->uv_inst || !g->motion_depth) return SVGF_INVALID_ARG;
    const size_t W = (size_t)c->W;
    // SVGF_PITCH_TEXTURE: the plane is a cudaTextureObject_t, nothing to check on the host
    if (g->normal_pitch != SVGF_PITCH_TEXTURE && ((g->normal_pitch && (g->normal_pitch < W * 8 || g->normal_pitch % 8)) || ((uintptr_t)g->normal_mat % 8)))
        return SVGF_INVALID_ARG;
    if (g->uv_pitch != SVGF_PITCH_TEXTURE && ((g->uv_pitch && (g->uv_pitch < W * 8 || g->uv_pitch % 8)) || ((uintptr_t)g->uv_inst % 8)))
        return SVGF_INVALID_ARG;
    if (g->motion_pitch != SVGF_PITCH_TEXTURE && ((g->motion_pitch && (g->motion_pitch < W * 16 || g->motion_pitch % 16)) || ((uintptr_t)g->motion_depth % 16)))
        return SVGF_INVALID_ARG;
    return SVGF_OK;
}

GBufView view(const svgf_ctx *c, const svgf_gbuffer *g) {
    GBufView v;
    v.normal = (const char *)g->normal_mat;
    v.uv = (const char *)g->uv_inst;
    v.motion = (const char *)g->motion_depth;
    v.normal_pitch = g->normal_pitch ? g->normal_pitch : (size_t)c->W * 8;
    v.uv_pitch = g->uv_pitch ? g->uv_pitch : (size_t)c->W * 8;
    v.motion_pitch = g->motion_pitch ? g->motion_pitch : (size_t)c->W * 16;
    v.tex = (g->normal_pitch == SVGF_PITCH_TEXTURE ? 1 : 0) | (g->uv_pitch == SVGF_PITCH_TEXTURE ? 2 : 0) | (g->motion_pitch == SVGF_PITCH_TEXTURE ? 4 : 0);
    return v;
}

inline dim3 grid_for(const svgf_ctx *c) { return dim3((c->W + 31) / 32, (c->H + 7) / 8); }

struct DeviceGuard {  // make the context's device current for the duration of a call
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
        if (prev == dev) prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Make sure a compact guide plane for `g` exists; returns its slot.  A cached plane is reused only when all three plane
// pointers and pitches match (svgf_ctx::GuideKey); contents changed in place under the same pointers need
// svgf_invalidate_guide.
svgf_status ensure_guide(svgf_ctx *c, const svgf_gbuffer *g, cudaStream_t s, int *slot) {
    for (int k = 0; k < 2; k++)
        if (c->guide_key[k].matches(g)) { *slot = k; return SVGF_OK; }
    const int k = 1 - c->guide_cur;
    build_guide_kernel<<<grid_for(c), 256, 0, s>>>(view(c, g), c->guide[k], c->W, c->H);
    c->launches++;
    SVGF_CUDA(c, cudaGetLastError());
    c->guide_key[k].set(g);
    c->guide_cur = k;
    *slot = k;
    return SVGF_OK;
}

template <bool F32>
svgf_status launch_temporal(svgf_ctx *c, const svgf_params *p, const svgf_gbuffer *cur, const svgf_gbuffer *prev,
                            const void *prev_colour, void *cur_colour, const uint8_t *hist_prev, uint8_t *hist_out,
                            void *cur_mom, const void *prev_mom, void *fused_var_out, cudaStream_t s) {
    using CT = typename ColourPlane<F32>::texel;
    using MT = typename MomentsPlane<F32>::texel;
    TemporalFused<F32> fused;
    fused.var_out = (CT *)fused_var_out;
    fused.worklist = c->worklist;
    fused.counter = c->work_counter + c->work_parity;   // zeroed by the previous frame's sparse variance pass (or svgf_create / svgf_reset)
    fused.zero_normal_shortcut = p->phi_normal > 0.0f;
    TemporalArgs a;
    a.W = c->W; a.H = c->H;
    a.depth_threshold = p->depth_threshold; a.normal_threshold = p->normal_threshold;
    a.history_cap = p->history_cap; a.alpha_min = p->alpha_min; a.moments_alpha_min = p->moments_alpha_min;
    a.vacuous_mesh_id = (p->mesh_id_mode == SVGF_MESH_ID_REFERENCE_VACUOUS);
    a.relative_depth = (p->depth_test_mode == SVGF_DEPTH_TEST_RELATIVE);
    a.force_fail = c->force_fail_next ? 1 : 0;
    // previous-frame guide: reuse the plane cached when `prev` was the current G-buffer
    int prev_slot = -1;
    if (!(p->flags & SVGF_FLAG_NO_GUIDE_CACHE))
        for (int k = 0; k < 2; k++)
            if (c->guide_key[k].matches(prev)) prev_slot = k;
    // current-frame guide goes into the other slot
    int cur_slot = (prev_slot >= 0) ? 1 - prev_slot : 1 - c->guide_cur;
    c->guide_key[cur_slot].clear();
    const dim3 grid = grid_for(c);
    const Guide pg = (prev_slot >= 0) ? c->guide[prev_slot] : Guide{nullptr, nullptr, nullptr, nullptr};
#define SVGF_LAUNCH_TEMPORAL(PG, BL)                                                                                              \
    temporal_kernel<F32, PG, BL><<<grid, 256, 0, s>>>(a, view(c, cur), view(c, prev), pg, c->guide[cur_slot], (const CT *)prev_colour, \
                                                      (CT *)cur_colour, hist_prev, hist_out, (MT *)cur_mom, (const MT *)prev_mom, fused)
    const bool bilinear = p->reproj_mode == SVGF_REPROJ_BILINEAR;
    if (prev_slot >= 0) { if (bilinear) SVGF_LAUNCH_TEMPORAL(true, true); else SVGF_LAUNCH_TEMPORAL(true, false); }
    else { if (bilinear) SVGF_LAUNCH_TEMPORAL(false, true); else SVGF_LAUNCH_TEMPORAL(false, false); }
#undef SVGF_LAUNCH_TEMPORAL
    c->launches++;
    SVGF_CUDA(c, cudaGetLastError());
    c->guide_key[cur_slot].set(cur);
    c->guide_cur = cur_slot;
    c->force_fail_next = false;
    return SVGF_OK;
}

SpatialArgs spatial_args(const svgf_ctx *c, const svgf_params *p, int level) {
    SpatialArgs a;
    a.W = c->W; a.H = c->H;
    a.phi_colour = p->phi_colour; a.phi_depth = p->phi_depth;
    a.nt = make_normal_term(p->phi_normal);
    a.step = 1 << level; a.level = level;
    a.var_blur = nullptr;
    return a;
}

template <bool F32>
svgf_status launch_variance(svgf_ctx *c, const svgf_params *p, int guide_slot, const void *in, const void *mom,
                            const uint8_t *hist, uint8_t *hist_publish, void *out, cudaStream_t s) {
    using CT = typename ColourPlane<F32>::texel;
    using MT = typename MomentsPlane<F32>::texel;
    const SpatialArgs a = spatial_args(c, p, 0);
    if (a.nt.series)
        variance_kernel<F32, true><<<grid_for(c), 256, 0, s>>>(a, c->guide[guide_slot], (const CT *)in, (const MT *)mom, hist,
                                                               hist_publish, (CT *)out);
    else
        variance_kernel<F32, false><<<grid_for(c), 256, 0, s>>>(a, c->guide[guide_slot], (const CT *)in, (const MT *)mom, hist,
                                                                hist_publish, (CT *)out);
    c->launches++;
    SVGF_CUDA(c, cudaGetLastError());
    return SVGF_OK;
}

// 7x7 estimate for the queued pixels; also publishes this frame's history lengths into the caller's plane and zeroes the
// next frame's worklist counter (see the kernel).
template <bool F32>
svgf_status launch_variance_sparse(svgf_ctx *c, const svgf_params *p, int guide_slot, const void *in, const void *mom,
                                   const uint8_t *hist, uint8_t *hist_publish, void *out, cudaStream_t s) {
    using CT = typename ColourPlane<F32>::texel;
    using MT = typename MomentsPlane<F32>::texel;
    const SpatialArgs a = spatial_args(c, p, 0);
    const int grid = c->num_sms * 8;
    const bool fold_publish = ((uintptr_t)hist_publish % 16) == 0;
    unsigned int *cur = c->work_counter + c->work_parity, *next = c->work_counter + (c->work_parity ^ 1);
    if (a.nt.series)
        variance_sparse_kernel<F32, true><<<grid, 256, 0, s>>>(a, c->guide[guide_slot], (const CT *)in, (const MT *)mom, hist,
                                                               c->worklist, cur, (CT *)out, fold_publish ? hist_publish : nullptr, next);
    else
        variance_sparse_kernel<F32, false><<<grid, 256, 0, s>>>(a, c->guide[guide_slot], (const CT *)in, (const MT *)mom, hist,
                                                                c->worklist, cur, (CT *)out, fold_publish ? hist_publish : nullptr, next);
    c->launches++;
    SVGF_CUDA(c, cudaGetLastError());
    c->work_parity ^= 1;
    if (!fold_publish) SVGF_CUDA(c, cudaMemcpyAsync(hist_publish, hist, (size_t)c->W * c->H, cudaMemcpyDeviceToDevice, s));
    return SVGF_OK;
}

// Levels 0 and 1 in one launch (svgf_kernels_fused.cuh).  Same preconditions as the packed single-level kernel.
template <bool F32>
bool fused01_applicable(const svgf_ctx *c, const svgf_params *p, const void *in, const void *out, const void *hist_colour) {
    const NormalTerm nt = make_normal_term(p->phi_normal);
    return (p->flags & SVGF_FLAG_FUSE_LEVELS_01) && !(p->flags & (SVGF_FLAG_BASIC_KERNELS | SVGF_FLAG_NO_LEVEL_FUSION)) && nt.series &&
           p->variance_prefilter == SVGF_VARIANCE_PREFILTER_NONE &&
           p->phi_depth > 0.0f && c->W % 2 == 0 && ((uintptr_t)in % 16) == 0 && ((uintptr_t)out % 16) == 0 &&
           (!hist_colour || ((uintptr_t)hist_colour % 16) == 0);
}
template <bool F32>
svgf_status launch_atrous_fused01(svgf_ctx *c, const svgf_params *p, int guide_slot, const void *in, void *out, void *hist_colour,
                                  cudaStream_t s) {
    const SpatialArgs a = spatial_args(c, p, 0);
    AtrousTiledArgs t;
    t.W = c->W; t.H = c->H; t.level = 0; t.tiles_x = t.tiles_y = 0;
    t.uniform_tiles = (p->flags & SVGF_FLAG_NO_UNIFORM_TILES) ? 0 : 1;
    t.var_blur = nullptr;
    t.yblock0 = 0; t.nyblocks = 0; t.yblock1 = 0; t.nyblocks1 = 0;
    t.kL_scale = kLog2e / p->phi_colour;
    t.kZ_scale = kLog2e / p->phi_depth;        // level 0; the kernel halves it for level 1
    t.k1 = a.nt.k1; t.k2 = a.nt.k2; t.k3 = a.nt.k3; t.k4 = a.nt.k4; t.k5 = a.nt.k5;
    if (p->phi_normal >= 100.0f) {             // same series selection as the single-level packed dispatch
        const NormalTerm e3 = economised_series3(p->phi_normal);
        t.k1 = e3.k1; t.k2 = e3.k2; t.k3 = e3.k3;
        return atrous_fused01(c, F32, 3, t, guide_slot, in, out, hist_colour, s);
    }
    return atrous_fused01(c, F32, 5, t, guide_slot, in, out, hist_colour, s);
}

// AtrousTiledArgs of one level for the tiled kernel families (packed / lattice / bulk / stream); returns the series
// terms the kernels are instantiated for: 3 = economised series (phi_normal >= 100, svgf_device.cuh economised_series3:
// weight error <= 0.56 / (phiN log2e)^3 <= 1.9e-7 absolute), 5 = Taylor (32 <= phi_normal < 100).
int tiled_args(const svgf_ctx *c, const svgf_params *p, int level, const float *var_blur, AtrousTiledArgs *t) {
    const NormalTerm nt = make_normal_term(p->phi_normal);
    t->W = c->W; t->H = c->H; t->level = level; t->tiles_x = t->tiles_y = 0;
    t->uniform_tiles = (p->flags & SVGF_FLAG_NO_UNIFORM_TILES) ? 0 : 1;
    t->var_blur = var_blur;
    t->yblock0 = 0; t->nyblocks = 0; t->yblock1 = 0; t->nyblocks1 = 0;
    t->kL_scale = kLog2e / p->phi_colour;
    t->kZ_scale = kLog2e / ((float)(1 << level) * p->phi_depth);
    t->k1 = nt.k1; t->k2 = nt.k2; t->k3 = nt.k3; t->k4 = nt.k4; t->k5 = nt.k5;
    if (p->phi_normal >= 100.0f) {
        const NormalTerm e3 = economised_series3(p->phi_normal);
        t->k1 = e3.k1; t->k2 = e3.k2; t->k3 = e3.k3;
        return 3;
    }
    return 5;
}

// Preconditions of the tiled fast paths: levels 0..4, series-mode normal term (phi_normal >= 32), phi_depth > 0 (null
// texels rely on |z - inf| * kZ = inf), 16-byte-aligned planes.  Everything else runs the per-pixel kernel - visibly:
// svgf_last_dispatch reports the family chosen for every level.
bool tiled_ok(const svgf_ctx *c, const svgf_params *p, int level) {
    return level <= 4 && p->phi_normal >= 32.0f && p->phi_depth > 0.0f && !(p->flags & SVGF_FLAG_BASIC_KERNELS);
}
bool pair_ok(const svgf_ctx *c, const void *in, const void *out, const void *hist_colour) {
    return c->W % 2 == 0 && ((uintptr_t)in % 16) == 0 && ((uintptr_t)out % 16) == 0 && (!hist_colour || ((uintptr_t)hist_colour % 16) == 0);
}

void note_dispatch(svgf_ctx *c, int level, int family) {
    if (level >= 0 && level < 16) {
        c->dispatch[level] = family;
        if (c->dispatch_n < level + 1) c->dispatch_n = level + 1;
    }
}

template <bool F32>
svgf_status launch_atrous_level(svgf_ctx *c, const svgf_params *p, int guide_slot, const void *in, void *out,
                                void *hist_colour, int level, cudaStream_t s) {
    using CT = typename ColourPlane<F32>::texel;
    SpatialArgs a = spatial_args(c, p, level);
    const bool prefilter = p->variance_prefilter == SVGF_VARIANCE_PREFILTER_GAUSS3;
    if (prefilter) {
        // blurred variance of this level's input, read by the level kernel for the centre pixel only
        if (!c->var_blur) SVGF_CUDA(c, cudaMalloc(&c->var_blur, (size_t)c->W * c->H * sizeof(float)));
        const dim3 g((c->W + 29) / 30, (c->H + 8 * kGauss3Rows - 1) / (8 * kGauss3Rows));
        variance_gauss3_kernel<F32><<<g, 256, 0, s>>>(c->W, c->H, (const CT *)in, c->var_blur);
        c->launches++;
        SVGF_CUDA(c, cudaGetLastError());
        a.var_blur = c->var_blur;
    }
    // (with the variance prefilter only the packed three-term kernel has a fast path: phi_normal >= 100)
    if (tiled_ok(c, p, level) && (F32 || c->W % 2 == 0) && ((uintptr_t)in % 16) == 0 && !(prefilter && p->phi_normal < 100.0f)) {
        AtrousTiledArgs t;
        const int terms = tiled_args(c, p, level, a.var_blur, &t);
        // A/B variants (measured slower): the warp-specialised streaming kernel and the persistent
        // bulk-copy (UBLKCP) scalar kernel; both take the Taylor series only
        const bool want_bulk = (p->flags & SVGF_FLAG_ATROUS_BULK) && !prefilter;
        const bool want_stream = (p->flags & SVGF_FLAG_ATROUS_STREAM) && !prefilter;
        if (want_bulk || want_stream) {
            const NormalTerm nt = make_normal_term(p->phi_normal);
            t.k1 = nt.k1; t.k2 = nt.k2; t.k3 = nt.k3; t.k4 = nt.k4; t.k5 = nt.k5;
        }
        if (want_stream && pair_ok(c, in, out, hist_colour)) {
            note_dispatch(c, level, kFamStream);
            return atrous_stream(c, F32, p->phi_normal >= 100.0f ? 4 : 5, t, guide_slot, in, out, hist_colour, s);
        }
        if (!want_bulk && pair_ok(c, in, out, hist_colour)) {
            note_dispatch(c, level, kFamPacked);
            return F32 ? atrous_packed_f32(c, terms, 3, t, guide_slot, in, out, hist_colour, s)
                       : atrous_packed_f16(c, terms, 3, t, guide_slot, in, out, hist_colour, s);
        }
        if (!prefilter) {
            const NormalTerm nt = make_normal_term(p->phi_normal);
            t.k1 = nt.k1; t.k2 = nt.k2; t.k3 = nt.k3; t.k4 = nt.k4; t.k5 = nt.k5;
            note_dispatch(c, level, kFamBulk);
            return atrous_tiled(c, F32, p->phi_normal >= 100.0f ? 4 : 5, t, guide_slot, in, out, hist_colour, s);
        }
        // prefilter with planes the packed kernel cannot take: the per-pixel kernel below
    }
    note_dispatch(c, level, kFamBasic);
    if (a.nt.series)
        atrous_kernel<F32, true><<<grid_for(c), 256, 0, s>>>(a, c->guide[guide_slot], (const CT *)in, (CT *)out, (CT *)hist_colour);
    else
        atrous_kernel<F32, false><<<grid_for(c), 256, 0, s>>>(a, c->guide[guide_slot], (const CT *)in, (CT *)out, (CT *)hist_colour);
    c->launches++;
    SVGF_CUDA(c, cudaGetLastError());
    return SVGF_OK;
}

// Staged run (the default for two or more consecutive levels): the first level is the packed kernel writing its result
// as lattice planes, every later level is the TMA-staged lattice kernel, and only the last one writes the storage
// format again.  Needs what the packed kernel needs, plus levels <= 4 throughout and no variance prefilter (its blur
// pass reads storage-format planes).
bool staged_applicable(const svgf_ctx *c, const svgf_params *p, const void *a, const void *b, const void *hist_colour, int first, int n) {
    if (n < 2 || first + n - 1 > 4 || !tiled_ok(c, p, first + n - 1)) return false;
    if (p->flags & (SVGF_FLAG_NO_STAGED_LEVELS | SVGF_FLAG_FUSE_LEVELS_01 | SVGF_FLAG_ATROUS_BULK | SVGF_FLAG_ATROUS_STREAM)) return false;
    if (p->variance_prefilter != SVGF_VARIANCE_PREFILTER_NONE) return false;
    return pair_ok(c, a, b, hist_colour) && !c->lat.failed;
}

// Number of buffer swaps run_atrous makes for levels first..first+n-1: one per launch (n, or n - 1 when levels 0 and 1
// go out as one fused launch).  svgf_frame needs it up front to aim the variance output so that the result lands in
// filter[0].  A staged run reports n: its result is written where n ping-pong hops would have left it.
int atrous_hops(const svgf_ctx *c, const svgf_params *p, const void *a, const void *b, const void *hist_colour, int first, int n) {
    const bool fuse = first == 0 && n >= 2 &&
                      (c->storage == SVGF_STORE_F32 ? fused01_applicable<true>(c, p, a, b, hist_colour)
                                                    : fused01_applicable<false>(c, p, a, b, hist_colour));
    return fuse ? n - 1 : n;
}

svgf_status run_atrous_staged(svgf_ctx *c, const svgf_params *p, int guide_slot, void *a, void *b, void *hist_colour, int first,
                              int n, void **result, cudaStream_t s) {
    const bool f32 = c->storage == SVGF_STORE_F32;
    void *final_dst = (n & 1) ? b : a;
    AtrousTiledArgs t;
    int terms = tiled_args(c, p, first, nullptr, &t);
    note_dispatch(c, first, kFamPackedStaged);
    svgf_status st = f32 ? atrous_packed_staged_f32(c, terms, t, guide_slot, a, 0, first == 0 ? hist_colour : nullptr, s)
                         : atrous_packed_staged_f16(c, terms, t, guide_slot, a, 0, first == 0 ? hist_colour : nullptr, s);
    if (st) return st;
    const bool pdl = !(p->flags & SVGF_FLAG_NO_DEPENDENT_LAUNCH);
    int src = 0;
    for (int l = first + 1; l < first + n; l++) {
        terms = tiled_args(c, p, l, nullptr, &t);
        void *out = (l == first + n - 1) ? final_dst : nullptr;
        note_dispatch(c, l, kFamLattice);
        st = f32 ? atrous_lattice_f32(c, terms, t, guide_slot, src, out, pdl, s) : atrous_lattice_f16(c, terms, t, guide_slot, src, out, pdl, s);
        if (st) return st;
        src = 1 - src;
    }
    if (result) *result = final_dst;
    return SVGF_OK;
}

// Levels first..first+n-1, ping-ponging a -> b -> a ...; *result = last output (or `a` when n == 0).
svgf_status run_atrous(svgf_ctx *c, const svgf_params *p, int guide_slot, void *a, void *b, void *hist_colour, int first,
                       int n, void **result, cudaStream_t s) {
    void *in = a, *out = b;
    const bool f32 = c->storage == SVGF_STORE_F32;
    c->dispatch_n = 0;
    if (staged_applicable(c, p, a, b, hist_colour, first, n)) {
        const svgf_status st = lattice_prepare(c, s);
        if (st == SVGF_OK) return run_atrous_staged(c, p, guide_slot, a, b, hist_colour, first, n, result, s);
        if (st != SVGF_UNSUPPORTED) return st;      // no tensor maps available: level by level below
    }
    if (first == 0 && n >= 2 && atrous_hops(c, p, a, b, hist_colour, first, n) == n - 1) {
        note_dispatch(c, 0, kFamFused01);
        note_dispatch(c, 1, kFamFused01);
        svgf_status st = f32 ? launch_atrous_fused01<true>(c, p, guide_slot, in, out, hist_colour, s)
                             : launch_atrous_fused01<false>(c, p, guide_slot, in, out, hist_colour, s);
        if (st) return st;
        void *t = in; in = out; out = t;
        first = 2; n -= 2;
    }
    for (int l = first; l < first + n; l++) {
        svgf_status st = (c->storage == SVGF_STORE_F32) ? launch_atrous_level<true>(c, p, guide_slot, in, out, hist_colour, l, s)
                                                        : launch_atrous_level<false>(c, p, guide_slot, in, out, hist_colour, l, s);
        if (st) return st;
        void *t = in; in = out; out = t;
    }
    if (result) *result = in;
    return SVGF_OK;
}

void prof_mark(svgf_ctx *c, int k, cudaStream_t s) {
    if (c->profiling && c->prof_frames < svgf_ctx::kMaxProf) cudaEventRecord(c->prof_ev[c->prof_frames * 4 + k], s);
}

}  // namespace

namespace svgf {
// One level of a staged run restricted to row blocks [yb0, yb0 + nyb) (+ [yb1, yb1 + nyb1)) of the level's tile grid (nyb == 0: the whole image);
// used by the band driver (svgf_band.cu), which runs the boundary rows of a level first.  kind 0: first level - packed
// kernel, storage-format `in` -> lattice planes sc[idx];  kind 1: lattice sc[idx] -> sc[1 - idx];  kind 2: lattice sc[idx]
// -> storage-format `out`.
svgf_status staged_level(svgf_ctx *c, const svgf_params *p, int guide_slot, int level, int kind, const void *in, int idx, void *out,
                         void *hist_colour, int yb0, int nyb, int yb1, int nyb1, bool pdl, cudaStream_t s) {
    const bool f32 = c->storage == SVGF_STORE_F32;
    AtrousTiledArgs t;
    const int terms = tiled_args(c, p, level, nullptr, &t);
    t.yblock0 = yb0; t.nyblocks = nyb; t.yblock1 = yb1; t.nyblocks1 = nyb1;     // the second range is honoured by the lattice kernel only
    if (kind == 0) {
        note_dispatch(c, level, kFamPackedStaged);
        return f32 ? atrous_packed_staged_f32(c, terms, t, guide_slot, in, idx, level == 0 ? hist_colour : nullptr, s)
                   : atrous_packed_staged_f16(c, terms, t, guide_slot, in, idx, level == 0 ? hist_colour : nullptr, s);
    }
    note_dispatch(c, level, kFamLattice);
    return f32 ? atrous_lattice_f32(c, terms, t, guide_slot, idx, kind == 2 ? out : nullptr, pdl, s)
               : atrous_lattice_f16(c, terms, t, guide_slot, idx, kind == 2 ? out : nullptr, pdl, s);
}
bool staged_run_possible(const svgf_ctx *c, const svgf_params *p, const void *a, const void *b, const void *hist_colour, int first, int n) {
    return staged_applicable(c, p, a, b, hist_colour, first, n);
}
}  // namespace svgf

extern "C" {

int svgf_abi_version(void) { return SVGF_ABI_VERSION; }

const char *svgf_status_string(svgf_status s) {
    switch (s) {
        case SVGF_OK: return "SVGF_OK";
        case SVGF_INVALID_ARG: return "SVGF_INVALID_ARG";
        case SVGF_UNSUPPORTED: return "SVGF_UNSUPPORTED";
        case SVGF_CUDA_ERROR: return "SVGF_CUDA_ERROR";
    }
    return "SVGF_?";
}

void svgf_default_params(svgf_params *p) {
    if (!p) return;
    std::memset(p, 0, sizeof(*p));
    p->history_cap = 24;          // reference src/App.h:112
    p->depth_threshold = 0.8f;    // :110
    p->normal_threshold = 0.9f;   // :111
    p->phi_colour = 10.0f;        // :113
    p->phi_normal = 128.0f;       // :114
    p->atrous_iterations = 5;     // BASELINE.json (reference default 3, :109)
    p->phi_depth = 1.0f;
    p->alpha_min = 0.0f;
    p->moments_alpha_min = 0.0f;
    p->mesh_id_mode = SVGF_MESH_ID_INTENDED;
    p->reproj_mode = SVGF_REPROJ_NEAREST_TRUNC;
    p->variance_prefilter = SVGF_VARIANCE_PREFILTER_NONE;
    p->flags = SVGF_FLAG_NONE;
    p->depth_test_mode = SVGF_DEPTH_TEST_ABSOLUTE;
}

svgf_status svgf_create(svgf_ctx **out, int device, int width, int height, svgf_storage storage) {
    if (!out || width <= 0 || height <= 0 || width > 32768 || height > 32768) return SVGF_INVALID_ARG;
    if (storage != SVGF_STORE_F16 && storage != SVGF_STORE_F32) return SVGF_INVALID_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return SVGF_CUDA_ERROR;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SVGF_CUDA_ERROR;
    if (prop.major != 10) return SVGF_UNSUPPORTED;  // sm_100a only: no other code path exists in this library
    DeviceGuard guard(device);
    if (!guard.ok) return SVGF_CUDA_ERROR;
    svgf_ctx *c = new (std::nothrow) svgf_ctx();
    if (!c) return SVGF_CUDA_ERROR;
    c->device = device; c->W = width; c->H = height; c->storage = storage;
    const size_t n = (size_t)width * height;
    cudaError_t e = cudaMalloc(&c->hist_shadow, n);
    for (int k = 0; k < 2 && e == cudaSuccess; k++) {
        e = cudaMalloc(&c->guide[k].n, n * sizeof(float4));
        if (e == cudaSuccess) e = cudaMalloc(&c->guide[k].dz, n * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&c->guide[k].mid, n * sizeof(unsigned short));
        if (e == cudaSuccess) e = cudaMalloc(&c->guide[k].seg, (size_t)((width + 31) / 32) * height * sizeof(float4));
    }
    c->num_sms = prop.multiProcessorCount;
    if (e == cudaSuccess) e = cudaMalloc(&c->worklist, n * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMalloc(&c->work_counter, 2 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(c->work_counter, 0, 2 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(c->hist_shadow, 0, n);
    if (e != cudaSuccess) { svgf_destroy(c); return SVGF_CUDA_ERROR; }
    *out = c;
    return SVGF_OK;
}

void svgf_destroy(svgf_ctx *c) {
    if (!c) return;
    DeviceGuard guard(c->device);
    cudaFree(c->hist_shadow);
    cudaFree(c->worklist);
    cudaFree(c->work_counter);
    cudaFree(c->var_blur);
    for (int k = 0; k < 2; k++) { cudaFree(c->guide[k].n); cudaFree(c->guide[k].dz); cudaFree(c->guide[k].mid); cudaFree(c->guide[k].seg); }
    lattice_destroy(c);
    if (c->prof_ev) {
        for (int i = 0; i < svgf_ctx::kMaxProf * 4; i++) cudaEventDestroy(c->prof_ev[i]);
        delete[] c->prof_ev;
    }
    if (c->hp.s_in) { cudaStreamSynchronize(c->hp.s_in); cudaStreamDestroy(c->hp.s_in); }
    if (c->hp.s_out) { cudaStreamSynchronize(c->hp.s_out); cudaStreamDestroy(c->hp.s_out); }
    for (int k = 0; k < svgf_ctx::HostPath::kRing; k++) {
        cudaFree(c->hp.normal[k]); cudaFree(c->hp.uv[k]); cudaFree(c->hp.motion[k]); cudaFree(c->hp.noisy[k]);
        if (c->hp.ev_in[k]) cudaEventDestroy(c->hp.ev_in[k]);
        if (c->hp.ev_done[k]) cudaEventDestroy(c->hp.ev_done[k]);
    }
    if (c->hp.ev_out) cudaEventDestroy(c->hp.ev_out);
    for (int k = 0; k < 2; k++) { cudaFree(c->hp.moments[k]); cudaFree(c->hp.filter[k]); }
    cudaFree(c->hp.history);
    delete c;
}

int svgf_last_cuda_error(const svgf_ctx *c) { return c ? c->last_err : 0; }
uint64_t svgf_launch_count(const svgf_ctx *c) { return c ? c->launches : 0; }

void svgf_invalidate_guide(svgf_ctx *c) {
    if (!c) return;
    c->guide_key[0].clear();
    c->guide_key[1].clear();
}

int svgf_last_dispatch(const svgf_ctx *c, int32_t *families, int max_levels) {
    if (!c) return 0;
    for (int i = 0; i < c->dispatch_n && i < max_levels; i++)
        if (families) families[i] = c->dispatch[i];
    return c->dispatch_n;
}

svgf_status svgf_reset(svgf_ctx *c, const svgf_frame_buffers *b, void *stream) {
    if (!c) return SVGF_INVALID_ARG;
    DeviceGuard guard(c->device);
    if (!guard.ok) return SVGF_CUDA_ERROR;
    cudaStream_t s = (cudaStream_t)stream;
    if (b) {
        for (int k = 0; k < 2; k++) {
            if (b->render[k]) SVGF_CUDA(c, cudaMemsetAsync(b->render[k], 0, colour_bytes(c), s));
            if (b->moments[k]) SVGF_CUDA(c, cudaMemsetAsync(b->moments[k], 0, moments_bytes(c), s));
        }
        if (b->history) SVGF_CUDA(c, cudaMemsetAsync(b->history, 0, (size_t)c->W * c->H, s));
    }
    SVGF_CUDA(c, cudaMemsetAsync(c->hist_shadow, 0, (size_t)c->W * c->H, s));
    SVGF_CUDA(c, cudaMemsetAsync(c->work_counter, 0, 2 * sizeof(unsigned int), s));
    svgf_invalidate_guide(c);
    c->force_fail_next = true;
    return SVGF_OK;
}

svgf_status svgf_temporal(svgf_ctx *c, const svgf_params *p, const svgf_gbuffer *cur, const svgf_gbuffer *prev,
                          const void *prev_colour, void *cur_colour, uint8_t *history, void *cur_moments,
                          const void *prev_moments, void *stream) {
    if (!c) return SVGF_INVALID_ARG;
    svgf_status st = check_params(p);
    if (st) return st;
    if ((st = check_gbuf(c, cur)) || (st = check_gbuf(c, prev))) return st;
    if (!prev_colour || !cur_colour || !history || !cur_moments || !prev_moments || prev_colour == cur_colour ||
        prev_moments == cur_moments)
        return SVGF_INVALID_ARG;
    DeviceGuard guard(c->device);
    if (!guard.ok) return SVGF_CUDA_ERROR;
    cudaStream_t s = (cudaStream_t)stream;
    st = (c->storage == SVGF_STORE_F32)
             ? launch_temporal<true>(c, p, cur, prev, prev_colour, cur_colour, history, c->hist_shadow, cur_moments, prev_moments, nullptr, s)
             : launch_temporal<false>(c, p, cur, prev, prev_colour, cur_colour, history, c->hist_shadow, cur_moments, prev_moments, nullptr, s);
    if (st) return st;
    // publish: the caller-visible plane now holds this frame's lengths
    SVGF_CUDA(c, cudaMemcpyAsync(history, c->hist_shadow, (size_t)c->W * c->H, cudaMemcpyDeviceToDevice, s));
    return SVGF_OK;
}

svgf_status svgf_variance(svgf_ctx *c, const svgf_params *p, const svgf_gbuffer *cur, const void *colour_in,
                          const void *moments, const uint8_t *history, void *colour_out, void *stream) {
    if (!c) return SVGF_INVALID_ARG;
    svgf_status st = check_params(p);
    if (st) return st;
    if ((st = check_gbuf(c, cur))) return st;
    if (!colour_in || !moments || !history || !colour_out || colour_in == colour_out) return SVGF_INVALID_ARG;
    DeviceGuard guard(c->device);
    if (!guard.ok) return SVGF_CUDA_ERROR;
    cudaStream_t s = (cudaStream_t)stream;
    int slot;
    if ((st = ensure_guide(c, cur, s, &slot))) return st;
    return (c->storage == SVGF_STORE_F32) ? launch_variance<true>(c, p, slot, colour_in, moments, history, nullptr, colour_out, s)
                                          : launch_variance<false>(c, p, slot, colour_in, moments, history, nullptr, colour_out, s);
}

svgf_status svgf_atrous(svgf_ctx *c, const svgf_params *p, const svgf_gbuffer *cur, void *buf_a, void *buf_b,
                        void *history_colour_out, int first_level, int num_levels, void **result, void *stream) {
    if (!c) return SVGF_INVALID_ARG;
    svgf_status st = check_params(p);
    if (st) return st;
    if ((st = check_gbuf(c, cur))) return st;
    if (!buf_a || !buf_b || buf_a == buf_b || first_level < 0 || num_levels < 0 || first_level + num_levels > 11)
        return SVGF_INVALID_ARG;
    if (history_colour_out == buf_a || history_colour_out == buf_b) return SVGF_INVALID_ARG;
    DeviceGuard guard(c->device);
    if (!guard.ok) return SVGF_CUDA_ERROR;
    cudaStream_t s = (cudaStream_t)stream;
    int slot;
    if ((st = ensure_guide(c, cur, s, &slot))) return st;
    return run_atrous(c, p, slot, buf_a, buf_b, history_colour_out, first_level, num_levels, result, s);
}

svgf_status svgf_frame(svgf_ctx *c, const svgf_params *p, const svgf_gbuffer gbuf[2], const svgf_frame_buffers *b,
                       void *stream) {
    if (!c || !gbuf || !b) return SVGF_INVALID_ARG;
    svgf_status st = check_params(p);
    if (st) return st;
    if (b->ping_pong != 0 && b->ping_pong != 1) return SVGF_INVALID_ARG;
    const int P = b->ping_pong, Q = 1 - P;
    if ((st = check_gbuf(c, &gbuf[P])) || (st = check_gbuf(c, &gbuf[Q]))) return st;
    for (int k = 0; k < 2; k++)
        if (!b->render[k] || !b->moments[k] || !b->filter[k]) return SVGF_INVALID_ARG;
    if (!b->history || b->render[0] == b->render[1] || b->moments[0] == b->moments[1] || b->filter[0] == b->filter[1])
        return SVGF_INVALID_ARG;
    DeviceGuard guard(c->device);
    if (!guard.ok) return SVGF_CUDA_ERROR;
    cudaStream_t s = (cudaStream_t)stream;
    const bool f32 = (c->storage == SVGF_STORE_F32);

    const int N = p->atrous_iterations;
    const int hops = atrous_hops(c, p, b->filter[0], b->filter[1], b->render[P], 0, N);
    void *v_out = b->filter[hops & 1], *other = b->filter[1 - (hops & 1)];
    const bool basic = (p->flags & SVGF_FLAG_BASIC_KERNELS) != 0;
    prof_mark(c, 0, s);
    // temporal: reads the caller-visible history plane (previous frame), writes the shadow plane.  Fused form: it
    // also writes the variance pass's output for every pixel that pass only copies or zeroes and queues the
    // short-history pixels; -> filter[hops & 1] so that the ping-pong launches end in filter[0] (the reference copies
    // instead, src/App.cu:510-513).
    st = f32 ? launch_temporal<true>(c, p, &gbuf[P], &gbuf[Q], b->render[Q], b->render[P], b->history, c->hist_shadow,
                                     b->moments[P], b->moments[Q], basic ? nullptr : v_out, s)
             : launch_temporal<false>(c, p, &gbuf[P], &gbuf[Q], b->render[Q], b->render[P], b->history, c->hist_shadow,
                                      b->moments[P], b->moments[Q], basic ? nullptr : v_out, s);
    if (st) return st;
    prof_mark(c, 1, s);
    const int slot = c->guide_cur;
    if (basic) {
        st = f32 ? launch_variance<true>(c, p, slot, b->render[P], b->moments[P], c->hist_shadow, b->history, v_out, s)
                 : launch_variance<false>(c, p, slot, b->render[P], b->moments[P], c->hist_shadow, b->history, v_out, s);
        if (st) return st;
    } else {
        // 7x7 estimate for the queued pixels only, then publish this frame's history lengths
        st = f32 ? launch_variance_sparse<true>(c, p, slot, b->render[P], b->moments[P], c->hist_shadow, b->history, v_out, s)
                 : launch_variance_sparse<false>(c, p, slot, b->render[P], b->moments[P], c->hist_shadow, b->history, v_out, s);
        if (st) return st;
    }
    prof_mark(c, 2, s);
    void *result = nullptr;
    st = run_atrous(c, p, slot, v_out, other, b->render[P], 0, N, &result, s);
    if (st) return st;
    prof_mark(c, 3, s);
    if (c->profiling && c->prof_frames < svgf_ctx::kMaxProf) c->prof_frames++;
    return (result == b->filter[0]) ? SVGF_OK : SVGF_CUDA_ERROR;  // invariant of the rotation above
}

svgf_status svgf_taa(svgf_ctx *c, const void *filtered, const void *taa_history, void *taa_out, void *stream) {
    if (!c || !filtered || !taa_history || !taa_out || taa_history == taa_out || filtered == taa_out) return SVGF_INVALID_ARG;
    DeviceGuard guard(c->device);
    if (!guard.ok) return SVGF_CUDA_ERROR;
    cudaStream_t s = (cudaStream_t)stream;
    const dim3 grid((c->W + kTaaBW - 1) / kTaaBW, (c->H + kTaaBH - 1) / kTaaBH);
    if (c->storage == SVGF_STORE_F32)
        taa_kernel<true><<<grid, kTaaBW * kTaaBH, 0, s>>>(c->W, c->H, (const float4 *)filtered, (const float4 *)taa_history, (float4 *)taa_out);
    else
        taa_kernel<false><<<grid, kTaaBW * kTaaBH, 0, s>>>(c->W, c->H, (const uint2 *)filtered, (const uint2 *)taa_history, (uint2 *)taa_out);
    c->launches++;
    SVGF_CUDA(c, cudaGetLastError());
    return SVGF_OK;
}

svgf_status svgf_demodulate(svgf_ctx *c, const void *albedo, void *colour, void *stream) {
    if (!c || !albedo || !colour || albedo == colour) return SVGF_INVALID_ARG;
    DeviceGuard guard(c->device);
    if (!guard.ok) return SVGF_CUDA_ERROR;
    const size_t n = (size_t)c->W * c->H;
    if (c->storage == SVGF_STORE_F32) demodulate_kernel<true><<<c->num_sms * 8, 256, 0, (cudaStream_t)stream>>>(n, (const float4 *)albedo, (float4 *)colour);
    else demodulate_kernel<false><<<c->num_sms * 8, 256, 0, (cudaStream_t)stream>>>(n, (const uint2 *)albedo, (uint2 *)colour);
    c->launches++;
    SVGF_CUDA(c, cudaGetLastError());
    return SVGF_OK;
}

svgf_status svgf_remodulate(svgf_ctx *c, const void *albedo, const void *filtered, void *out, void *stream) {
    if (!c || !albedo || !filtered || !out || albedo == out) return SVGF_INVALID_ARG;
    DeviceGuard guard(c->device);
    if (!guard.ok) return SVGF_CUDA_ERROR;
    const size_t n = (size_t)c->W * c->H;
    if (c->storage == SVGF_STORE_F32)
        remodulate_kernel<true><<<c->num_sms * 8, 256, 0, (cudaStream_t)stream>>>(n, (const float4 *)albedo, (const float4 *)filtered, (float4 *)out);
    else
        remodulate_kernel<false><<<c->num_sms * 8, 256, 0, (cudaStream_t)stream>>>(n, (const uint2 *)albedo, (const uint2 *)filtered, (uint2 *)out);
    c->launches++;
    SVGF_CUDA(c, cudaGetLastError());
    return SVGF_OK;
}

// ---- stage profiling -----------------------------------------------------------------------------------------
// Between svgf_profile_begin and svgf_profile_end every svgf_frame records four CUDA events on the caller's
// stream (before temporal, before variance, before a-trous, after a-trous).  svgf_profile_end synchronises
// the last event and returns the summed GPU milliseconds per stage and the number of frames captured.
svgf_status svgf_profile_begin(svgf_ctx *c) {
    if (!c) return SVGF_INVALID_ARG;
    DeviceGuard guard(c->device);
    if (!c->prof_ev) {
        c->prof_ev = new (std::nothrow) cudaEvent_t[svgf_ctx::kMaxProf * 4];
        if (!c->prof_ev) return SVGF_CUDA_ERROR;
        for (int i = 0; i < svgf_ctx::kMaxProf * 4; i++) SVGF_CUDA(c, cudaEventCreate(&c->prof_ev[i]));
    }
    c->prof_frames = 0;
    c->profiling = true;
    return SVGF_OK;
}

svgf_status svgf_profile_end(svgf_ctx *c, double stage_ms[3], int *frames) {
    if (!c || !c->profiling) return SVGF_INVALID_ARG;
    DeviceGuard guard(c->device);
    c->profiling = false;
    double acc[3] = {0, 0, 0};
    if (c->prof_frames > 0) SVGF_CUDA(c, cudaEventSynchronize(c->prof_ev[(c->prof_frames - 1) * 4 + 3]));
    for (int f = 0; f < c->prof_frames; f++)
        for (int k = 0; k < 3; k++) {
            float ms = 0.f;
            SVGF_CUDA(c, cudaEventElapsedTime(&ms, c->prof_ev[f * 4 + k], c->prof_ev[f * 4 + k + 1]));
            acc[k] += ms;
        }
    if (stage_ms) for (int k = 0; k < 3; k++) stage_ms[k] = acc[k];
    if (frames) *frames = c->prof_frames;
    return SVGF_OK;
}

// ---- host-buffer path ----------------------------------------------------------------------------------------
// Frame t:   copy-in stream : wait(kernels of frame t-2 done)  H2D inputs -> ring slot t%3          record ev_in[t%3]
//            caller's stream: wait(ev_in[t%3])  noisy -> RenderBuffer[P]  svgf_frame              record ev_done[t%3]
//                             wait(ev_out)                       (so a sync of the caller's stream covers the result)
//            copy-out stream: wait(ev_done[t%3])  D2H result (and history)                         record ev_out
// Slot t%3 was the *previous* G-buffer of frame t-2 (its last reader), hence the first wait.  The kernels of frame
// t+1 are ordered after the D2H of frame t by the caller's stream itself, so FilterBuffer[0] is never overwritten
// under the copy.  Nothing here waits for work the caller queued earlier on `stream`: the inputs are host memory.
svgf_status svgf_frame_host(svgf_ctx *c, const svgf_params *p, const void *h_normal, const void *h_uv, const void *h_motion,
                            const void *h_colour, void *h_result, uint8_t *h_history_out, int reset, void *stream) {
    if (!c || !h_normal || !h_uv || !h_motion || !h_colour) return SVGF_INVALID_ARG;
    svgf_status st = check_params(p);
    if (st) return st;
    DeviceGuard guard(c->device);
    if (!guard.ok) return SVGF_CUDA_ERROR;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)c->W * c->H;
    svgf_ctx::HostPath &hp = c->hp;
    constexpr int R = svgf_ctx::HostPath::kRing;
    if (!hp.ready) {
        // a retry after a failed allocation keeps what is already there (every pointer / handle starts out null)
        auto dmalloc = [&](void **p, size_t bytes) { return *p ? cudaSuccess : cudaMalloc(p, bytes); };
        auto mkevent = [&](cudaEvent_t *e) { return *e ? cudaSuccess : cudaEventCreateWithFlags(e, cudaEventDisableTiming); };
        for (int k = 0; k < R; k++) {
            SVGF_CUDA(c, dmalloc(&hp.normal[k], n * 8));
            SVGF_CUDA(c, dmalloc(&hp.uv[k], n * 8));
            SVGF_CUDA(c, dmalloc(&hp.motion[k], n * 16));
            SVGF_CUDA(c, dmalloc(&hp.noisy[k], colour_bytes(c)));
            SVGF_CUDA(c, mkevent(&hp.ev_in[k]));
            SVGF_CUDA(c, mkevent(&hp.ev_done[k]));
        }
        for (int k = 0; k < 2; k++) {
            SVGF_CUDA(c, dmalloc(&hp.moments[k], moments_bytes(c)));
            SVGF_CUDA(c, dmalloc(&hp.filter[k], colour_bytes(c)));
        }
        SVGF_CUDA(c, dmalloc((void **)&hp.history, n));
        SVGF_CUDA(c, mkevent(&hp.ev_out));
        if (!hp.s_in) SVGF_CUDA(c, cudaStreamCreateWithFlags(&hp.s_in, cudaStreamNonBlocking));
        if (!hp.s_out) SVGF_CUDA(c, cudaStreamCreateWithFlags(&hp.s_out, cudaStreamNonBlocking));
        hp.frame = 0;
        hp.ready = true;
        reset = 1;
    }
    svgf_frame_buffers b;
    // RenderBuffer[P] IS the ring slot the noisy radiance was uploaded into (the temporal pass works in place and level 0
    // leaves the next frame's colour history there); RenderBuffer[1 - P] is the previous frame's slot.  No device copy.
    for (int k = 0; k < 2; k++) { b.moments[k] = hp.moments[k]; b.filter[k] = hp.filter[k]; }
    b.history = hp.history;
    if (reset) {
        if (hp.frame) {   // a sequence is in flight: drain it before its ring slots are reused from frame 0
            SVGF_CUDA(c, cudaStreamSynchronize(hp.s_in));
            SVGF_CUDA(c, cudaStreamSynchronize(s));
            SVGF_CUDA(c, cudaStreamSynchronize(hp.s_out));
        }
        hp.frame = 0;
        hp.ping_pong = 0;
        b.render[0] = nullptr;            // about to be overwritten by the upload
        b.render[1] = hp.noisy[R - 1];    // "previous colour" of the first frame (slot of frame -1)
        if ((st = svgf_reset(c, &b, s))) return st;
    }
    const uint64_t t = hp.frame;
    const int k = (int)(t % R), kprev = (int)((t + R - 1) % R);
    const int P = hp.ping_pong;
    b.ping_pong = P;

    // copy-in
    if (t >= 2) SVGF_CUDA(c, cudaStreamWaitEvent(hp.s_in, hp.ev_done[(t - 2) % R], 0));
    SVGF_CUDA(c, cudaMemcpyAsync(hp.normal[k], h_normal, n * 8, cudaMemcpyHostToDevice, hp.s_in));
    SVGF_CUDA(c, cudaMemcpyAsync(hp.uv[k], h_uv, n * 8, cudaMemcpyHostToDevice, hp.s_in));
    SVGF_CUDA(c, cudaMemcpyAsync(hp.motion[k], h_motion, n * 16, cudaMemcpyHostToDevice, hp.s_in));
    SVGF_CUDA(c, cudaMemcpyAsync(hp.noisy[k], h_colour, colour_bytes(c), cudaMemcpyHostToDevice, hp.s_in));
    SVGF_CUDA(c, cudaEventRecord(hp.ev_in[k], hp.s_in));

    // kernels, on the caller's stream
    SVGF_CUDA(c, cudaStreamWaitEvent(s, hp.ev_in[k], 0));
    b.render[P] = hp.noisy[k];
    b.render[1 - P] = hp.noisy[kprev];
    svgf_gbuffer g[2];
    const int slot_of[2] = {P == 0 ? k : kprev, P == 0 ? kprev : k};   // g[P] = this frame's slot, g[1-P] = last frame's
    for (int q = 0; q < 2; q++) {
        g[q].position_id = nullptr; g[q].position_pitch = 0;
        g[q].normal_mat = hp.normal[slot_of[q]]; g[q].normal_pitch = 0;
        g[q].uv_inst = hp.uv[slot_of[q]]; g[q].uv_pitch = 0;
        g[q].motion_depth = hp.motion[slot_of[q]]; g[q].motion_pitch = 0;
    }
    // the contents of ring slot k have just changed under the same pointer
    for (int q = 0; q < 2; q++)
        if (c->guide_key[q].motion == hp.motion[k]) c->guide_key[q].clear();
    if ((st = svgf_frame(c, p, g, &b, s))) return st;
    SVGF_CUDA(c, cudaEventRecord(hp.ev_done[k], s));

    // copy-out
    if (h_result || h_history_out) {
        SVGF_CUDA(c, cudaStreamWaitEvent(hp.s_out, hp.ev_done[k], 0));
        if (h_result) SVGF_CUDA(c, cudaMemcpyAsync(h_result, hp.filter[0], colour_bytes(c), cudaMemcpyDeviceToHost, hp.s_out));
        if (h_history_out) SVGF_CUDA(c, cudaMemcpyAsync(h_history_out, hp.history, n, cudaMemcpyDeviceToHost, hp.s_out));
        SVGF_CUDA(c, cudaEventRecord(hp.ev_out, hp.s_out));
        SVGF_CUDA(c, cudaStreamWaitEvent(s, hp.ev_out, 0));
    }
    hp.ping_pong = 1 - P;
    hp.frame = t + 1;
    return SVGF_OK;
}

}  // extern "C"
