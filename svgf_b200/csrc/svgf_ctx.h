// svgf_ctx.h — the context behind include/svgf.h and the entry points of the a-trous kernel families.
// Every family is compiled in its own translation unit (svgf_tu_*.cu) so that the library builds in parallel; the
// entry points are plain functions selected at run time by svgf_api.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/svgf.h"
#include "svgf_device.cuh"
#include "svgf_kernels_tiled.cuh"   // AtrousTiledArgs

struct svgf_ctx {
    int device = 0, W = 0, H = 0;
    svgf_storage storage = SVGF_STORE_F16;
    uint8_t *hist_shadow = nullptr;       // this frame's history lengths until published (D3)
    svgf::Guide guide[2] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};  // compact guide planes, ping-pong
    int num_sms = 148;
    unsigned int *worklist = nullptr;     // indices of short-history pixels queued by the fused temporal pass
    unsigned int *work_counter = nullptr;
    float *var_blur = nullptr;            // 3x3-blurred variance of the current a-trous input (GAUSS3 prefilter), allocated on first use
    const void *guide_key[2] = {nullptr, nullptr};  // motion_depth pointer of the G-buffer each plane was built from
    int guide_cur = 0;                    // slot of the most recently built guide
    bool force_fail_next = false;         // set by svgf_reset
    int last_err = 0;
    uint64_t launches = 0;
    // stage profiling (svgf_profile_*): events around temporal / variance / a-trous inside svgf_frame
    bool profiling = false;
    static constexpr int kMaxProf = 4096;
    cudaEvent_t *prof_ev = nullptr;       // 4 events per frame
    int prof_frames = 0;
    // device-resident state of the host-buffer path (svgf_frame_host): a 3-slot ring of staged inputs filled on a
    // copy-in stream, results drained on a copy-out stream, so that the PCIe transfers of frames t+1 and t-1
    // overlap the kernels of frame t
    struct HostPath {
        static constexpr int kRing = 3;
        void *normal[kRing] = {}, *uv[kRing] = {}, *motion[kRing] = {}, *noisy[kRing] = {};
        void *render[2] = {nullptr, nullptr}, *moments[2] = {nullptr, nullptr}, *filter[2] = {nullptr, nullptr};
        uint8_t *history = nullptr;
        cudaStream_t s_in = nullptr, s_out = nullptr;
        cudaEvent_t ev_in[kRing] = {}, ev_done[kRing] = {}, ev_out = nullptr;
        uint64_t frame = 0;                 // frames submitted since the ring was created
        int ping_pong = 0;
        bool ready = false;
    } hp;
};


inline svgf_status svgf_cuda_fail(svgf_ctx *c, cudaError_t e) {
    if (c) c->last_err = (int)e;
    return SVGF_CUDA_ERROR;
}
#define SVGF_CUDA(c, x)                                         \
    do {                                                        \
        cudaError_t e_ = (x);                                   \
        if (e_ != cudaSuccess) return svgf_cuda_fail((c), e_);  \
    } while (0)

namespace svgf {
// one a-trous level (a.level = 0..4) / levels 0+1 fused; terms = series terms of the normal weight (3, 4 or 5);
// rows = outputs per thread and column of the packed kernel (3, or 4 with terms == 3)
svgf_status atrous_packed_f16(svgf_ctx *c, int terms, int rows, const AtrousTiledArgs &a, int guide_slot, const void *in, void *out, void *hist_colour, cudaStream_t s);
svgf_status atrous_packed_f32(svgf_ctx *c, int terms, int rows, const AtrousTiledArgs &a, int guide_slot, const void *in, void *out, void *hist_colour, cudaStream_t s);
svgf_status atrous_tiled(svgf_ctx *c, bool f32, int terms, const AtrousTiledArgs &a, int guide_slot, const void *in, void *out, void *hist_colour, cudaStream_t s);
svgf_status atrous_stream(svgf_ctx *c, bool f32, int terms, const AtrousTiledArgs &a, int guide_slot, const void *in, void *out, void *hist_colour, cudaStream_t s);
svgf_status atrous_fused01(svgf_ctx *c, bool f32, int terms, const AtrousTiledArgs &a, int guide_slot, const void *in, void *out, void *hist_colour, cudaStream_t s);
}  // namespace svgf
