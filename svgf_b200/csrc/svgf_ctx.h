// svgf_ctx.h — the context behind include/svgf.h and the entry points of the a-trous kernel families.
// Every family is compiled in its own translation unit (svgf_tu_*.cu) so that the library builds in parallel; the
// entry points are plain functions selected at run time by svgf_api.cu.
#pragma once
#include <cuda.h>   // CUtensorMap (type only: the driver entry point is fetched at run time, no libcuda link)
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>

#include "../../include/svgf.h"
#include "svgf_device.cuh"
#include "svgf_kernels_tiled.cuh"   // AtrousTiledArgs

struct svgf_ctx {
    int device = 0, W = 0, H = 0;
    svgf_storage storage = SVGF_STORE_F16;
    uint8_t *hist_shadow = nullptr;       // this frame's history lengths until published (D3)
    svgf::Guide guide[2] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};  // compact guide planes, ping-pong
    int num_sms = 148;
    unsigned int *worklist = nullptr;     // indices of short-history pixels queued by the fused temporal pass
    unsigned int *work_counter = nullptr; // two counters used alternately: the sparse variance pass of frame t zeroes frame t+1's
    int work_parity = 0;
    float *var_blur = nullptr;            // 3x3-blurred variance of the current a-trous input (GAUSS3 prefilter), allocated on first use
    // identity of the G-buffer each guide plane was built from: the three plane pointers, their pitches and the caller's
    // generation counter (svgf_gbuffer has none, so svgf_invalidate_guide / a changed pointer or pitch are the signals)
    struct GuideKey {
        const void *motion = nullptr, *normal = nullptr, *uv = nullptr;
        size_t motion_pitch = 0, normal_pitch = 0, uv_pitch = 0;
        bool matches(const svgf_gbuffer *g) const {
            return motion && g->motion_depth == motion && g->normal_mat == normal && g->uv_inst == uv && g->motion_pitch == motion_pitch &&
                   g->normal_pitch == normal_pitch && g->uv_pitch == uv_pitch;
        }
        void set(const svgf_gbuffer *g) {
            motion = g->motion_depth; normal = g->normal_mat; uv = g->uv_inst;
            motion_pitch = g->motion_pitch; normal_pitch = g->normal_pitch; uv_pitch = g->uv_pitch;
        }
        void clear() { motion = nullptr; }
    } guide_key[2];
    int guide_cur = 0;                    // slot of the most recently built guide
    bool force_fail_next = false;         // set by svgf_reset
    int last_err = 0;
    uint64_t launches = 0;
    // lattice planes of the TMA-staged a-trous levels (svgf_kernels_lattice.cuh), allocated on first use
    struct Lattice {
        bool ready = false, failed = false;
        int pitch_pairs = 0, rows = 0;
        size_t npairs = 0;
        svgf::LatticeColour sc[2] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
        svgf::LatticeNormals sn = {nullptr, nullptr};
        // tensor maps per level 1..4 (index 0 unused) and plane: sc[0].c0 c1 lz, sc[1].c0 c1 lz, sn.n0, sn.n1
        CUtensorMap map[5][8];
    } lat;
    // which kernel family ran each a-trous level of the last svgf_frame / svgf_atrous call (svgf_last_dispatch)
    int dispatch[16] = {};
    int dispatch_n = 0;
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
// Raises a kernel's dynamic shared-memory limit once per device.  `done` is a per-instantiation bit set indexed by device
// ordinal (ordinals >= 64 simply repeat the call, which is harmless); concurrent contexts may race to set the same bit
// with the same value.
template <typename K> inline cudaError_t configure_smem_once(std::atomic<unsigned long long> &done, int device, K kern, size_t bytes) {
    const unsigned long long bit = device < 64 ? (1ull << device) : 0ull;
    if (bit && (done.load(std::memory_order_acquire) & bit)) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess && bit) done.fetch_or(bit, std::memory_order_release);
    return e;
}

// svgf_dispatch_family values (include/svgf.h)
enum { kFamBasic = 1, kFamPacked = 2, kFamPackedStaged = 3, kFamLattice = 4, kFamBulk = 5, kFamStream = 6, kFamFused01 = 7 };

// svgf_tma.cu: allocate the lattice planes and encode their tensor maps (idempotent)
svgf_status lattice_prepare(svgf_ctx *c, cudaStream_t s);
void lattice_destroy(svgf_ctx *c);
// first level of a staged run: the packed kernel writing lattice planes (set `dst`) instead of `out`
svgf_status atrous_packed_staged_f16(svgf_ctx *c, int terms, const AtrousTiledArgs &a, int guide_slot, const void *in, int dst, void *hist_colour, cudaStream_t s);
svgf_status atrous_packed_staged_f32(svgf_ctx *c, int terms, const AtrousTiledArgs &a, int guide_slot, const void *in, int dst, void *hist_colour, cudaStream_t s);
// lattice level a.level (1..4): planes sc[src] -> sc[1 - src], or -> `out` in the storage format when out != nullptr;
// pdl: launch with programmatic stream serialisation (the kernel's prologue overlaps the previous level's tail)
svgf_status atrous_lattice_f16(svgf_ctx *c, int terms, const AtrousTiledArgs &a, int guide_slot, int src, void *out, bool pdl, cudaStream_t s);
svgf_status atrous_lattice_f32(svgf_ctx *c, int terms, const AtrousTiledArgs &a, int guide_slot, int src, void *out, bool pdl, cudaStream_t s);

// svgf_api.cu, for the band driver (svgf_band.cu)
svgf_status staged_level(svgf_ctx *c, const svgf_params *p, int guide_slot, int level, int kind, const void *in, int idx, void *out,
                         void *hist_colour, int yb0, int nyb, int yb1, int nyb1, bool pdl, cudaStream_t s);
bool staged_run_possible(const svgf_ctx *c, const svgf_params *p, const void *a, const void *b, const void *hist_colour, int first, int n);

// one a-trous level (a.level = 0..4) / levels 0+1 fused; terms = series terms of the normal weight (3, 4 or 5);
// rows = outputs per thread and column of the packed kernel (3, or 4 with terms == 3)
svgf_status atrous_packed_f16(svgf_ctx *c, int terms, int rows, const AtrousTiledArgs &a, int guide_slot, const void *in, void *out, void *hist_colour, cudaStream_t s);
svgf_status atrous_packed_f32(svgf_ctx *c, int terms, int rows, const AtrousTiledArgs &a, int guide_slot, const void *in, void *out, void *hist_colour, cudaStream_t s);
svgf_status atrous_tiled(svgf_ctx *c, bool f32, int terms, const AtrousTiledArgs &a, int guide_slot, const void *in, void *out, void *hist_colour, cudaStream_t s);
svgf_status atrous_stream(svgf_ctx *c, bool f32, int terms, const AtrousTiledArgs &a, int guide_slot, const void *in, void *out, void *hist_colour, cudaStream_t s);
svgf_status atrous_fused01(svgf_ctx *c, bool f32, int terms, const AtrousTiledArgs &a, int guide_slot, const void *in, void *out, void *hist_colour, cudaStream_t s);
}  // namespace svgf
