// ref_harness.cu — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Host harness around the REFERENCE's own kernels (filter::TemporalFilter / FilterMoments / FilterKernel,
// /root/reference/src/Filter.cuh:359-624).  oracle/Makefile compiles Filter.cuh from where it lies under
// /root/reference (six by-value signature edits applied by sed into a temp dir, see the Makefile; math
// untouched) together with this file into oracle/_ref/libsvgf_refkernels.so.  No reference source is
// copied into the repository.
//
// The harness plays the role of application::{TemporalFilter,FilterMoments,WaveletFilter}
// (src/App.cu:469-514): same launch shapes (16x16 blocks, grid (W/16+1, H/16+1)), same argument order,
// G-buffers fed through cudaArray-backed texture objects created like src/CudaUtil.h:84-95 (zeroed
// cudaTextureDesc: point filter, un-normalised coordinates, element read mode).
//
// Uses: (1) validating oracle/svgf_oracle.cpp against the real reference math on a B200
// (tests/test_reference_kernels.py); (2) the "reference" arm of bench.py (--impl reference).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>

#include "Filter.cuh"  // the reference's, patched for compilation only (oracle/Makefile)

namespace {

struct GBufArrays {
    cudaArray_t arr[4] = {nullptr, nullptr, nullptr, nullptr};  // Position, Normal, UV, Motion (src/App.h:33-39)
    cudaTextureObject_t tex[4] = {0, 0, 0, 0};
};

cudaError_t make_gbuf(GBufArrays &g, int W, int H) {
    for (int i = 0; i < 4; i++) {
        cudaChannelFormatDesc d = (i == 0 || i == 3) ? cudaCreateChannelDesc<float4>() : cudaCreateChannelDesc<ushort4>();
        cudaError_t e = cudaMallocArray(&g.arr[i], &d, W, H);
        if (e != cudaSuccess) return e;
        cudaResourceDesc rd;
        std::memset(&rd, 0, sizeof(rd));
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = g.arr[i];
        cudaTextureDesc td;
        std::memset(&td, 0, sizeof(td));  // src/CudaUtil.h:88-95
        e = cudaCreateTextureObject(&g.tex[i], &rd, &td, nullptr);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

void free_gbuf(GBufArrays &g) {
    for (int i = 0; i < 4; i++) {
        if (g.tex[i]) cudaDestroyTextureObject(g.tex[i]);
        if (g.arr[i]) cudaFreeArray(g.arr[i]);
    }
}

}  // namespace

struct svgf_ref_ctx {
    int W, H;
    GBufArrays gb[2];
    filter::half4 *render[2];
    filter::half2 *moments[2];
    filter::half4 *filt[2];
    uint8_t *history;
    int ping_pong;
    cudaEvent_t ev0, ev1;
};

struct svgf_ref_params {
    int SpatialFilterSteps;
    float DepthThreshold, NormalThreshold;
    int HistoryLength;
    float PhiColour, PhiNormal;
    int moments_quirk;  // 1 = pass MomentsBuffer[0] to FilterMoments like src/App.cu:484; 0 = MomentsBuffer[PingPongInx]
};

#define CK(x)                                  \
    do {                                       \
        cudaError_t e_ = (x);                  \
        if (e_ != cudaSuccess) return (int)e_; \
    } while (0)

extern "C" {

int svgf_ref_create(svgf_ref_ctx **out, int W, int H) {
    svgf_ref_ctx *c = new svgf_ref_ctx();
    c->W = W; c->H = H; c->ping_pong = 0;
    const size_t n = (size_t)W * H;
    for (int i = 0; i < 2; i++) {
        CK(make_gbuf(c->gb[i], W, H));
        CK(cudaMalloc(&c->render[i], n * sizeof(filter::half4)));
        CK(cudaMalloc(&c->moments[i], n * sizeof(filter::half2)));
        CK(cudaMalloc(&c->filt[i], n * sizeof(filter::half4)));
    }
    CK(cudaMalloc(&c->history, n));
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
    *out = c;
    return 0;
}

void svgf_ref_destroy(svgf_ref_ctx *c) {
    if (!c) return;
    for (int i = 0; i < 2; i++) {
        free_gbuf(c->gb[i]);
        cudaFree(c->render[i]); cudaFree(c->moments[i]); cudaFree(c->filt[i]);
    }
    cudaFree(c->history);
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
    delete c;
}

// Zero all state (the first-frame definition D12 shared with the oracle and the product).
int svgf_ref_reset(svgf_ref_ctx *c) {
    const size_t n = (size_t)c->W * c->H;
    for (int i = 0; i < 2; i++) {
        CK(cudaMemset(c->render[i], 0, n * sizeof(filter::half4)));
        CK(cudaMemset(c->moments[i], 0, n * sizeof(filter::half2)));
        CK(cudaMemset(c->filt[i], 0, n * sizeof(filter::half4)));
        // zero prev G-buffer planes
        void *z = nullptr;
        CK(cudaMalloc(&z, n * 16));
        CK(cudaMemset(z, 0, n * 16));
        CK(cudaMemcpy2DToArray(c->gb[i].arr[0], 0, 0, z, (size_t)c->W * 16, (size_t)c->W * 16, c->H, cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy2DToArray(c->gb[i].arr[1], 0, 0, z, (size_t)c->W * 8, (size_t)c->W * 8, c->H, cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy2DToArray(c->gb[i].arr[2], 0, 0, z, (size_t)c->W * 8, (size_t)c->W * 8, c->H, cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy2DToArray(c->gb[i].arr[3], 0, 0, z, (size_t)c->W * 16, (size_t)c->W * 16, c->H, cudaMemcpyDeviceToDevice));
        CK(cudaFree(z));
    }
    CK(cudaMemset(c->history, 0, n));
    c->ping_pong = 0;
    return 0;
}

// kind: cudaMemcpyHostToDevice (1) or cudaMemcpyDeviceToDevice (3).  Dense planes (pitch = W * texel).
int svgf_ref_set_gbuffer(svgf_ref_ctx *c, int slot, const void *normal, const void *uv, const void *motion, int kind) {
    const size_t W = c->W;
    if (normal) CK(cudaMemcpy2DToArray(c->gb[slot].arr[1], 0, 0, normal, W * 8, W * 8, c->H, (cudaMemcpyKind)kind));
    if (uv) CK(cudaMemcpy2DToArray(c->gb[slot].arr[2], 0, 0, uv, W * 8, W * 8, c->H, (cudaMemcpyKind)kind));
    if (motion) CK(cudaMemcpy2DToArray(c->gb[slot].arr[3], 0, 0, motion, W * 16, W * 16, c->H, (cudaMemcpyKind)kind));
    return 0;
}

// which: 0 render[slot], 1 moments[slot], 2 filter[slot], 3 history
static void *plane(svgf_ref_ctx *c, int which, int slot, size_t *bytes) {
    const size_t n = (size_t)c->W * c->H;
    switch (which) {
        case 0: *bytes = n * 8; return c->render[slot];
        case 1: *bytes = n * 4; return c->moments[slot];
        case 2: *bytes = n * 8; return c->filt[slot];
        case 3: *bytes = n; return c->history;
    }
    *bytes = 0;
    return nullptr;
}
int svgf_ref_set_plane(svgf_ref_ctx *c, int which, int slot, const void *src, int kind) {
    size_t b; void *d = plane(c, which, slot, &b);
    if (!d) return -1;
    CK(cudaMemcpy(d, src, b, (cudaMemcpyKind)kind));
    return 0;
}
int svgf_ref_get_plane(svgf_ref_ctx *c, int which, int slot, void *dst, int kind) {
    size_t b; void *s = plane(c, which, slot, &b);
    if (!s) return -1;
    CK(cudaMemcpy(dst, s, b, (cudaMemcpyKind)kind));
    return 0;
}
void svgf_ref_set_ping_pong(svgf_ref_ctx *c, int p) { c->ping_pong = p & 1; }
int svgf_ref_get_ping_pong(svgf_ref_ctx *c) { return c->ping_pong; }

static gpupt::cudaFramebuffer fb(const GBufArrays &g) { return {g.tex[0], g.tex[1], g.tex[2], g.tex[3]}; }

// application::TemporalFilter, src/App.cu:469-478
int svgf_ref_temporal(svgf_ref_ctx *c, const svgf_ref_params *p) {
    dim3 block(16, 16), grid(c->W / 16 + 1, c->H / 16 + 1);
    const int P = c->ping_pong;
    filter::TemporalFilter<<<grid, block>>>(c->render[1 - P], c->render[P], fb(c->gb[P]), fb(c->gb[1 - P]), c->history,
                                            c->moments[P], c->moments[1 - P], c->W, c->H, p->DepthThreshold,
                                            p->NormalThreshold, p->HistoryLength);
    return (int)cudaGetLastError();
}

// application::FilterMoments, src/App.cu:480-489
int svgf_ref_variance(svgf_ref_ctx *c, const svgf_ref_params *p) {
    dim3 block(16, 16), grid(c->W / 16 + 1, c->H / 16 + 1);
    const int P = c->ping_pong;
    filter::FilterMoments<<<grid, block>>>(c->render[P], c->filt[0], c->moments[p->moments_quirk ? 0 : P], c->gb[P].tex[3],
                                           c->gb[P].tex[1], c->history, c->W, c->H, p->PhiColour, p->PhiNormal);
    return (int)cudaGetLastError();
}

// application::WaveletFilter, src/App.cu:491-514
int svgf_ref_wavelet(svgf_ref_ctx *c, const svgf_ref_params *p) {
    dim3 block(16, 16), grid(c->W / 16 + 1, c->H / 16 + 1);
    const int P = c->ping_pong;
    int pp = 0;
    for (int i = 0; i < p->SpatialFilterSteps; i++) {
        filter::FilterKernel<<<grid, block>>>(c->filt[pp], c->gb[P].tex[3], c->gb[P].tex[1], c->history, c->filt[1 - pp],
                                              c->render[P], c->W, c->H, 1 << i, p->PhiColour, p->PhiNormal, i);
        pp = 1 - pp;
    }
    if (p->SpatialFilterSteps % 2 != 0)
        CK(cudaMemcpy(c->filt[0], c->filt[1], (size_t)c->W * c->H * sizeof(filter::half4), cudaMemcpyDeviceToDevice));
    return (int)cudaGetLastError();
}

// One à-trous level on its own (for per-level validation of the oracle): filt[0] -> filt[1]
int svgf_ref_atrous_level(svgf_ref_ctx *c, const svgf_ref_params *p, int level) {
    dim3 block(16, 16), grid(c->W / 16 + 1, c->H / 16 + 1);
    const int P = c->ping_pong;
    filter::FilterKernel<<<grid, block>>>(c->filt[0], c->gb[P].tex[3], c->gb[P].tex[1], c->history, c->filt[1], c->render[P],
                                          c->W, c->H, 1 << level, p->PhiColour, p->PhiNormal, level);
    return (int)cudaGetLastError();
}

// application::TAA, src/App.cu:516-522: FilterBuffer[0] -> FilterBuffer[1], which is also read as the history (D13: racy
// unless the history is a fixed point of the kernel - tests/test_taa.py iterates it to one)
int svgf_ref_taa(svgf_ref_ctx *c) {
    dim3 block(16, 16), grid(c->W / 16 + 1, c->H / 16 + 1);
    filter::TAAFilterKernel<<<grid, block>>>(c->filt[0], c->filt[1], c->W, c->H);
    return (int)cudaGetLastError();
}

// Render()'s filter stages (src/App.cu:552-556) + EndFrame's flip (src/App.cu:374)
int svgf_ref_frame(svgf_ref_ctx *c, const svgf_ref_params *p, int flip) {
    int e;
    if ((e = svgf_ref_temporal(c, p))) return e;
    if ((e = svgf_ref_variance(c, p))) return e;
    if ((e = svgf_ref_wavelet(c, p))) return e;
    if (flip) c->ping_pong = 1 - c->ping_pong;
    return 0;
}

// End-to-end with HOST buffers: upload this frame's G-buffer + noisy colour, run the stages, download the
// result and flip.  Mirrors svgf_frame_host of the product ABI.
int svgf_ref_frame_host(svgf_ref_ctx *c, const svgf_ref_params *p, const void *h_normal, const void *h_uv,
                        const void *h_motion, const void *h_colour, void *h_result, uint8_t *h_history_out) {
    const int P = c->ping_pong;
    int e;
    if ((e = svgf_ref_set_gbuffer(c, P, h_normal, h_uv, h_motion, (int)cudaMemcpyHostToDevice))) return e;
    if ((e = svgf_ref_set_plane(c, 0, P, h_colour, (int)cudaMemcpyHostToDevice))) return e;
    if ((e = svgf_ref_frame(c, p, 0))) return e;
    if (h_result && (e = svgf_ref_get_plane(c, 2, 0, h_result, (int)cudaMemcpyDeviceToHost))) return e;
    if (h_history_out && (e = svgf_ref_get_plane(c, 3, 0, h_history_out, (int)cudaMemcpyDeviceToHost))) return e;
    c->ping_pong = 1 - c->ping_pong;
    return 0;
}

// Device-resident timing: run `iters` frames back to back on whatever the buffers hold (flipping the
// ping-pong each frame like EndFrame) and return the elapsed GPU milliseconds measured with CUDA events.
int svgf_ref_time_frames(svgf_ref_ctx *c, const svgf_ref_params *p, int iters, float *ms_total) {
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(c->ev0));
    for (int i = 0; i < iters; i++) {
        int e = svgf_ref_frame(c, p, 1);
        if (e) return e;
    }
    CK(cudaEventRecord(c->ev1));
    CK(cudaEventSynchronize(c->ev1));
    CK(cudaEventElapsedTime(ms_total, c->ev0, c->ev1));
    return 0;
}

int svgf_ref_sync(void) { return (int)cudaDeviceSynchronize(); }

}  // extern "C"
