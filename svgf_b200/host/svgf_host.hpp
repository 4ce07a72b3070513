// svgf_host.hpp — header-only C++ host side above the C ABI of include/svgf.h.
//
// Mirrors the filter-relevant members of the reference's `application` class so that code written against the
// reference reads the same here:
//   buffers      RenderBuffer[2], MomentsBuffer[2], FilterBuffer[2], HistoryLengthBuffer, Framebuffer[2], PingPongInx
//                (reference src/App.h:129-141, allocated in ResizeRenderTextures() src/App.cu:742-778)
//   tunables     SpatialFilterSteps, DepthThreshold, NormalThreshold, HistoryLength, PhiColour, PhiNormal
//                (src/App.h:109-114)
//   stages       TemporalFilter(), FilterMoments(), WaveletFilter()  (src/App.cu:469-514), called in that order by
//                Render() (src/App.cu:552-556);  FilterFrame() is the three as one svgf_frame call
//   EndFrame()   flips PingPongInx (src/App.cu:374)
// `buffer` is the reference's cudaMalloc/cudaFree RAII (src/Buffer.cpp:12-56).  Failures throw svgf::error where the
// reference asserts (src/App.cu:41-48).  No OpenGL: the G-buffer planes are linear device memory in the reference's
// texel formats (INTEGRATION.md shows the cudaArray -> linear copy for a GL-interop caller).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>

#include "../../include/svgf.h"

namespace svgf {

struct error : std::runtime_error {
    svgf_status status;
    int cuda_error;
    error(svgf_status s, const char *where, int ce = 0)
        : std::runtime_error(std::string(where) + ": " + svgf_status_string(s) + (ce ? " (cudaError " + std::to_string(ce) + ")" : "")),
          status(s), cuda_error(ce) {}
};

inline void cuda_check(cudaError_t e, const char *where) {
    if (e != cudaSuccess) throw error(SVGF_CUDA_ERROR, where, (int)e);
}

// reference `buffer` (src/Buffer.h / src/Buffer.cpp:12-56): owns `Size` bytes of device memory
struct buffer {
    void *Data = nullptr;
    size_t Size = 0;
    explicit buffer(size_t bytes) : Size(bytes) {
        cuda_check(cudaMalloc(&Data, bytes), "buffer::buffer");
        cuda_check(cudaMemset(Data, 0, bytes), "buffer::buffer");
    }
    ~buffer() { cudaFree(Data); }
    buffer(const buffer &) = delete;
    buffer &operator=(const buffer &) = delete;
};

// reference `framebuffer` as consumed by the filter (cudaFramebuffer, src/App.h:41-44): three linear planes
struct framebuffer {
    std::shared_ptr<buffer> Normal, UV, Motion;   // ushort4 fp16 bits | ushort4 fp16 bits | float4
    framebuffer(int W, int H)
        : Normal(std::make_shared<buffer>((size_t)W * H * 8)), UV(std::make_shared<buffer>((size_t)W * H * 8)),
          Motion(std::make_shared<buffer>((size_t)W * H * 16)) {}
    svgf_gbuffer view() const { return svgf_gbuffer{nullptr, 0, Normal->Data, 0, UV->Data, 0, Motion->Data, 0}; }
};

class filter_stage {
public:
    // tunables under the reference's names (src/App.h:109-114); SpatialFilterSteps is 3 there, 5 in BASELINE.json
    int SpatialFilterSteps = 5;
    float DepthThreshold = 0.8f;
    float NormalThreshold = 0.9f;
    int HistoryLength = 24;
    float PhiColour = 10.0f;
    float PhiNormal = 128.0f;

    int RenderWidth, RenderHeight;
    int PingPongInx = 0;
    std::shared_ptr<buffer> RenderBuffer[2], MomentsBuffer[2], FilterBuffer[2], HistoryLengthBuffer;
    std::shared_ptr<framebuffer> Framebuffer[2];
    cudaStream_t Stream = nullptr;   // the reference uses the default stream

    filter_stage(int W, int H, int device = 0, svgf_storage storage = SVGF_STORE_F16) : RenderWidth(W), RenderHeight(H), Storage(storage) {
        cuda_check(cudaSetDevice(device), "cudaSetDevice");
        ResizeRenderTextures(device);
    }
    ~filter_stage() { svgf_destroy(Ctx); }
    filter_stage(const filter_stage &) = delete;
    filter_stage &operator=(const filter_stage &) = delete;

    // application::ResizeRenderTextures(), src/App.cu:742-778 (filter buffers only)
    void ResizeRenderTextures(int device) {
        const size_t N = (size_t)RenderWidth * RenderHeight, c = (Storage == SVGF_STORE_F32) ? 16 : 8;
        for (int k = 0; k < 2; k++) {
            Framebuffer[k] = std::make_shared<framebuffer>(RenderWidth, RenderHeight);
            RenderBuffer[k] = std::make_shared<buffer>(N * c);
            MomentsBuffer[k] = std::make_shared<buffer>(N * c / 2);
            FilterBuffer[k] = std::make_shared<buffer>(N * c);
        }
        HistoryLengthBuffer = std::make_shared<buffer>(N);
        if (Ctx) svgf_destroy(Ctx);
        Ctx = nullptr;
        check(svgf_create(&Ctx, device, RenderWidth, RenderHeight, Storage), "svgf_create");
        Reset();
    }

    void Reset() {
        svgf_frame_buffers b = bufs();
        check(svgf_reset(Ctx, &b, Stream), "svgf_reset");
        PingPongInx = 0;
    }

    // application::TemporalFilter(), src/App.cu:469-478
    void TemporalFilter() {
        const svgf_params p = params();
        const svgf_gbuffer cur = Framebuffer[PingPongInx]->view(), prev = Framebuffer[1 - PingPongInx]->view();
        check(svgf_temporal(Ctx, &p, &cur, &prev, RenderBuffer[1 - PingPongInx]->Data, RenderBuffer[PingPongInx]->Data,
                            (uint8_t *)HistoryLengthBuffer->Data, MomentsBuffer[PingPongInx]->Data,
                            MomentsBuffer[1 - PingPongInx]->Data, Stream), "svgf_temporal");
    }
    // application::FilterMoments(), src/App.cu:480-489 (moments_index = 0 reproduces src/App.cu:484)
    void FilterMoments(int moments_index = -1) {
        const svgf_params p = params();
        const svgf_gbuffer cur = Framebuffer[PingPongInx]->view();
        check(svgf_variance(Ctx, &p, &cur, RenderBuffer[PingPongInx]->Data,
                            MomentsBuffer[moments_index < 0 ? PingPongInx : moments_index]->Data,
                            (const uint8_t *)HistoryLengthBuffer->Data, FilterBuffer[0]->Data, Stream), "svgf_variance");
    }
    // application::WaveletFilter(), src/App.cu:491-514
    void WaveletFilter() {
        const svgf_params p = params();
        const svgf_gbuffer cur = Framebuffer[PingPongInx]->view();
        void *result = nullptr;
        check(svgf_atrous(Ctx, &p, &cur, FilterBuffer[0]->Data, FilterBuffer[1]->Data, RenderBuffer[PingPongInx]->Data, 0,
                          SpatialFilterSteps, &result, Stream), "svgf_atrous");
        if (result != FilterBuffer[0]->Data)   // odd step count: src/App.cu:510-513
            cuda_check(cudaMemcpyAsync(FilterBuffer[0]->Data, result, FilterBuffer[0]->Size, cudaMemcpyDeviceToDevice, Stream),
                       "WaveletFilter copy");
    }
    // the three stages of Render() (src/App.cu:552-556) as one fused call
    void FilterFrame() {
        const svgf_params p = params();
        const svgf_gbuffer g[2] = {Framebuffer[0]->view(), Framebuffer[1]->view()};
        const svgf_frame_buffers b = bufs();
        check(svgf_frame(Ctx, &p, g, &b, Stream), "svgf_frame");
    }
    // application::TAA(), src/App.cu:516-522: FilterBuffer[0] -> TAABuffer[PingPongInx]; the history is the previous frame's
    // resolve, TAABuffer[1 - PingPongInx] (the reference uses FilterBuffer[1] for both and races on it, include/svgf.h)
    void TAA() {
        for (auto &t : TAABuffer)
            if (!t) t = std::make_shared<buffer>(FilterBuffer[0]->Size);
        check(svgf_taa(Ctx, FilterBuffer[0]->Data, TAABuffer[1 - PingPongInx]->Data, TAABuffer[PingPongInx]->Data, Stream), "svgf_taa");
    }
    std::shared_ptr<buffer> TAABuffer[2];
    // application::EndFrame(), src/App.cu:374
    void EndFrame() { PingPongInx = 1 - PingPongInx; }

    svgf_ctx *context() const { return Ctx; }
    svgf_params params() const {
        svgf_params p;
        svgf_default_params(&p);
        p.atrous_iterations = SpatialFilterSteps; p.depth_threshold = DepthThreshold; p.normal_threshold = NormalThreshold;
        p.history_cap = HistoryLength; p.phi_colour = PhiColour; p.phi_normal = PhiNormal;
        p.mesh_id_mode = MeshIdMode; p.flags = Flags;
        p.reproj_mode = ReprojMode; p.variance_prefilter = VariancePrefilter; p.depth_test_mode = DepthTestMode;
        return p;
    }
    int MeshIdMode = SVGF_MESH_ID_INTENDED;
    uint32_t Flags = SVGF_FLAG_NONE;
    // switches the reference has no counterpart for (include/svgf.h); the defaults are the reference's behaviour
    int ReprojMode = SVGF_REPROJ_NEAREST_TRUNC;
    int VariancePrefilter = SVGF_VARIANCE_PREFILTER_NONE;
    int DepthTestMode = SVGF_DEPTH_TEST_ABSOLUTE;

private:
    svgf_storage Storage;
    svgf_ctx *Ctx = nullptr;
    svgf_frame_buffers bufs() const {
        svgf_frame_buffers b;
        for (int k = 0; k < 2; k++) { b.render[k] = RenderBuffer[k]->Data; b.moments[k] = MomentsBuffer[k]->Data; b.filter[k] = FilterBuffer[k]->Data; }
        b.history = (uint8_t *)HistoryLengthBuffer->Data;
        b.ping_pong = PingPongInx;
        return b;
    }
    void check(svgf_status s, const char *where) const {
        if (s != SVGF_OK) throw error(s, where, Ctx ? svgf_last_cuda_error(Ctx) : 0);
    }
};

}  // namespace svgf
