// svgf_host_demo.cpp — drives svgf::filter_stage the way application::Render() drives the reference's filter
// (src/App.cu:548-556 + EndFrame :366-375): per frame  Rasterize+Trace (here: the procedural generator writing the
// same texel formats)  ->  TemporalFilter  ->  FilterMoments  ->  WaveletFilter  ->  EndFrame.
//
//   svgf_host_demo [W H frames] [--staged|--fused] [--f32]
// prints one JSON line: GPU ms/frame of the filter stages (CUDA events on the launching stream) and FNV-1a hashes
// of the final FilterBuffer[0] and HistoryLengthBuffer (tests/test_host_cpp.py compares them with the Python mirror).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../csrc/synth_scene.h"
#include "svgf_host.hpp"

extern "C" int svgf_synth_frame_device(const svgf_synth_cfg *cfg, void *position, void *normal, void *uv, void *motion,
                                       void *colour, void *stream);

static uint64_t fnv1a(const std::vector<unsigned char> &v) {
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : v) { h ^= c; h *= 1099511628211ull; }
    return h;
}

int main(int argc, char **argv) {
    int W = 640, H = 360, frames = 8;
    bool fused = true, f32 = false;
    int pos = 0;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--staged")) fused = false;
        else if (!strcmp(argv[i], "--fused")) fused = true;
        else if (!strcmp(argv[i], "--f32")) f32 = true;
        else { int v = atoi(argv[i]); if (pos == 0) W = v; else if (pos == 1) H = v; else frames = v; pos++; }
    }
    try {
        svgf::filter_stage app(W, H, 0, f32 ? SVGF_STORE_F32 : SVGF_STORE_F16);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        float total_ms = 0.f;
        for (int t = 0; t < frames; t++) {
            const int P = app.PingPongInx;
            svgf_synth_cfg cfg{W, H, 0u, t, 3.25f, 0.5f, 32, f32 ? 1 : 0};
            const int rc = svgf_synth_frame_device(&cfg, nullptr, app.Framebuffer[P]->Normal->Data, app.Framebuffer[P]->UV->Data,
                                                   app.Framebuffer[P]->Motion->Data, app.RenderBuffer[P]->Data, app.Stream);
            if (rc) { fprintf(stderr, "generator failed: cudaError %d\n", rc); return 2; }
            cudaEventRecord(e0, app.Stream);
            if (fused) app.FilterFrame();
            else { app.TemporalFilter(); app.FilterMoments(); app.WaveletFilter(); }
            cudaEventRecord(e1, app.Stream);
            svgf::cuda_check(cudaEventSynchronize(e1), "frame");
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (t >= frames / 2) total_ms += ms;
            app.EndFrame();
        }
        std::vector<unsigned char> res(app.FilterBuffer[0]->Size), hist(app.HistoryLengthBuffer->Size);
        svgf::cuda_check(cudaMemcpy(res.data(), app.FilterBuffer[0]->Data, res.size(), cudaMemcpyDeviceToHost), "readback");
        svgf::cuda_check(cudaMemcpy(hist.data(), app.HistoryLengthBuffer->Data, hist.size(), cudaMemcpyDeviceToHost), "readback");
        printf("{\"width\": %d, \"height\": %d, \"frames\": %d, \"mode\": \"%s\", \"storage\": \"%s\", \"ms_per_frame\": %.4f, "
               "\"result_fnv1a\": \"%016llx\", \"history_fnv1a\": \"%016llx\", \"launches\": %llu}\n",
               W, H, frames, fused ? "fused" : "staged", f32 ? "f32" : "f16", total_ms / (frames - frames / 2),
               (unsigned long long)fnv1a(res), (unsigned long long)fnv1a(hist), (unsigned long long)svgf_launch_count(app.context()));
    } catch (const svgf::error &e) {
        fprintf(stderr, "svgf error: %s\n", e.what());
        return 1;
    }
    return 0;
}
