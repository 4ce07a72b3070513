// svgf_tu_stream.cu — instantiations and launch of the streaming a-trous kernel (measured variant)
#include "svgf_ctx.h"
#include "svgf_kernels_stream.cuh"

namespace svgf {
namespace {
template <bool F32, int STEP, int TERMS>
svgf_status launch_atrous_stream(svgf_ctx *c, const AtrousTiledArgs &t, int guide_slot, const void *in, void *out, void *hist_colour,
                                 cudaStream_t s) {
    using CT = typename ColourPlane<F32>::texel;
    using G = StreamGeom<STEP>;
    auto kern = atrous_stream_kernel<F32, STEP, TERMS>;
    static std::atomic<unsigned long long> configured{0};
    const size_t smem = G::smem_bytes;
    SVGF_CUDA(c, configure_smem_once(configured, c->device, kern, smem));
    AtrousStreamArgs a;
    a.t = t;
    a.n_strips = (c->W + kStripW - 1) / kStripW;
    a.rows_a = c->H / STEP;
    a.rows_b = c->H % STEP;
    // the (strip, phase, row) stream is cut into equal contiguous ranges, one per CTA; two CTAs per SM
    const long long T = (long long)a.n_strips * c->H;
    long long grid = 2LL * c->num_sms;
    if (grid > T / 8) grid = T / 8 > 0 ? T / 8 : 1;
    kern<<<(unsigned int)grid, kStreamThreads, smem, s>>>(a, c->guide[guide_slot].n, c->guide[guide_slot].dz, (const CT *)in,
                                                                    (CT *)out, (CT *)hist_colour);
    c->launches++;
    SVGF_CUDA(c, cudaGetLastError());
    return SVGF_OK;
}
template <bool F32, int TERMS>
svgf_status dispatch_atrous_stream(svgf_ctx *c, const AtrousTiledArgs &a, int guide_slot, const void *in, void *out,
                                   void *hist_colour, cudaStream_t s) {
    switch (a.level) {
        case 0: return launch_atrous_stream<F32, 1, TERMS>(c, a, guide_slot, in, out, hist_colour, s);
        case 1: return launch_atrous_stream<F32, 2, TERMS>(c, a, guide_slot, in, out, hist_colour, s);
        case 2: return launch_atrous_stream<F32, 4, TERMS>(c, a, guide_slot, in, out, hist_colour, s);
        case 3: return launch_atrous_stream<F32, 8, TERMS>(c, a, guide_slot, in, out, hist_colour, s);
        case 4: return launch_atrous_stream<F32, 16, TERMS>(c, a, guide_slot, in, out, hist_colour, s);
    }
    return SVGF_UNSUPPORTED;
}

}  // namespace

svgf_status atrous_stream(svgf_ctx *c, bool f32, int terms, const AtrousTiledArgs &a, int guide_slot, const void *in, void *out, void *hist_colour,
                  cudaStream_t s) {
    if (terms == 4) return f32 ? dispatch_atrous_stream<true, 4>(c, a, guide_slot, in, out, hist_colour, s) : dispatch_atrous_stream<false, 4>(c, a, guide_slot, in, out, hist_colour, s);
    if (terms == 5) return f32 ? dispatch_atrous_stream<true, 5>(c, a, guide_slot, in, out, hist_colour, s) : dispatch_atrous_stream<false, 5>(c, a, guide_slot, in, out, hist_colour, s);
    return SVGF_UNSUPPORTED;
}
}  // namespace svgf
