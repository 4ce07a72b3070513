// svgf_tu_fused.cu — instantiations and launch of the fused level-0+1 a-trous kernel (measured variant)
#include "svgf_ctx.h"
#include "svgf_kernels_fused.cuh"

namespace svgf {
namespace {
template <bool F32, int TERMS>
svgf_status launch_atrous_fused01_t(svgf_ctx *c, const AtrousTiledArgs &t, int guide_slot, const void *in, void *out, void *hist_colour,
                                    cudaStream_t s) {
    using CT = typename ColourPlane<F32>::texel;
    auto kern = atrous_fused01_kernel<F32, TERMS>;
    static std::atomic<unsigned long long> configured{0};
    SVGF_CUDA(c, configure_smem_once(configured, c->device, kern, FusedGeom::smem_bytes));
    const dim3 grid((c->W + kFzW - 1) / kFzW, (c->H + kFzH - 1) / kFzH);
    kern<<<grid, kFzThreads, FusedGeom::smem_bytes, s>>>(t, c->guide[guide_slot].n, c->guide[guide_slot].dz, (const CT *)in, (CT *)out,
                                                         (CT *)hist_colour);
    c->launches++;
    SVGF_CUDA(c, cudaGetLastError());
    return SVGF_OK;
}
}  // namespace

svgf_status atrous_fused01(svgf_ctx *c, bool f32, int terms, const AtrousTiledArgs &a, int guide_slot, const void *in, void *out, void *hist_colour,
                           cudaStream_t s) {
    if (terms == 3) return f32 ? launch_atrous_fused01_t<true, 3>(c, a, guide_slot, in, out, hist_colour, s) : launch_atrous_fused01_t<false, 3>(c, a, guide_slot, in, out, hist_colour, s);
    if (terms == 5) return f32 ? launch_atrous_fused01_t<true, 5>(c, a, guide_slot, in, out, hist_colour, s) : launch_atrous_fused01_t<false, 5>(c, a, guide_slot, in, out, hist_colour, s);
    return SVGF_UNSUPPORTED;
}
}  // namespace svgf
