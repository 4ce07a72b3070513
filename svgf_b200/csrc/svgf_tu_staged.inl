// svgf_tu_staged.inl — the packed kernel as the FIRST level of a staged run (writes lattice planes), ONE storage type
// (included by svgf_tu_staged_f16.cu / svgf_tu_staged_f32.cu with SVGF_TU_F32 defined)
#include "svgf_ctx.h"
#include "svgf_kernels_packed.cuh"

namespace svgf {
namespace {
template <bool F32, int STEP, int TERMS>
svgf_status launch_staged(svgf_ctx *c, AtrousTiledArgs a, int guide_slot, const void *in, int dst, void *hist_colour, cudaStream_t s) {
    using CT = typename ColourPlane<F32>::texel;
    using G = PackedGeom<STEP>;
    auto kern = atrous_packed_kernel<F32, STEP, TERMS, kPkRows, false, false, true>;
    static std::atomic<unsigned long long> configured{0};
    SVGF_CUDA(c, configure_smem_once(configured, c->device, kern, G::smem_bytes));
    const int all_yblocks = (c->H + G::tile_rows * STEP - 1) / (G::tile_rows * STEP);
    const dim3 grid((c->W + kTileW - 1) / kTileW, (a.nyblocks > 0 ? a.nyblocks : all_yblocks) * STEP);
    kern<<<grid, kPkThreads, G::smem_bytes, s>>>(a, c->guide[guide_slot].n, c->guide[guide_slot].dz, (const CT *)in, (CT *)nullptr,
                                                 (CT *)hist_colour, c->lat.sc[dst], c->lat.sn, c->lat.pitch_pairs);
    c->launches++;
    SVGF_CUDA(c, cudaGetLastError());
    return SVGF_OK;
}
template <bool F32, int TERMS>
svgf_status dispatch_staged(svgf_ctx *c, const AtrousTiledArgs &a, int guide_slot, const void *in, int dst, void *hist_colour, cudaStream_t s) {
    switch (a.level) {
        case 0: return launch_staged<F32, 1, TERMS>(c, a, guide_slot, in, dst, hist_colour, s);
        case 1: return launch_staged<F32, 2, TERMS>(c, a, guide_slot, in, dst, hist_colour, s);
        case 2: return launch_staged<F32, 4, TERMS>(c, a, guide_slot, in, dst, hist_colour, s);
        case 3: return launch_staged<F32, 8, TERMS>(c, a, guide_slot, in, dst, hist_colour, s);
    }
    return SVGF_UNSUPPORTED;
}
}  // namespace

svgf_status SVGF_TU_STAGED_ENTRY(svgf_ctx *c, int terms, const AtrousTiledArgs &a, int guide_slot, const void *in, int dst, void *hist_colour, cudaStream_t s) {
    constexpr bool F32 = SVGF_TU_F32;
    if (terms == 3) return dispatch_staged<F32, 3>(c, a, guide_slot, in, dst, hist_colour, s);
    if (terms == 5) return dispatch_staged<F32, 5>(c, a, guide_slot, in, dst, hist_colour, s);
    return SVGF_UNSUPPORTED;
}
}  // namespace svgf
