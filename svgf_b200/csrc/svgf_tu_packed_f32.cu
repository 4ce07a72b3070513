// svgf_tu_packed_f32.cu — instantiations and launch of the packed FP32x2 a-trous kernel, fp32 storage
#include "svgf_ctx.h"
#include "svgf_kernels_packed.cuh"

namespace svgf {
namespace {
template <bool F32, int STEP, int TERMS, int R, bool PREF, bool HC = false>
svgf_status launch_atrous_packed(svgf_ctx *c, AtrousTiledArgs a, int guide_slot, const void *in, void *out, void *hist_colour,
                                 cudaStream_t s) {
    using CT = typename ColourPlane<F32>::texel;
    using G = PackedGeom<STEP>;
    auto kern = atrous_packed_kernel<F32, STEP, TERMS, R, PREF>;
    static std::atomic<unsigned long long> configured{0};
    SVGF_CUDA(c, configure_smem_once(configured, c->device, kern, G::smem_bytes));
    const int all_yblocks = (c->H + G::tile_rows * STEP - 1) / (G::tile_rows * STEP);
    const dim3 grid((c->W + kTileW - 1) / kTileW, (a.nyblocks > 0 ? a.nyblocks : all_yblocks) * STEP);
    kern<<<grid, kPkPairs * (G::tile_rows / R), G::smem_bytes, s>>>(a, c->guide[guide_slot].n, c->guide[guide_slot].dz, (const CT *)in, (CT *)out,
                                                 (CT *)hist_colour, LatticeColour{nullptr, nullptr, nullptr}, LatticeNormals{nullptr, nullptr}, 0);
    c->launches++;
    SVGF_CUDA(c, cudaGetLastError());
    return SVGF_OK;
}
template <bool F32, int TERMS, int R, bool PREF = false>
svgf_status dispatch_atrous_packed(svgf_ctx *c, const AtrousTiledArgs &a, int guide_slot, const void *in, void *out,
                                   void *hist_colour, cudaStream_t s) {
    switch (a.level) {
        case 0: return launch_atrous_packed<F32, 1, TERMS, R, PREF>(c, a, guide_slot, in, out, hist_colour, s);
        case 1: return launch_atrous_packed<F32, 2, TERMS, R, PREF>(c, a, guide_slot, in, out, hist_colour, s);
        case 2: return launch_atrous_packed<F32, 4, TERMS, R, PREF>(c, a, guide_slot, in, out, hist_colour, s);
        case 3: return launch_atrous_packed<F32, 8, TERMS, R, PREF>(c, a, guide_slot, in, out, hist_colour, s);
        case 4: return launch_atrous_packed<F32, 16, TERMS, R, PREF>(c, a, guide_slot, in, out, hist_colour, s);
    }
    return SVGF_UNSUPPORTED;
}

}  // namespace

svgf_status atrous_packed_f32(svgf_ctx *c, int terms, int rows, const AtrousTiledArgs &a, int guide_slot, const void *in, void *out, void *hist_colour,
                              cudaStream_t s) {
    // rows = 4 (192 threads, 164 registers) and rows = 6 (128 threads, 238 registers) were measured on B200 and lost:
    // 0.926 / 1.234 ms per 4K frame of a-trous against 0.761 ms for rows = 3 - fewer resident warps cost more than the
    // shared-memory traffic they save; instantiate dispatch_atrous_packed<.., 3, 4> here to repeat it
    if (rows != kPkRows) return SVGF_UNSUPPORTED;
    if (a.var_blur)   // variance prefilter: instantiated for the three-term series (phi_normal >= 100) only
        return terms == 3 ? dispatch_atrous_packed<true, 3, kPkRows, true>(c, a, guide_slot, in, out, hist_colour, s) : SVGF_UNSUPPORTED;
    switch (terms) {
        case 3: return dispatch_atrous_packed<true, 3, kPkRows>(c, a, guide_slot, in, out, hist_colour, s);
        case 5: return dispatch_atrous_packed<true, 5, kPkRows>(c, a, guide_slot, in, out, hist_colour, s);
    }
    return SVGF_UNSUPPORTED;
}
}  // namespace svgf
