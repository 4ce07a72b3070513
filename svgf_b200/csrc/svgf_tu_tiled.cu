// svgf_tu_tiled.cu — instantiations and launch of the persistent bulk-copy a-trous kernel (measured variant, odd-width fallback)
#include "svgf_ctx.h"
#include "svgf_kernels_tiled.cuh"

namespace svgf {
namespace {
template <bool F32, int STEP, int RG, int TERMS>
svgf_status launch_atrous_tiled(svgf_ctx *c, AtrousTiledArgs a, int guide_slot, const void *in, void *out, void *hist_colour,
                                cudaStream_t s) {
    using CT = typename ColourPlane<F32>::texel;
    using G = TileGeom<F32, STEP, RG>;
    auto kern = atrous_tiled_kernel<F32, STEP, RG, TERMS>;
    static std::atomic<unsigned long long> configured{0};
    static std::atomic<int> ctas_per_sm{0};   // a property of the kernel and of sm_100a, identical on every device
    SVGF_CUDA(c, configure_smem_once(configured, c->device, kern, G::smem_bytes));
    int cps = ctas_per_sm.load(std::memory_order_relaxed);
    if (cps == 0) {
        SVGF_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cps, kern, G::threads, G::smem_bytes));
        if (cps < 1) return SVGF_UNSUPPORTED;
        ctas_per_sm.store(cps, std::memory_order_relaxed);
    }
    a.tiles_x = (c->W + kTileW - 1) / kTileW;
    a.tiles_y = ((c->H + G::tile_rows * STEP - 1) / (G::tile_rows * STEP)) * STEP;
    const int n_tiles = a.tiles_x * a.tiles_y;
    const int grid = n_tiles < c->num_sms * cps ? n_tiles : c->num_sms * cps;
    kern<<<grid, G::threads, G::smem_bytes, s>>>(a, c->guide[guide_slot].n, c->guide[guide_slot].dz, (const CT *)in, (CT *)out,
                                                 (CT *)hist_colour);
    c->launches++;
    SVGF_CUDA(c, cudaGetLastError());
    return SVGF_OK;
}

// Row groups per CTA: 2 (256 threads, two CTAs per SM) while the tile fits twice; 4 (512 threads, one CTA per SM,
// less vertical halo) for the wide-halo levels — chosen so every instantiation fits the 227 KB of shared memory.
template <bool F32, int TERMS>
svgf_status dispatch_atrous_tiled(svgf_ctx *c, const AtrousTiledArgs &a, int guide_slot, const void *in, void *out,
                                  void *hist_colour, cudaStream_t s) {
    switch (a.level) {
        case 0: return launch_atrous_tiled<F32, 1, 2, TERMS>(c, a, guide_slot, in, out, hist_colour, s);
        case 1: return launch_atrous_tiled<F32, 2, 2, TERMS>(c, a, guide_slot, in, out, hist_colour, s);
        case 2: return launch_atrous_tiled<F32, 4, 2, TERMS>(c, a, guide_slot, in, out, hist_colour, s);
        case 3: return launch_atrous_tiled<F32, 8, 4, TERMS>(c, a, guide_slot, in, out, hist_colour, s);
        case 4: return launch_atrous_tiled<F32, 16, (F32 ? 2 : 4), TERMS>(c, a, guide_slot, in, out, hist_colour, s);
    }
    return SVGF_UNSUPPORTED;
}

}  // namespace

svgf_status atrous_tiled(svgf_ctx *c, bool f32, int terms, const AtrousTiledArgs &a, int guide_slot, const void *in, void *out, void *hist_colour,
                  cudaStream_t s) {
    if (terms == 4) return f32 ? dispatch_atrous_tiled<true, 4>(c, a, guide_slot, in, out, hist_colour, s) : dispatch_atrous_tiled<false, 4>(c, a, guide_slot, in, out, hist_colour, s);
    if (terms == 5) return f32 ? dispatch_atrous_tiled<true, 5>(c, a, guide_slot, in, out, hist_colour, s) : dispatch_atrous_tiled<false, 5>(c, a, guide_slot, in, out, hist_colour, s);
    return SVGF_UNSUPPORTED;
}
}  // namespace svgf
