// svgf_tu_lattice.inl — instantiations and launch of the TMA-staged lattice level for ONE storage type
// (included by svgf_tu_lattice_f16.cu / svgf_tu_lattice_f32.cu with SVGF_TU_F32 defined)
#include "svgf_ctx.h"
#include "svgf_kernels_lattice.cuh"

namespace svgf {
namespace {
template <bool F32, int STEP, int TERMS, bool LAST>
svgf_status launch_lattice(svgf_ctx *c, const AtrousTiledArgs &t, int guide_slot, int src, void *out, bool pdl, cudaStream_t s) {
    using CT = typename ColourPlane<F32>::texel;
    using G = LatGeom<STEP>;
    auto kern = atrous_lattice_kernel<F32, STEP, TERMS, LAST>;
    static std::atomic<unsigned long long> configured{0};
    SVGF_CUDA(c, configure_smem_once(configured, c->device, kern, G::smem_bytes));
    const svgf_ctx::Lattice &L = c->lat;
    LatticeArgs a;
    a.W = c->W; a.H = c->H;
    a.pitch_pairs = L.pitch_pairs;
    a.segs_x = (c->W + 31) / 32;
    a.kL_scale = t.kL_scale; a.kZ_scale = t.kZ_scale;
    a.k1 = t.k1; a.k2 = t.k2; a.k3 = t.k3; a.k4 = t.k4; a.k5 = t.k5;
    a.uniform_tiles = t.uniform_tiles;
    a.yblock0 = t.yblock0; a.nyblocks0 = t.nyblocks1 > 0 ? t.nyblocks : 0; a.yblock1 = t.yblock1;
    const CUtensorMap *m = L.map[t.level];
    const int q = src * 3;
    cudaLaunchConfig_t cfg = {};
    const int all_yblocks = (c->H + G::block_rows * STEP - 1) / (G::block_rows * STEP);
    cfg.gridDim = dim3((c->W + kTileW - 1) / kTileW, (t.nyblocks > 0 ? t.nyblocks + t.nyblocks1 : all_yblocks) * G::subtiles * STEP);
    cfg.blockDim = dim3(G::threads);
    cfg.dynamicSmemBytes = G::smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    SVGF_CUDA(c, cudaLaunchKernelEx(&cfg, kern, m[q + 0], m[q + 1], m[q + 2], m[6], m[7], a, (const float *)c->guide[guide_slot].dz,
                                    (const float4 *)c->guide[guide_slot].seg, L.sc[1 - src], (CT *)out));
    c->launches++;
    return SVGF_OK;
}
template <bool F32, int TERMS, bool LAST>
svgf_status dispatch_lattice(svgf_ctx *c, const AtrousTiledArgs &a, int guide_slot, int src, void *out, bool pdl, cudaStream_t s) {
    switch (a.level) {
        case 1: return launch_lattice<F32, 2, TERMS, LAST>(c, a, guide_slot, src, out, pdl, s);
        case 2: return launch_lattice<F32, 4, TERMS, LAST>(c, a, guide_slot, src, out, pdl, s);
        case 3: return launch_lattice<F32, 8, TERMS, LAST>(c, a, guide_slot, src, out, pdl, s);
        case 4: return launch_lattice<F32, 16, TERMS, LAST>(c, a, guide_slot, src, out, pdl, s);
    }
    return SVGF_UNSUPPORTED;
}
}  // namespace

svgf_status SVGF_TU_LATTICE_ENTRY(svgf_ctx *c, int terms, const AtrousTiledArgs &a, int guide_slot, int src, void *out, bool pdl, cudaStream_t s) {
    constexpr bool F32 = SVGF_TU_F32;
    if (terms == 3) return out ? dispatch_lattice<F32, 3, true>(c, a, guide_slot, src, out, pdl, s) : dispatch_lattice<F32, 3, false>(c, a, guide_slot, src, out, pdl, s);
    if (terms == 5) return out ? dispatch_lattice<F32, 5, true>(c, a, guide_slot, src, out, pdl, s) : dispatch_lattice<F32, 5, false>(c, a, guide_slot, src, out, pdl, s);
    return SVGF_UNSUPPORTED;
}
}  // namespace svgf
