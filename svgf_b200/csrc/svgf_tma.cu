// svgf_tma.cu — host side of the TMA-staged a-trous levels: the context-owned lattice planes (svgf_kernels_lattice.cuh)
// and their tensor maps.  cuTensorMapEncodeTiled is a driver-API function; it is fetched through the runtime
// (cudaGetDriverEntryPoint), so the library links against libcudart only.
#include <cuda.h>
#include <cuda_runtime.h>

#include "svgf_ctx.h"

namespace svgf {
namespace {

using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// Null texels everywhere (colour 0, z = +inf, zero normal); the producers only ever write texels of the image, so the
// padding keeps these values for the life of the context.
__global__ void __launch_bounds__(256)
lattice_clear_kernel(LatticeColour a, LatticeColour b, LatticeNormals n, size_t npairs) {
    const float inf = __int_as_float(0x7f800000);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f), nul = make_float4(0.f, 0.f, inf, inf);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npairs; i += (size_t)gridDim.x * blockDim.x) {
        a.c0[i] = z4; a.c1[i] = z4; a.lz[i] = nul;
        b.c0[i] = z4; b.c1[i] = z4; b.lz[i] = nul;
        n.n0[i] = z4; n.n1[i] = make_float2(0.f, 0.f);
    }
}

// One plane at one dilation: the padded plane [rows][pitch_pairs] of `elems_per_pair` 8-byte elements viewed as
// {pitch_pairs * elems_per_pair, step, rows / step}; box = the tile's staged rows (16) of one phase x the tile's pairs.
bool encode_plane(CUtensorMap *map, void *base, int pitch_pairs, int rows, int elems_per_pair, int step) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return false;
    const cuuint64_t row_elems = (cuuint64_t)pitch_pairs * elems_per_pair;
    const cuuint64_t dims[3] = {row_elems, (cuuint64_t)step, (cuuint64_t)(rows / step)};
    const cuuint64_t strides[2] = {row_elems * 8, row_elems * 8 * (cuuint64_t)step};   // bytes, dims 1 and 2
    const cuuint32_t box[3] = {(cuuint32_t)((kTileW / 2 + 2 * step) * elems_per_pair), 1u, (cuuint32_t)(3 * kLatRowGroups + 4)};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

void lattice_destroy(svgf_ctx *c) {
    svgf_ctx::Lattice &L = c->lat;
    for (int k = 0; k < 2; k++) { cudaFree(L.sc[k].c0); cudaFree(L.sc[k].c1); cudaFree(L.sc[k].lz); }
    cudaFree(L.sn.n0); cudaFree(L.sn.n1);
    L = svgf_ctx::Lattice();
}

svgf_status lattice_prepare(svgf_ctx *c, cudaStream_t s) {
    svgf_ctx::Lattice &L = c->lat;
    if (L.ready) return SVGF_OK;
    if (L.failed) return SVGF_UNSUPPORTED;
    // rows cover the tallest tile grid of any level (12 * 16-row blocks) plus the padding; both multiples of 16, so
    // every dilation 1..16 divides them
    const int Wt = (c->W + kTileW - 1) / kTileW * kTileW, Ht = (c->H + 191) / 192 * 192;
    L.pitch_pairs = (kLatPadX + Wt + kLatPadX) / 2;
    L.rows = kLatPadY + Ht + kLatPadY;
    L.npairs = (size_t)L.pitch_pairs * L.rows;
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < 2 && e == cudaSuccess; k++) {
        e = cudaMalloc(&L.sc[k].c0, L.npairs * 16);
        if (e == cudaSuccess) e = cudaMalloc(&L.sc[k].c1, L.npairs * 16);
        if (e == cudaSuccess) e = cudaMalloc(&L.sc[k].lz, L.npairs * 16);
    }
    if (e == cudaSuccess) e = cudaMalloc(&L.sn.n0, L.npairs * 16);
    if (e == cudaSuccess) e = cudaMalloc(&L.sn.n1, L.npairs * 8);
    if (e != cudaSuccess) {
        lattice_destroy(c);
        L.failed = true;
        return svgf_cuda_fail(c, e);
    }
    lattice_clear_kernel<<<c->num_sms * 8, 256, 0, s>>>(L.sc[0], L.sc[1], L.sn, L.npairs);   // one-time setup: not counted as a launch
    e = cudaGetLastError();
    bool ok = e == cudaSuccess;
    for (int level = 1; level <= 4 && ok; level++) {
        const int step = 1 << level;
        void *planes[8] = {L.sc[0].c0, L.sc[0].c1, L.sc[0].lz, L.sc[1].c0, L.sc[1].c1, L.sc[1].lz, L.sn.n0, L.sn.n1};
        for (int q = 0; q < 8 && ok; q++) ok = encode_plane(&L.map[level][q], planes[q], L.pitch_pairs, L.rows, q == 7 ? 1 : 2, step);
    }
    if (!ok) {   // no tensor maps (driver too old?): the packed kernel keeps serving every level
        lattice_destroy(c);
        L.failed = true;
        return e != cudaSuccess ? svgf_cuda_fail(c, e) : SVGF_UNSUPPORTED;
    }
    L.ready = true;
    return SVGF_OK;
}

}  // namespace svgf
