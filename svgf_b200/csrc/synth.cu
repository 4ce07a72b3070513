// synth.cu — C entry points of the procedural input generator (libsvgf_synth.so): the same per-pixel
// function (synth_scene.h) run either by a CUDA kernel into device planes or by an OpenMP loop into host
// planes.  Compiled with -fmad=false / -ffp-contract=off so both produce identical bits.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <omp.h>
#include <string.h>

#include "synth_scene.h"

namespace {

struct Planes {
    float4 *position;   // may be null
    ushort4 *normal;
    ushort4 *uv;
    float4 *motion;
    void *colour;       // half4 bits (ushort4) or float4
};

SYNTH_HD void store_texel(const Planes &pl, size_t i, const synth::Texel &t, int storage) {
    if (pl.position) pl.position[i] = make_float4(t.pos[0], t.pos[1], t.pos[2], t.pos[3]);
    if (pl.normal) pl.normal[i] = make_ushort4(t.nrm[0], t.nrm[1], t.nrm[2], t.nrm[3]);
    if (pl.uv) pl.uv[i] = make_ushort4(t.uv[0], t.uv[1], t.uv[2], t.uv[3]);
    if (pl.motion) pl.motion[i] = make_float4(t.mot[0], t.mot[1], t.mot[2], t.mot[3]);
    if (pl.colour) {
        if (storage == 1) {
            ((float4 *)pl.colour)[i] = make_float4(t.col[0], t.col[1], t.col[2], t.col[3]);
        } else {
            ((ushort4 *)pl.colour)[i] = make_ushort4(synth::f2h_bits(t.col[0]), synth::f2h_bits(t.col[1]),
                                                     synth::f2h_bits(t.col[2]), synth::f2h_bits(t.col[3]));
        }
    }
}

__global__ void __launch_bounds__(256) synth_kernel(svgf_synth_cfg cfg, Planes pl) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= cfg.width || y >= cfg.height) return;
    const synth::Texel t = synth::shade_pixel(cfg, x, y);
    store_texel(pl, (size_t)y * cfg.width + x, t, cfg.storage);
}

bool cfg_ok(const svgf_synth_cfg *c) {
    return c && c->width > 0 && c->height > 0 && c->frame >= 0 && (c->storage == 0 || c->storage == 1);
}

}  // namespace

extern "C" {

// Host planes, dense (pitch = W * texel).  Any plane pointer may be NULL to skip it.  threads <= 0: all cores.
int svgf_synth_frame_host(const svgf_synth_cfg *cfg, void *position, void *normal, void *uv, void *motion, void *colour,
                          int threads) {
    if (!cfg_ok(cfg)) return 1;
    Planes pl{(float4 *)position, (ushort4 *)normal, (ushort4 *)uv, (float4 *)motion, colour};
    const svgf_synth_cfg c = *cfg;
#pragma omp parallel for schedule(dynamic, 8) num_threads(threads > 0 ? threads : omp_get_max_threads())
    for (int y = 0; y < c.height; y++)
        for (int x = 0; x < c.width; x++) store_texel(pl, (size_t)y * c.width + x, synth::shade_pixel(c, x, y), c.storage);
    return 0;
}

// Rows [y0, y1) of the frame only, into host planes of (y1 - y0) x W texels (bounded CPU samples of large frames).
int svgf_synth_rows_host(const svgf_synth_cfg *cfg, int y0, int y1, void *position, void *normal, void *uv, void *motion,
                         void *colour, int threads) {
    if (!cfg_ok(cfg) || y0 < 0 || y1 > cfg->height || y0 >= y1) return 1;
    Planes pl{(float4 *)position, (ushort4 *)normal, (ushort4 *)uv, (float4 *)motion, colour};
    const svgf_synth_cfg c = *cfg;
#pragma omp parallel for schedule(dynamic, 8) num_threads(threads > 0 ? threads : omp_get_max_threads())
    for (int y = y0; y < y1; y++)
        for (int x = 0; x < c.width; x++)
            store_texel(pl, (size_t)(y - y0) * c.width + x, synth::shade_pixel(c, x, y), c.storage);
    return 0;
}

// Device planes; asynchronous on `stream` (cudaStream_t).  Returns a cudaError_t as int.
int svgf_synth_frame_device(const svgf_synth_cfg *cfg, void *position, void *normal, void *uv, void *motion, void *colour,
                            void *stream) {
    if (!cfg_ok(cfg)) return 1;
    Planes pl{(float4 *)position, (ushort4 *)normal, (ushort4 *)uv, (float4 *)motion, colour};
    dim3 grid((cfg->width + 31) / 32, (cfg->height + 7) / 8);
    synth_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*cfg, pl);
    return (int)cudaGetLastError();
}

// ---- texture-backed G-buffers: what the reference's GL interop hands the filter (src/CudaUtil.h:68-99: one cudaArray per
// attachment, a texture object with a zeroed cudaTextureDesc - element reads, point filter, un-normalised coordinates).
// Here the arrays are plain cudaMallocArray allocations filled from linear device planes; no GL involved.
struct svgf_synth_texgbuf {
    int W, H;
    cudaArray_t arr[3];               // normal, uv, motion
    cudaTextureObject_t tex[3];
};

int svgf_synth_texgbuf_create(int W, int H, svgf_synth_texgbuf **out) {
    if (!out || W <= 0 || H <= 0) return 1;
    svgf_synth_texgbuf *t = new svgf_synth_texgbuf();
    t->W = W; t->H = H;
    for (int i = 0; i < 3; i++) { t->arr[i] = nullptr; t->tex[i] = 0; }
    for (int i = 0; i < 3; i++) {
        const cudaChannelFormatDesc d = (i == 2) ? cudaCreateChannelDesc<float4>() : cudaCreateChannelDesc<ushort4>();
        cudaError_t e = cudaMallocArray(&t->arr[i], &d, W, H);
        if (e == cudaSuccess) {
            cudaResourceDesc rd;
            memset(&rd, 0, sizeof(rd));
            rd.resType = cudaResourceTypeArray;
            rd.res.array.array = t->arr[i];
            cudaTextureDesc td;
            memset(&td, 0, sizeof(td));                  // src/CudaUtil.h:88-95
            td.readMode = cudaReadModeElementType;
            e = cudaCreateTextureObject(&t->tex[i], &rd, &td, nullptr);
        }
        if (e != cudaSuccess) {
            for (int k = 0; k < 3; k++) { if (t->tex[k]) cudaDestroyTextureObject(t->tex[k]); if (t->arr[k]) cudaFreeArray(t->arr[k]); }
            delete t;
            return (int)e;
        }
    }
    *out = t;
    return 0;
}

void svgf_synth_texgbuf_destroy(svgf_synth_texgbuf *t) {
    if (!t) return;
    for (int k = 0; k < 3; k++) { if (t->tex[k]) cudaDestroyTextureObject(t->tex[k]); if (t->arr[k]) cudaFreeArray(t->arr[k]); }
    delete t;
}

// Dense linear DEVICE planes -> the arrays (asynchronous on `stream`).
int svgf_synth_texgbuf_upload(svgf_synth_texgbuf *t, const void *normal, const void *uv, const void *motion, void *stream) {
    if (!t || !normal || !uv || !motion) return 1;
    const size_t W = (size_t)t->W;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemcpy2DToArrayAsync(t->arr[0], 0, 0, normal, W * 8, W * 8, t->H, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpy2DToArrayAsync(t->arr[1], 0, 0, uv, W * 8, W * 8, t->H, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpy2DToArrayAsync(t->arr[2], 0, 0, motion, W * 16, W * 16, t->H, cudaMemcpyDeviceToDevice, s);
    return (int)e;
}

// tex[0..2] = cudaTextureObject_t of normal, uv, motion
void svgf_synth_texgbuf_objects(const svgf_synth_texgbuf *t, unsigned long long tex[3]) {
    for (int k = 0; k < 3; k++) tex[k] = t ? (unsigned long long)t->tex[k] : 0ull;
}

}  // extern "C"
