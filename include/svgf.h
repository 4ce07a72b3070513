/*
 * svgf.h — C ABI of the B200-native SVGF filter (libsvgf_b200.so).
 *
 * Drop-in boundary for the one hot path of jacquespillet/SVGF: the three private stage methods of
 * `application` that launch the Filter.cuh kernels (reference src/App.cu:469-514, called in order from
 * Render() src/App.cu:552-556, ping-pong flip in EndFrame() src/App.cu:366-375).  Every entry point below
 * names the reference call site / kernel it replaces.
 *
 * Conventions
 *  - All image pointers are caller-owned DEVICE memory, row-major, dense (`y*W + x`) for the colour /
 *    moments / history planes exactly like the reference's `buffer`s (src/App.cu:763-773); G-buffer planes
 *    are pitch-linear memory (row pitch in bytes, 0 = dense) or the reference's own texture objects over the GL
 *    attachments (svgf_gbuffer below).
 *  - Texel formats are the reference's (src/App.cu:746-752, resources/shaders/GBuffer.frag:62-87):
 *      position_id  : float4  world xyz, w = primitive index            (never read by the filter; may be NULL)
 *      normal_mat   : ushort4 fp16 bit patterns: unit normal xyz, w = material index
 *      uv_inst      : ushort4 fp16 bit patterns: barycentrics xyz, w = instance index ("mesh id")
 *      motion_depth : float4  xy = motion in pixels (cur -> prev), z = linear depth (0 = background),
 *                             w = max(|dFdx z|, |dFdy z|)
 *    colour planes  : half4 (SVGF_STORE_F16, reference layout, src/Filter.cuh:15) or float4 (SVGF_STORE_F32)
 *                     = (r, g, b, variance);  moments planes: half2 / float2 = (E[L], E[L^2]);
 *    history        : uint8 per pixel (src/App.cu:773).
 *  - Calls are asynchronous on the caller's stream (a `cudaStream_t` passed as void*; NULL = default
 *    stream); nothing synchronises the host.  Errors are returned, never asserted
 *    (the reference asserts, src/App.cu:41-48).
 *  - One context per (device, resolution, storage); a context is not thread-safe, different contexts are.
 *  - There is no CPU fallback: every entry point launches sm_100a kernels or fails.
 */
#ifndef SVGF_B200_H
#define SVGF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVGF_ABI_VERSION 2

typedef struct svgf_ctx svgf_ctx;

typedef enum svgf_status {
    SVGF_OK = 0,
    SVGF_INVALID_ARG = 1,
    SVGF_UNSUPPORTED = 2,
    SVGF_CUDA_ERROR = 3
} svgf_status;

/* Storage of colour+variance / moments planes.  F16 is the reference layout (src/Filter.cuh:15-16). */
typedef enum svgf_storage { SVGF_STORE_F16 = 0, SVGF_STORE_F32 = 1 } svgf_storage;

/* Mesh-id consistency test of the reprojection (src/Filter.cuh:245-247).
 * INTENDED decodes uv_inst.w (fp16) to an int instance index as GBuffer.frag:77 wrote it;
 * REFERENCE_VACUOUS reproduces what the reference's type-punned float4 fetch does on hardware
 * (both ids read as denormals -> 0 -> the test always passes). */
typedef enum svgf_mesh_id_mode { SVGF_MESH_ID_INTENDED = 0, SVGF_MESH_ID_REFERENCE_VACUOUS = 1 } svgf_mesh_id_mode;

/* History fetch.  NEAREST_TRUNC is what the reference does (src/Filter.cuh:231-232: one texel at
 * coord + ivec2(motion), C truncation).  BILINEAR is the SVGF paper's form, which the reference leaves out: the four
 * texels around p = coord + motion (x0 = floor(p.x), y0 = floor(p.y)), visited in the order (x0,y0) (x0+1,y0) (x0,y0+1)
 * (x0+1,y0+1), each subjected to the reference's own consistency tests (inside the image, depth, mesh id, normal;
 * src/Filter.cuh:235-252) against the current pixel; colour, moments and history length are the bilinear-weighted
 * means over the texels that pass (sum of w*v divided by sum of w, FP32, in that order); the reprojection fails when
 * the passing weights sum to less than 0.01.  The fetched history length is int(mean + 0.5). */
typedef enum svgf_reproj_mode { SVGF_REPROJ_NEAREST_TRUNC = 0, SVGF_REPROJ_BILINEAR = 1 } svgf_reproj_mode;

/* Variance pre-filter of the a-trous levels.  NONE is what the reference does (src/Filter.cuh:547,562: the centre
 * texel's own variance scales the luminance edge-stopping term).  GAUSS3 is the SVGF paper's form, which the reference
 * leaves out: that variance is first blurred with the 3x3 kernel (1 2 1; 2 4 2; 1 2 1) / 16 over the level's INPUT
 * plane at +-1 pixel (not dilated), separably - rows y-1, y, y+1 combined first, then columns - with coordinates
 * clamped to the image; everything else is unchanged. */
typedef enum svgf_variance_prefilter { SVGF_VARIANCE_PREFILTER_NONE = 0, SVGF_VARIANCE_PREFILTER_GAUSS3 = 1 } svgf_variance_prefilter;

/* Depth consistency test of the reprojection.  ABSOLUTE is what the reference does (src/Filter.cuh:242: |z_prev - z_cur| >
 * DepthThreshold, world units).  RELATIVE is the form the reference left commented out one line above (src/Filter.cuh:241):
 * |z_prev - z_cur| / (dz_cur + 1e-2) > DepthThreshold with dz_cur the current pixel's depth derivative (motion_depth.w; 0 for
 * background) - the threshold then scales with the slope of the surface. */
typedef enum svgf_depth_test_mode { SVGF_DEPTH_TEST_ABSOLUTE = 0, SVGF_DEPTH_TEST_RELATIVE = 1 } svgf_depth_test_mode;

/* svgf_params.flags */
#define SVGF_FLAG_NONE 0u
/* svgf_frame only: do not reuse the compact guide plane cached from the previous svgf_frame call for the
 * previous-frame consistency tests; always re-read prev_gbuf. */
#define SVGF_FLAG_NO_GUIDE_CACHE 1u
/* svgf_frame / svgf_atrous: run every a-trous level as its own launch (disables two-level fusion). */
#define SVGF_FLAG_NO_LEVEL_FUSION 2u
/* svgf_frame / svgf_atrous: run a-trous levels 0 and 1 as ONE launch with the level-0 result kept in shared memory
 * (BASELINE configs[2] "two-level-fused" variant).  Results are bit-identical to two launches.  Off by default: the
 * level is bound by arithmetic, not HBM, and the fused tile evaluates level 0 1.5 x redundantly on its apron - measured
 * slower on B200.  SVGF_FLAG_NO_LEVEL_FUSION overrides it. */
#define SVGF_FLAG_FUSE_LEVELS_01 16u
/* Run every stage with the simple one-thread-per-pixel kernels instead of the tiled ones (A/B baseline). */
#define SVGF_FLAG_BASIC_KERNELS 4u
/* a-trous: do not take the uniform-normal tile shortcut (tiles whose staged texels all carry one normal vector
 * evaluate the normal weight once per tile; results are bit-identical either way - the flag exists for A/B timing
 * and for the parity tests that pin that identity). */
#define SVGF_FLAG_NO_UNIFORM_TILES 8u

/* svgf_frame / svgf_atrous: run every a-trous level from and to storage-format planes.  By default two or more consecutive
 * levels run STAGED: the first level (packed kernel) writes its result as context-owned, pre-transformed "lattice planes"
 * (fp32, clamped, pixel-pair interleaved, luminance attached - exactly what the next level's imageLoad would have produced,
 * src/Filter.cuh:78-83), the following levels load their tiles from those with tensor-map TMA and only the last level
 * writes the storage format again.  Results are BIT-identical to the level-by-level path (same arithmetic in the same
 * order; the intermediate planes hold exactly what a storage-format round trip would have produced); the flag exists for
 * A/B timing and for the tests that pin that identity. */
#define SVGF_FLAG_NO_STAGED_LEVELS 32u
/* A/B variants of the a-trous level, both measured slower than the default: the persistent bulk-copy
 * (cp.async.bulk) scalar kernel and the warp-specialised streaming kernel. */
#define SVGF_FLAG_ATROUS_BULK 64u
#define SVGF_FLAG_ATROUS_STREAM 128u
/* Staged a-trous levels: launch every level as an ordinary stream-ordered kernel instead of with programmatic dependent
 * launch (the default lets a level's prologue - barrier setup, tensor-map prefetch, uniform-tile test - overlap the
 * previous level's tail). */
#define SVGF_FLAG_NO_DEPENDENT_LAUNCH 256u

/* svgf_band_frame only, DIAGNOSTICS: skip every neighbour exchange (pixels near the band edges come out wrong).  Shows what a
 * band's frame costs without communication and cross-rank waiting. */
#define SVGF_FLAG_BAND_NO_EXCHANGE 512u

/* Kernel family that ran an a-trous level (svgf_last_dispatch). */
typedef enum svgf_dispatch_family {
    SVGF_FAMILY_NONE = 0,
    SVGF_FAMILY_BASIC = 1,          /* one thread per pixel: levels > 4, phi_normal < 32, phi_depth == 0, odd widths with the
                                       prefilter, unaligned planes, SVGF_FLAG_BASIC_KERNELS - several times slower */
    SVGF_FAMILY_PACKED = 2,         /* shared-memory tiles, packed FP32x2 arithmetic, storage format in and out */
    SVGF_FAMILY_PACKED_STAGED = 3,  /* same, writing lattice planes for the level that follows */
    SVGF_FAMILY_LATTICE = 4,        /* TMA-staged tiles from lattice planes */
    SVGF_FAMILY_BULK = 5,           /* persistent cp.async.bulk kernel (odd widths; SVGF_FLAG_ATROUS_BULK) */
    SVGF_FAMILY_STREAM = 6,         /* SVGF_FLAG_ATROUS_STREAM */
    SVGF_FAMILY_FUSED01 = 7         /* SVGF_FLAG_FUSE_LEVELS_01 */
} svgf_dispatch_family;

/* Tunables.  Defaults (svgf_default_params) are the reference's members src/App.h:109-114, GUI ranges
 * src/GUI.cpp:988-993.  phi_depth / alpha_min / moments_alpha_min are additions whose defaults
 * reproduce the reference exactly. */
typedef struct svgf_params {
    int32_t history_cap;         /* HistoryLength      = 24   (1..255; stored into uint8, src/Filter.cuh:400) */
    float depth_threshold;       /* DepthThreshold     = 0.8  absolute, world units (src/Filter.cuh:242) */
    float normal_threshold;      /* NormalThreshold    = 0.9  (src/Filter.cuh:252) */
    float phi_colour;            /* PhiColour          = 10   (src/Filter.cuh:460,562) */
    float phi_normal;            /* PhiNormal          = 128  (src/Filter.cuh:419) */
    int32_t atrous_iterations;   /* SpatialFilterSteps = 3 in the reference; 5 in BASELINE (0..10) */
    float phi_depth;             /* multiplies both depth phis (src/Filter.cuh:461,563); 1 = reference */
    float alpha_min;             /* colour alpha  = max(1/h, alpha_min);          0 = reference (src/Filter.cuh:381) */
    float moments_alpha_min;     /* moments alpha = max(1/h, moments_alpha_min);  0 = reference */
    int32_t mesh_id_mode;        /* svgf_mesh_id_mode */
    int32_t reproj_mode;         /* svgf_reproj_mode */
    int32_t variance_prefilter;  /* svgf_variance_prefilter */
    uint32_t flags;              /* SVGF_FLAG_* */
    int32_t depth_test_mode;     /* svgf_depth_test_mode */
} svgf_params;

/* One G-buffer = the reference's `cudaFramebuffer` (src/App.h:41-44; attachment order src/App.h:33-39).  Each plane is either
 *  - pitch-linear device memory (pitch in bytes, 0 = dense), or
 *  - a cudaTextureObject_t, exactly what the reference passes (src/App.cu:473-475,485-486,504: CudaMappings[i]->TexObj,
 *    created over the GL attachment's cudaArray by src/CudaUtil.h:68-99): store the handle in the pointer field
 *    ((const void *)(uintptr_t)tex) and set the plane's pitch to SVGF_PITCH_TEXTURE.  The texture must be what CudaUtil.h
 *    creates: element read mode, point filtering, un-normalised coordinates, channel format float4 (position, motion) /
 *    ushort4 (normal, uv).  The kernels fetch it with tex2D at integer texel coordinates like src/Filter.cuh:182-207.
 * The G-buffer is read once per frame (by the temporal pass, which compacts it into the context's guide planes), so
 * texture-backed planes cost no extra pass and no copy. */
#define SVGF_PITCH_TEXTURE ((size_t)-1)
typedef struct svgf_gbuffer {
    const void *position_id;  size_t position_pitch;   /* float4;  may be NULL (not read by any filter kernel) */
    const void *normal_mat;   size_t normal_pitch;     /* ushort4 (fp16 bits) */
    const void *uv_inst;      size_t uv_pitch;         /* ushort4 (fp16 bits) */
    const void *motion_depth; size_t motion_pitch;     /* float4 */
} svgf_gbuffer;

/* The reference's ping-pong buffer set (src/App.h:136-139) for svgf_frame. */
typedef struct svgf_frame_buffers {
    void *render[2];     /* RenderBuffer[0..1]  colour planes; [ping_pong] holds this frame's noisy radiance on entry */
    void *moments[2];    /* MomentsBuffer[0..1] */
    void *filter[2];     /* FilterBuffer[0..1]; the result lands in filter[0], filter[1] is scratch */
    uint8_t *history;    /* HistoryLengthBuffer */
    int32_t ping_pong;   /* PingPongInx: index of the CURRENT frame's render/moments/G-buffer */
} svgf_frame_buffers;

/* ABI version of the loaded library (== SVGF_ABI_VERSION it was built with). */
int svgf_abi_version(void);
const char *svgf_status_string(svgf_status s);

/* Reference defaults: src/App.h:109-114 with atrous_iterations = 5 (BASELINE.json) and the added knobs neutral. */
void svgf_default_params(svgf_params *p);

/* Replaces application::ResizeRenderTextures()'s filter part (src/App.cu:742-778): the context owns only
 * scratch — the history shadow plane (race-free snapshot semantics for src/Filter.cuh:255 vs :400), two
 * compact guide planes; the lattice planes and tensor maps of the staged a-trous levels are allocated on first use.
 * `device` is a CUDA ordinal. */
svgf_status svgf_create(svgf_ctx **out, int device, int width, int height, svgf_storage storage);
void svgf_destroy(svgf_ctx *ctx);

/* First-frame semantics, undefined in the reference (raw cudaMalloc, src/Buffer.cpp:20): zero-fills both
 * render/moments planes, the history plane and the context's cached previous-frame guide, so that every
 * pixel of the next frame fails reprojection (h = 1, alpha = 1). */
svgf_status svgf_reset(svgf_ctx *ctx, const svgf_frame_buffers *bufs, void *stream);

/* Replaces application::TemporalFilter() (src/App.cu:469-478) -> filter::TemporalFilter (src/Filter.cuh:359-404).
 * cur_colour is updated in place to (accumulated rgb, variance); history is updated in place but with
 * snapshot semantics (reads see the previous frame's values). */
svgf_status svgf_temporal(svgf_ctx *ctx, const svgf_params *params,
                          const svgf_gbuffer *cur_gbuf, const svgf_gbuffer *prev_gbuf,
                          const void *prev_colour, void *cur_colour, uint8_t *history,
                          void *cur_moments, const void *prev_moments, void *stream);

/* Replaces application::FilterMoments() (src/App.cu:480-489) -> filter::FilterMoments (src/Filter.cuh:430-525).
 * `moments` is explicit: the reference always passes MomentsBuffer[0] (src/App.cu:484). */
svgf_status svgf_variance(svgf_ctx *ctx, const svgf_params *params, const svgf_gbuffer *cur_gbuf,
                          const void *colour_in, const void *moments, const uint8_t *history,
                          void *colour_out, void *stream);

/* Replaces application::WaveletFilter()'s loop body (src/App.cu:497-508) -> filter::FilterKernel
 * (src/Filter.cuh:527-624) for levels first_level .. first_level+num_levels-1 (Step = 1 << level).
 * Ping-pongs between buf_a (input of first_level) and buf_b; the final level's output pointer is returned in
 * *result (either buf_a or buf_b).  history_colour_out receives the level-0 output (non-background
 * pixels only, src/Filter.cuh:554-558,619-622) when level 0 is among the levels run; may be NULL otherwise. */
svgf_status svgf_atrous(svgf_ctx *ctx, const svgf_params *params, const svgf_gbuffer *cur_gbuf,
                        void *buf_a, void *buf_b, void *history_colour_out,
                        int first_level, int num_levels, void **result, void *stream);

/* Replaces the stage sequence of application::Render() (src/App.cu:552-556): temporal -> variance ->
 * atrous_iterations levels.  gbuf[ping_pong] is the current G-buffer, gbuf[1-ping_pong] the previous.
 * After the call: filter[0] = final result; render[ping_pong] = next frame's colour history (level-0 output;
 * temporal output for background pixels or when atrous_iterations == 0); moments[ping_pong], history
 * updated.  Stages may be fused internally; the buffers named here match the unfused sequence.
 * The caller flips ping_pong afterwards (EndFrame, src/App.cu:374). */
svgf_status svgf_frame(svgf_ctx *ctx, const svgf_params *params, const svgf_gbuffer gbuf[2],
                       const svgf_frame_buffers *bufs, void *stream);

/* Replaces application::TAA() (src/App.cu:516-522) -> filter::TAAFilterKernel (src/Filter.cuh:288-357): temporal
 * anti-aliasing of the filtered frame against the previous resolve (3x3 neighbourhood clamp in PAL-YUV, :267-285) and
 * sRGB encoding (:145-148); the output alpha is 1.  `filtered` is svgf_frame's result (filter[0]); `taa_history` is the
 * PREVIOUS call's taa_out (all zeros before the first frame: alpha 0 makes the blend take none of the current texel, as
 * in the reference); `taa_out` receives this frame's display image.  The reference passes one plane in both roles and
 * reads texels other threads may already have overwritten (src/Filter.cuh:299 vs :355); here the two must be distinct
 * planes that the caller swaps every frame (snapshot semantics). */
svgf_status svgf_taa(svgf_ctx *ctx, const void *filtered, const void *taa_history, void *taa_out, void *stream);

/* Albedo demodulation, the step of the SVGF paper the reference leaves out (README.md:14,172-174): filter the untextured
 * illumination so that textures are not blurred.  Two element-wise passes around the frame, both optional (nothing else
 * in the path changes, so without them the reference's behaviour is untouched):
 *   svgf_demodulate : colour.rgb <- colour.rgb / max(albedo.rgb, 1e-3)   before svgf_frame / svgf_temporal (in place; w kept)
 *   svgf_remodulate : out.rgb    <- filtered.rgb * albedo.rgb            after the last a-trous level (w copied)
 * `albedo` is a plane in the storage texel format, rgb = surface albedo of the pixel (w ignored).  IEEE division and
 * multiplication, one rounding each, then the storage format's rounding.  Like every plane of the reference the
 * illumination is clamped to [0,1] by the passes in between (src/Filter.cuh:63-83): scale HDR input accordingly. */
svgf_status svgf_demodulate(svgf_ctx *ctx, const void *albedo, void *colour, void *stream);
svgf_status svgf_remodulate(svgf_ctx *ctx, const void *albedo, const void *filtered, void *out, void *stream);

/* The context caches a compact "guide" plane (depth, depth derivative, normal, mesh id: 22 B/px) per
 * G-buffer, keyed by the three plane pointers and pitches of the svgf_gbuffer: svgf_temporal / svgf_frame build it for the current
 * G-buffer, the following svgf_variance / svgf_atrous calls naming the same G-buffer reuse it, and the next
 * frame's temporal pass reads it as the previous-frame guide (the reference's two ping-ponged framebuffers,
 * src/App.cu:745-746, satisfy this by construction).  Call this after modifying a G-buffer in place outside
 * that order; svgf_reset implies it. */
void svgf_invalidate_guide(svgf_ctx *ctx);

/* Stage timing for benchmarks: between begin and end every svgf_frame records four CUDA events on the
 * caller's stream (before temporal / variance / a-trous, after a-trous; up to 4096 frames).  end
 * synchronises the last event and returns summed GPU milliseconds per stage {temporal, variance, atrous}. */
svgf_status svgf_profile_begin(svgf_ctx *ctx);
svgf_status svgf_profile_end(svgf_ctx *ctx, double stage_ms[3], int *frames);

/* Last CUDA error code seen by this context (cudaError_t as int), 0 if none. */
int svgf_last_cuda_error(const svgf_ctx *ctx);

/* Number of kernel launches issued through this context so far (for bench accounting). */
uint64_t svgf_launch_count(const svgf_ctx *ctx);

/* Which kernel family (svgf_dispatch_family) ran each a-trous level of the most recent svgf_frame / svgf_atrous call:
 * families[level] for level < min(return value, max_levels); returns the number of levels recorded.  Makes the slow
 * per-pixel fallbacks (SVGF_FAMILY_BASIC) visible instead of silent. */
int svgf_last_dispatch(const svgf_ctx *ctx, int32_t *families, int max_levels);

/* Host-buffer path (what the end-to-end benchmark times, and what a caller without device memory of its own
 * uses): one frame's inputs host->device, svgf_frame, the result device->host.  `h_*` are pinned (or pageable)
 * HOST pointers in the same texel formats; the previous frame's state stays resident in the context between
 * calls (reset=1 starts a new sequence).
 * The call is asynchronous and PIPELINED across calls: inputs are staged into a 3-slot device ring on a
 * context-owned copy-in stream, the kernels run on `stream`, the result leaves on a context-owned copy-out
 * stream, and `stream` is made to wait for that copy - so the PCIe transfers of frames t+1 and t-1 overlap the
 * kernels of frame t, and synchronising `stream` (or an event recorded on it after the call) guarantees that
 * h_result / h_history_out of every earlier call are complete and that the h_* inputs may be reused.  Host inputs
 * must stay valid and unmodified until then. */
svgf_status svgf_frame_host(svgf_ctx *ctx, const svgf_params *params,
                            const void *h_normal_mat, const void *h_uv_inst, const void *h_motion_depth,
                            const void *h_noisy_colour, void *h_result, uint8_t *h_history_out,
                            int reset, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SVGF_B200_H */
