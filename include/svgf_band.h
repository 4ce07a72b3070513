/*
 * svgf_band.h — one large frame filtered by N GPUs in horizontal bands (BASELINE config 4), on top of
 * the C ABI of svgf.h.  One process (or thread) per GPU; rank g owns image rows [y0, y1) and works on a LOCAL image of rows
 * [ly0, ly1) = the band plus an apron of up to 32 rows on each side (clipped to the frame).  All planes handed to
 * svgf_band_frame - G-buffers, RenderBuffer / MomentsBuffer / FilterBuffer, HistoryLengthBuffer - are local images of
 * width x (ly1 - ly0); the caller fills the G-buffer and noisy radiance of its local rows (apron included: they are
 * inputs, nobody exchanges them) and reads the result from filter[0], rows [y0 - ly0, y1 - ly0).
 *
 * The path has no all-reduce.  Neighbouring ranks exchange, per frame, on a driver-owned side stream - by NCCL send/recv
 * (svgf_band_create), by peer-memory pulls over NVLink with one process per GPU and no NCCL (svgf_band_create_ipc), or by
 * peer copies when one process drives every band (svgf_band_create_group); the schedule and the results are the same:
 *   - before a-trous level 3 the 16 band rows of level 2's output next to each boundary, before level 4 the 32 rows of
 *     level 3's (2 * 2^i rows, reference src/Filter.cuh:571-576).  The level that PRODUCES those rows runs its boundary row
 *     blocks first, the exchange is posted, and the interior row blocks run while the rows are in flight;
 *   - the halos of levels 0..2 (2 + 4 + 8 rows), the 7x7 window of the variance pass and the reach of the motion vectors are
 *     covered by computing those stages a few rows into the apron instead (no exchange);
 *   - once per frame, overlapped with levels 1..4, the apron rows of the next frame's previous-frame state (colour
 *     history, moments, history lengths).
 * The stitched bands are BIT-identical to the whole frame filtered on one GPU (tests/test_band_driver.py, bench.py).
 *
 * Limits: 2..5 a-trous levels on the staged kernel path (even width, phi_normal >= 32, phi_depth > 0, no variance
 * prefilter), vertical motion of at most 15 pixels per frame, bands of at least 32 rows.  Anything else returns
 * SVGF_UNSUPPORTED (svgf_b200/bands.py drives the general case level by level).
 */
#ifndef SVGF_BAND_H
#define SVGF_BAND_H

#include "svgf.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SVGF_BAND_APRON 32
#define SVGF_BAND_UNIQUE_ID_BYTES 128

typedef struct svgf_band svgf_band;

/* ncclGetUniqueId: call on ONE rank and distribute the 128 bytes to the others (any transport) before svgf_band_create. */
svgf_status svgf_band_unique_id(void *id_out);

/* Collective over all `world` ranks (ncclCommInitRank).  row_bounds: world + 1 increasing row indices from 0 to
 * full_height (rank g owns [row_bounds[g], row_bounds[g + 1])), or NULL for equal heights. */
svgf_status svgf_band_create(svgf_band **out, int device, int rank, int world, int width, int full_height, svgf_storage storage,
                             const void *unique_id, const int32_t *row_bounds);
void svgf_band_destroy(svgf_band *b);

/* rows[0..3] = y0, y1 (owned rows), ly0, ly1 (rows of the local image). */
void svgf_band_rows(const svgf_band *b, int32_t rows[4]);

/* svgf_reset on the local planes (collective only in the sense that every rank must call it for the same frame). */
svgf_status svgf_band_reset(svgf_band *b, const svgf_frame_buffers *bufs, void *stream);

/* svgf_frame for this rank's band.  Asynchronous on `stream`; the exchanges run on a driver-owned stream ordered against it
 * by events.  After the call (in stream order): filter[0] band rows hold the result; render[ping_pong], moments[ping_pong]
 * and history hold the next frame's previous-frame state for the whole local image once the state exchange posted by this
 * call has completed - the next svgf_band_frame waits for it, svgf_band_sync does so explicitly. */
svgf_status svgf_band_frame(svgf_band *b, const svgf_params *params, const svgf_gbuffer gbuf[2], const svgf_frame_buffers *bufs,
                            void *stream);

/* The same frame with ALL bands inside one process - one thread driving several GPUs, or (tests) every band on one GPU: no
 * NCCL; a band's aprons are refreshed by cudaMemcpyPeerAsync from its neighbours' planes on the band's own side stream,
 * ordered by events.  svgf_band_create_group creates the `world` bands (devices[g] = CUDA device of band g; destroy each
 * with svgf_band_destroy; svgf_band_rows / reset / sync / launch_count work per band).  svgf_band_group_frame issues one
 * frame of every band: gbufs[2 * g + k] = G-buffer k of band g, bufs[g], streams[g].  The schedule per band is that of
 * svgf_band_frame (same plan, same kernels, same row ranges); the bands advance together from exchange to exchange.
 * Results are bit-identical to svgf_band_frame over NCCL and to the whole frame on one GPU. */
svgf_status svgf_band_create_group(svgf_band **out, const int32_t *devices, int world, int width, int full_height, svgf_storage storage,
                                   const int32_t *row_bounds);
svgf_status svgf_band_group_frame(svgf_band *const *bands, int world, const svgf_params *params, const svgf_gbuffer *gbufs,
                                  const svgf_frame_buffers *bufs, void *const *streams);

/* One process per GPU WITHOUT NCCL: the peer-memory transport.  Every rank exports its exchange planes and a few flag words
 * as CUDA IPC handles (svgf_band_ipc_export: SVGF_BAND_IPC_BYTES opaque bytes, to be handed to the two neighbouring ranks by
 * any means - torch.distributed.all_gather_object, MPI, a file) and maps its neighbours' (svgf_band_ipc_connect: the blob of
 * rank - 1 and of rank + 1, NULL where there is none).  svgf_band_frame then exchanges with two launches per exchange on
 * the side stream: a one-thread kernel publishes this band's "rows ready" number and waits for the neighbours', a grid pulls
 * their rows over NVLink (peer loads) into the local aprons and acknowledges, so that a producer never overwrites rows a
 * slower neighbour has not fetched.  Same plan, same kernels and bit-identical results as the NCCL transport; no SM-resident
 * proxy, no host involvement per exchange.  A neighbour that never arrives turns into a CUDA error after ~10 s (the wait
 * traps), not a hang. */
#define SVGF_BAND_IPC_BYTES 640
svgf_status svgf_band_create_ipc(svgf_band **out, int device, int rank, int world, int width, int full_height, svgf_storage storage,
                                 const int32_t *row_bounds);
svgf_status svgf_band_ipc_export(svgf_band *b, void *blob_out);
svgf_status svgf_band_ipc_connect(svgf_band *b, const void *up_blob, const void *down_blob);

/* Makes `stream` wait for every exchange this driver has posted (before reading state planes from outside). */
svgf_status svgf_band_sync(svgf_band *b, void *stream);

/* The schedule of a-trous levels 1 .. levels-1 for one band (pure function: no device, no communicator; svgf_band_frame
 * executes exactly this list).  band_lo / band_hi: local rows of the owned band [band_lo, band_hi) inside a local image of
 * local_rows rows.  Steps, in issue order:
 *   LAUNCH     level `level` over row blocks [yblock0, yblock0 + nyblocks) of its tile grid (a block = 12 * 2^level rows)
 *              and, when nyblocks1 > 0, [yblock1, yblock1 + nyblocks1) in the same launch (the two boundary strips)
 *   EXCHANGE   post the exchange of `rows` band rows of level `level`'s output with each neighbour (for level + 1)
 *   WAIT_HALO  wait for the exchange of level `level`'s `rows`-row halo: every later launch of the level may read apron rows
 *              (launches of the level BEFORE it cover only row blocks further than `rows` rows from a neighbour's edge)
 * Returns the number of steps (<= max_steps), or -1 for unsupported arguments. */
enum { SVGF_BAND_STEP_LAUNCH = 0, SVGF_BAND_STEP_EXCHANGE = 1, SVGF_BAND_STEP_WAIT_HALO = 2 };
typedef struct svgf_band_step {
    int32_t kind, level, yblock0, nyblocks, rows, yblock1, nyblocks1;
} svgf_band_step;
int svgf_band_plan(int rank, int world, int band_lo, int band_hi, int local_rows, int levels, svgf_band_step *steps, int max_steps);

/* Kernel launches issued by this band's context (bench accounting) and the last CUDA / NCCL error code seen. */
uint64_t svgf_band_launch_count(const svgf_band *b);
int svgf_band_last_error(const svgf_band *b);

#ifdef __cplusplus
}
#endif
#endif /* SVGF_BAND_H */
