/*
 * svgf_oracle.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar CPU restatement of the reference's SVGF filter math (jacquespillet/SVGF @ b1328b8,
 * src/Filter.cuh:18-83,182-263,359-624 and the host sequencing of src/App.cu:469-514,366-375), used only
 * as the parity checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs.  Nothing under svgf_b200/ may include, link or load this.
 *
 * Parity pinning: the reference ships no tests, golden vectors or fixtures for this path,
 * so this oracle is pinned by (1) hand-derived known-answer tests in tests/test_oracle_kat.py and
 * (2) the reference's OWN kernels, compiled from /root/reference/src/Filter.cuh by oracle/Makefile into
 * oracle/_ref/libsvgf_refkernels.so and compared against this restatement on a B200
 * (tests/test_reference_kernels.py, -m gpu).  Until (2) has run green on a GPU box the status is
 * "parity pinned by KATs only".
 *
 * All pointers are HOST memory.  Structs are shared with the product ABI (include/svgf.h) so the same
 * params / G-buffer descriptors drive both sides; only the type definitions are shared, no code.
 */
#ifndef SVGF_ORACLE_H
#define SVGF_ORACLE_H

#include "../include/svgf.h"

#ifdef __cplusplus
extern "C" {
#endif

/* float <-> half bit conversions (RNE, bit-exact with CUDA __float2half / __half2float, src/Filter.cuh:18-52) */
uint16_t svgf_oracle_f2h(float f);
float svgf_oracle_h2f(uint16_t h);

/* Number of OpenMP threads the row loops will use (0 = library default = all cores). */
void svgf_oracle_set_threads(int n);
int svgf_oracle_get_threads(void);

/* src/Filter.cuh:359-404 (+ LoadPreviousData :225-258).  Snapshot semantics for the history plane (D3):
 * history_prev is read, history_out is written; they must not alias. */
int svgf_oracle_temporal(const svgf_params *p, int W, int H, int storage,
                         const svgf_gbuffer *cur, const svgf_gbuffer *prev,
                         const void *prev_colour, void *cur_colour /* in place */,
                         const uint8_t *history_prev, uint8_t *history_out,
                         void *cur_moments, const void *prev_moments);

/* src/Filter.cuh:430-525 */
int svgf_oracle_variance(const svgf_params *p, int W, int H, int storage, const svgf_gbuffer *cur,
                         const void *colour_in, const void *moments, const uint8_t *history, void *colour_out);

/* src/Filter.cuh:527-624, one level: Step = 1 << level, Iteration = level.
 * history_colour_out is written iff level == 0 (and only for non-background pixels). */
int svgf_oracle_atrous_level(const svgf_params *p, int W, int H, int storage, const svgf_gbuffer *cur,
                             const void *in, void *out, void *history_colour_out, int level);

/* src/App.cu:552-556 + 491-514: temporal -> variance -> N levels; result copied into filter[0].
 * history is one plane updated with snapshot semantics (an internal copy is taken). */
int svgf_oracle_frame(const svgf_params *p, int W, int H, int storage, const svgf_gbuffer gbuf[2],
                      const svgf_frame_buffers *bufs);

#ifdef __cplusplus
}
#endif
#endif
