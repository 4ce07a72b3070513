// svgf_kernels_stream.cuh — the a-trous level as a STREAMING stencil: warp-specialised, register sliding window.
//
// ncu on the tiled/packed kernel (profiles/atrous_packed_r01.*) shows the level bound by the shared-memory data
// pipe (63 % of peak wavefronts, L1TEX 78 %) ahead of the FP32 pipe (56 %) and far ahead of HBM (19 %): with a
// 3-row register block every staged texel is re-read from shared memory for 2.06 taps only.  This kernel removes
// that re-reading:
//   * a consumer thread owns ONE pixel-pair column of a 256-pixel strip and marches down the rows of one lattice
//     phase (rows y = phase + STEP*j).  It keeps FIVE outputs in flight (rows j-2 .. j+2): every texel of tap row j
//     is read from shared memory once and applied to all five (dy = +2 .. -2) — 5 taps per read instead of 2,
//     and no vertical halo is ever re-loaded (each row is staged once per strip);
//   * one PRODUCER warp per CTA stages rows into a ring of shared-memory slots (coalesced 16-byte loads, fp16->fp32
//     decode, the reference's [0,1] clamp, un-contracted luminance, pixel-pair interleave for the packed FP32x2
//     arithmetic, null texels outside the image) and hands them to the four consumer warps through full/empty
//     mbarriers — global-memory latency never appears on the consumers' critical path and there is no CTA-wide
//     barrier anywhere;
//   * work is the stream of all (strip, phase, row) triples cut into gridDim.x equal contiguous ranges, so every
//     CTA gets the same number of rows (+-1) at every dilation and a CTA's range is at most a few vertically
//     contiguous pieces (each piece costs 4 halo rows);
//   * the five in-flight outputs live in registers with static indices: the row loop is unrolled by five and the
//     accumulator slot of output r is r mod 5 (no register rotation); to keep that unrolled body inside the
//     instruction cache the loop over the four off-centre tap columns is rolled, with the per-column kernel-weight
//     and 1/length constants read from a small __constant__ table (dx = 0 keeps immediates);
//   * the two roles are two warpgroups with different register budgets (setmaxnreg: consumers 200, producers 56),
//     two CTAs per SM: every SM sub-partition runs two consumer warps and two (mostly sleeping) producer warps;
//   * level 0 (STEP = 1) has odd tap offsets: the producer stages a second, one-pixel-shifted copy of every row so
//     that all five tap columns are aligned pair loads there too.
// Arithmetic is the packed FP32x2 tap of svgf_kernels_packed.cuh (same instructions, same roundings).
#pragma once
#include "svgf_device.cuh"
#include "svgf_kernels_packed.cuh"

namespace svgf {

constexpr int kStripW = 256;                    // output pixels per strip row
constexpr int kStreamConsumers = kStripW / 2;   // warpgroup 0: one pixel pair per consumer thread
constexpr int kStreamProducers = 128;           // warpgroup 1
constexpr int kStreamThreads = kStreamConsumers + kStreamProducers;
// register split between the two warpgroups (setmaxnreg): 2 CTAs/SM -> per SM sub-partition 2 consumer + 2 producer
// warps: 2*32*200 + 2*32*56 = 16384 = the sub-partition's register file
constexpr int kConsumerRegs = 200, kProducerRegs = 56;

template <int STEP> struct StreamGeom {
    static constexpr int halo_pairs = STEP;                       // 2*STEP pixels on each side
    static constexpr int np = kStripW / 2 + 2 * halo_pairs;       // pixel pairs per staged row
    static constexpr int copies = (STEP == 1) ? 2 : 1;            // STEP 1: aligned pairs + pairs shifted by one pixel
    static constexpr int row_f4 = np * 5 * copies;                // float4 per row slot: C0 C1 G0 G1 LD (x copies)
    static constexpr int ring = (STEP == 1) ? 5 : 8;              // staged rows resident per CTA (3 are in use)
    static constexpr size_t smem_bytes = (size_t)ring * row_f4 * 16 + 2 * ring * 8;
};

struct AtrousStreamArgs {
    AtrousTiledArgs t;       // W, H, scales, series coefficients, level
    int n_strips;            // ceil(W / 256)
    int rows_a, rows_b;      // H = rows_a * STEP + rows_b
};

// -log2(kernel weight) and 1/length of tap (|dx|, |dy|), for the rolled tap-column loop.  Entry (0,0) is +inf: the
// centre tap then gets the weight 2^-inf = 0 exactly, i.e. it is skipped like the reference does (src/Filter.cuh:584)
__constant__ float kTapCk[3][3] = {{__builtin_huge_valf(), 0.584962458f, 2.58496246f}, {0.584962458f, 1.16992489f, 3.16992489f}, {2.58496246f, 3.16992489f, 5.16992489f}};
__constant__ float kTapCinv[3][3] = {{1.0f, 1.0f, 0.5f}, {1.0f, 0.70710678f, 0.44721360f}, {0.5f, 0.44721360f, 0.35355339f}};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// producer-side wait: the producers run several rows ahead and spend most of their time here, so back off with
// nanosleep instead of spinning through the issue slots the consumer warps of the same sub-partition need
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done) __nanosleep(400);
    } while (!done);
}
// "slot is free again" goes through hardware named barriers (ids 1 .. ring): the consumers ARRIVE when they are done with
// a row, the producers SYNC before they refill the slot and block in the barrier unit — no polling, no issue slots.
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// position of stream index k (0 <= k < H) inside a strip: phases in order, rows of a phase in order
template <int STEP>
__device__ __forceinline__ void stream_decode(int k, int rows_a, int rows_b, int &ph, int &j, int &rows) {
    const int big = rows_b * (rows_a + 1);
    if (k < big) { ph = k / (rows_a + 1); j = k - ph * (rows_a + 1); rows = rows_a + 1; }
    else { const int kk = k - big; const int p = kk / rows_a; ph = rows_b + p; j = kk - p * rows_a; rows = rows_a; }
}

struct StagedPx { float r, g, b, v, z, nx, ny, nz, l; };

template <bool F32>
__device__ __forceinline__ StagedPx stage_px(typename ColourPlane<F32>::texel c, float4 g) {
    const float4 d = ColourPlane<F32>::decode(c);
    StagedPx s;
    s.r = __saturatef(d.x); s.g = __saturatef(d.y); s.b = __saturatef(d.z); s.v = __saturatef(d.w);      // :543,:586
    s.z = g.x; s.nx = g.y; s.ny = g.z; s.nz = g.w;
    s.l = luminance(s.r, s.g, s.b);
    return s;
}
__device__ __forceinline__ void store_pair(float4 *row, int np, int p, const StagedPx &a, const StagedPx &b, float dz0, float dz1) {
    row[p] = make_float4(a.r, b.r, a.g, b.g);
    row[np + p] = make_float4(a.b, b.b, a.v, b.v);
    row[2 * np + p] = make_float4(a.z, b.z, a.nx, b.nx);
    row[3 * np + p] = make_float4(a.ny, b.ny, a.nz, b.nz);
    row[4 * np + p] = make_float4(a.l, b.l, dz0, dz1);
}
__device__ __forceinline__ PkTap load_tap(const float4 *row, int np, int si) {
    const float4 c0 = row[si], c1 = row[np + si], g0 = row[2 * np + si], g1 = row[3 * np + si];
    const float2 l = *reinterpret_cast<const float2 *>(row + 4 * np + si);
    PkTap t;
    t.r = make_float2(c0.x, c0.y); t.g = make_float2(c0.z, c0.w); t.b = make_float2(c1.x, c1.y); t.v = make_float2(c1.z, c1.w);
    t.z = make_float2(g0.x, g0.y); t.nx = make_float2(g0.z, g0.w); t.ny = make_float2(g1.x, g1.y); t.nz = make_float2(g1.z, g1.w);
    t.l = l;
    return t;
}

template <bool F32, int STEP, int TERMS>
__global__ void __launch_bounds__(kStreamThreads, 2)
atrous_stream_kernel(AtrousStreamArgs a, const float4 *__restrict__ guide_n, const float *__restrict__ guide_dz,
                     const typename ColourPlane<F32>::texel *__restrict__ in, typename ColourPlane<F32>::texel *__restrict__ out,
                     typename ColourPlane<F32>::texel *__restrict__ hist_colour) {
    using G = StreamGeom<STEP>;
    using CT = typename ColourPlane<F32>::texel;
    constexpr int kRing = G::ring;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *ring = reinterpret_cast<float4 *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t)kRing * G::row_f4);
    uint64_t *empty = full + kRing;

    const int tid = threadIdx.x;
    const int W = a.t.W, H = a.t.H;
    if (tid == 0) {
        for (int s = 0; s < kRing; s++) { mbar_init(full + s, kStreamProducers / 32); mbar_init(empty + s, kStreamConsumers / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // this CTA's contiguous range of the (strip, phase, row) stream
    const long long T = (long long)a.n_strips * H;
    const long long i0 = T * blockIdx.x / gridDim.x, i1 = T * (blockIdx.x + 1) / gridDim.x;

    if (tid >= kStreamConsumers) {
        // ================================ producer warpgroup ================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kProducerRegs));
        const int pt = tid - kStreamConsumers;
        const float4 null_g = make_float4(__int_as_float(0x7f800000), 0.f, 0.f, 0.f);   // null texel: z = +inf, zero normal
        unsigned int n = 0;    // staged rows so far
        for (long long i = i0; i < i1;) {
            const int xs = (int)(i / H), k = (int)(i - (long long)xs * H);
            int ph, j0, rows;
            stream_decode<STEP>(k, a.rows_a, a.rows_b, ph, j0, rows);
            const long long left = i1 - i;
            const int L = (int)((rows - j0) < left ? (rows - j0) : left);
            const int x0 = xs * kStripW;
            constexpr int kAhead = 6;     // rows prefetched into L2 ahead of the row being staged
            constexpr bool kTwo = (G::np > kStreamProducers);   // threads below np - 128 stage a second pair
            auto prefetch_row = [&](int q) {
                const int gy = ph + (j0 + q) * STEP;
                if (gy < 0 || gy >= H) return;
#pragma unroll
                for (int h = 0; h < (kTwo ? 2 : 1); h++) {
                    const int p = pt + h * kStreamProducers;
                    const int gx = x0 - 2 * STEP + 2 * p;
                    if (p < G::np && gx >= 0 && gx < W) {
                        const size_t gi = (size_t)gy * W + gx;
                        prefetch_l2(in + gi);
                        if (F32) prefetch_l2(in + gi + 1);
                        prefetch_l2(guide_n + gi);
                        prefetch_l2(guide_n + gi + 1);
                        prefetch_l2(guide_dz + gi);
                    }
                }
            };
            for (int q = -2; q < -2 + kAhead && q < L + 2; q++) prefetch_row(q);
#pragma unroll 1
            for (int q = -2; q < L + 2; q++, n++) {
                if (q + kAhead < L + 2) prefetch_row(q + kAhead);
                const int gy = ph + (j0 + q) * STEP;
                const bool row_ok = (gy >= 0 && gy < H);
                const int slot = n % kRing;
                float4 *row = ring + (size_t)slot * G::row_f4;
                // all global loads of this thread's (up to two) pairs first ...
                float4 rg0[2], rg1[2], rg2[2];
                CT rc0[2], rc1[2], rc2[2];
                float2 rdz[2];
#pragma unroll
                for (int h = 0; h < (kTwo ? 2 : 1); h++) {
                    const int p = pt + h * kStreamProducers;
                    const int gx = x0 - 2 * STEP + 2 * p;
                    rg0[h] = rg1[h] = rg2[h] = null_g;
                    rc0[h] = rc1[h] = rc2[h] = CT();
                    rdz[h] = make_float2(0.f, 0.f);
                    if (p < G::np && row_ok && gx >= 0 && gx < W) {      // W is even: a pair is inside or outside
                        const size_t gi = (size_t)gy * W + gx;
                        if (F32) {
                            rc0[h] = __ldg(in + gi);
                            rc1[h] = __ldg(in + gi + 1);
                        } else {
                            const uint4 t = __ldg(reinterpret_cast<const uint4 *>(in + gi));
                            *reinterpret_cast<uint2 *>(&rc0[h]) = make_uint2(t.x, t.y);
                            *reinterpret_cast<uint2 *>(&rc1[h]) = make_uint2(t.z, t.w);
                        }
                        rg0[h] = __ldg(guide_n + gi);
                        rg1[h] = __ldg(guide_n + gi + 1);
                        rdz[h] = __ldg(reinterpret_cast<const float2 *>(guide_dz + gi));
                    }
                    if (STEP == 1 && p < G::np && row_ok && gx + 2 >= 0 && gx + 2 < W) {   // first pixel of the next pair, for the shifted copy
                        const size_t gi = (size_t)gy * W + gx + 2;
                        rc2[h] = __ldg(in + gi);
                        rg2[h] = __ldg(guide_n + gi);
                    }
                }
                // ... then the slot must be free ...
                if (n >= kRing) named_sync(1 + slot, kStreamThreads);
                // ... convert and store
#pragma unroll
                for (int h = 0; h < (kTwo ? 2 : 1); h++) {
                    const int p = pt + h * kStreamProducers;
                    if (p < G::np) {
                        const StagedPx s0 = stage_px<F32>(rc0[h], rg0[h]), s1 = stage_px<F32>(rc1[h], rg1[h]);
                        store_pair(row, G::np, p, s0, s1, rdz[h].x, rdz[h].y);
                        if (STEP == 1) {
                            const StagedPx s2 = stage_px<F32>(rc2[h], rg2[h]);
                            store_pair(row + 5 * G::np, G::np, p, s1, s2, 0.f, 0.f);
                        }
                    }
                }
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(full + slot);
            }
            i += L;
        }
        return;
    }

    // ================================ consumer warpgroup ================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kConsumerRegs));
    const int lane = tid & 31;
    const int pc = tid + G::halo_pairs;      // this thread's pair column inside a staged row
    PkCoef kc;
    kc.k1 = a.t.k1; kc.k2 = a.t.k2; kc.k3 = a.t.k3; kc.k4 = a.t.k4; kc.k5 = a.t.k5;

    // five outputs in flight; output row r lives in slot r mod 5 (static indices: the row loop is unrolled by 5)
    PkAcc A[5];
    PkCentreN C[5];
    unsigned int live = 0;       // bits 2s, 2s+1: the two pixels of the output in slot s are non-background
#pragma unroll
    for (int s = 0; s < 5; s++) {
        A[s].S = A[s].r = A[s].g = A[s].b = A[s].v = make_float2(0.f, 0.f);
        C[s].nlc = C[s].nzc = C[s].nx = C[s].ny = C[s].nz = C[s].kL = C[s].kZ = make_float2(0.f, 0.f);
    }

    unsigned int nb = 0;         // staged-row index of this piece's first row (q = -2)
    auto row_ptr = [&](unsigned int n) { return ring + (size_t)(n % kRing) * G::row_f4; };
    auto wait_row = [&](unsigned int n) { mbar_wait(full + (n % kRing), (n / kRing) & 1); };

    for (long long i = i0; i < i1;) {
        const int xs = (int)(i / H), k = (int)(i - (long long)xs * H);
        int ph, j0, rows;
        stream_decode<STEP>(k, a.rows_a, a.rows_b, ph, j0, rows);
        const long long left = i1 - i;
        const int L = (int)((rows - j0) < left ? (rows - j0) : left);
        const int gx = xs * kStripW + 2 * tid;
        live = 0;
        // start output `r` (centre = staged row r) in slot `slot`
        wait_row(nb);
        wait_row(nb + 1);
        wait_row(nb + 2);
#define SVGF_START_OUTPUT(U, R)                                                                                              \
        {                                                                                                                    \
            const float4 *row_ = row_ptr(nb + (R) + 2);                                                                      \
            const float4 c0 = row_[pc], c1 = row_[G::np + pc], g0 = row_[2 * G::np + pc], g1 = row_[3 * G::np + pc],         \
                         ld = row_[4 * G::np + pc];                                                                          \
            A[U].S = f2bc(1.0f);                                                                                             \
            A[U].r = make_float2(c0.x, c0.y); A[U].g = make_float2(c0.z, c0.w);                                              \
            A[U].b = make_float2(c1.x, c1.y); A[U].v = make_float2(c1.z, c1.w);                                              \
            C[U].nlc = make_float2(-ld.x, -ld.y);                                                                            \
            C[U].nzc = make_float2(-g0.x, -g0.y); C[U].nx = make_float2(g0.z, g0.w);                                         \
            C[U].ny = make_float2(g1.x, g1.y); C[U].nz = make_float2(g1.z, g1.w);                                            \
            C[U].kL = make_float2(a.t.kL_scale * rsqrtf(1e-10f + c1.z), a.t.kL_scale * rsqrtf(1e-10f + c1.w));               \
            C[U].kZ = make_float2(__fdividef(a.t.kZ_scale, fmaxf(ld.z, 1e-6f)), __fdividef(a.t.kZ_scale, fmaxf(ld.w, 1e-6f))); \
            const bool valid_ = ((R) < L) && (gx < W);                                                                       \
            live &= ~(3u << (2 * (U)));                                                                                      \
            if (valid_ && g0.x != kBackgroundZ) live |= 1u << (2 * (U));                                                     \
            if (valid_ && g0.y != kBackgroundZ) live |= 2u << (2 * (U));                                                     \
        }
        // prologue: output r lives in slot (r + 4) mod 5: output 0 starts in slot 4, outputs 1, 2, .. follow inside the loop
        SVGF_START_OUTPUT(4, 0)

#pragma unroll 1
        for (int qb = -2; qb < L + 2; qb += 5) {
#pragma unroll
            for (int u = 0; u < 5; u++) {
                const int q = qb + u;
                if (q >= L + 2) break;
                // in flight: outputs q-2 .. q+2; q = u - 2 (mod 5), so output q + 2 - kk sits in slot (u + 4 - kk) mod 5
                // ---- 1. taps of row q onto the five outputs in flight (dy = q - r = kk - 2) ----
                const int gyq = ph + (j0 + q) * STEP;
                if (gyq >= 0 && gyq < H && __any_sync(0xffffffffu, live != 0)) {
                    const float4 *row = row_ptr(nb + q + 2);
                    // tap columns -2/-1 then +2/+1: two columns (ten independent taps) per trip of a rolled loop with
                    // compile-time constants; the centre column follows with its centre tap left out
#pragma unroll 1
                    for (int h = 0; h < 2; h++) {
                        const int dxa = h ? 2 : -2, dxb = h ? 1 : -1;
                        int sa, sb;
                        if (STEP > 1) { sa = pc + dxa * (STEP / 2); sb = pc + dxb * (STEP / 2); }
                        else { sa = pc + dxa / 2; sb = 5 * G::np + pc + (dxb - 1) / 2; }      // odd offsets: the shifted copy
                        const PkTap ta = load_tap(row, G::np, sa);
                        const PkTap tb = load_tap(row, G::np, sb);
#pragma unroll
                        for (int kk = 0; kk < 5; kk++) {
                            const int ay = kk < 2 ? 2 - kk : kk - 2;
                            pk_tap2n<TERMS>(A[(u + 4 - kk) % 5], C[(u + 4 - kk) % 5], ta, tap_neg_log2_kernel(2, ay), tap_inv_len(2, ay), kc);
                            pk_tap2n<TERMS>(A[(u + 4 - kk) % 5], C[(u + 4 - kk) % 5], tb, tap_neg_log2_kernel(1, ay), tap_inv_len(1, ay), kc);
                        }
                    }
                    {
                        const PkTap t = load_tap(row, G::np, pc);
#pragma unroll
                        for (int kk = 0; kk < 5; kk++) {
                            if (kk == 2) continue;                                           // :584
                            const int ay = kk < 2 ? 2 - kk : kk - 2;
                            pk_tap2n<TERMS>(A[(u + 4 - kk) % 5], C[(u + 4 - kk) % 5], t, tap_neg_log2_kernel(0, ay), tap_inv_len(0, ay), kc);
                        }
                    }
                }
                // ---- 2. row q is no longer needed ----
                named_arrive(1 + (int)((nb + q + 2) % kRing), kStreamThreads);
                // ---- 3. output r = q - 2 (slot u) is complete: normalise and store (:615-622) ----
                {
                    const int s = u;
                    const int r = q - 2;
                    if (r >= 0 && r < L && gx < W) {
                        const int gy = ph + (j0 + r) * STEP;
                        const size_t gi = (size_t)gy * W + gx;
                        const bool l0 = (live >> (2 * s)) & 1u, l1 = (live >> (2 * s + 1)) & 1u;
                        const float i0 = __frcp_rn(A[s].S.x), i1 = __frcp_rn(A[s].S.y);
                        float4 o0 = make_float4(A[s].r.x * i0, A[s].g.x * i0, A[s].b.x * i0, A[s].v.x * (i0 * i0));
                        float4 o1 = make_float4(A[s].r.y * i1, A[s].g.y * i1, A[s].b.y * i1, A[s].v.y * (i1 * i1));
                        if (!(l0 && l1)) {       // background: the clamped centre passes through (:556); re-read it (rare path)
                            if (!l0) o0 = clamp01(ColourPlane<F32>::decode(__ldg(in + gi)));
                            if (!l1) o1 = clamp01(ColourPlane<F32>::decode(__ldg(in + gi + 1)));
                        }
                        const CT e0 = ColourPlane<F32>::encode(o0), e1 = ColourPlane<F32>::encode(o1);
                        const bool h0 = (a.t.level == 0) && hist_colour && l0, h1 = (a.t.level == 0) && hist_colour && l1;
                        if (F32) {
                            out[gi] = e0; out[gi + 1] = e1;
                            if (h0) hist_colour[gi] = e0;
                            if (h1) hist_colour[gi + 1] = e1;
                        } else {
                            const uint2 u0 = *reinterpret_cast<const uint2 *>(&e0), u1 = *reinterpret_cast<const uint2 *>(&e1);
                            *reinterpret_cast<uint4 *>(out + gi) = make_uint4(u0.x, u0.y, u1.x, u1.y);
                            if (h0 && h1) *reinterpret_cast<uint4 *>(hist_colour + gi) = make_uint4(u0.x, u0.y, u1.x, u1.y);
                            else {
                                if (h0) hist_colour[gi] = e0;
                                if (h1) hist_colour[gi + 1] = e1;
                            }
                        }
                    }
                }
                // ---- 4. the freed slot u takes output q + 3 (its centre is row q + 3), used from the next iteration on ----
                if (q + 3 <= L + 1) {
                    wait_row(nb + q + 5);
                    SVGF_START_OUTPUT(u, q + 3)
                } else {
                    live &= ~(3u << (2 * u));
                }
            }
        }
#undef SVGF_START_OUTPUT
        nb += (unsigned int)(L + 4);
        i += L;
    }
}

}  // namespace svgf
