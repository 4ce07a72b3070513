// svgf_kernels_lattice.cuh — a-trous levels 1..4 on PRE-TRANSFORMED planes staged by TMA (cp.async.bulk.tensor).
//
// The packed kernel (svgf_kernels_packed.cuh) spends a quarter of its issue slots turning storage texels into the
// pair-interleaved fp32 tile its tap loop reads (fp16 -> fp32, the reference's [0,1] clamp, luminance, pair
// interleave, bounds tests, five STS per pair), and its 2 x 8 warps per SM idle while that happens.  Here the PRODUCER
// of a level's input writes it in exactly the tile's shared-memory format, so that a tile is five tensor-map TMA loads
// issued by one thread and the tap loop starts as soon as the bytes land:
//
//   "lattice planes" (context-owned, one float4/float2 per horizontal PIXEL PAIR, padded by 32 null texels on every
//   side so that no tile ever leaves the allocation - no bounds logic anywhere on the load side):
//     c0 {r0 r1 g0 g1}   c1 {b0 b1 v0 v1}   lz {l0 l1 z0 z1}      colour set, ping-ponged between levels
//     n0 {nx0 nx1 ny0 ny1}   n1 {nz0 nz1}                          normals, written once per frame
//   Values are what the reference's next level would have loaded (src/Filter.cuh:78-83,543,586): the previous level's
//   result rounded through the storage format (fp16 storage: one __float2half_rn, exactly the reference's store),
//   clamped to [0,1], with the luminance (:260-263, un-contracted) computed once per pixel.  z is GetDepth's value
//   (:199-207; +inf in the padding: |zc - inf| * kZ = inf drives the weight to exactly 0 like the reference's skipped
//   out-of-image taps, :579).
//
//   A tile is one row phase of the level's lattice: 16 rows spaced STEP apart.  The padded plane is described to TMA as a
//   3-D tensor {row bytes / 8, STEP, rows / STEP} of 8-byte elements, so those 16 rows are ONE box {tile pairs x 2, 1, 16}
//   at coordinates (x, phase, row block): no element strides (limited to 8 by the hardware), no per-row copies.
//
// Uniform-normal tiles: the temporal pass publishes, per 32-pixel row segment, whether all its non-background texels
// carry one normal vector (svgf_device.cuh SegmentState).  A tile whose 6 x 16 covering segments agree evaluates the
// normal weight once (12 packed operations per pair of taps instead of 19) and never loads the normal planes at all.
// Background texels are wildcards: their weight is 0 through z = 1e30 whatever the normal term says.
#pragma once
#include <cuda.h>   // CUtensorMap (type only; the encode function is fetched at run time, svgf_tma.cu)

#include "svgf_device.cuh"
#include "svgf_kernels_packed.cuh"

namespace svgf {

template <int STEP> struct LatGeom {
    static constexpr int tile_rows = kPkRows * kLatRowGroups;  // 12 output rows (6 with two row groups)
    static constexpr int block_rows = kPkRows * kPkRowGroups;  // 12: lattice rows of a launch's row block (what the band driver counts in)
    static constexpr int subtiles = block_rows / tile_rows;    // tiles stacked in a row block
    static constexpr int threads = kPkPairs * kLatRowGroups;
    static constexpr int rows = tile_rows + 4;                 // 16 staged rows
    static constexpr int pairs = kPkPairs + 2 * STEP;          // 64 + halo of 2*STEP pixels = STEP pairs on each side
    static constexpr int npairs = pairs * rows;
    static constexpr uint32_t plane16 = (uint32_t)npairs * 16u, plane8 = (uint32_t)npairs * 8u;
    static constexpr uint32_t off_c0 = 0, off_c1 = plane16, off_lz = 2 * plane16, off_n0 = 3 * plane16, off_n1 = 4 * plane16;
#if SVGF_EXP == 3
    static constexpr uint32_t off_misc = 3 * plane16;            // TIMING PROBE (tools/build_exp.sh 3): no room for the normal planes, every tile forced uniform, 3 CTAs/SM
#else
    static constexpr uint32_t off_misc = 4 * plane16 + plane8;   // two mbarriers + the uniform-tile reduction records
#endif
    static constexpr size_t smem_bytes = (size_t)off_misc + 256 + 128;   // + slack to align the base to 128 bytes
    static_assert(plane16 % 128 == 0 && plane8 % 16 == 0, "TMA destinations (multiples of plane16) must stay 128-byte aligned");
};

struct LatticeArgs {
    int W, H;
    int pitch_pairs;            // pixel pairs per padded plane row
    int segs_x;                 // 32-pixel segments per image row (segment map pitch)
    float kL_scale, kZ_scale;   // log2e / phi_colour, log2e / (STEP * phi_depth)
    float k1, k2, k3, k4, k5;   // normal-term series coefficients
    int uniform_tiles;          // 0: never take the uniform-normal shortcut (SVGF_FLAG_NO_UNIFORM_TILES)
    int yblock0, nyblocks0, yblock1;   // row blocks of this launch: [yblock0, yblock0 + nyblocks0) then [yblock1, ...) (band driver:
                                       // boundary rows first, svgf_band.cu); nyblocks0 == 0: one range from yblock0
};

// ---- TMA / mbarrier primitives (PTX ISA: cp.async.bulk.tensor, mbarrier) -----------------------------------------
__device__ __forceinline__ void tma_load_3d(void *dst_smem, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_descriptor(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// mbarrier wait that cannot hang the GPU: a transaction count that never completes (a driver that rejects the tensor map,
// a byte-count bug) traps after ~2 s instead of spinning until the watchdog, so the host sees a CUDA error.
__device__ __forceinline__ void mbar_wait_or_trap(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    long long t0 = 0;
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        const long long now = clock64();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 4000000000LL) __trap();
    }
}
// Programmatic dependent launch: a level is launched while its predecessor still runs; everything before this wait
// (barrier init, descriptor prefetch, the segment-map test, depth-derivative loads - none of which the predecessor
// writes) overlaps the predecessor's tail.
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dependency_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

struct LtCentre { float2 nlc, nzc, kL, kZ; };          // centre luminance and depth NEGATED (one FADD2 per difference)
struct LtNormal { float2 nx, ny, nz; };

// one tap row (pair `si`) applied to the outputs it serves.  UNIF: un = 1 - sat(n.n) and pn = P(un) of the tile's one normal
// are two tile constants; the exponent is still formed by the same single fma(-u, p, -base) as the general form, so a
// pixel gets the same bits whichever form its tile takes (and whichever way the image is cut into tiles or bands).
template <int TERMS, bool UNIF>
__device__ __forceinline__ float2 lt_weight(const LtCentre &C, const LtNormal &N, const float4 &lz, const float4 &n0, const float2 &n1,
                                            float ck, float cinv, const PkCoef &k, float un, float pn) {
    const float2 ql = make_float2(lz.x, lz.y), qz = make_float2(lz.z, lz.w);
    float2 base = __ffma2_rn(f2abs(__fadd2_rn(ql, C.nlc)), C.kL, f2bc(ck));
    const float2 tz = __fmul2_rn(f2abs(__fadd2_rn(qz, C.nzc)), C.kZ);
    base = __ffma2_rn(tz, f2bc(cinv), base);
    float2 e;
    if (UNIF) {
        e = __ffma2_rn(f2bc(-un), f2bc(pn), f2neg(base));
    } else {
        float2 d = __fmul2_rn(N.nx, make_float2(n0.x, n0.y));            // (x*x' + y*y') + z*z', reference dot order
        d = __ffma2_rn(N.ny, make_float2(n0.z, n0.w), d);
        d = __ffma2_rn(N.nz, n1, d);
        float2 u = __fadd2_rn(f2bc(1.0f), f2neg(d));
        u = make_float2(fmaxf(u.x, 0.0f), fmaxf(u.y, 0.0f));              // d > 1 saturates to 1 (u = 0); d < 0: weight ~ 2^-96
        float2 p;
        if (TERMS == 5) { p = __ffma2_rn(u, f2bc(k.k5), f2bc(k.k4)); p = __ffma2_rn(u, p, f2bc(k.k3)); p = __ffma2_rn(u, p, f2bc(k.k2)); }
        else p = __ffma2_rn(u, f2bc(k.k3), f2bc(k.k2));
        p = __ffma2_rn(u, p, f2bc(k.k1));
        e = __ffma2_rn(f2neg(u), p, f2neg(base));
    }
    return make_float2(fast_exp2(e.x), fast_exp2(e.y));
}
template <int TERMS, bool UNIF>
__device__ __forceinline__ void lt_tap(PkAcc &A, const LtCentre &C, const LtNormal &N, const float4 &c0, const float4 &c1, const float4 &lz,
                                       const float4 &n0, const float2 &n1, float ck, float cinv, const PkCoef &k, float un, float pn) {
    const float2 w = lt_weight<TERMS, UNIF>(C, N, lz, n0, n1, ck, cinv, k, un, pn);
    A.S = __fadd2_rn(A.S, w);
    A.r = __ffma2_rn(w, make_float2(c0.x, c0.y), A.r);
    A.g = __ffma2_rn(w, make_float2(c0.z, c0.w), A.g);
    A.b = __ffma2_rn(w, make_float2(c1.x, c1.y), A.b);
    A.v = __ffma2_rn(__fmul2_rn(w, w), make_float2(c1.z, c1.w), A.v);
}

template <int STEP, int TERMS, bool UNIF>
__device__ __forceinline__ void lt_all_taps(PkAcc (&A)[kPkRows], const LtCentre (&C)[kPkRows], const LtNormal (&N)[kPkRows],
                                            const float4 *sC0, const float4 *sC1, const float4 *sLZ, const float4 *sN0,
                                            const float2 *sN1, int row0, int pcol, const PkCoef &k, float un, float pn) {
    using G = LatGeom<STEP>;
    static_assert(STEP % 2 == 0, "odd tap offsets break the pixel pairs: level 0 stays with the packed kernel");
#pragma unroll
    for (int dx = -2; dx <= 2; dx++) {
#pragma unroll
        for (int t = -2; t < kPkRows + 2; t++) {
            const int si = (row0 + t) * G::pairs + pcol + dx * (STEP / 2);
            const float4 c0 = sC0[si], c1 = sC1[si], lz = sLZ[si];
            float4 n0 = make_float4(0.f, 0.f, 0.f, 0.f);
            float2 n1 = make_float2(0.f, 0.f);
            if (!UNIF) { n0 = sN0[si]; n1 = sN1[si]; }
#pragma unroll
            for (int j = 0; j < kPkRows; j++) {
                const int dy = t - j;
                if (dy < -2 || dy > 2 || (dx == 0 && dy == 0)) continue;
                const int ax = dx < 0 ? -dx : dx, ay = dy < 0 ? -dy : dy;
                lt_tap<TERMS, UNIF>(A[j], C[j], N[j], c0, c1, lz, n0, n1, tap_neg_log2_kernel(ax, ay), tap_inv_len(ax, ay), k, un, pn);
            }
        }
    }
}

// 32-pixel row segments of the guide (svgf_device.cuh): w = 0 only background / outside, 1 one normal (xyz), 2 mixed
constexpr float kSegWild = 0.0f, kSegUniform = 1.0f, kSegMixed = 2.0f;

// LAST: the level's result is the caller's plane in the storage format (`out`); otherwise it is the next lattice
// level's input (`dst`).
template <bool F32, int STEP, int TERMS, bool LAST>
#if SVGF_EXP == 3
__global__ void __launch_bounds__(kPkPairs * kLatRowGroups, 3)
#else
__global__ void __launch_bounds__(kPkPairs * kLatRowGroups, 8 / kLatRowGroups)
#endif
atrous_lattice_kernel(const __grid_constant__ CUtensorMap mC0, const __grid_constant__ CUtensorMap mC1,
                      const __grid_constant__ CUtensorMap mLZ, const __grid_constant__ CUtensorMap mN0,
                      const __grid_constant__ CUtensorMap mN1, LatticeArgs a, const float *__restrict__ guide_dz,
                      const float4 *__restrict__ seg, LatticeColour dst, typename ColourPlane<F32>::texel *__restrict__ out) {
    using G = LatGeom<STEP>;
    using CT = typename ColourPlane<F32>::texel;
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    // 128-byte alignment of the TMA destinations, computed on the shared-space address so that every access below stays
    // an LDS / STS (a pointer rebuilt from an integer would turn them into generic loads)
    unsigned char *smem = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
    float4 *sC0 = reinterpret_cast<float4 *>(smem + G::off_c0);
    float4 *sC1 = reinterpret_cast<float4 *>(smem + G::off_c1);
    float4 *sLZ = reinterpret_cast<float4 *>(smem + G::off_lz);
    float4 *sN0 = reinterpret_cast<float4 *>(smem + G::off_n0);
    float2 *sN1 = reinterpret_cast<float2 *>(smem + G::off_n1);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + G::off_misc);          // [0] colour planes, [1] normal planes

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * kTileW;
    // blockIdx.y = ((row block, sub-tile), phase): a row block is 12 lattice rows of every phase, cut into `subtiles` tiles;
    // the launch covers row blocks [yblock0, yblock0 + nyblocks0) and then [yblock1, ...) (one range when nyblocks0 == 0)
    const int by = blockIdx.y / STEP, phase = blockIdx.y % STEP;
    const int yb = by / G::subtiles, sub = by % G::subtiles;
    const int yblock = (a.nyblocks0 == 0 || yb < a.nyblocks0) ? a.yblock0 + yb : a.yblock1 + (yb - a.nyblocks0);
    const int lrow0 = yblock * G::block_rows + sub * G::tile_rows;  // first lattice row (of this phase) of the tile
    const int y0 = lrow0 * STEP + phase;

    const int cx = x0 - 2 * STEP + kLatPadX;                       // 8-byte elements == pixels for the 16-byte-per-pair planes
    const int cy = lrow0 - 2 + kLatPadY / STEP;                     // lattice row of the first staged row (kLatPadY % STEP == 0)
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        grid_dependency_wait();   // the previous level's planes are complete from here on (only this thread reads them)
        mbar_expect_tx(&bar[0], 3 * G::plane16);
        tma_load_3d(sC0, &mC0, cx, phase, cy, &bar[0]);
        tma_load_3d(sC1, &mC1, cx, phase, cy, &bar[0]);
        tma_load_3d(sLZ, &mLZ, cx, phase, cy, &bar[0]);
    }

    // ---- uniform-normal test, PER WARP: a warp filters 64 x 3 outputs and taps (64 + 4 STEP) x 7 texels, which 4 x 7 of
    //      the guide's 32-pixel row segments cover - one map entry per lane (guide data: not written by the previous level;
    //      the loads run while the colour planes are already in flight).  A warp whose segments agree on one normal takes
    //      the uniform form whatever the rest of the tile looks like; the normal planes are loaded iff some warp needs them.
    const int pcx = tid & (kPkPairs - 1), tg = tid / kPkPairs;
    bool uniform_n;
    float3 nref;
    {
        const int lane = tid & 31, half = (tid >> 5) & 1;              // which 64-pixel half of the tile row this warp owns
        const int r = lane >> 2, sx = (x0 >> 5) + 2 * half - 1 + (lane & 3);
        const int gy = y0 + (tg * kPkRows + r - 2) * STEP;
        float4 sg = make_float4(0.f, 0.f, 0.f, kSegWild);
        if (r < kPkRows + 4 && sx >= 0 && sx < a.segs_x && gy >= 0 && gy < a.H) sg = __ldg(seg + (size_t)gy * a.segs_x + sx);
        const unsigned has = __ballot_sync(0xffffffffu, sg.w == kSegUniform);
        const int leader = has ? (__ffs(has) - 1) : 0;
        nref = make_float3(__shfl_sync(0xffffffffu, sg.x, leader), __shfl_sync(0xffffffffu, sg.y, leader), __shfl_sync(0xffffffffu, sg.z, leader));
        const bool ok = sg.w == kSegWild || (sg.w == kSegUniform && sg.x == nref.x && sg.y == nref.y && sg.z == nref.z);
        uniform_n = __all_sync(0xffffffffu, ok) && a.uniform_tiles != 0;
#if SVGF_EXP == 3 || SVGF_EXP == 4
        uniform_n = true;       // TIMING PROBES ONLY: wrong pixels on mixed tiles
#endif
        if (!has) nref = make_float3(0.f, 0.f, 0.f);                    // nothing but background: no tap carries weight
    }
    // depth derivatives of this thread's outputs (guide data as well)
    const int gx = x0 + 2 * pcx;
    const int pcol = pcx + STEP;
    const int row0 = tg * kPkRows + 2;
    float2 dz[kPkRows];
#pragma unroll
    for (int j = 0; j < kPkRows; j++) {
        const int gy = y0 + (tg * kPkRows + j) * STEP;
        dz[j] = (gx < a.W && gy < a.H) ? __ldg(reinterpret_cast<const float2 *>(guide_dz + (size_t)gy * a.W + gx)) : make_float2(0.f, 0.f);
    }
    const bool tile_needs_normals = __syncthreads_or(!uniform_n) != 0;   // also publishes the barrier init
    if (tile_needs_normals && tid == 0) {
        mbar_expect_tx(&bar[1], G::plane16 + G::plane8);
        tma_load_3d(sN0, &mN0, cx, phase, cy, &bar[1]);
        tma_load_3d(sN1, &mN1, cx >> 1, phase, cy, &bar[1]);
    }
    grid_dependency_launch();

    PkCoef k;
    k.k1 = a.k1; k.k2 = a.k2; k.k3 = a.k3; k.k4 = a.k4; k.k5 = a.k5;
    float un, pn;
    pk_normal_term<TERMS>(nref.x, nref.y, nref.z, nref.x, nref.y, nref.z, k, un, pn);

    mbar_wait_or_trap(&bar[0], 0);

    LtCentre C[kPkRows];
    LtNormal N[kPkRows];
    PkAcc A[kPkRows];
    bool live0[kPkRows], live1[kPkRows];
    bool any_live = false;
#pragma unroll
    for (int j = 0; j < kPkRows; j++) {
        const int si = (row0 + j) * G::pairs + pcol;
        const float4 c0 = sC0[si], c1 = sC1[si], lz = sLZ[si];
        const int gy = y0 + (tg * kPkRows + j) * STEP;
        A[j].S = f2bc(1.0f);                                                       // :567-568
        A[j].r = make_float2(c0.x, c0.y); A[j].g = make_float2(c0.z, c0.w);
        A[j].b = make_float2(c1.x, c1.y); A[j].v = make_float2(c1.z, c1.w);
        C[j].nlc = make_float2(-lz.x, -lz.y);
        C[j].nzc = make_float2(-lz.z, -lz.w);
        const bool inside = (gx < a.W) && (gy < a.H);
        live0[j] = inside && (lz.z != kBackgroundZ);                               // :554: background passes through
        live1[j] = inside && (lz.w != kBackgroundZ);
        any_live |= live0[j] | live1[j];
        C[j].kL = make_float2(a.kL_scale * fast_rsqrt(1e-10f + c1.z), a.kL_scale * fast_rsqrt(1e-10f + c1.w));   // :562
        C[j].kZ = make_float2(__fdividef(a.kZ_scale, fmaxf(dz[j].x, 1e-6f)), __fdividef(a.kZ_scale, fmaxf(dz[j].y, 1e-6f)));   // :563
        N[j].nx = N[j].ny = N[j].nz = make_float2(0.f, 0.f);
    }

    if (uniform_n) {
        if (__any_sync(0xffffffffu, any_live)) lt_all_taps<STEP, TERMS, true>(A, C, N, sC0, sC1, sLZ, sN0, sN1, row0, pcol, k, un, pn);
    } else {
        mbar_wait_or_trap(&bar[1], 0);
#pragma unroll
        for (int j = 0; j < kPkRows; j++) {
            const int si = (row0 + j) * G::pairs + pcol;
            const float4 n0 = sN0[si];
            const float2 n1 = sN1[si];
            N[j].nx = make_float2(n0.x, n0.y); N[j].ny = make_float2(n0.z, n0.w); N[j].nz = n1;
        }
        if (__any_sync(0xffffffffu, any_live)) lt_all_taps<STEP, TERMS, false>(A, C, N, sC0, sC1, sLZ, sN0, sN1, row0, pcol, k, 0.f, 0.f);
    }

    // ---- normalise and store both pixels of the pair (:615-618) ----
#pragma unroll
    for (int j = 0; j < kPkRows; j++) {
        const int gy = y0 + (tg * kPkRows + j) * STEP;
        if (gx >= a.W || gy >= a.H) continue;
        const int si = (row0 + j) * G::pairs + pcol;
        const float i0 = fast_rcp(A[j].S.x), i1 = fast_rcp(A[j].S.y);
        float4 o0 = make_float4(A[j].r.x * i0, A[j].g.x * i0, A[j].b.x * i0, A[j].v.x * (i0 * i0));
        float4 o1 = make_float4(A[j].r.y * i1, A[j].g.y * i1, A[j].b.y * i1, A[j].v.y * (i1 * i1));
        if (!(live0[j] && live1[j])) {
            const float4 c0 = sC0[si], c1 = sC1[si];
            if (!live0[j]) o0 = make_float4(c0.x, c0.z, c1.x, c1.z);                 // :556 (clamped centre)
            if (!live1[j]) o1 = make_float4(c0.y, c0.w, c1.y, c1.w);
        }
        if (LAST) {
            const size_t gi = (size_t)gy * a.W + gx;
            const CT e0 = ColourPlane<F32>::encode(o0), e1 = ColourPlane<F32>::encode(o1);
            if (F32) {
                out[gi] = e0; out[gi + 1] = e1;
            } else {
                const uint2 u0 = *reinterpret_cast<const uint2 *>(&e0), u1 = *reinterpret_cast<const uint2 *>(&e1);
                *reinterpret_cast<uint4 *>(out + gi) = make_uint4(u0.x, u0.y, u1.x, u1.y);
            }
        } else {
            const float4 lz = sLZ[si];
            lattice_store_pair<F32>(dst, lattice_index(gx, gy, a.pitch_pairs), o0, o1, make_float2(lz.z, lz.w));
        }
    }
}

}  // namespace svgf
