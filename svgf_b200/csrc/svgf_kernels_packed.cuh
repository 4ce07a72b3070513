// svgf_kernels_packed.cuh — the a-trous level with packed FP32x2 arithmetic (Blackwell FFMA2 / FADD2 / FMUL2).
//
// The level is bound by instruction issue before it is bound by the FP32 pipe (ncu on the scalar tiled kernel:
// issue slots 65 % busy with 1.5 eligible warps per scheduler, FMA pipe 45 %; tools/microbench.cu: FFMA2 retires
// two FMAs per issue slot at the same FMA throughput).  So each thread filters TWO horizontally adjacent outputs
// (x, x+1) at once and every FP32 instruction of the tap evaluation is the packed .f32x2 form: ~11 issue slots
// per tap instead of ~28.
//
// What makes the packing free of shuffles and moves:
//  * shared memory holds PIXEL PAIRS interleaved per channel — {r0,r1,g0,g1} {b0,b1,v0,v1} {z0,z1,nx0,nx1}
//    {ny0,ny1,nz0,nz1} {l0,l1} — so one LDS.128 delivers two ready-made 64-bit register pairs; the taps of outputs
//    (x, x+1) at offset dx*STEP are the stored pair at column x + dx*STEP whenever dx*STEP is even;
//  * the per-output quantities (centre luminance, depth, normal, the two edge-stopping scales, the five sums) are
//    natural pairs; per-tap constants (kernel weight, 1/length) are immediates broadcast by the instruction;
//  * level 0 (STEP = 1) has odd offsets: the dx = +-1 taps of outputs (x, x+1) are packed across the two directions -
//    (x <- x+1, x+1 <- x) is the centre pair with its halves swapped, (x <- x-1, x+1 <- x+2) is {left pair.hi, right
//    pair.lo} - at the price of two register moves per operand pair instead of a scalar evaluation per pixel.
// Everything else follows svgf_kernels_tiled.cuh: one lattice row phase per tile, contiguous 128-pixel x range,
// R outputs per thread and column for register-level tap reuse, per-pixel work (decode, clamp, luminance) done once
// at staging, null texels (z = +inf) outside the image.  Arithmetic and rounding are identical to the scalar
// kernels (fma.rn.f32x2 is two independent fma.rn.f32).
#pragma once
#include "svgf_device.cuh"
#include "svgf_kernels_tiled.cuh"

namespace svgf {

constexpr int kPkRows = 3;        // outputs per thread and column (R)
constexpr int kPkRowGroups = 4;   // row groups per CTA
constexpr int kPkPairs = kTileW / 2;
constexpr int kPkThreads = kPkPairs * kPkRowGroups;   // 256

template <int STEP> struct PackedGeom {
    static constexpr int tile_rows = kPkRows * kPkRowGroups;       // 12
    static constexpr int cols = kTileW + 4 * STEP;
    static constexpr int pairs = cols / 2;
    static constexpr int rows = tile_rows + 4;
    static constexpr int npairs = pairs * rows;
    static constexpr size_t smem_bytes = (size_t)npairs * 72;
};

__device__ __forceinline__ float2 f2abs(float2 a) { return make_float2(fabsf(a.x), fabsf(a.y)); }
__device__ __forceinline__ float2 f2neg(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 f2bc(float a) { return make_float2(a, a); }

struct PkCentre { float2 lc, zc, nx, ny, nz, kL, kZ; };
struct PkAcc { float2 S, r, g, b, v; };
struct PkTap { float2 r, g, b, v, l, z, nx, ny, nz; };
struct PkCoef { float k1, k2, k3, k4, k5; };

// one tap for both outputs of the pair; ck = -log2(kernel weight), cinv = 1/length(xx,yy).
// UNIF: every normal the tile can see is the same vector (see the kernel), so u = 1 - sat(n.n) and the series value p
// are two tile constants (un, pn) and the dot product, the clamp and the series drop out of the tap; the exponent is
// still formed by the same single fma(-u, p, -base), so the weight has the same bits as in the general form.
template <int TERMS, bool UNIF = false>
__device__ __forceinline__ void pk_tap2(PkAcc &A, const PkCentre &C, const PkTap &q, float ck, float cinv, const PkCoef &k,
                                        float un = 0.f, float pn = 0.f) {
    float2 base = __ffma2_rn(f2abs(__fadd2_rn(q.l, f2neg(C.lc))), C.kL, f2bc(ck));
    const float2 tz = __fmul2_rn(f2abs(__fadd2_rn(q.z, f2neg(C.zc))), C.kZ);
    base = __ffma2_rn(tz, f2bc(cinv), base);
    float2 e;
    if (UNIF) {
        e = __ffma2_rn(f2bc(-un), f2bc(pn), f2neg(base));
    } else {
        float2 d = __fmul2_rn(C.nx, q.nx);                      // (x*x' + y*y') + z*z', reference dot order
        d = __ffma2_rn(C.ny, q.ny, d);
        d = __ffma2_rn(C.nz, q.nz, d);
        float2 u = __fadd2_rn(f2bc(1.0f), f2neg(d));
        u = make_float2(fmaxf(u.x, 0.0f), fmaxf(u.y, 0.0f));    // d > 1 saturates to 1 (u = 0); d < 0 gives u > 1: weight ~ 2^-96
        float2 p;
        if (TERMS == 5) { p = __ffma2_rn(u, f2bc(k.k5), f2bc(k.k4)); p = __ffma2_rn(u, p, f2bc(k.k3)); p = __ffma2_rn(u, p, f2bc(k.k2)); }
        else if (TERMS == 4) { p = __ffma2_rn(u, f2bc(k.k4), f2bc(k.k3)); p = __ffma2_rn(u, p, f2bc(k.k2)); }
        else p = __ffma2_rn(u, f2bc(k.k3), f2bc(k.k2));
        p = __ffma2_rn(u, p, f2bc(k.k1));
        e = __ffma2_rn(f2neg(u), p, f2neg(base));
    }
    const float2 w = make_float2(fast_exp2(e.x), fast_exp2(e.y));
    A.S = __fadd2_rn(A.S, w);
    A.r = __ffma2_rn(w, q.r, A.r);
    A.g = __ffma2_rn(w, q.g, A.g);
    A.b = __ffma2_rn(w, q.b, A.b);
    A.v = __ffma2_rn(__fmul2_rn(w, w), q.v, A.v);
}

// u and p of one normal pair, scalar, same operations and roundings as the packed tap
template <int TERMS>
__device__ __forceinline__ void pk_normal_term(float nx, float ny, float nz, float qnx, float qny, float qnz, const PkCoef &k,
                                               float &u, float &p) {
    const float d = fmaf(nz, qnz, fmaf(ny, qny, nx * qnx));
    u = fmaxf(1.0f - d, 0.0f);
    if (TERMS == 5) { p = fmaf(u, k.k5, k.k4); p = fmaf(u, p, k.k3); p = fmaf(u, p, k.k2); }
    else if (TERMS == 4) { p = fmaf(u, k.k4, k.k3); p = fmaf(u, p, k.k2); }
    else p = fmaf(u, k.k3, k.k2);
    p = fmaf(u, p, k.k1);
}

// same tap with the centre's luminance and depth stored NEGATED (nlc = -lc, nzc = -zc): inside a rolled loop the
// compiler otherwise hoists the two negations into extra live registers per output
struct PkCentreN { float2 nlc, nzc, nx, ny, nz, kL, kZ; };
template <int TERMS>
__device__ __forceinline__ void pk_tap2n(PkAcc &A, const PkCentreN &C, const PkTap &q, float ck, float cinv, const PkCoef &k) {
    float2 base = __ffma2_rn(f2abs(__fadd2_rn(q.l, C.nlc)), C.kL, f2bc(ck));
    const float2 tz = __fmul2_rn(f2abs(__fadd2_rn(q.z, C.nzc)), C.kZ);
    base = __ffma2_rn(tz, f2bc(cinv), base);
    float2 d = __fmul2_rn(C.nx, q.nx);                      // (x*x' + y*y') + z*z', reference dot order
    d = __ffma2_rn(C.ny, q.ny, d);
    d = __ffma2_rn(C.nz, q.nz, d);
    float2 u = __fadd2_rn(f2bc(1.0f), f2neg(d));
    u = make_float2(fmaxf(u.x, 0.0f), fmaxf(u.y, 0.0f));
    float2 p;
    if (TERMS == 5) { p = __ffma2_rn(u, f2bc(k.k5), f2bc(k.k4)); p = __ffma2_rn(u, p, f2bc(k.k3)); p = __ffma2_rn(u, p, f2bc(k.k2)); }
    else if (TERMS == 4) { p = __ffma2_rn(u, f2bc(k.k4), f2bc(k.k3)); p = __ffma2_rn(u, p, f2bc(k.k2)); }
    else p = __ffma2_rn(u, f2bc(k.k3), f2bc(k.k2));
    p = __ffma2_rn(u, p, f2bc(k.k1));
    const float2 e = __ffma2_rn(f2neg(u), p, f2neg(base));
    const float2 w = make_float2(fast_exp2(e.x), fast_exp2(e.y));
    A.S = __fadd2_rn(A.S, w);
    A.r = __ffma2_rn(w, q.r, A.r);
    A.g = __ffma2_rn(w, q.g, A.g);
    A.b = __ffma2_rn(w, q.b, A.b);
    A.v = __ffma2_rn(__fmul2_rn(w, w), q.v, A.v);
}

// Colour planes of the tile.  HC = false: two float4 planes {r0,r1,g0,g1} {b0,b1,v0,v1}.  HC = true (fp16 storage only):
// ONE uint4 plane of half2 pairs {r0r1, g0g1, b0b1, v0v1} - the clamped values are fp16 numbers, so the narrowing is
// exact and every result stays bit-identical, while a tap reads 16 instead of 32 bytes of colour (the level sits on the
// shared-memory pipe: L1TEX 79-87 %, profiles/atrous_r01s8.*); the price is 8 half->float conversions per tap row.
__device__ __forceinline__ float2 pk_h2f(unsigned int h) { return __half22float2(*reinterpret_cast<const __half2 *>(&h)); }
__device__ __forceinline__ unsigned int pk_f2h(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const unsigned int *>(&h);
}
template <bool HC>
__device__ __forceinline__ void pk_load_colour(const float4 *sC0, const float4 *sC1, int si, float2 &r, float2 &g, float2 &b, float2 &v) {
    if (HC) {
        const uint4 t = reinterpret_cast<const uint4 *>(sC0)[si];
        r = pk_h2f(t.x); g = pk_h2f(t.y); b = pk_h2f(t.z); v = pk_h2f(t.w);
    } else {
        const float4 c0 = sC0[si], c1 = sC1[si];
        r = make_float2(c0.x, c0.y); g = make_float2(c0.z, c0.w); b = make_float2(c1.x, c1.y); v = make_float2(c1.z, c1.w);
    }
}

__device__ __forceinline__ constexpr float tap_inv_len(int ax, int ay) {
    const int l2 = ax * ax + ay * ay;   // 1 2 4 5 8
    return l2 == 1 ? 1.0f : l2 == 2 ? 0.70710678f : l2 == 4 ? 0.5f : l2 == 5 ? 0.44721360f : 0.35355339f;
}

// all 24 taps of the kPkRows outputs of one thread (both pixels of the pair)
// PITCH = pixel pairs between consecutive LATTICE rows of the shared-memory planes (the tile's own width for the
// single-level kernel; the fused two-level kernel passes its staging pitch, doubled for the dilated level)
// R = outputs per thread and column: a staged tap row serves up to min(R, 5) of them, so shared-memory traffic per
// output falls as (R + 4) / R while the per-thread state grows by 24 registers per row.
template <int STEP, int TERMS, bool UNIF, int PITCH = PackedGeom<STEP>::pairs, int R = kPkRows, bool HC = false>
__device__ __forceinline__ void pk_all_taps(PkAcc (&A)[R], const PkCentre (&C)[R], const float4 *sC0, const float4 *sC1,
                                            const float4 *sG0, const float4 *sG1, const float2 *sL, int row0, int pcol,
                                            const PkCoef &k, float un, float pn) {
    struct G { enum { pairs = PITCH }; };
    if (STEP > 1) {
#pragma unroll
        for (int dx = -2; dx <= 2; dx++) {
#pragma unroll
            for (int t = -2; t < R + 2; t++) {
                const int si = (row0 + t) * G::pairs + pcol + dx * (STEP / 2);
                float4 g0, g1 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (UNIF) { const float2 zz = *reinterpret_cast<const float2 *>(&sG0[si]); g0 = make_float4(zz.x, zz.y, 0.f, 0.f); }
                else { g0 = sG0[si]; g1 = sG1[si]; }
                PkTap q;
                pk_load_colour<HC>(sC0, sC1, si, q.r, q.g, q.b, q.v);
                q.z = make_float2(g0.x, g0.y); q.nx = make_float2(g0.z, g0.w); q.ny = make_float2(g1.x, g1.y); q.nz = make_float2(g1.z, g1.w);
                q.l = sL[si];
#pragma unroll
                for (int j = 0; j < R; j++) {
                    const int dy = t - j;
                    if (dy < -2 || dy > 2 || (dx == 0 && dy == 0)) continue;
                    const int ax = dx < 0 ? -dx : dx, ay = dy < 0 ? -dy : dy;
                    pk_tap2<TERMS, UNIF>(A[j], C[j], q, tap_neg_log2_kernel(ax, ay), tap_inv_len(ax, ay), k, un, pn);
                }
            }
        }
    } else {
        // level 0: the six columns x-2 .. x+3 are the three aligned pairs m = -1, 0, +1.  Pair m serves dx = 2m packed
        // (outputs x and x+1 tap columns x+2m and x+1+2m).  The odd offsets stay packed as well, by pairing the two outputs
        // with DIFFERENT dx of the same |dx| = 1 (same kernel weight and length): (x <- x+1, x+1 <- x) is the centre pair
        // with its halves swapped, (x <- x-1, x+1 <- x+2) is {pair(-1).hi, pair(+1).lo}.  Each lane of a packed operation
        // is the scalar operation on the same operands, so the results are those of the scalar form.
        auto load_pair = [&](int si, PkTap &q) {
            float4 g0, g1 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (UNIF) { const float2 zz = *reinterpret_cast<const float2 *>(&sG0[si]); g0 = make_float4(zz.x, zz.y, 0.f, 0.f); }
            else { g0 = sG0[si]; g1 = sG1[si]; }
            pk_load_colour<HC>(sC0, sC1, si, q.r, q.g, q.b, q.v);
            q.z = make_float2(g0.x, g0.y); q.nx = make_float2(g0.z, g0.w); q.ny = make_float2(g1.x, g1.y); q.nz = make_float2(g1.z, g1.w);
            q.l = sL[si];
        };
        auto halves = [](const float2 &lo_src, const float2 &hi_src) { return make_float2(lo_src.y, hi_src.x); };   // {a.hi, b.lo}
        auto cross = [&](const PkTap &a, const PkTap &b) {
            PkTap q;
            q.r = halves(a.r, b.r); q.g = halves(a.g, b.g); q.b = halves(a.b, b.b); q.v = halves(a.v, b.v);
            q.l = halves(a.l, b.l); q.z = halves(a.z, b.z);
            q.nx = halves(a.nx, b.nx); q.ny = halves(a.ny, b.ny); q.nz = halves(a.nz, b.nz);
            return q;
        };
        auto row_taps = [&](const PkTap &q, int t, int ax, bool skip_centre) {
#pragma unroll
            for (int j = 0; j < R; j++) {
                const int dy = t - j;
                if (dy < -2 || dy > 2 || (skip_centre && dy == 0)) continue;
                const int ay = dy < 0 ? -dy : dy;
                pk_tap2<TERMS, UNIF>(A[j], C[j], q, tap_neg_log2_kernel(ax, ay), tap_inv_len(ax, ay), k, un, pn);
            }
        };
#pragma unroll
        for (int t = -2; t < R + 2; t++) {
            const int si = (row0 + t) * G::pairs + pcol;
            PkTap q0, qa, qb;
            load_pair(si, q0);
            row_taps(q0, t, 0, true);                       // dx = 0
            row_taps(cross(q0, q0), t, 1, false);           // x <- x+1, x+1 <- x
            load_pair(si - 1, qa);
            row_taps(qa, t, 2, false);                      // dx = -2
            load_pair(si + 1, qb);
            row_taps(qb, t, 2, false);                      // dx = +2
            row_taps(cross(qa, qb), t, 1, false);           // x <- x-1, x+1 <- x+2
        }
    }
}

// PREF: the centre variance comes from the pre-blurred plane a.var_blur (SVGF_VARIANCE_PREFILTER_GAUSS3); a template
// parameter because even a never-taken branch here costs the register-capped default kernel 4 %
// STAGED: this level feeds a lattice level (svgf_kernels_lattice.cuh): instead of `out` in the storage format it writes
// the successor's pre-transformed input planes `lc` (result rounded through the storage format, clamped, luminance, depth)
// and, once per frame, the pair-interleaved normal planes `ln` - all from values the thread already holds.
template <bool F32, int STEP, int TERMS, int R = kPkRows, bool PREF = false, bool HC = false, bool STAGED = false>
__global__ void __launch_bounds__(kPkPairs * (PackedGeom<STEP>::tile_rows / R), 2)
atrous_packed_kernel(AtrousTiledArgs a, const float4 *__restrict__ guide_n, const float *__restrict__ guide_dz,
                     const typename ColourPlane<F32>::texel *__restrict__ in, typename ColourPlane<F32>::texel *__restrict__ out,
                     typename ColourPlane<F32>::texel *__restrict__ hist_colour, LatticeColour lc, LatticeNormals ln, int lat_pitch_pairs) {
    using G = PackedGeom<STEP>;
    constexpr int kThreads = kPkPairs * (G::tile_rows / R);   // R = 3: 256 threads, R = 4: 192
    static_assert(G::tile_rows % R == 0, "row groups must tile the 12 rows");
    static_assert(!(HC && F32), "half-precision colour tiles are exact for fp16 storage only");
    using CT = typename ColourPlane<F32>::texel;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *sC0 = reinterpret_cast<float4 *>(smem_raw);        // r0 r1 g0 g1
    float4 *sC1 = sC0 + G::npairs;                             // b0 b1 v0 v1
    float4 *sG0 = sC1 + G::npairs;                             // z0 z1 nx0 nx1
    float4 *sG1 = sG0 + G::npairs;                             // ny0 ny1 nz0 nz1
    float2 *sL = reinterpret_cast<float2 *>(sG1 + G::npairs);  // l0 l1

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * kTileW;
    const int yblock = blockIdx.y / STEP + a.yblock0, phase = blockIdx.y % STEP;
    const int y0 = yblock * (G::tile_rows * STEP) + phase;

    // ---- stage the tile, one pixel pair per thread and iteration; all global loads first ----
    // Uniform-normal tiles: when every in-image texel the tile stages (halo included) carries the same non-zero
    // normal as the tile's first pixel - any planar surface: floors, walls, box faces - the normal weight of all
    // 24 x 1536 (centre, tap) pairs is one number, and the taps run the UNIF form (12 instead of 19 packed
    // operations, no normal loads).  Compared as floats: equal values give equal dot products (+0 == -0 included),
    // NaN never matches.  Texels outside the image are excluded: their weight is 0 through z = +inf either way.
    const float4 nref = __ldg(guide_n + (size_t)min(y0, a.H - 1) * a.W + x0);
    bool same_n = a.uniform_tiles && (nref.y != 0.0f || nref.z != 0.0f || nref.w != 0.0f);
    constexpr int kIters = (G::npairs + kThreads - 1) / kThreads;
    CT rc0[kIters], rc1[kIters];
    float4 rg0[kIters], rg1[kIters];
#pragma unroll
    for (int i = 0; i < kIters; i++) {
        const int idx = tid + i * kThreads;
        const int r = idx / G::pairs, pc = idx - r * G::pairs;
        const int gx = x0 - 2 * STEP + 2 * pc, gy = y0 + (r - 2) * STEP;
        rg0[i] = rg1[i] = make_float4(__int_as_float(0x7f800000), 0.f, 0.f, 0.f);   // null texel: z = +inf, zero normal
        rc0[i] = rc1[i] = CT();
        if (idx < G::npairs && gx >= 0 && gx < a.W && gy >= 0 && gy < a.H) {       // W is even: a pair is inside or outside
            const size_t gi = (size_t)gy * a.W + gx;
            if (F32) {
                rc0[i] = __ldg(in + gi);
                rc1[i] = __ldg(in + gi + 1);
            } else {
                const uint4 t = __ldg(reinterpret_cast<const uint4 *>(in + gi));
                *reinterpret_cast<uint2 *>(&rc0[i]) = make_uint2(t.x, t.y);
                *reinterpret_cast<uint2 *>(&rc1[i]) = make_uint2(t.z, t.w);
            }
            rg0[i] = __ldg(guide_n + gi);
            rg1[i] = __ldg(guide_n + gi + 1);
            same_n &= (rg0[i].y == nref.y) & (rg0[i].z == nref.z) & (rg0[i].w == nref.w) & (rg1[i].y == nref.y) & (rg1[i].z == nref.z) &
                      (rg1[i].w == nref.w);
        }
    }
#pragma unroll
    for (int i = 0; i < kIters; i++) {
        const int idx = tid + i * kThreads;
        if (idx < G::npairs) {
            const float4 c0 = ColourPlane<F32>::decode(rc0[i]), c1 = ColourPlane<F32>::decode(rc1[i]);
            const float r0 = __saturatef(c0.x), g0 = __saturatef(c0.y), b0 = __saturatef(c0.z), v0 = __saturatef(c0.w);   // :543,:586
            const float r1 = __saturatef(c1.x), g1 = __saturatef(c1.y), b1 = __saturatef(c1.z), v1 = __saturatef(c1.w);
            if (HC) {
                reinterpret_cast<uint4 *>(sC0)[idx] = make_uint4(pk_f2h(r0, r1), pk_f2h(g0, g1), pk_f2h(b0, b1), pk_f2h(v0, v1));
            } else {
                sC0[idx] = make_float4(r0, r1, g0, g1);
                sC1[idx] = make_float4(b0, b1, v0, v1);
            }
            sG0[idx] = make_float4(rg0[i].x, rg1[i].x, rg0[i].y, rg1[i].y);
            sG1[idx] = make_float4(rg0[i].z, rg1[i].z, rg0[i].w, rg1[i].w);
            sL[idx] = make_float2(luminance(r0, g0, b0), luminance(r1, g1, b1));
        }
    }
    const bool uniform_n = __syncthreads_and(same_n) != 0;

    const int pcx = tid & (kPkPairs - 1), tg = tid / kPkPairs;
    const int gx = x0 + 2 * pcx;
    const int pcol = pcx + STEP;                      // pair column in shared memory (halo of 2*STEP pixels = STEP pairs)
    const int row0 = tg * R + 2;
    PkCoef k;
    k.k1 = a.k1; k.k2 = a.k2; k.k3 = a.k3; k.k4 = a.k4; k.k5 = a.k5;
    float un, pn;
    pk_normal_term<TERMS>(nref.y, nref.z, nref.w, nref.y, nref.z, nref.w, k, un, pn);

    PkCentre C[R];
    PkAcc A[R];
    bool live0[R], live1[R];
    bool any_live = false;
#pragma unroll
    for (int j = 0; j < R; j++) {
        const int si = (row0 + j) * G::pairs + pcol;
        const float4 g0 = sG0[si], g1 = sG1[si];
        const int gy = y0 + (tg * R + j) * STEP;
        A[j].S = f2bc(1.0f);                                                       // :567-568
        pk_load_colour<HC>(sC0, sC1, si, A[j].r, A[j].g, A[j].b, A[j].v);
        C[j].lc = sL[si];
        C[j].zc = make_float2(g0.x, g0.y); C[j].nx = make_float2(g0.z, g0.w);
        C[j].ny = make_float2(g1.x, g1.y); C[j].nz = make_float2(g1.z, g1.w);
        const bool inside = (gx < a.W) && (gy < a.H);
        live0[j] = inside && (g0.x != kBackgroundZ);                               // :554: background passes through
        live1[j] = inside && (g0.y != kBackgroundZ);
        any_live |= live0[j] | live1[j];
        float2 var = A[j].v;                                                        // :547 / the pre-blurred plane (GAUSS3)
        if (PREF && inside) var = __ldg(reinterpret_cast<const float2 *>(a.var_blur + (size_t)gy * a.W + gx));
        C[j].kL = make_float2(a.kL_scale * fast_rsqrt(1e-10f + var.x), a.kL_scale * fast_rsqrt(1e-10f + var.y));   // :562
        float2 dz = make_float2(0.f, 0.f);
        if (inside) dz = __ldg(reinterpret_cast<const float2 *>(guide_dz + (size_t)gy * a.W + gx));
        C[j].kZ = make_float2(__fdividef(a.kZ_scale, fmaxf(dz.x, 1e-6f)), __fdividef(a.kZ_scale, fmaxf(dz.y, 1e-6f)));   // :563
    }

    if (__any_sync(0xffffffffu, any_live)) {
        if (uniform_n) pk_all_taps<STEP, TERMS, true, G::pairs, R, HC>(A, C, sC0, sC1, sG0, sG1, sL, row0, pcol, k, un, pn);
        else pk_all_taps<STEP, TERMS, false, G::pairs, R, HC>(A, C, sC0, sC1, sG0, sG1, sL, row0, pcol, k, 0.f, 0.f);
    }

    // ---- normalise and store both pixels of the pair (:615-622) ----
#pragma unroll
    for (int j = 0; j < R; j++) {
        const int gy = y0 + (tg * R + j) * STEP;
        if (gx >= a.W || gy >= a.H) continue;
        const size_t gi = (size_t)gy * a.W + gx;
        const int si = (row0 + j) * G::pairs + pcol;
        float2 cr, cg, cb, cv;
        pk_load_colour<HC>(sC0, sC1, si, cr, cg, cb, cv);
        const float i0 = fast_rcp(A[j].S.x), i1 = fast_rcp(A[j].S.y);
        float4 o0 = make_float4(A[j].r.x * i0, A[j].g.x * i0, A[j].b.x * i0, A[j].v.x * (i0 * i0));
        float4 o1 = make_float4(A[j].r.y * i1, A[j].g.y * i1, A[j].b.y * i1, A[j].v.y * (i1 * i1));
        if (!live0[j]) o0 = make_float4(cr.x, cg.x, cb.x, cv.x);                     // :556 (clamped centre)
        if (!live1[j]) o1 = make_float4(cr.y, cg.y, cb.y, cv.y);
        const CT e0 = ColourPlane<F32>::encode(o0), e1 = ColourPlane<F32>::encode(o1);
        const bool h0 = (a.level == 0) && hist_colour && live0[j], h1 = (a.level == 0) && hist_colour && live1[j];
        if (STAGED) {
            const size_t li = lattice_index(gx, gy, lat_pitch_pairs);
            lattice_store_pair<F32>(lc, li, o0, o1, C[j].zc);
            ln.n0[li] = make_float4(C[j].nx.x, C[j].nx.y, C[j].ny.x, C[j].ny.y);
            ln.n1[li] = C[j].nz;
        }
        if (F32) {
            if (!STAGED) { out[gi] = e0; out[gi + 1] = e1; }
            if (h0) hist_colour[gi] = e0;
            if (h1) hist_colour[gi + 1] = e1;
        } else {
            const uint2 u0 = *reinterpret_cast<const uint2 *>(&e0), u1 = *reinterpret_cast<const uint2 *>(&e1);
            if (!STAGED) *reinterpret_cast<uint4 *>(out + gi) = make_uint4(u0.x, u0.y, u1.x, u1.y);
            if (h0 && h1) *reinterpret_cast<uint4 *>(hist_colour + gi) = make_uint4(u0.x, u0.y, u1.x, u1.y);
            else {
                if (h0) hist_colour[gi] = e0;
                if (h1) hist_colour[gi + 1] = e1;
            }
        }
    }
}

}  // namespace svgf
