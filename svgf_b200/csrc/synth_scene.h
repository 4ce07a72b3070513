// synth_scene.h — procedural G-buffer + 1-spp-style radiance generator, one code path for host and device.
//
// Stands in for the two producers the filter consumes in the reference, which cannot run on the benchmark
// box (no GL / OptiX): the rasteriser (resources/shaders/GBuffer.frag:62-87, src/App.cu:378-413) and the
// 1-spp path tracer (src/Raygen.cu:456-506).  It emits exactly the reference's texel formats
// (src/App.cu:746-752):
//   position float4 (world xyz, primitive id)            GBuffer.frag:66,79
//   normal   ushort4 = fp16 bits (unit normal, material) GBuffer.frag:65,78,85
//   uv       ushort4 = fp16 bits (barycentric-like uvw, instance id) GBuffer.frag:64,77,86
//   motion   float4  ((prevPx - curPx), linear distance to camera, max(|dz/dx|,|dz/dy|)) GBuffer.frag:67-73,81-82
//   colour   half4 / float4 (noisy radiance, 1)          src/Raygen.cu:505-506
// Background pixels are all-zero in every G-buffer plane (glClearColor(0,0,0,1), src/App.cu:383 -> depth 0).
//
// Only + - * / sqrtf floorf and integer hashing are used, and the file is compiled with FMA contraction off
// on both sides (-fmad=false / -ffp-contract=off), so the CUDA and CPU generators are bit-identical
// (tests/test_synth.py checks that on the GPU box).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SYNTH_HD __host__ __device__ __forceinline__
#else
#define SYNTH_HD inline
#endif

struct svgf_synth_cfg {
    int32_t width, height;
    uint32_t seed;      // scene layout + noise stream (config 5 uses seeds 0..63)
    int32_t frame;      // frame index t >= 0
    float pan_px;       // horizontal camera pan in pixels/frame at the focal depth (3.25 in BASELINE config 2)
    float vert_px;      // vertical pan in pixels/frame at the focal depth (0.5)
    int32_t half_period;  // frames until the pan reverses direction (32)
    int32_t storage;    // 0: colour as half4, 1: colour as float4
};

namespace synth {

struct V3 { float x, y, z; };
SYNTH_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
SYNTH_HD V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
SYNTH_HD V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
SYNTH_HD V3 mul(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
SYNTH_HD float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
SYNTH_HD float len(V3 a) { return sqrtf(dot(a, a)); }
SYNTH_HD float absf(float a) { return a < 0.0f ? -a : a; }
SYNTH_HD float maxf(float a, float b) { return a < b ? b : a; }
SYNTH_HD float minf(float a, float b) { return b < a ? b : a; }

SYNTH_HD uint32_t pcg(uint32_t v) {
    uint32_t s = v * 747796405u + 2891336453u;
    uint32_t w = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u;
    return (w >> 22u) ^ w;
}
SYNTH_HD float u01(uint32_t h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }

// float -> fp16 bits, round-to-nearest-even (same result as __float2half_rn; written out so host and device agree).
SYNTH_HD uint16_t f2h_bits(float f) {
    union { float f; uint32_t u; } c;
    c.f = f;
    const uint32_t x = c.u;
    const uint32_t sign = (x >> 16) & 0x8000u;
    const uint32_t ax = x & 0x7fffffffu;
    if (ax > 0x7f800000u) return (uint16_t)(sign | 0x7fffu);
    if (ax >= 0x47800000u) return (uint16_t)(sign | 0x7c00u);
    if (ax < 0x33000000u) return (uint16_t)sign;
    const int e = (int)(ax >> 23) - 127;
    const uint32_t m = (ax & 0x7fffffu) | 0x800000u;
    uint32_t h;
    if (e < -14) {
        const int sh = (-14 - e) + 13;
        const uint32_t q = m >> sh, rem = m & ((1u << sh) - 1u), half = 1u << (sh - 1);
        h = q + ((rem > half || (rem == half && (q & 1u))) ? 1u : 0u);
    } else {
        const uint32_t q = ((uint32_t)(e + 15) << 10) | ((m >> 13) & 0x3ffu), rem = m & 0x1fffu;
        h = q + ((rem > 0x1000u || (rem == 0x1000u && (q & 1u))) ? 1u : 0u);
    }
    return (uint16_t)(sign | h);
}

constexpr int kSpheres = 8;
constexpr int kBoxes = 2;
constexpr float kFocalDepth = 8.0f;
constexpr float kTan30 = 0.57735026919f;

struct Camera { V3 pos; float focal; };

SYNTH_HD int tri_wave(int t, int half_period) {
    if (half_period <= 0) return t;
    const int p = 2 * half_period;
    const int m = t % p;
    return m <= half_period ? m : p - m;
}

// Camera translates (pan), never rotates: nearer objects move faster on screen => disocclusion bands.
SYNTH_HD Camera camera_at(const svgf_synth_cfg &c, int t) {
    Camera cam;
    cam.focal = (0.5f * (float)c.height) / kTan30;  // 60 degree vertical FOV (src/Scene.h:42)
    const float k = (float)tri_wave(t < 0 ? 0 : t, c.half_period);
    cam.pos = v3(-1.0f + (c.pan_px * kFocalDepth / cam.focal) * k, 0.25f + (c.vert_px * kFocalDepth / cam.focal) * k, 8.0f);
    return cam;
}

struct Hit {
    float t;      // ray parameter along the un-normalised direction
    V3 n;         // unit normal
    int inst;     // instance id (0 = background)
    int prim;     // primitive id
};

SYNTH_HD void sphere_of(uint32_t seed, int i, V3 &c, float &r) {
    const uint32_t h0 = pcg(seed * 0x9E3779B9u + 0x1000u + (uint32_t)i);
    const uint32_t h1 = pcg(h0), h2 = pcg(h1), h3 = pcg(h2);
    // depths 3..15 from the camera plane z = 8  =>  z in [5, -7]
    const float depth = 3.0f + 12.0f * ((float)i + u01(h0)) * (1.0f / (float)kSpheres);
    r = 0.45f + 0.9f * u01(h3);
    c = v3((u01(h1) - 0.5f) * 1.1f * depth, -2.0f + r + 2.5f * u01(h2), 8.0f - depth);
}

SYNTH_HD void box_of(uint32_t seed, int i, V3 &lo, V3 &hi) {
    const uint32_t h0 = pcg(seed * 0x85EBCA6Bu + 0x2000u + (uint32_t)i);
    const uint32_t h1 = pcg(h0), h2 = pcg(h1);
    const float depth = 6.0f + 5.0f * (float)i + 2.0f * u01(h0);
    const float cx = (i == 0 ? -1.0f : 1.0f) * (1.5f + 2.0f * u01(h1));
    const float s = 0.8f + 0.7f * u01(h2);
    lo = v3(cx - s, -2.0f, 8.0f - depth - s);
    hi = v3(cx + s, -2.0f + 2.0f * s, 8.0f - depth + s);
}

SYNTH_HD Hit trace(const svgf_synth_cfg &c, V3 o, V3 d) {
    Hit h;
    h.t = 1e30f; h.n = v3(0, 0, 0); h.inst = 0; h.prim = 0;
    // ground plane y = -2
    if (d.y < 0.0f) {
        const float t = (-2.0f - o.y) / d.y;
        const float z = o.z + t * d.z;
        if (t > 0.0f && z > -8.0f && t < h.t) { h.t = t; h.n = v3(0, 1, 0); h.inst = 1; h.prim = 1; }
    }
    // back wall z = -8, up to y = 6.3 (above it: background)
    if (d.z < 0.0f) {
        const float t = (-8.0f - o.z) / d.z;
        const float y = o.y + t * d.y;
        if (t > 0.0f && y >= -2.0f && y < 6.3f && t < h.t) { h.t = t; h.n = v3(0, 0, 1); h.inst = 2; h.prim = 2; }
    }
    const float dd = dot(d, d);
    for (int i = 0; i < kSpheres; i++) {
        V3 ctr; float r;
        sphere_of(c.seed, i, ctr, r);
        const V3 oc = sub(o, ctr);
        const float b = dot(oc, d);
        const float cc = dot(oc, oc) - r * r;
        const float disc = b * b - dd * cc;
        if (disc > 0.0f) {
            const float t = (-b - sqrtf(disc)) / dd;
            if (t > 0.0f && t < h.t) {
                const V3 p = add(o, mul(d, t));
                h.t = t; h.n = mul(sub(p, ctr), 1.0f / r); h.inst = 3 + i; h.prim = 16 + i;
                const float nl = len(h.n);
                h.n = mul(h.n, 1.0f / nl);
            }
        }
    }
    for (int i = 0; i < kBoxes; i++) {
        V3 lo, hi;
        box_of(c.seed, i, lo, hi);
        // slab test; d components can be 0 -> +-inf, handled by min/max ordering below
        float tmin = 0.0f, tmax = 1e30f;
        int axis = -1; float sgn = 0.0f;
        const float oo[3] = {o.x, o.y, o.z}, ddv[3] = {d.x, d.y, d.z};
        const float l3[3] = {lo.x, lo.y, lo.z}, h3[3] = {hi.x, hi.y, hi.z};
        bool miss = false;
        for (int a = 0; a < 3; a++) {
            if (ddv[a] == 0.0f) {
                if (oo[a] < l3[a] || oo[a] > h3[a]) miss = true;
                continue;
            }
            float t0 = (l3[a] - oo[a]) / ddv[a], t1 = (h3[a] - oo[a]) / ddv[a];
            float s = -1.0f;
            if (t0 > t1) { const float tt = t0; t0 = t1; t1 = tt; s = 1.0f; }
            if (t0 > tmin) { tmin = t0; axis = a; sgn = s; }
            if (t1 < tmax) tmax = t1;
        }
        if (!miss && axis >= 0 && tmin < tmax && tmin > 0.0f && tmin < h.t) {
            h.t = tmin;
            h.n = v3(axis == 0 ? sgn : 0.0f, axis == 1 ? sgn : 0.0f, axis == 2 ? sgn : 0.0f);
            h.inst = 3 + kSpheres + i;
            h.prim = 32 + 6 * i + 2 * axis + (sgn > 0.0f ? 1 : 0);
        }
    }
    return h;
}

struct Texel {
    float pos[4];
    uint16_t nrm[4];
    uint16_t uv[4];
    float mot[4];
    float col[4];
};

SYNTH_HD float fractf(float v) { return v - floorf(v); }

SYNTH_HD V3 albedo_of(uint32_t seed, int inst) {
    const uint32_t h = pcg(seed * 0xC2B2AE35u + 0x3000u + (uint32_t)inst);
    return v3(0.25f + 0.75f * u01(h), 0.25f + 0.75f * u01(pcg(h)), 0.25f + 0.75f * u01(pcg(pcg(h))));
}

SYNTH_HD Texel shade_pixel(const svgf_synth_cfg &c, int x, int y) {
    Texel o;
    const Camera cam = camera_at(c, c.frame);
    const Camera pcam = camera_at(c, c.frame - 1);
    const float fx = (float)x + 0.5f, fy = (float)y + 0.5f;
    const float hw = 0.5f * (float)c.width, hh = 0.5f * (float)c.height;
    const V3 d = v3(fx - hw, fy - hh, -cam.focal);
    const Hit h = trace(c, cam.pos, d);

    // noise stream: counter-based hash of (seed, frame, pixel[, channel])
    const uint32_t pix = (uint32_t)y * (uint32_t)c.width + (uint32_t)x;
    const uint32_t hn = pcg(pcg(c.seed ^ 0x5F6Fu) + pcg((uint32_t)c.frame * 0x9E3779B9u + pcg(pix)));
    const bool lit = u01(hn) >= 0.5f;
    const float g0 = lit ? 2.0f * (0.5f + u01(pcg(hn + 1u))) : 0.0f;
    const float g1 = lit ? 2.0f * (0.5f + u01(pcg(hn + 2u))) : 0.0f;
    const float g2 = lit ? 2.0f * (0.5f + u01(pcg(hn + 3u))) : 0.0f;

    if (h.inst == 0) {
        for (int k = 0; k < 4; k++) { o.pos[k] = 0.0f; o.nrm[k] = 0; o.uv[k] = 0; o.mot[k] = 0.0f; }
        // sky radiance (envmap): smooth vertical gradient, still 1-spp noisy
        const float v = fy / (float)c.height;
        o.col[0] = minf(0.35f * v * g0, 10.0f);
        o.col[1] = minf(0.55f * v * g1, 10.0f);
        o.col[2] = minf(0.90f * v * g2, 10.0f);
        o.col[3] = 1.0f;
        return o;
    }

    const V3 P = add(cam.pos, mul(d, h.t));
    const float dl = len(d);
    const float depth = h.t * dl;  // distance(CameraPosition, worldPos), GBuffer.frag:72
    o.pos[0] = P.x; o.pos[1] = P.y; o.pos[2] = P.z; o.pos[3] = (float)h.prim;
    o.nrm[0] = f2h_bits(h.n.x); o.nrm[1] = f2h_bits(h.n.y); o.nrm[2] = f2h_bits(h.n.z);
    o.nrm[3] = f2h_bits((float)(h.inst % 5));  // material index
    const float u = fractf(P.x * 0.37f + P.z * 0.11f), v = fractf(P.y * 0.41f + P.z * 0.23f) * (1.0f - u);
    o.uv[0] = f2h_bits(u); o.uv[1] = f2h_bits(v); o.uv[2] = f2h_bits(1.0f - u - v);
    o.uv[3] = f2h_bits((float)h.inst);          // instance index = "mesh id"

    // motion = previous pixel position of this world point - current pixel position (GBuffer.frag:67-70)
    const V3 rel = sub(P, pcam.pos);
    const float ppx = hw + pcam.focal * rel.x / (-rel.z), ppy = hh + pcam.focal * rel.y / (-rel.z);
    o.mot[0] = ppx - fx;
    o.mot[1] = ppy - fy;
    o.mot[2] = depth;
    // depth derivative: screen-space differences on the primitive's tangent plane, like dFdx/dFdy on a triangle
    const float pn = dot(sub(P, cam.pos), h.n);
    const V3 dx1 = v3(d.x + 1.0f, d.y, d.z), dy1 = v3(d.x, d.y + 1.0f, d.z);
    const float ndx = dot(dx1, h.n), ndy = dot(dy1, h.n);
    const float zx = absf(ndx) > 1e-6f ? absf((pn / ndx) * len(dx1) - depth) : 1e3f;
    const float zy = absf(ndy) > 1e-6f ? absf((pn / ndy) * len(dy1) - depth) : 1e3f;
    o.mot[3] = minf(maxf(zx, zy), 1e3f);

    // smooth shading in [0, 1.2] x 1-spp-style noise (E[g] = 1); values > 1 exercise the reference's [0,1] clamp
    const V3 ldir = v3(0.3713907f, 0.7427814f, 0.5570860f);
    const float diff = maxf(0.0f, dot(h.n, ldir));
    V3 alb = albedo_of(c.seed, h.inst);
    if (h.inst <= 2) {  // checker on ground and wall
        const float a = (h.inst == 1) ? P.x : P.x, b = (h.inst == 1) ? P.z : P.y;
        const int chk = ((int)floorf(a * 0.8f) + (int)floorf(b * 0.8f)) & 1;
        alb = mul(alb, chk ? 1.0f : 0.45f);
    }
    const float shade = 0.15f + 1.05f * diff;
    o.col[0] = minf(alb.x * shade * g0, 10.0f);  // Params.Clamp = 10, src/Tracing.h:33
    o.col[1] = minf(alb.y * shade * g1, 10.0f);
    o.col[2] = minf(alb.z * shade * g2, 10.0f);
    o.col[3] = 1.0f;
    return o;
}

}  // namespace synth
