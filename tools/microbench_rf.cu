// Does register-file operand bandwidth limit FFMA / FFMA2 on B200?  Measures warp-instruction issue rate per SM
// sub-partition for FMA-pipe instructions with 1, 2 or 3 DISTINCT register operands (the a-trous tap is made of
// 3-operand FFMA2s).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb_rf tools/microbench_rf.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;
constexpr int N = 8;   // independent chains per thread

// x = x*a + b, a/b uniform scalars (compiler may use immediate/uniform operands)
__global__ void k_ffma_1reg(float *out, float a, float b) {
    float x[N];
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < N; i++) x[i] = fmaf(x[i], a, b);
    float s = 0;
#pragma unroll
    for (int i = 0; i < N; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// x = y*z + x with y, z distinct per chain and varying per thread: 3 distinct register operands
__global__ void k_ffma_3reg(float *out, const float *in) {
    float x[N], y[N], z[N];
#pragma unroll
    for (int i = 0; i < N; i++) { x[i] = in[threadIdx.x + i]; y[i] = in[threadIdx.x + 32 + i]; z[i] = in[threadIdx.x + 64 + i]; }
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < N; i++) x[i] = fmaf(y[i], z[(i + 1) % N], x[i]);
    float s = 0;
#pragma unroll
    for (int i = 0; i < N; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2_1reg(float *out, float a, float b) {
    float2 x[N];
    const float2 aa = make_float2(a, a), bb = make_float2(b, b);
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < N; i++) x[i] = __ffma2_rn(x[i], aa, bb);
    float s = 0;
#pragma unroll
    for (int i = 0; i < N; i++) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2_3reg(float *out, const float2 *in) {
    float2 x[N], y[N], z[N];
#pragma unroll
    for (int i = 0; i < N; i++) { x[i] = in[threadIdx.x + i]; y[i] = in[threadIdx.x + 32 + i]; z[i] = in[threadIdx.x + 64 + i]; }
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < N; i++) x[i] = __ffma2_rn(y[i], z[(i + 1) % N], x[i]);
    float s = 0;
#pragma unroll
    for (int i = 0; i < N; i++) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 3 register operands but one of them shared by consecutive instructions (operand-reuse cache candidate):
// x[i] = y[i]*q + x[i] with q the same register for all i (like one staged tap applied to several outputs)
__global__ void k_ffma2_shared(float *out, const float2 *in) {
    float2 x[N], y[N];
    float2 q = in[threadIdx.x + 200];
#pragma unroll
    for (int i = 0; i < N; i++) { x[i] = in[threadIdx.x + i]; y[i] = in[threadIdx.x + 32 + i]; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < N; i++) x[i] = __ffma2_rn(y[i], q, x[i]);
        q.x += 1e-9f;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < N; i++) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 2-operand packed ops: x = x + y (FADD2), x = x * y (FMUL2)
__global__ void k_fadd2(float *out, const float2 *in) {
    float2 x[N], y[N];
#pragma unroll
    for (int i = 0; i < N; i++) { x[i] = in[threadIdx.x + i]; y[i] = in[threadIdx.x + 32 + i]; }
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < N; i++) x[i] = __fadd2_rn(x[i], y[(i + 1) % N]);
    float s = 0;
#pragma unroll
    for (int i = 0; i < N; i++) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// dependent chain latency: one chain per thread, one warp per SM sub-partition
__global__ void k_lat_ffma2(float *out, const float2 *in, long long *cyc) {
    float2 x = in[threadIdx.x], y = in[threadIdx.x + 32], z = in[threadIdx.x + 64];
    const long long t0 = clock64();
#pragma unroll 16
    for (int it = 0; it < 1024; it++) x = __ffma2_rn(x, y, z);
    const long long t1 = clock64();
    out[threadIdx.x] = x.x + x.y;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lat_ffma(float *out, const float *in, long long *cyc) {
    float x = in[threadIdx.x], y = in[threadIdx.x + 32], z = in[threadIdx.x + 64];
    const long long t0 = clock64();
#pragma unroll 16
    for (int it = 0; it < 1024; it++) x = fmaf(x, y, z);
    const long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lat_ex2(float *out, const float *in, long long *cyc) {
    float x = in[threadIdx.x];
    const long long t0 = clock64();
#pragma unroll 16
    for (int it = 0; it < 1024; it++) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); x = y; }
    const long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <typename F> float time_it(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    int sms, clk;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float *out, *in; long long *cyc;
    cudaMalloc(&out, 148 * 16 * 1024 * sizeof(float));
    cudaMalloc(&in, 4096 * sizeof(float2));
    cudaMalloc(&cyc, 8);
    cudaMemset(in, 0, 4096 * sizeof(float2));
    for (int wps : {1, 2, 4, 8}) {          // warps per SM sub-partition
        const int threads = 128 * wps > 1024 ? 1024 : 128 * wps, blocks = sms * (128 * wps / threads);
        const double winstr = (double)blocks * threads / 32 * ITERS * N;      // warp instructions
        auto rate = [&](float ms) { return (ms * 1e-3) * (clk * 1e3) * sms * 4 / winstr; };   // cycles per warp-instr per SMSP
        printf("---- %d warp(s) per SM sub-partition: cycles per warp-instruction per sub-partition ----\n", wps);
        printf("FFMA  1 reg operand  : %.2f\n", rate(time_it([&] { k_ffma_1reg<<<blocks, threads>>>(out, 1.0001f, 0.5f); })));
        printf("FFMA  3 reg operands : %.2f\n", rate(time_it([&] { k_ffma_3reg<<<blocks, threads>>>(out, in); })));
        printf("FFMA2 1 reg operand  : %.2f\n", rate(time_it([&] { k_ffma2_1reg<<<blocks, threads>>>(out, 1.0001f, 0.5f); })));
        printf("FFMA2 3 reg operands : %.2f\n", rate(time_it([&] { k_ffma2_3reg<<<blocks, threads>>>(out, (const float2 *)in); })));
        printf("FFMA2 3 reg, 1 shared: %.2f\n", rate(time_it([&] { k_ffma2_shared<<<blocks, threads>>>(out, (const float2 *)in); })));
        printf("FADD2 2 reg operands : %.2f\n", rate(time_it([&] { k_fadd2<<<blocks, threads>>>(out, (const float2 *)in); })));
    }
    long long h;
    k_lat_ffma<<<1, 32>>>(out, in, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("dependent FFMA  latency: %.2f cycles\n", h / 1024.0);
    k_lat_ffma2<<<1, 32>>>(out, (const float2 *)in, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("dependent FFMA2 latency: %.2f cycles\n", h / 1024.0);
    k_lat_ex2<<<1, 32>>>(out, in, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("dependent MUFU.EX2 latency: %.2f cycles\n", h / 1024.0);
    return 0;
}
