// Micro-benchmarks of the B200 pipes the a-trous inner loop leans on: FFMA vs packed FFMA2, MUFU.EX2, LDS.128.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/microbench tools/microbench.cu && /tmp/microbench
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

__global__ void k_ffma(float *out, float a, float b) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = fmaf(x[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2(float *out, float a, float b) {
    float2 x[8];
    const float2 aa = make_float2(a, a), bb = make_float2(b, b);
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = __ffma2_rn(x[i], aa, bb);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ex2(float *out, float a) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 0.001f + i * 0.1f;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[i])); x[i] = y * a; }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: 18 FFMA + 1 EX2 per "tap", like the inner loop
__global__ void k_mix(float *out, float a, float b) {
    float x[6];
#pragma unroll
    for (int i = 0; i < 6; i++) x[i] = threadIdx.x * 0.001f + i * 0.1f;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int i = 0; i < 6; i++) x[i] = fmaf(x[i], a, b);
        float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[0])); x[0] = y;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_lds128(float *out) {
    __shared__ float4 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(i, i, i, i);
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    int idx = threadIdx.x;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float4 v = sm[(idx + i * 32) & 1023];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        idx += 1;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
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
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float *out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
    const int blocks = sms * 4, threads = 512;   // 2048 threads / SM
    const double n = (double)blocks * threads * ITERS * 8;
    float ms;
    ms = time_it([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
    printf("FFMA  : %.3f ms  %.1f lane-FMA/clk/SM (clk %d kHz)\n", ms, n / (ms * 1e-3) / sms / (clk * 1e3), clk);
    ms = time_it([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
    printf("FFMA2 : %.3f ms  %.1f lane-FMA/clk/SM (2 per instr)\n", ms, 2 * n / (ms * 1e-3) / sms / (clk * 1e3));
    ms = time_it([&] { k_ex2<<<blocks, threads>>>(out, 0.999f); });
    printf("EX2+FMUL: %.3f ms  %.1f lane-EX2/clk/SM\n", ms, n / (ms * 1e-3) / sms / (clk * 1e3));
    ms = time_it([&] { k_mix<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
    printf("18 FFMA + 1 EX2: %.3f ms  %.2f clk per warp-tap per SMSP\n", ms,
           (ms * 1e-3) * (clk * 1e3) / ((double)blocks * threads / 32 / sms / 4 * ITERS));
    ms = time_it([&] { k_lds128<<<blocks, threads>>>(out); });
    printf("LDS.128 (+4 FADD): %.3f ms  %.1f B/clk/SM\n", ms, n * 16 / (ms * 1e-3) / sms / (clk * 1e3));
    return 0;
}
