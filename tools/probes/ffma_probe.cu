// ffma_probe.cu -- FP32 pipe throughput on sm_100a: scalar FFMA (3 register operands) vs packed fma.rn.f32x2.
// Decides whether an FMA-heavy HBM kernel (the 16-tap upfirdn2d blur in 2-byte storage needs 26e12 FMA/s at the
// HBM roofline) has to be written on packed math.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a ffma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NACC = 16;     // independent accumulators per thread
constexpr int ITERS = 4096;

__global__ void k_ffma(float *out, float a, float b) {
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    float m0 = a, m1 = b;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fmaf(acc[i], (i & 1) ? m0 : m1, (i & 2) ? m1 : m0);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// window value x tap -> accumulator: the stencil's operand pattern (accumulator is also the addend)
__global__ void k_ffma_acc(float *out, float a, float b) {
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    float w[4] = {a, b, a + 1.f, b + 1.f}, k[4] = {b, a, b * 2.f, a * 2.f};
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fmaf(w[i & 3], k[(i >> 2) & 3], acc[i]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}

__global__ void k_ffma2_acc(float *out, float a, float b) {
    float2 acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i);
    float2 w[4] = {{a, b}, {b, a}, {a + 1.f, b}, {b + 1.f, a}}, k[4] = {{b, b}, {a, a}, {b * 2.f, b * 2.f}, {a * 2.f, a * 2.f}};
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = ffma2(w[i & 3], k[(i >> 2) & 3], acc[i]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static void run(const char *name, F kern, double fma_per_thread, int threads, int blocks_per_sm) {
    int dev = 0, sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    float *out;
    const int grid = sms * blocks_per_sm;
    cudaMalloc(&out, sizeof(float) * grid * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) kern<<<grid, threads>>>(out, 1.0001f, 0.9999f);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) kern<<<grid, threads>>>(out, 1.0001f, 0.9999f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = fma_per_thread * threads * grid * 10;
    const double per_s = fma / (ms * 1e-3);
    printf("%-14s %4d thr x %d/SM: %.2f TFMA/s = %.1f FMA/clk/SM at the nominal %d MHz (%s)\n", name, threads, blocks_per_sm,
           per_s * 1e-12, per_s / sms / (khz * 1e3), khz / 1000, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    for (int thr : {128, 256, 512}) {
        run("ffma", k_ffma, (double)NACC * ITERS, thr, 2);
        run("ffma_acc", k_ffma_acc, (double)NACC * ITERS, thr, 2);
        run("ffma2_acc", k_ffma2_acc, 2.0 * NACC * ITERS, thr, 2);
    }
    return 0;
}
