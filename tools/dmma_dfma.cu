// Microbenchmark: do the FP64 tensor-core instruction DMMA (mma.sync.m8n8k4.f64) and the
// FP64 FMA pipe issue concurrently on sm_100a, or do they share the pipe?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/dmma_dfma tools/dmma_dfma.cu && /tmp/dmma_dfma
// Three kernels, same grid (148 x 4 CTAs of 256 threads), each warp runs ITERS iterations of
//   A: 8 independent DFMA chains            (8 DFMA per iteration)
//   B: 4 independent DMMA accumulators      (4 DMMA per iteration; one DMMA = 256 FMA per warp = 8 DFMA-equivalents)
//   C: both interleaved                     (8 DFMA + 4 DMMA per iteration)
// If the pipes were independent, t(C) ~ max(t(A), t(B)); if shared, t(C) ~ t(A) + t(B).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, double a, double b) {
    double f[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { f[i] = threadIdx.x * 1e-3 + i; c[i] = i; }
    for (int it = 0; it < ITERS; ++it) {
        if (MODE != 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = fma(f[i], a, b);
        }
        if (MODE != 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) dmma(c[2 * i], c[2 * i + 1], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += f[i] + c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
float run(double* out, int grid) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(out, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<grid, 256>>>(out, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * 4;
    double* out;
    cudaMalloc(&out, sizeof(double) * grid * 256);
    const float ta = run<0>(out, grid), tb = run<1>(out, grid), tc = run<2>(out, grid);
    const double warps = (double)grid * 8, fma_a = warps * ITERS * 8 * 32, fma_b = warps * ITERS * 4 * 256;
    printf("{\"sms\": %d, \"dfma_only_ms\": %.4f, \"dmma_only_ms\": %.4f, \"both_ms\": %.4f, "
           "\"dfma_tflops\": %.2f, \"dmma_tflops\": %.2f, \"both_over_sum\": %.3f, \"both_over_max\": %.3f}\n",
           sms, ta, tb, tc, 2 * fma_a / (ta * 1e-3) / 1e12, 2 * fma_b / (tb * 1e-3) / 1e12, tc / (ta + tb), tc / (ta > tb ? ta : tb));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
