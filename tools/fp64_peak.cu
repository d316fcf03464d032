// FP64 roofline microbenchmark for B200: DFMA (vector pipe), DMMA (mma.sync m8n8k4 f64) and both interleaved.
// Prints one JSON object.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; } } while (0)

constexpr int ITERS = 4096;

__global__ void dfma_kernel(double* out, double a, double b) {
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = a + i + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fma(x[i], b, a);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void dmma_kernel(double* out, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = threadIdx.x; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void mixed_kernel(double* out, double a, double b) {
    double c[4][2], x[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) { c[i][0] = i; c[i][1] = threadIdx.x; }
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = a + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            dmma(c[i][0], c[i][1], a, b);
            x[2 * i] = fma(x[2 * i], b, a);
            x[2 * i + 1] = fma(x[2 * i + 1], b, a);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
static double time_kernel(K k, int grid, int block, double* out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) k<<<grid, block>>>(out, 1.0000001, 0.9999999);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k<<<grid, block>>>(out, 1.0000001, 0.9999999);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best * 1e-3;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    double* out;
    CK(cudaMalloc(&out, (size_t)sms * 16 * 1024 * sizeof(double)));
    printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);
    const int warps_list[] = {4, 8, 16, 32};
    for (int wi = 0; wi < 4; ++wi) {
        const int block = 256, per_sm = warps_list[wi] * 32 / block;
        const int grid = sms * (per_sm > 0 ? per_sm : 1);
        const int blk = per_sm > 0 ? block : warps_list[wi] * 32;
        const double threads = (double)grid * blk;
        double t = time_kernel(dfma_kernel, grid, blk, out);
        printf(", \"dfma_tflops_w%d\": %.3f", warps_list[wi], threads * 16.0 * ITERS * 2.0 / t * 1e-12);
        t = time_kernel(dmma_kernel, grid, blk, out);
        printf(", \"dmma_tflops_w%d\": %.3f", warps_list[wi], threads / 32.0 * 8.0 * ITERS * 512.0 / t * 1e-12);
        t = time_kernel(mixed_kernel, grid, blk, out);
        printf(", \"mixed_tflops_w%d\": %.3f", warps_list[wi], (threads / 32.0 * 4.0 * ITERS * 512.0 + threads * 8.0 * ITERS * 2.0) / t * 1e-12);
    }
    // sustained DFMA (about 2 s) to see the power-capped figure
    {
        const int grid = sms * 4, blk = 256;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0);
        int n = 0;
        for (; n < 400; ++n) dfma_kernel<<<grid, blk>>>(out, 1.0000001, 0.9999999);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf(", \"dfma_tflops_sustained\": %.3f, \"sustained_seconds\": %.2f", (double)grid * blk * 16.0 * ITERS * 2.0 * n / (ms * 1e-3) * 1e-12, ms * 1e-3);
    }
    printf("}\n");
    return 0;
}
