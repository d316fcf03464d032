// Latency / throughput probes for the FP64 path on B200 (single warp unless noted).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma_chain(double* out, long long* cyc, double a, double b) {
    double x = a;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) {
        x = fma(x, b, a); x = fma(x, b, a); x = fma(x, b, a); x = fma(x, b, a);
        x = fma(x, b, a); x = fma(x, b, a); x = fma(x, b, a); x = fma(x, b, a);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = x;
}
// 3 distinct register operands per DFMA, 25 independent accumulators (like the rank-1 update)
__global__ void k_dfma_3op(double* out, long long* cyc, const double* in) {
    double acc[25], c[5], r[5];
    for (int i = 0; i < 5; ++i) { c[i] = in[i + threadIdx.x]; r[i] = in[8 + i + threadIdx.x]; }
    for (int i = 0; i < 25; ++i) acc[i] = in[16 + i];
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 128; ++it) {
#pragma unroll
        for (int a = 0; a < 5; ++a)
#pragma unroll
            for (int b = 0; b < 5; ++b) acc[a * 5 + b] = fma(c[a], r[b], acc[a * 5 + b]);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    double s = 0; for (int i = 0; i < 25; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_shfl_chain(double* out, long long* cyc, double a) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) { x = __shfl_xor_sync(0xffffffffu, x, 1); x = __shfl_xor_sync(0xffffffffu, x, 2); x = __shfl_xor_sync(0xffffffffu, x, 4); x = __shfl_xor_sync(0xffffffffu, x, 1); }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = x;
}
__global__ void k_bar64(long long* cyc) {
    const int bar = 1 + (threadIdx.x >> 6);
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; ++i) asm volatile("bar.sync %0, 64;" ::"r"(bar) : "memory");
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds_chain(long long* cyc, int* out) {
    __shared__ int s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i * 17 + 5) & 1023;
    __syncthreads();
    int x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; ++i) x = s[x];
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = x;
}
__global__ void k_rcp_chain(double* out, long long* cyc, double a) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) { x = 1.0 / x + 1.5; }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = x;
}
int main() {
    double *out, *in; long long* cyc; int* io;
    cudaMalloc(&out, 1 << 22); cudaMalloc(&in, 4096); cudaMalloc(&cyc, 8 * 4096); cudaMalloc(&io, 4096); cudaMemset(in, 0, 4096);
    long long h[4096];
    k_dfma_chain<<<1, 32>>>(out, cyc, 1.0000001, 0.999999); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("{\"dfma_dependent_latency_cycles\": %.2f", h[0] / 2048.0);
    for (int warps = 1; warps <= 16; warps *= 2) {
        k_dfma_3op<<<1, 32 * warps>>>(out, cyc, in); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        printf(", \"dfma_3op_cycles_per_warp_instr_%dwarps_1sm\": %.3f", warps, h[0] / (128.0 * 25) / warps * 4);   // per SMSP issue-cycle, normalised: cycles*4SMSP/(instr*warps)
    }
    k_shfl_chain<<<1, 32>>>(out, cyc, 1.0); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf(", \"shfl64_dependent_latency_cycles\": %.2f", h[0] / 1024.0);
    k_bar64<<<1, 64>>>(cyc); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf(", \"bar_sync_64_cycles\": %.2f", h[0] / 1024.0);
    k_bar64<<<1, 384>>>(cyc); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf(", \"bar_sync_64_cycles_6groups\": %.2f", h[0] / 1024.0);
    k_lds_chain<<<1, 32>>>(cyc, io); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf(", \"lds_dependent_latency_cycles\": %.2f", h[0] / 1024.0);
    k_rcp_chain<<<1, 32>>>(out, cyc, 1.0); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf(", \"double_div_plus_add_latency_cycles\": %.2f", h[0] / 256.0);
    printf("}\n");
    return 0;
}
