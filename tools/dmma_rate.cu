// How fast can W warps per SM sub-partition stream independent DMMA m8n8k4?  (cycles per DMMA per sub-partition; the pipe's
// rate is 16.)  Decides whether one warp alone can keep the FP64 tensor pipe busy (it cannot: see profiles/r02_dmma_rate.txt).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void probe(double* out, long long* cyc, int iters) {
    const int lane = threadIdx.x & 31;
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = i + lane; c[i][1] = i * lane; }
    const double a = 1.0 + 1e-3 * lane, b = 1.0 - 1e-3 * lane;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int NACC>
void run(int warps_per_smsp, double* out, long long* cyc) {
    const int iters = 2000;
    const int threads = warps_per_smsp * 4 * 32;
    probe<NACC><<<148, threads>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    probe<NACC><<<148, threads>>>(out, cyc, iters);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double m = 0; for (int i = 0; i < 148; ++i) m += h[i];
    m /= 148;
    printf("%d warp(s)/sub-partition, %2d independent accumulators: %.2f cycles per DMMA per warp, %.2f per sub-partition\n", warps_per_smsp, NACC,
           m / iters / NACC, m / iters / NACC / warps_per_smsp);
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 4096 * 8); cudaMalloc(&cyc, 148 * 8);
    for (int w = 1; w <= 4; ++w) { run<25>(w, out, cyc); run<8>(w, out, cyc); run<2>(w, out, cyc); run<1>(w, out, cyc); }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
