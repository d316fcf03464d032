// Probe for the DMMA-based block Gauss-Jordan design: per "block step" a warp does 10 fragment LDS.64,
// 50 DMMA m8n8k4 on 25 resident 8x8 tiles, PANEL dependent-ish DFMA work and 2 group barriers.
// Reports SM cycles per block step for 12 warps/SM (3 per SMSP); ideal DMMA-only = 3*50*16 = 2400.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int PANEL, bool BAR>
__global__ void __launch_bounds__(384, 1) probe(double* out, long long* cyc, int iters) {
    __shared__ double sm[12][2][64 * 4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bar = 1 + (warp >> 1);
    double c[25][2];
#pragma unroll
    for (int i = 0; i < 25; ++i) { c[i][0] = i + lane; c[i][1] = i * lane; }
    for (int i = lane; i < 512; i += 32) (&sm[warp][0][0])[i] = 1e-3 * (i + warp);
    __syncthreads();
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = 1.0 + 1e-3 * (lane + i);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const double* L = &sm[warp][it & 1][0];
        const double* U = &sm[warp ^ 1][it & 1][0];
        double u[5][2];
#pragma unroll
        for (int tb = 0; tb < 5; ++tb) { u[tb][0] = U[tb * 32 + lane]; u[tb][1] = U[160 + tb * 32 + lane - 96 * (tb > 2)]; }
#pragma unroll
        for (int ta = 0; ta < 5; ++ta) {
            const double l0 = L[ta * 32 + lane], l1 = L[(ta * 32 + lane + 64) & 255];
#pragma unroll
            for (int tb = 0; tb < 5; ++tb) {
                dmma(c[ta * 5 + tb][0], c[ta * 5 + tb][1], l0, u[tb][0]);
                dmma(c[ta * 5 + tb][0], c[ta * 5 + tb][1], l1, u[tb][1]);
            }
            if (ta == 1 && BAR) asm volatile("bar.sync %0, 64;" ::"r"(bar) : "memory");
            // panel-like work: PANEL/5 dependent-chain DFMAs per ta over 8 chains
#pragma unroll
            for (int q = 0; q < PANEL / 5; ++q) x[q & 7] = fma(x[q & 7], x[(q + 3) & 7], 0.25);
        }
        if (lane < 20) sm[warp][(it & 1) ^ 1][lane * 4] = x[0] * 1e-9;
        if (BAR) asm volatile("bar.sync %0, 64;" ::"r"(bar) : "memory");
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 25; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 123.456) out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int PANEL, bool BAR>
void run(const char* name, double* out, long long* cyc) {
    const int iters = 2000;
    probe<PANEL, BAR><<<148, 384>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    probe<PANEL, BAR><<<148, 384>>>(out, cyc, iters);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double m = 0; for (int i = 0; i < 148; ++i) m += h[i];
    printf("%s: %.1f SM cycles per block step (ideal DMMA only 2400, +panel %d DFMA -> %d)\n", name, m / 148 / iters, PANEL, 2400 + PANEL * 6);
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 4096 * 8); cudaMalloc(&cyc, 148 * 8);
    run<0, false>("dmma only, no barrier", out, cyc);
    run<0, true>("dmma only, 2 barriers", out, cyc);
    run<60, true>("panel 60", out, cyc);
    run<120, true>("panel 120", out, cyc);
    run<240, true>("panel 240", out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
