// Generic-size (m > 40) MVAR path: the same mathematics as mvar_kernels.cu, organised for matrices that do not
// fit one 40 x 40 register tile (BASELINE cfg5: 2 x 64 channels, p = 15, F = 512, 100 trials per window).
//
//   LWR (ar_coeff, src/mtmvar.py:90-123) as batched block GEMMs.  With SA = [A_1 ... A_k] (m x km) and
//   SB = [B_k ... B_1] (reversed, right-aligned in a p-block buffer) one order is
//        D   = Gamma(k+1) - SA * [Gamma(k); ...; Gamma(1)]          (one GEMM of depth k m)
//        Kf  = D Vb^-1,  Kb = D^T Vf^-1                              (2 SPD inverses + 2 GEMMs)
//        SA' = SA - Kf SB,   SB' = SB - Kb SA                        (2 GEMMs, block-aligned thanks to the reversal)
//        Vf -= Kf D^T,  Vb -= Kb D;   SA gets Kf appended, SB gets Kb prepended
//   A(f)^-1 (mvar_transfer_function, src/mtmvar.py:155-160) by Gauss-Jordan with partial pivoting on an
//   L2-resident scratch matrix per CTA.
// These kernels favour simplicity; the register-tile kernels remain the fast path for m <= 40.
#include <cstdio>

#include "hs_tile.cuh"
#include "hs_internal.h"
#include "mvar_launch.h"

namespace hs {

// ------------------------------------------------------------------------------------------------
// Batched GEMM  C[b] = beta * C[b] + alpha * op(A[b]) * op(B[b]),  op = identity or transpose, row-major.
// One 64-thread tile group per 40 x 40 block of C, 40-deep panels staged k-major in shared memory.
// ------------------------------------------------------------------------------------------------
struct GemmArgs {
    const double* A; const double* B; double* C;
    long long sA, sB, sC;        // batch strides (elements)
    int lda, ldb, ldc;
    int M, N, K;
    int tA, tB;                  // 1: operand stored transposed (A is K x M / B is N x K)
    double alpha, beta;
};

#ifdef HS_EXPERIMENT      // round 1's register-tile DFMA version (HS_BGEMM_DFMA=1): 10.5 ms of GEMMs per cfg5 fit
__global__ void __launch_bounds__(64) bgemm_kernel(const GemmArgs g) {
    __shared__ double pa[kPadMax * kPadMax], pb[kPadMax * kPadMax];
    const Group grp = make_group();
    const int tiles_n = (g.N + kPadMax - 1) / kPadMax;
    const int bi = blockIdx.x / tiles_n, bj = blockIdx.x % tiles_n;
    const int i0 = bi * kPadMax, j0 = bj * kPadMax;
    const double* A = g.A + (long long)blockIdx.y * g.sA;
    const double* B = g.B + (long long)blockIdx.y * g.sB;
    double* C = g.C + (long long)blockIdx.y * g.sC;
    double acc[kTileMax][kTileMax];
#pragma unroll
    for (int a = 0; a < kTileMax; ++a)
#pragma unroll
        for (int b = 0; b < kTileMax; ++b) acc[a][b] = 0.0;
    for (int k0 = 0; k0 < g.K; k0 += kPadMax) {
        __syncthreads();
        for (int e = threadIdx.x; e < kPadMax * kPadMax; e += 64) {
            const int kk = e / kPadMax, c = e - kk * kPadMax;       // pa[kk][c] = opA[i0+c][k0+kk], pb[kk][c] = opB[k0+kk][j0+c]
            double va = 0.0, vb = 0.0;
            if (k0 + kk < g.K) {
                if (i0 + c < g.M) va = g.tA ? A[(long long)(k0 + kk) * g.lda + i0 + c] : A[(long long)(i0 + c) * g.lda + k0 + kk];
                if (j0 + c < g.N) vb = g.tB ? B[(long long)(j0 + c) * g.ldb + k0 + kk] : B[(long long)(k0 + kk) * g.ldb + j0 + c];
            }
            pa[e] = va;
            pb[e] = vb;
        }
        __syncthreads();
        tile_mac<kTileMax, false>(acc, pa, kPadMax, pb, kPadMax, min(kPadMax, g.K - k0), grp);
    }
#pragma unroll
    for (int a = 0; a < kTileMax; ++a) {
        const int i = i0 + grp.tr + 8 * a;
#pragma unroll
        for (int b = 0; b < kTileMax; ++b) {
            const int j = j0 + grp.tc + 8 * b;
            if (i < g.M && j < g.N) {
                double* c = C + (long long)i * g.ldc + j;
                *c = (g.beta == 0.0 ? 0.0 : g.beta * *c) + g.alpha * acc[a][b];
            }
        }
    }
}

#endif

// The same batched GEMM on the FP64 tensor pipe: 64 x 64 block of C per CTA, 4 warps (2 x 2, 32 x 32 outputs = 16 accumulator tiles
// each), K in chunks of 32 staged as As[i][k] / Bs[j][k] (row stride 36 = 4 mod 16 doubles: conflict-free m8n8k4 fragments for both).
constexpr int kGB = 64, kGK = 32, kGLd = kGK + 4;
__device__ __forceinline__ void dmma884_b(double& c0, double& c1, const double a, const double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(128) bgemm_mma_kernel(const GemmArgs g) {
    __shared__ double As[kGB * kGLd], Bs[kGB * kGLd];
    const int tiles_n = (g.N + kGB - 1) / kGB;
    const int i0 = (blockIdx.x / tiles_n) * kGB, j0 = (blockIdx.x % tiles_n) * kGB;
    const double* A = g.A + (long long)blockIdx.y * g.sA;
    const double* B = g.B + (long long)blockIdx.y * g.sB;
    double* C = g.C + (long long)blockIdx.y * g.sC;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g4 = lane >> 2, t4 = lane & 3;
    const int wr = warp >> 1, wc = warp & 1;
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    for (int k0 = 0; k0 < g.K; k0 += kGK) {
        __syncthreads();
        // stage along the contiguous direction of each operand
        for (int e = threadIdx.x; e < kGB * kGK; e += 128) {
            int i, k;
            if (g.tA) { i = e % kGB; k = e / kGB; } else { k = e % kGK; i = e / kGK; }
            double v = 0.0;
            if (i0 + i < g.M && k0 + k < g.K) v = g.tA ? A[(long long)(k0 + k) * g.lda + i0 + i] : A[(long long)(i0 + i) * g.lda + k0 + k];
            As[i * kGLd + k] = v;
        }
        for (int e = threadIdx.x; e < kGB * kGK; e += 128) {
            int j, k;
            if (g.tB) { k = e % kGK; j = e / kGK; } else { j = e % kGB; k = e / kGB; }
            double v = 0.0;
            if (j0 + j < g.N && k0 + k < g.K) v = g.tB ? B[(long long)(j0 + j) * g.ldb + k0 + k] : B[(long long)(k0 + k) * g.ldb + j0 + j];
            Bs[j * kGLd + k] = v;
        }
        __syncthreads();
        const double* pa = As + (32 * wr + g4) * kGLd + t4;
        const double* pb = Bs + (32 * wc + g4) * kGLd + t4;
#pragma unroll
        for (int kk = 0; kk < kGK; kk += 4) {
            double a[4], b[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                a[q] = pa[q * 8 * kGLd + kk];
                b[q] = pb[q * 8 * kGLd + kk];
            }
#pragma unroll
            for (int qa = 0; qa < 4; ++qa)
#pragma unroll
                for (int qb = 0; qb < 4; ++qb) dmma884_b(acc[qa][qb][0], acc[qa][qb][1], a[qa], b[qb]);
        }
    }
#pragma unroll
    for (int qa = 0; qa < 4; ++qa) {
        const int i = i0 + 32 * wr + 8 * qa + g4;
#pragma unroll
        for (int qb = 0; qb < 4; ++qb) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = j0 + 32 * wc + 8 * qb + 2 * t4 + h;
                if (i < g.M && j < g.N) {
                    double* c = C + (long long)i * g.ldc + j;
                    *c = (g.beta == 0.0 ? 0.0 : g.beta * *c) + g.alpha * acc[qa][qb][h];
                }
            }
        }
    }
}

static int bgemm(const GemmArgs& g, int batch, cudaStream_t st) {
#ifdef HS_EXPERIMENT
    if (exp_env_int("HS_BGEMM_DFMA", 0) == 1) {
        dim3 grid0(((g.M + kPadMax - 1) / kPadMax) * ((g.N + kPadMax - 1) / kPadMax), batch);
        bgemm_kernel<<<grid0, 64, 0, st>>>(g);
        return check_launch("bgemm_kernel");
    }
#endif
    dim3 grid(((g.M + kGB - 1) / kGB) * ((g.N + kGB - 1) / kGB), batch);
    bgemm_mma_kernel<<<grid, 128, 0, st>>>(g);
    return check_launch("bgemm_mma_kernel");
}

// ------------------------------------------------------------------------------------------------
// Batched inverse of m x m real matrices (the residual covariances, SPD) by Gauss-Jordan with partial
// pivoting inside shared memory (m <= 160), one CTA per matrix.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) binv_kernel(const double* __restrict__ V, double* __restrict__ X, int m, int* status, int flag) {
    extern __shared__ __align__(16) double sm[];
    double* a = sm;                         // m x m
    double* col = sm + (size_t)m * m;       // m
    __shared__ int s_piv;
    __shared__ double s_red[256];
    __shared__ int s_idx[256];
    int* perm = reinterpret_cast<int*>(col + m);   // rowmap / colmap  (2m ints)
    int* used = perm + 2 * m;
    const double* v = V + (size_t)blockIdx.x * m * m;
    for (int e = threadIdx.x; e < m * m; e += 256) a[e] = v[e];
    for (int e = threadIdx.x; e < m; e += 256) used[e] = 0;
    __syncthreads();
    for (int k = 0; k < m; ++k) {
        double best = -1.0;
        int bi = 1 << 20;
        for (int i = threadIdx.x; i < m; i += 256) {
            const double mag = fabs(a[i * m + k]);
            if (!used[i] && (mag > best || bi == (1 << 20))) { if (mag > best) best = mag; bi = i; }
        }
        s_red[threadIdx.x] = best;
        s_idx[threadIdx.x] = bi;
        __syncthreads();
        for (int off = 128; off > 0; off >>= 1) {
            if (threadIdx.x < off) {
                const double ob = s_red[threadIdx.x + off];
                const int oi = s_idx[threadIdx.x + off];
                if (ob > s_red[threadIdx.x] || (ob == s_red[threadIdx.x] && oi < s_idx[threadIdx.x])) { s_red[threadIdx.x] = ob; s_idx[threadIdx.x] = oi; }
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            s_piv = s_idx[0];
            used[s_idx[0]] = 1;
            perm[s_idx[0]] = k;            // rowmap[r] = k
            perm[m + k] = s_idx[0];        // colmap[k] = r
            if (!(s_red[0] > 0.0)) atomicOr(status + blockIdx.x, flag);
        }
        __syncthreads();
        const int r = s_piv;
        const double inv = 1.0 / a[r * m + k];
        for (int i = threadIdx.x; i < m; i += 256) col[i] = (i == r) ? 0.0 : a[i * m + k];
        __syncthreads();
        for (int j = threadIdx.x; j < m; j += 256) a[r * m + j] = (j == k) ? inv : a[r * m + j] * inv;
        __syncthreads();
        for (int e = threadIdx.x; e < m * m; e += 256) {
            const int i = e / m, j = e - i * m;
            if (i != r) a[e] = ((j == k) ? 0.0 : a[e]) - col[i] * a[r * m + j];
        }
        __syncthreads();
    }
    double* x = X + (size_t)blockIdx.x * m * m;
    for (int e = threadIdx.x; e < m * m; e += 256) {
        const int i = e / m, j = e - i * m;
        x[(size_t)perm[i] * m + perm[m + j]] = a[e];
    }
}


// ------------------------------------------------------------------------------------------------
// Inverse of symmetric positive definite m x m matrices (the residual covariances Vf / Vb of the LWR recursion), m <= 128:
// in-place Gauss-Jordan WITHOUT pivoting (stable for SPD matrices; the m <= 40 path does the same) with the matrix in REGISTERS:
// 32 x 32 threads, thread (ty, tx) owns the cyclic 4 x 4 sub-matrix rows ty + 32 a, columns tx + 32 b.  Per pivot the owners of
// column k and row k publish them to a double-buffered shared-memory pair, ONE barrier, then every thread updates its 16
// entries: 128 barriers per inverse (binv_kernel: pivot search + index arithmetic, ~1.2 ms per 128 x 128 matrix; this: ~15 us).
// ------------------------------------------------------------------------------------------------
constexpr int kSpdE = 4;
__global__ void __launch_bounds__(1024, 1) binv_spd_kernel(const double* __restrict__ V, double* __restrict__ X, const int m, int* status, const int flag) {
    __shared__ double colb[2][32 * kSpdE], rowb[2][32 * kSpdE];
    __shared__ int bad;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const double* v = V + (size_t)blockIdx.x * m * m;
    double a[kSpdE][kSpdE];
#pragma unroll
    for (int ia = 0; ia < kSpdE; ++ia)
#pragma unroll
        for (int ib = 0; ib < kSpdE; ++ib) {
            const int i = ty + 32 * ia, j = tx + 32 * ib;
            a[ia][ib] = (i < m && j < m) ? v[(size_t)i * m + j] : (i == j ? 1.0 : 0.0);       // identity padding
        }
    if (threadIdx.x == 0) bad = 0;
    for (int k = 0; k < m; ++k) {
        const int buf = k & 1, ka = k >> 5, kt = k & 31;
        if (tx == kt) {
#pragma unroll
            for (int ia = 0; ia < kSpdE; ++ia)
#pragma unroll
                for (int ib = 0; ib < kSpdE; ++ib)
                    if (ib == ka) colb[buf][ty + 32 * ia] = a[ia][ib];
        }
        if (ty == kt) {
#pragma unroll
            for (int ia = 0; ia < kSpdE; ++ia)
#pragma unroll
                for (int ib = 0; ib < kSpdE; ++ib)
                    if (ia == ka) rowb[buf][tx + 32 * ib] = a[ia][ib];
        }
        __syncthreads();
        const double pv = rowb[buf][k];
        if (threadIdx.x == 0 && !(pv > 0.0)) bad = 1;          // not positive definite (or NaN)
        const double inv = 1.0 / pv;
        double cv[kSpdE], rv[kSpdE];
#pragma unroll
        for (int q = 0; q < kSpdE; ++q) {
            cv[q] = colb[buf][ty + 32 * q];
            rv[q] = rowb[buf][tx + 32 * q] * inv;
        }
#pragma unroll
        for (int ia = 0; ia < kSpdE; ++ia) {
            const int i = ty + 32 * ia;
#pragma unroll
            for (int ib = 0; ib < kSpdE; ++ib) {
                const int j = tx + 32 * ib;
                double x;
                if (i == k) x = (j == k) ? inv : rv[ib];
                else if (j == k) x = -cv[ia] * inv;
                else x = fma(-cv[ia], rv[ib], a[ia][ib]);
                a[ia][ib] = x;
            }
        }
        // no second barrier: the next pivot publishes into the other buffer
    }
    __syncthreads();
    if (threadIdx.x == 0 && bad) atomicOr(status + blockIdx.x, flag);
    double* x = X + (size_t)blockIdx.x * m * m;
#pragma unroll
    for (int ia = 0; ia < kSpdE; ++ia)
#pragma unroll
        for (int ib = 0; ib < kSpdE; ++ib) {
            const int i = ty + 32 * ia, j = tx + 32 * ib;
            if (i < m && j < m) x[(size_t)i * m + j] = a[ia][ib];
        }
}

static int launch_binv(const double* V, double* X, int m, int n, int* status, size_t inv_smem, cudaStream_t st) {
    if (m <= 32 * kSpdE) {
        binv_spd_kernel<<<n, 1024, 0, st>>>(V, X, m, status, HS_STATUS_SINGULAR_YW);
        return check_launch("binv_spd_kernel");
    }
    binv_kernel<<<n, 256, inv_smem, st>>>(V, X, m, status, HS_STATUS_SINGULAR_YW);
    return check_launch("binv_kernel");
}

__global__ void transpose_stack_kernel(const double* __restrict__ R, int m, int p, double* __restrict__ Grev, double* __restrict__ Vf,
                                       double* __restrict__ Vb) {
    // Grev[w] = [Gamma(p); ...; Gamma(1)] (p m x m), Gamma(l) = R(l)^T;  Vf = Vb = Gamma(0)
    const int w = blockIdx.y;
    const size_t mm = (size_t)m * m;
    const double* Rw = R + (size_t)w * (p + 1) * mm;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < (size_t)(p + 1) * mm; e += (size_t)gridDim.x * blockDim.x) {
        const int l = (int)(e / mm);
        const size_t ij = e - (size_t)l * mm;
        const int i = (int)(ij / m), j = (int)(ij % m);
        const double v = Rw[(size_t)l * mm + (size_t)j * m + i];
        if (l == 0) {
            Vf[(size_t)w * mm + ij] = v;
            Vb[(size_t)w * mm + ij] = v;
        } else {
            Grev[((size_t)w * p + (p - l)) * mm + ij] = v;
        }
    }
}

__global__ void copy_block_kernel(const double* __restrict__ src, long long s_src, int ld_src, double* __restrict__ dst, long long s_dst,
                                  int ld_dst, int rows, int cols) {
    const double* s = src + (long long)blockIdx.y * s_src;
    double* d = dst + (long long)blockIdx.y * s_dst;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)rows * cols; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e / cols), j = (int)(e % cols);
        d[(long long)i * ld_dst + j] = s[(long long)i * ld_src + j];
    }
}

__global__ void sa_to_coeffs_kernel(const double* __restrict__ SA, int m, int p, double* __restrict__ A) {
    // A[w][i][j][k] = A_{k+1}[i][j] = SA[w][i][k m + j]
    const int w = blockIdx.y;
    const size_t tot = (size_t)m * m * p;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (size_t)gridDim.x * blockDim.x) {
        const size_t ij = e / p;
        const int k = (int)(e - ij * p);
        const int i = (int)(ij / m), j = (int)(ij % m);
        A[(size_t)w * tot + e] = SA[(size_t)w * tot + (size_t)i * m * p + (size_t)k * m + j];
    }
}

size_t lwr_generic_ws_doubles(int n_win, int m, int p) {
    const size_t mm = (size_t)m * m;
    // Grev p, SA 2p (ping-pong), SB 2p, Vf, Vb, Xf, Xb, D, Kf, Kb
    return (size_t)n_win * mm * ((size_t)5 * p + 7);
}

int launch_lwr_generic(const K4Params& P, cudaStream_t st) {
    const int m = P.m, p = P.p, nw = P.n_win;
    if (m > 160) return set_error(HS_ERR_UNSUPPORTED, "yw_solve: m=%d > 160 (shared-memory inverse)", m);
    const long long mm = (long long)m * m;
    double* w = P.ws;
    double* Grev = w; w += (size_t)nw * p * mm;
    double* SA[2]; SA[0] = w; w += (size_t)nw * p * mm; SA[1] = w; w += (size_t)nw * p * mm;
    double* SB[2]; SB[0] = w; w += (size_t)nw * p * mm; SB[1] = w; w += (size_t)nw * p * mm;
    double* Vf = w; w += nw * mm;
    double* Vb = w; w += nw * mm;
    double* Xf = w; w += nw * mm;
    double* Xb = w; w += nw * mm;
    double* D = w; w += nw * mm;
    double* Kf = w; w += nw * mm;
    double* Kb = w; w += nw * mm;
    const int mp = m * p;
    dim3 g1(64, nw);
    transpose_stack_kernel<<<g1, 256, 0, st>>>(P.R, m, p, Grev, Vf, Vb);
    int rc = check_launch("transpose_stack_kernel");
    if (rc) return rc;
    const size_t inv_smem = (mm + m) * sizeof(double) + 3 * m * sizeof(int) + 16;
    cudaError_t ce = cudaFuncSetAttribute(binv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)inv_smem);
    if (ce != cudaSuccess) return set_error(HS_ERR_CUDA, "binv: %s", cudaGetErrorString(ce));
    int cur = 0;
    for (int k = 0; k < p; ++k) {
        const bool last = (k == p - 1);
        // D = Gamma(k+1) - SA[:, :km] * Grev[(p-k)m : pm, :]
        copy_block_kernel<<<dim3(16, nw), 256, 0, st>>>(Grev + (size_t)(p - k - 1) * mm, (long long)p * mm, m, D, (long long)mm, m, m, m);
        if ((rc = check_launch("copy_block_kernel"))) return rc;
        if (k > 0) {
            GemmArgs g{SA[cur], Grev + (size_t)(p - k) * mm, D, (long long)p * mm, (long long)p * mm, (long long)mm, mp, m, m, m, m, k * m, 0, 0, -1.0, 1.0};
            if ((rc = bgemm(g, nw, st))) return rc;
        }
        // inverses
        if ((rc = launch_binv(Vb, Xb, m, nw, P.status, inv_smem, st))) return rc;
        if (!last) {
            if ((rc = launch_binv(Vf, Xf, m, nw, P.status, inv_smem, st))) return rc;
        }
        {   // Kf = D * Xb
            GemmArgs g{D, Xb, Kf, (long long)mm, (long long)mm, (long long)mm, m, m, m, m, m, m, 0, 0, 1.0, 0.0};
            if ((rc = bgemm(g, nw, st))) return rc;
        }
        if (!last) {   // Kb = D^T * Xf
            GemmArgs g{D, Xf, Kb, (long long)mm, (long long)mm, (long long)mm, m, m, m, m, m, m, 1, 0, 1.0, 0.0};
            if ((rc = bgemm(g, nw, st))) return rc;
        }
        const int nxt = cur ^ 1;
        // SA' = SA - Kf * SB   (SB blocks (p-k)..(p-1) hold B_k..B_1)
        if (k > 0) {
            copy_block_kernel<<<dim3(32, nw), 256, 0, st>>>(SA[cur], (long long)p * mm, mp, SA[nxt], (long long)p * mm, mp, m, k * m);
            if ((rc = check_launch("copy_block_kernel"))) return rc;
            GemmArgs g{Kf, SB[cur] + (size_t)(p - k) * m, SA[nxt], (long long)mm, (long long)p * mm, (long long)p * mm, m, mp, mp, m, k * m, m, 0, 0, -1.0, 1.0};
            if ((rc = bgemm(g, nw, st))) return rc;
        }
        // append Kf as A_{k+1}
        copy_block_kernel<<<dim3(16, nw), 256, 0, st>>>(Kf, (long long)mm, m, SA[nxt] + (size_t)k * m, (long long)p * mm, mp, m, m);
        if ((rc = check_launch("copy_block_kernel"))) return rc;
        if (!last) {
            // SB' = SB - Kb * SA  on blocks (p-k)..(p-1); prepend Kb at block p-k-1
            if (k > 0) {
                copy_block_kernel<<<dim3(32, nw), 256, 0, st>>>(SB[cur] + (size_t)(p - k) * m, (long long)p * mm, mp, SB[nxt] + (size_t)(p - k) * m,
                                                                 (long long)p * mm, mp, m, k * m);
                if ((rc = check_launch("copy_block_kernel"))) return rc;
                GemmArgs g{Kb, SA[cur], SB[nxt] + (size_t)(p - k) * m, (long long)mm, (long long)p * mm, (long long)p * mm, m, mp, mp, m, k * m, m, 0, 0, -1.0, 1.0};
                if ((rc = bgemm(g, nw, st))) return rc;
            }
            copy_block_kernel<<<dim3(16, nw), 256, 0, st>>>(Kb, (long long)mm, m, SB[nxt] + (size_t)(p - k - 1) * m, (long long)p * mm, mp, m, m);
            if ((rc = check_launch("copy_block_kernel"))) return rc;
            // Vb -= Kb * D
            GemmArgs g{Kb, D, Vb, (long long)mm, (long long)mm, (long long)mm, m, m, m, m, m, m, 0, 0, -1.0, 1.0};
            if ((rc = bgemm(g, nw, st))) return rc;
        }
        {   // Vf -= Kf * D^T
            GemmArgs g{Kf, D, Vf, (long long)mm, (long long)mm, (long long)mm, m, m, m, m, m, m, 0, 1, -1.0, 1.0};
            if ((rc = bgemm(g, nw, st))) return rc;
        }
        if (P.Vall) {
            copy_block_kernel<<<dim3(16, nw), 256, 0, st>>>(Vf, (long long)mm, m, P.Vall + (size_t)k * mm, (long long)p * mm, m, m, m);
            if ((rc = check_launch("copy_block_kernel"))) return rc;
        }
        cur = nxt;
    }
    sa_to_coeffs_kernel<<<dim3(64, nw), 256, 0, st>>>(SA[cur], m, p, P.A);
    if ((rc = check_launch("sa_to_coeffs_kernel"))) return rc;
    copy_block_kernel<<<dim3(16, nw), 256, 0, st>>>(Vf, (long long)mm, m, P.V, (long long)mm, m, m, m);
    return check_launch("copy_block_kernel");
}

// ------------------------------------------------------------------------------------------------
// Generic A(f)^-1: persistent CTAs, one m x m complex matrix at a time in a private global scratch slot
// (L2-resident), Gauss-Jordan with partial pivoting, then |H|^2 / H / row sums.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transfer_generic_kernel(const K5Params P, double2* __restrict__ scratch) {
    const int m = P.m, p = P.p, F = P.F;
    double2* a = scratch + (size_t)blockIdx.x * m * m;
    extern __shared__ __align__(16) unsigned char smg[];
    double2* col = reinterpret_cast<double2*>(smg);                 // m
    double2* row = col + m;                                         // m
    int* rowmap = reinterpret_cast<int*>(row + m);
    int* colmap = rowmap + m;
    int* used = colmap + m;
    __shared__ double s_red[256];
    __shared__ int s_idx[256];
    __shared__ int s_piv;
    const long long n_mat = (long long)P.n_win * F;
    for (long long mat = blockIdx.x; mat < n_mat; mat += gridDim.x) {
        const int w = (int)(mat / F), f = (int)(mat % F);
        const double* Aw = P.A + (size_t)w * m * m * p;
        __syncthreads();
        for (int e = threadIdx.x; e < m * m; e += 256) {
            const int i = e / m, j = e - i * m;
            double re = (i == j) ? 1.0 : 0.0, im = 0.0;
            for (int k = 0; k < p; ++k) {
                const double2 z = __ldg(&P.z[(size_t)k * F + f]);
                const double c = Aw[(size_t)e * p + k];
                re = fma(-c, z.x, re);
                im = fma(-c, z.y, im);
            }
            a[e] = make_double2(re, im);
            if (P.Af) P.Af[((size_t)w * m * m + e) * F + f] = make_double2(re, im);
        }
        for (int e = threadIdx.x; e < m; e += 256) used[e] = 0;
        __syncthreads();
        for (int k = 0; k < m; ++k) {
            double best = -1.0;
            int bi = 1 << 20;
            for (int i = threadIdx.x; i < m; i += 256) {
                const double2 v = a[(size_t)i * m + k];
                const double mag = fma(v.x, v.x, v.y * v.y);
                if (!used[i] && (mag > best || bi == (1 << 20))) { if (mag > best) best = mag; bi = i; }
            }
            s_red[threadIdx.x] = best;
            s_idx[threadIdx.x] = bi;
            __syncthreads();
            for (int off = 128; off > 0; off >>= 1) {
                if (threadIdx.x < off) {
                    const double ob = s_red[threadIdx.x + off];
                    const int oi = s_idx[threadIdx.x + off];
                    if (ob > s_red[threadIdx.x] || (ob == s_red[threadIdx.x] && oi < s_idx[threadIdx.x])) { s_red[threadIdx.x] = ob; s_idx[threadIdx.x] = oi; }
                }
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                s_piv = s_idx[0];
                used[s_idx[0]] = 1;
                rowmap[s_idx[0]] = k;
                colmap[k] = s_idx[0];
                if (!(s_red[0] > 0.0)) atomicOr(&P.status[w], 1);
            }
            __syncthreads();
            const int r = s_piv;
            const double2 pv = a[(size_t)r * m + k];
            const double d = 1.0 / fma(pv.x, pv.x, pv.y * pv.y);
            const double2 iv = make_double2(pv.x * d, -pv.y * d);
            for (int i = threadIdx.x; i < m; i += 256) col[i] = (i == r) ? make_double2(0.0, 0.0) : a[(size_t)i * m + k];
            for (int j = threadIdx.x; j < m; j += 256) {
                const double2 x = (j == k) ? make_double2(1.0, 0.0) : a[(size_t)r * m + j];
                row[j] = make_double2(fma(x.x, iv.x, -x.y * iv.y), fma(x.x, iv.y, x.y * iv.x));
            }
            __syncthreads();
            for (int e = threadIdx.x; e < m * m; e += 256) {
                const int i = e / m, j = e - i * m;
                double2 x = a[e];
                if (i == r) {
                    x = row[j];
                } else {
                    if (j == k) x = make_double2(0.0, 0.0);
                    const double2 c = col[i], rv = row[j];
                    x.x = fma(-c.x, rv.x, fma(c.y, rv.y, x.x));
                    x.y = fma(-c.x, rv.y, fma(-c.y, rv.x, x.y));
                }
                a[e] = x;
            }
            __syncthreads();
        }
        // outputs: inverse[rowmap[i]][colmap[j]] = a[i][j]
        for (int e = threadIdx.x; e < m * m; e += 256) {
            const int i = e / m, j = e - i * m;
            const double2 v = a[e];
            const size_t o = ((size_t)w * m * m + (size_t)rowmap[i] * m + colmap[j]) * F + f;
            if (P.dtf) P.dtf[o] = fma(v.x, v.x, v.y * v.y);
            if (P.H) P.H[o] = v;
        }
        if (P.rowpart) {
            __syncthreads();
            for (int i = threadIdx.x; i < m; i += 256) {
                // row sums in output-row order: row rowmap[i] of H is storage row i
                double acc = 0.0;
                for (int j = 0; j < m; ++j) {
                    const double2 v = a[(size_t)i * m + j];
                    acc += fma(v.x, v.x, v.y * v.y);
                }
                P.rowpart[((size_t)w * F + f) * m + rowmap[i]] = acc;
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Blocked generic A(f)^-1 (40 < m <= 160; cfg5: m = 128): Gauss-Jordan with PARTIAL PIVOTING in blocks of 16 pivots, the
// rank-16 updates on the FP64 tensor pipe.  One CTA of 16 warps per matrix (persistent over matrices); the matrix lives in a
// private global scratch slot (256 KB at m = 128: L2 resident, too large for shared memory next to the panels).
// Per block of columns K = {k0 .. k0+15}:
//   1. the column panel C = S[:, K] is copied to shared memory (twice: one copy is destroyed by the pivot search);
//   2. pivot rows r_0 .. r_15 by LU-style elimination with partial pivoting INSIDE the panel (the panel holds every row, so these are
//      the pivots unblocked partial pivoting would choose: the Gauss-Jordan updates of already used rows do not touch unused rows);
//   3. D = C[R, :], P = D^-1 (16 x 16, elimination in pivot order), L = C P for the other rows, L[r_t, :] = e_t - P[t, :];
//   4. the panel columns of S get the identity pattern (1 at (r_t, k0 + t)), U = S[R, :] is gathered;
//   5. S <- S - L U for ALL rows: 4 DMMA m8n8k4 per (8 x 8 tile, 4 pivots) -- afterwards the pivot rows hold P U, the panel columns -L.
// The matrix is read and written once per 16 pivots instead of once per pivot (round 1's kernel: 64 MB of L2 traffic per matrix).
// Same result mapping as the unblocked kernel: inverse[rowmap[i]][colmap[j]] = a[i][j].
// ------------------------------------------------------------------------------------------------
constexpr int kNB = 16;
constexpr int kLLd = kNB + 4;           // row stride of the L panel in doubles (= 4 mod 16: conflict-free A fragments)
constexpr int kBlkThreads = 512;

__device__ __forceinline__ void dmma884_g(double& c0, double& c1, const double a, const double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 cmul_g(const double2 a, const double2 b) { return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x)); }

struct BlkSmem {
    static __host__ __device__ int mp(int m) { return (m + kNB - 1) / kNB * kNB; }
    static __host__ __device__ int ldu(int m) { return mp(m) + 4; }
    static __host__ __device__ size_t bytes(int m) {
        const size_t Mp = mp(m);
        return 2 * Mp * kNB * sizeof(double2)            // Cp, C0
               + 2 * Mp * kLLd * sizeof(double)          // Lre, Lim
               + 2 * (size_t)kNB * ldu(m) * sizeof(double)   // Ure, Uim
               + 2 * kNB * kNB * sizeof(double2)         // D, P
               + (3 * Mp + kNB) * sizeof(int) + 64;
    }
};

__global__ void __launch_bounds__(kBlkThreads, 1) transfer_blocked_kernel(const K5Params P, double2* __restrict__ scratch) {
    extern __shared__ __align__(16) unsigned char smb[];
    const int m = P.m, p = P.p, F = P.F;
    const int Mp = BlkSmem::mp(m), ldu = BlkSmem::ldu(m);
    double2* Cp = reinterpret_cast<double2*>(smb);                       // [Mp][16]
    double2* C0 = Cp + (size_t)Mp * kNB;                                 // [Mp][16]
    double* Lre = reinterpret_cast<double*>(C0 + (size_t)Mp * kNB);      // [Mp][kLLd]
    double* Lim = Lre + (size_t)Mp * kLLd;
    double* Ure = Lim + (size_t)Mp * kLLd;                               // [16][ldu]
    double* Uim = Ure + (size_t)kNB * ldu;
    double2* Dm = reinterpret_cast<double2*>(Uim + (size_t)kNB * ldu);   // [16][16]
    double2* Pm = Dm + kNB * kNB;
    int* rowmap = reinterpret_cast<int*>(Pm + kNB * kNB);                // [Mp]
    int* colmap = rowmap + Mp;
    int* used = colmap + Mp;
    int* piv = used + Mp;                                                // [16] pivot rows of the current block
    __shared__ double s_red[kBlkThreads / 32];
    __shared__ int s_idx[kBlkThreads / 32];
    double2* a = scratch + (size_t)blockIdx.x * Mp * Mp;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g4 = lane >> 2, t4 = lane & 3;
    const long long n_mat = (long long)P.n_win * F;
    for (long long mat = blockIdx.x; mat < n_mat; mat += gridDim.x) {
        const int w = (int)(mat / F), f = (int)(mat % F);
        const double* Aw = P.A + (size_t)w * m * m * p;
        __syncthreads();
        // z_k(f) of this bin: the D block is free until step 3 of the first panel
        double2* zf = Dm;
        for (int k = tid; k < p && k < kNB * kNB; k += kBlkThreads) zf[k] = __ldg(&P.z[(size_t)k * F + f]);
        __syncthreads();
        // ---- A(f) = I - sum_k A_k z_k(f), padded to Mp x Mp with the identity
        for (int e = tid; e < Mp * Mp; e += kBlkThreads) {
            const int i = e / Mp, j = e - i * Mp;
            double re = (i == j) ? 1.0 : 0.0, im = 0.0;
            if (i < m && j < m) {
                const double* cp = Aw + ((size_t)i * m + j) * p;
                for (int k = 0; k < p; ++k) {
                    const double2 z = (p <= kNB * kNB) ? zf[k] : __ldg(&P.z[(size_t)k * F + f]);
                    const double c = cp[k];
                    re = fma(-c, z.x, re);
                    im = fma(-c, z.y, im);
                }
                if (P.Af) P.Af[((size_t)w * m * m + (size_t)i * m + j) * F + f] = make_double2(re, im);
            }
            a[e] = make_double2(re, im);
        }
        for (int e = tid; e < Mp; e += kBlkThreads) used[e] = 0;
        __syncthreads();
        for (int k0 = 0; k0 < Mp; k0 += kNB) {
            // ---- 1. column panel -> shared memory
            for (int e = tid; e < Mp * kNB; e += kBlkThreads) {
                const int i = e / kNB, q = e - i * kNB;
                const double2 v = a[(size_t)i * Mp + k0 + q];
                Cp[e] = v;
                C0[e] = v;
            }
            __syncthreads();
            // ---- 2. pivot rows by partial pivoting inside the panel: thread i < Mp keeps row i of the panel in registers, the first
            //         ceil(Mp / 32) warps run the 16 elimination steps with two named barriers each (the other warps wait below)
            const int n_pw = (Mp + 31) >> 5;                      // panel warps
            if (warp < n_pw) {
                const bool has_row = tid < Mp;
                double2 r[kNB];
#pragma unroll
                for (int c = 0; c < kNB; ++c) r[c] = has_row ? Cp[tid * kNB + c] : make_double2(0.0, 0.0);
                bool mine_used = has_row ? (used[tid] != 0) : true;
#pragma unroll
                for (int q = 0; q < kNB; ++q) {
                    double best = mine_used ? -1.0 : fma(r[q].x, r[q].x, r[q].y * r[q].y);
                    int bi = mine_used ? (1 << 20) : tid;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        const double ob = __shfl_xor_sync(0xffffffffu, best, off);
                        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                    }
                    if (lane == 0) { s_red[warp] = best; s_idx[warp] = bi; }
                    asm volatile("bar.sync 3, %0;" ::"r"(n_pw * 32) : "memory");
                    double bb = s_red[0];
                    int pr = s_idx[0];
                    for (int u = 1; u < n_pw; ++u)
                        if (s_red[u] > bb || (s_red[u] == bb && s_idx[u] < pr)) { bb = s_red[u]; pr = s_idx[u]; }
                    if (tid == pr) {                               // the pivot row publishes itself
                        mine_used = true;
                        used[tid] = 1;
                        piv[q] = tid;
                        rowmap[tid] = k0 + q;
                        colmap[k0 + q] = tid;
                        if (!(bb > 0.0)) atomicOr(&P.status[w], 1);
#pragma unroll
                        for (int c = 0; c < kNB; ++c) Dm[c] = r[c];      // row buffer (D is rebuilt in step 3)
                    }
                    asm volatile("bar.sync 3, %0;" ::"r"(n_pw * 32) : "memory");
                    if (!mine_used) {
                        const double2 pv = Dm[q];
                        const double d = 1.0 / fma(pv.x, pv.x, pv.y * pv.y);
                        const double2 l = cmul_g(r[q], make_double2(pv.x * d, -pv.y * d));
#pragma unroll
                        for (int c = 0; c < kNB; ++c) {
                            if (c > q) {
                                const double2 rv = Dm[c];
                                r[c].x = fma(-l.x, rv.x, fma(l.y, rv.y, r[c].x));
                                r[c].y = fma(-l.x, rv.y, fma(-l.y, rv.x, r[c].y));
                            }
                        }
                    }
                    // the next step's publication of Dm happens after the next first barrier: everyone has read this one by then
                }
            }
            __syncthreads();
            // ---- 3. D = C0[R, :] and P = D^-1 (Gauss-Jordan in pivot order, one thread per entry of the 16 x 16 block)
            if (tid < kNB * kNB) {
                const int t = tid / kNB, sidx = tid % kNB;
                Dm[tid] = C0[piv[t] * kNB + sidx];
                Pm[tid] = make_double2(t == sidx ? 1.0 : 0.0, 0.0);
            }
            __syncthreads();
            for (int sidx = 0; sidx < kNB; ++sidx) {
                double2 dnew = make_double2(0.0, 0.0), pnew = make_double2(0.0, 0.0);
                if (tid < kNB * kNB) {
                    const int t = tid / kNB, c = tid % kNB;
                    const double2 pvt = Dm[sidx * kNB + sidx];
                    const double dd = 1.0 / fma(pvt.x, pvt.x, pvt.y * pvt.y);
                    const double2 iv = make_double2(pvt.x * dd, -pvt.y * dd);
                    const double2 rd = cmul_g(Dm[sidx * kNB + c], iv), rp = cmul_g(Pm[sidx * kNB + c], iv);      // scaled pivot row
                    if (t == sidx) {
                        dnew = rd;
                        pnew = rp;
                    } else {
                        const double2 l = Dm[t * kNB + sidx];
                        const double2 d0 = Dm[t * kNB + c], p0 = Pm[t * kNB + c];
                        dnew = make_double2(fma(-l.x, rd.x, fma(l.y, rd.y, d0.x)), fma(-l.x, rd.y, fma(-l.y, rd.x, d0.y)));
                        pnew = make_double2(fma(-l.x, rp.x, fma(l.y, rp.y, p0.x)), fma(-l.x, rp.y, fma(-l.y, rp.x, p0.y)));
                    }
                }
                __syncthreads();
                if (tid < kNB * kNB) { Dm[tid] = dnew; Pm[tid] = pnew; }
                __syncthreads();
            }
            // ---- 3b. L = C0 P (rows outside R), L[r_t, :] = e_t - P[t, :]
            for (int e = tid; e < Mp * kNB; e += kBlkThreads) {
                const int i = e / kNB, q = e - i * kNB;
                double2 acc = make_double2(0.0, 0.0);
                const int rm = rowmap[i] - k0;               // 0..15 when row i is one of this block's pivot rows
                if (used[i] && rm >= 0 && rm < kNB && piv[rm] == i) {
                    const double2 pq = Pm[rm * kNB + q];
                    acc = make_double2((rm == q ? 1.0 : 0.0) - pq.x, -pq.y);
                } else {
#pragma unroll
                    for (int sidx = 0; sidx < kNB; ++sidx) {
                        const double2 c = C0[i * kNB + sidx], pq = Pm[sidx * kNB + q];
                        acc.x = fma(c.x, pq.x, fma(-c.y, pq.y, acc.x));
                        acc.y = fma(c.x, pq.y, fma(c.y, pq.x, acc.y));
                    }
                }
                Lre[i * kLLd + q] = acc.x;
                Lim[i * kLLd + q] = acc.y;
                // ---- 4a. identity pattern into the panel columns of S
                a[(size_t)i * Mp + k0 + q] = make_double2((used[i] && rm == q && piv[q] == i) ? 1.0 : 0.0, 0.0);
            }
            __syncthreads();
            // ---- 4b. U = S[R, :]
            for (int e = tid; e < kNB * Mp; e += kBlkThreads) {
                const int t = e / Mp, j = e - t * Mp;
                const double2 v = a[(size_t)piv[t] * Mp + j];
                Ure[t * ldu + j] = v.x;
                Uim[t * ldu + j] = v.y;
            }
            __syncthreads();
            // ---- 5. S <- S - L U on the tensor pipe: warp = tile row(s), 16 DMMAs per 8 x 8 complex tile
            const int n_t = Mp >> 3;
            for (int ta = warp; ta < n_t; ta += kBlkThreads / 32) {
                double nlr[4], pli[4], nli[4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const double lr = Lre[(8 * ta + g4) * kLLd + 4 * kk + t4], li = Lim[(8 * ta + g4) * kLLd + 4 * kk + t4];
                    nlr[kk] = -lr;
                    pli[kk] = li;
                    nli[kk] = -li;
                }
                double2* rowp = a + (size_t)(8 * ta + g4) * Mp + 2 * t4;
                double2 c0 = rowp[0], c1 = rowp[1];
                for (int tb = 0; tb < n_t; ++tb) {
                    double2 n0 = c0, n1 = c1;
                    if (tb + 1 < n_t) { n0 = rowp[8 * (tb + 1)]; n1 = rowp[8 * (tb + 1) + 1]; }      // next tile's entries in flight
                    double cr0 = c0.x, cr1 = c1.x, ci0 = c0.y, ci1 = c1.y;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const double ur = Ure[(4 * kk + t4) * ldu + 8 * tb + g4], ui = Uim[(4 * kk + t4) * ldu + 8 * tb + g4];
                        dmma884_g(cr0, cr1, nlr[kk], ur);
                        dmma884_g(cr0, cr1, pli[kk], ui);
                        dmma884_g(ci0, ci1, nlr[kk], ui);
                        dmma884_g(ci0, ci1, nli[kk], ur);
                    }
                    rowp[8 * tb] = make_double2(cr0, ci0);
                    rowp[8 * tb + 1] = make_double2(cr1, ci1);
                    c0 = n0;
                    c1 = n1;
                }
            }
            __syncthreads();
        }
        // ---- outputs: inverse[rowmap[i]][colmap[j]] = a[i][j]
        for (int e = tid; e < Mp * Mp; e += kBlkThreads) {
            const int i = e / Mp, j = e - i * Mp;
            const int oi = rowmap[i], oj = colmap[j];
            if (oi < m && oj < m) {
                const double2 v = a[e];
                const size_t o = ((size_t)w * m * m + (size_t)oi * m + oj) * F + f;
                if (P.dtf) P.dtf[o] = fma(v.x, v.x, v.y * v.y);
                if (P.H) P.H[o] = v;
            }
        }
        if (P.rowpart) {
            for (int i = tid; i < Mp; i += kBlkThreads) {
                const int oi = rowmap[i];
                if (oi < m) {
                    double acc = 0.0;
                    for (int j = 0; j < Mp; ++j) {
                        if (colmap[j] < m) {
                            const double2 v = a[(size_t)i * Mp + j];
                            acc += fma(v.x, v.x, v.y * v.y);
                        }
                    }
                    P.rowpart[((size_t)w * F + f) * m + oi] = acc;
                }
            }
        }
    }
}

int transfer_generic_grid() { return device_sm_count() * 2; }
size_t transfer_generic_scratch_bytes(int m) {
    const size_t Mp = BlkSmem::mp(m);
    return (size_t)transfer_generic_grid() * Mp * Mp * sizeof(double2);       // covers both kernels (Mp >= m)
}

int launch_transfer_generic(const K5Params& P, void* scratch, cudaStream_t st) {
    const size_t smem_blk = BlkSmem::bytes(P.m);
    if (smem_blk <= 220 * 1024 && exp_env_int("HS_K5_GENERIC_UNBLOCKED", 0) == 0) {
        cudaError_t e = cudaFuncSetAttribute(transfer_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_blk);
        if (e != cudaSuccess) return set_error(HS_ERR_CUDA, "transfer_blocked: %s", cudaGetErrorString(e));
        long long n_mat = (long long)P.n_win * P.F;
        int grid = device_sm_count();
        if (n_mat < grid) grid = (int)n_mat;
        transfer_blocked_kernel<<<grid, kBlkThreads, smem_blk, st>>>(P, reinterpret_cast<double2*>(scratch));
        return check_launch("transfer_blocked_kernel");
    }
    const size_t smem = (size_t)2 * P.m * sizeof(double2) + 3 * P.m * sizeof(int) + 16;
    long long n_mat = (long long)P.n_win * P.F;
    int grid = transfer_generic_grid();
    if (n_mat < grid) grid = (int)n_mat;
    transfer_generic_kernel<<<grid, 256, smem, st>>>(P, reinterpret_cast<double2*>(scratch));
    return check_launch("transfer_generic_kernel");
}

}  // namespace hs
