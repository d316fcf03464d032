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

static int bgemm(const GemmArgs& g, int batch, cudaStream_t st) {
    dim3 grid(((g.M + kPadMax - 1) / kPadMax) * ((g.N + kPadMax - 1) / kPadMax), batch);
    bgemm_kernel<<<grid, 64, 0, st>>>(g);
    return check_launch("bgemm_kernel");
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
        binv_kernel<<<nw, 256, inv_smem, st>>>(Vb, Xb, m, P.status, HS_STATUS_SINGULAR_YW);
        if ((rc = check_launch("binv_kernel"))) return rc;
        if (!last) {
            binv_kernel<<<nw, 256, inv_smem, st>>>(Vf, Xf, m, P.status, HS_STATUS_SINGULAR_YW);
            if ((rc = check_launch("binv_kernel"))) return rc;
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

int transfer_generic_grid() { return device_sm_count() * 2; }
size_t transfer_generic_scratch_bytes(int m) { return (size_t)transfer_generic_grid() * m * m * sizeof(double2); }

int launch_transfer_generic(const K5Params& P, void* scratch, cudaStream_t st) {
    const size_t smem = (size_t)2 * P.m * sizeof(double2) + 3 * P.m * sizeof(int) + 16;
    long long n_mat = (long long)P.n_win * P.F;
    int grid = transfer_generic_grid();
    if (n_mat < grid) grid = (int)n_mat;
    transfer_generic_kernel<<<grid, 256, smem, st>>>(P, reinterpret_cast<double2*>(scratch));
    return check_launch("transfer_generic_kernel");
}

}  // namespace hs
