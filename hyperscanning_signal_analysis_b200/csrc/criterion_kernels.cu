// Model-order criteria for batches of windows (reference mvar_criterion, src/mtmvar.py:551-601).
//
// The reference refits the model for every order 1..P and takes  ln det V_k + penalty(k)  (:577-588), argmin (:590).
// K4 (the LWR recursion run once to order P) already delivers V_k of EVERY order (d_Vall); what is left per
// (window, order) is ln det of an m x m symmetric positive definite matrix and, per window, the first minimum.
//   criterion_kernel : one CTA per (window, order): V_k staged in shared memory (row stride m + 1), right-looking LDL^T
//                      without pivoting (V_k is SPD), ln det = sum ln d_i  -- the value np.log(np.linalg.det(V)) takes
//                      whenever det is representable, and still finite when det under/overflows;
//   argmin_kernel    : one thread per window, first index of the minimum like np.argmin (NaN wins like in NumPy).
#include <cmath>

#include "hs_internal.h"

namespace hs {

constexpr int kCritThreads = 128;

__global__ void __launch_bounds__(kCritThreads) criterion_kernel(const double* __restrict__ Vall, const int P, const int m, const double pen_unit,
                                                                 double* __restrict__ crit, double* __restrict__ logdet) {
    extern __shared__ double a[];                // [m][m + 1]
    __shared__ double dsum;
    const int ld = m + 1;
    const int w = blockIdx.x / P, k = blockIdx.x - w * P;
    const double* V = Vall + (size_t)blockIdx.x * m * m;
    for (int e = threadIdx.x; e < m * m; e += kCritThreads) {
        const int i = e / m, j = e - i * m;
        a[i * ld + j] = V[e];
    }
    if (threadIdx.x == 0) dsum = 0.0;
    __syncthreads();
    for (int c = 0; c < m; ++c) {
        const double d = a[c * ld + c];
        if (threadIdx.x == 0) dsum += log(d);      // d <= 0 (not positive definite): NaN / -inf, as log(det) would give
        const double rd = 1.0 / d;
        const int rem = m - 1 - c;
        // trailing lower triangle (i >= j > c):  a[i][j] -= a[i][c] a[j][c] / d
        for (int e = threadIdx.x; e < rem * rem; e += kCritThreads) {
            const int i = c + 1 + e / rem, j = c + 1 + e % rem;
            if (j <= i) a[i * ld + j] = fma(-a[i * ld + c] * rd, a[j * ld + c], a[i * ld + j]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (logdet) logdet[blockIdx.x] = dsum;
        crit[blockIdx.x] = dsum + pen_unit * (double)(k + 1);
    }
}

__global__ void argmin_kernel(const double* __restrict__ crit, const int n_win, const int P, int* __restrict__ popt) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_win) return;
    const double* c = crit + (size_t)w * P;
    int best = 0;
    double bv = c[0];
    for (int k = 1; k < P; ++k) {
        const double v = c[k];
        if (bv != bv) break;                       // np.argmin returns the first NaN
        if (v < bv || v != v) { bv = v; best = k; }
    }
    popt[w] = best + 1;                            // model_order_range starts at 1 (mtmvar.py:573)
}

// Generalised partial directed coherence (reference gen_partial_directed_coherence, src/mtmvar.py:449-468):
//   gpdc[i][j][f] = (|A_ij(f)| / sigma_i) / sqrt(sum_k |A_kj(f)|^2 / sigma_k^2),  sigma_k^2 = V_kk;  0 where the denominator is 0.
// One thread per (window, target j, bin f): bins are the fastest index, so the column walks are coalesced.
__global__ void __launch_bounds__(256) gpdc_kernel(const double2* __restrict__ Af, const double* __restrict__ V, const int m, const int F,
                                                   double* __restrict__ out) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y, w = blockIdx.z;
    if (f >= F) return;
    const size_t base = (size_t)w * m * m * F;
    const double* Vw = V + (size_t)w * m * m;
    double den = 0.0;
    for (int k = 0; k < m; ++k) {
        const double2 a = Af[base + ((size_t)k * m + j) * F + f];
        const double mag = hypot(a.x, a.y);                   // np.abs
        den += (mag * mag) / Vw[(size_t)k * m + k];
    }
    den = sqrt(den);
    for (int i = 0; i < m; ++i) {
        const double2 a = Af[base + ((size_t)i * m + j) * F + f];
        const double num = hypot(a.x, a.y) / sqrt(Vw[(size_t)i * m + i]);
        out[base + ((size_t)i * m + j) * F + f] = (den != 0.0) ? num / den : 0.0;
    }
}

}  // namespace hs

using namespace hs;

extern "C" int hs_gpdc_f64(const void* d_Af, const double* d_V, int n_win, int m, int F, double* d_gpdc, void* stream) {
    if (n_win <= 0) return HS_OK;
    if (!d_Af || !d_V || !d_gpdc) return set_error(HS_ERR_INVALID, "hs_gpdc_f64: null pointer");
    if (m < 1 || F < 1 || n_win > 65535 || m > 65535) return set_error(HS_ERR_INVALID, "hs_gpdc_f64: bad sizes");
    dim3 grid((F + 255) / 256, m, n_win);
    gpdc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const double2*>(d_Af), d_V, m, F, d_gpdc);
    return check_launch("gpdc_kernel");
}

extern "C" int hs_mvar_criterion_f64(const double* d_Vall, int n_win, int P, int m, int n_samples, int crit_type, double* d_crit,
                                     double* d_logdet, int32_t* d_popt, void* stream) {
    if (n_win <= 0) return HS_OK;
    if (!d_Vall || !d_crit || !d_popt) return set_error(HS_ERR_INVALID, "hs_mvar_criterion_f64: null pointer");
    if (P < 1 || m < 1 || n_samples < 2) return set_error(HS_ERR_INVALID, "hs_mvar_criterion_f64: bad sizes");
    // penalty per unit of model order, the reference's expressions (mtmvar.py:579-586)
    const double n = (double)n_samples, m2 = (double)m * (double)m;
    double pen;
    if (crit_type == 0) pen = 2.0 * m2 / n;                              // AIC
    else if (crit_type == 1) pen = 2.0 * log(log(n)) * m2 / n;           // HQ
    else if (crit_type == 2) pen = log(n) * m2 / n;                      // SC
    else return set_error(HS_ERR_INVALID, "Invalid criterion type. Choose from 'AIC', 'HQ', 'SC'.");
    const size_t smem = (size_t)m * (m + 1) * sizeof(double);
    if (smem > 200 * 1024) return set_error(HS_ERR_UNSUPPORTED, "hs_mvar_criterion_f64: m=%d does not fit shared memory", m);
    cudaStream_t st = (cudaStream_t)stream;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(criterion_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(HS_ERR_CUDA, "hs_mvar_criterion_f64: %s", cudaGetErrorString(e));
    }
    criterion_kernel<<<n_win * P, kCritThreads, smem, st>>>(d_Vall, P, m, pen, d_crit, d_logdet);
    int rc = check_launch("criterion_kernel");
    if (rc) return rc;
    argmin_kernel<<<(n_win + 127) / 128, 128, 0, st>>>(d_crit, n_win, P, d_popt);
    return check_launch("argmin_kernel");
}
