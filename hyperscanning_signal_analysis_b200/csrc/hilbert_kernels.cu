// Hilbert envelope |analytic signal| of real signals, any FFT length N whose prime factors are <= 31
// (scipy.fft.next_fast_len returns 11-smooth lengths).
//
// Replaces  np.abs(scipy.signal.hilbert(x, N=fast_len)[:orig_len])  of EEG_IBI_FFDTF_Pipeline._compute_asymmetry
// (/root/reference src/eeg_alpha_ibi_ffdtf.py:352-356).  scipy.signal.hilbert:  X = fft(x, N) (zero padded),
// X[k] *= h[k]  with  h[0] = 1, h[1..N/2-1] = 2, h[N/2] = 1 (N even), h[k] = 2 for k <= (N-1)/2 (N odd), 0 elsewhere,
// analytic = ifft(X).  |ifft(Y)| = |fft(conj(Y))| / N, so both transforms are the same forward FFT.
//
// The FFT is a hand-written mixed-radix Stockham autosort transform in global memory (the data of one task is a
// few hundred KB: L2 resident), one launch per radix pass:
//      v_r   = in[j + r N/R] * exp(-2 pi i r (j mod Ns) / (Ns R)),           j < N/R,  r < R
//      out[(j / Ns) Ns R + (j mod Ns) + q Ns] = sum_r v_r exp(-2 pi i q r / R),       q < R
// with Ns the product of the radices already applied.  No bit reversal, natural order in and out.
#include <cmath>
#include <vector>

#include "fft_internal.h"
#include "hs_internal.h"

namespace hs {

namespace {

constexpr int kFftThreads = 128;
constexpr int kMaxRadix = 31;

__device__ __forceinline__ double2 cmulf(const double2 a, const double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}

__global__ void hilbert_load_kernel(const double* __restrict__ x, const long long n, const long long sig_stride, const long long N,
                                    double2* __restrict__ z) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (t < N) z[(size_t)s * N + t] = make_double2(t < n ? x[(size_t)s * sig_stride + t] : 0.0, 0.0);
}

template <int R>
__global__ void __launch_bounds__(kFftThreads) fft_pass_kernel(const double2* __restrict__ in, double2* __restrict__ out, const long long N,
                                                               const long long Ns, const int r_rt) {
    const int RR = R > 0 ? R : r_rt;
    __shared__ double2 wr[kMaxRadix];          // exp(-2 pi i k / R)
    if (threadIdx.x < RR) {
        double sn, cs;
        sincospi(-2.0 * (double)threadIdx.x / (double)RR, &sn, &cs);
        wr[threadIdx.x] = make_double2(cs, sn);
    }
    __syncthreads();
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nb = N / RR;
    if (j >= nb) return;
    const size_t base = (size_t)blockIdx.y * N;
    const long long k = j % Ns;
    const double ang = -2.0 * (double)k / ((double)Ns * (double)RR);
    double2 v[R > 0 ? R : kMaxRadix];
#pragma unroll
    for (int r = 0; r < (R > 0 ? R : kMaxRadix); ++r) {
        if (r < RR) {
            double sn, cs;
            sincospi(ang * (double)r, &sn, &cs);
            v[r] = cmulf(in[base + j + (long long)r * nb], make_double2(cs, sn));
        }
    }
    const size_t obase = base + (size_t)(j / Ns) * Ns * RR + k;
#pragma unroll
    for (int q = 0; q < (R > 0 ? R : kMaxRadix); ++q) {
        if (q < RR) {
            double2 acc = v[0];
            int idx = 0;
#pragma unroll
            for (int r = 1; r < (R > 0 ? R : kMaxRadix); ++r) {
                if (r < RR) {
                    idx += q;
                    if (idx >= RR) idx -= RR;
                    const double2 w = wr[idx];
                    acc.x = fma(v[r].x, w.x, fma(-v[r].y, w.y, acc.x));
                    acc.y = fma(v[r].x, w.y, fma(v[r].y, w.x, acc.y));
                }
            }
            out[obase + (size_t)q * Ns] = acc;
        }
    }
}

// X[k] <- conj(h[k] X[k])
__global__ void hilbert_mask_kernel(double2* __restrict__ z, const long long N) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    double h;
    if ((N & 1) == 0) h = (k == 0 || k == N / 2) ? 1.0 : (k < N / 2 ? 2.0 : 0.0);
    else h = (k == 0) ? 1.0 : (k <= (N - 1) / 2 ? 2.0 : 0.0);
    double2* p = z + (size_t)blockIdx.y * N + k;
    const double2 v = *p;
    *p = make_double2(h * v.x, -h * v.y);
}

__global__ void hilbert_abs_kernel(const double2* __restrict__ z, const long long N, const long long n, double* __restrict__ env,
                                   const long long env_stride, double2* __restrict__ analytic) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double2 v = z[(size_t)blockIdx.y * N + t];
    const double s = 1.0 / (double)N;
    const double re = v.x * s, im = -v.y * s;          // ifft(Y) = conj(fft(conj(Y))) / N
    if (env) env[(size_t)blockIdx.y * env_stride + t] = hypot(re, im);
    if (analytic) analytic[(size_t)blockIdx.y * n + t] = make_double2(re, im);
}

int fft_passes(double2*& a, double2*& b, int n_sig, long long N, const std::vector<int>& radices, cudaStream_t st) {
    long long Ns = 1;
    for (int R : radices) {
        const long long nb = N / R;
        dim3 grid((unsigned)((nb + kFftThreads - 1) / kFftThreads), n_sig);
        switch (R) {
            case 2: fft_pass_kernel<2><<<grid, kFftThreads, 0, st>>>(a, b, N, Ns, R); break;
            case 3: fft_pass_kernel<3><<<grid, kFftThreads, 0, st>>>(a, b, N, Ns, R); break;
            case 4: fft_pass_kernel<4><<<grid, kFftThreads, 0, st>>>(a, b, N, Ns, R); break;
            case 5: fft_pass_kernel<5><<<grid, kFftThreads, 0, st>>>(a, b, N, Ns, R); break;
            case 7: fft_pass_kernel<7><<<grid, kFftThreads, 0, st>>>(a, b, N, Ns, R); break;
            case 11: fft_pass_kernel<11><<<grid, kFftThreads, 0, st>>>(a, b, N, Ns, R); break;
            default: fft_pass_kernel<0><<<grid, kFftThreads, 0, st>>>(a, b, N, Ns, R); break;
        }
        int rc = check_launch("fft_pass_kernel");
        if (rc) return rc;
        Ns *= R;
        double2* t = a; a = b; b = t;
    }
    return HS_OK;
}

bool factorize(long long N, std::vector<int>& radices) {
    radices.clear();
    while (N % 4 == 0) { radices.push_back(4); N /= 4; }
    for (int p = 2; p <= kMaxRadix; ++p)
        while (N % p == 0) { radices.push_back(p); N /= p; }
    return N == 1;
}

}  // namespace

bool fft_smooth(long long N) {
    std::vector<int> r;
    return N >= 1 && factorize(N, r);
}

int fft_forward_batched(double2*& a, double2*& b, int batch, long long N, cudaStream_t st) {
    std::vector<int> radices;
    if (!factorize(N, radices)) return set_error(HS_ERR_UNSUPPORTED, "fft: N = %lld has a prime factor > %d", N, kMaxRadix);
    // the batch index is the grid's y dimension (<= 65535 rows per launch)
    for (int b0 = 0; b0 < batch; b0 += 65535) {
        const int nb = batch - b0 < 65535 ? batch - b0 : 65535;
        double2* aa = a + (size_t)b0 * N;
        double2* bb = b + (size_t)b0 * N;
        int rc = fft_passes(aa, bb, nb, N, radices, st);
        if (rc) return rc;
    }
    if (radices.size() & 1) { double2* t = a; a = b; b = t; }
    return HS_OK;
}

}  // namespace hs

using namespace hs;

extern "C" {

size_t hs_hilbert_ws_bytes(int n_sig, int64_t N) {
    if (n_sig <= 0 || N <= 0) return 256;
    return 2 * (((size_t)n_sig * (size_t)N * 16 + 255) / 256 * 256) + 256;
}

int hs_hilbert_f64(const double* d_x, int n_sig, int64_t n, int64_t sig_stride, int64_t N, double* d_env, int64_t env_stride,
                   void* d_analytic, void* d_ws, void* stream) {
    if (!d_x || !d_ws || (!d_env && !d_analytic)) return set_error(HS_ERR_INVALID, "hs_hilbert_f64: null pointer");
    if (n < 1 || N < 1) return set_error(HS_ERR_INVALID, "hs_hilbert_f64: N must be positive.");
    if (n_sig <= 0) return HS_OK;
    std::vector<int> radices;
    if (!factorize(N, radices)) return set_error(HS_ERR_UNSUPPORTED, "hs_hilbert_f64: N = %lld has a prime factor > %d", (long long)N, kMaxRadix);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t half = ((size_t)n_sig * (size_t)N * 16 + 255) / 256 * 256;
    double2* a = reinterpret_cast<double2*>(d_ws);
    double2* b = reinterpret_cast<double2*>(reinterpret_cast<unsigned char*>(d_ws) + half);
    const long long nv = n < N ? n : N;          // scipy: fft(x, N) truncates when N < len(x)
    dim3 gN((unsigned)((N + 255) / 256), n_sig);
    hilbert_load_kernel<<<gN, 256, 0, st>>>(d_x, nv, sig_stride, N, a);
    int rc = check_launch("hilbert_load_kernel");
    if (rc) return rc;
    rc = fft_passes(a, b, n_sig, N, radices, st);
    if (rc) return rc;
    hilbert_mask_kernel<<<gN, 256, 0, st>>>(a, N);
    rc = check_launch("hilbert_mask_kernel");
    if (rc) return rc;
    rc = fft_passes(a, b, n_sig, N, radices, st);
    if (rc) return rc;
    dim3 gn((unsigned)((nv + 255) / 256), n_sig);
    hilbert_abs_kernel<<<gn, 256, 0, st>>>(a, N, nv, d_env, env_stride, reinterpret_cast<double2*>(d_analytic));
    return check_launch("hilbert_abs_kernel");
}

}
