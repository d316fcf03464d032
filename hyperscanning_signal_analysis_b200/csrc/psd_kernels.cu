// K6: multitaper PSD on sm_100a, FP64 -- three paths behind hs_mt_psd_f64 (include/hs_b200.h):
//   (1) mt_psd_r16_kernel   n = 4096 / 8192: one CTA per (signal, taper-pair group), 16 points per thread in REGISTERS,
//                           radix-16 stages with one shared-memory exchange between stages (3 exchanges + read-out per
//                           transform instead of 7 in-place passes), tapers multiplied into the first stage's loads;
//   (2) mt_psd_kernel       n = n1 * 2^a whose buffer fits shared memory and n1 <= 64: direct DFT over the odd factor x
//                           radix-4 DIF passes in shared memory (round 1's kernel);
//   (3) general             ANY other length: batched global-memory Stockham FFT (the mixed-radix passes of
//                           hilbert_kernels.cu) when every prime factor is <= 31, otherwise Bluestein's chirp-z through a
//                           power-of-two transform of length >= 2n - 1.  Real movie segments have arbitrary lengths
//                           (reference src/io_utils.py:131-151), and MNE accepts any n_times (src/psd.py:30-32).
// All paths pack two tapers into one complex transform and use the same normalisation.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fft_internal.h"
#include "hs_internal.h"

namespace hs {

// =====================================================================================
// K6  multitaper PSD (mne.time_frequency.psd_array_multitaper defaults reached through
//     compute_psd_multitaper, src/psd.py:30-32; restated in SURVEY.md A.6):
//       x0 = x - mean(x);  X_k = rfft(x0 * taper_k);  DC (and Nyquist for even n) / sqrt(2);
//       psd = sum_k |w_k X_k|^2 * 2 / sum_k w_k^2,   w_k = sqrt(eigval_k).
//     No cuFFT: a hand-written transform in shared memory.  n = N1 * N2 with N2 the power-of-two part:
//       1. direct DFT of length N1 over the stride-N2 subsequences (reads x0*taper straight from global),
//       2. twiddle W_n^(t2 k1),
//       3. N1 in-place radix-2 DIF FFTs of length N2 (output left in bit-reversed order and read that way).
//     Two tapers are packed into one complex transform (z = x0 h_a + i x0 h_b) and separated in the power sum:
//       |X_a[k]|^2 = |Z[k] + conj Z[n-k]|^2 / 4,   |X_b[k]|^2 = |Z[k] - conj Z[n-k]|^2 / 4.
//     The (signals x tapers x n) complex intermediate never leaves the SM.
// =====================================================================================
constexpr size_t kPsdSmemMax = 200 * 1024;
constexpr int kPsdThreads = 512;

// Taper-pair groups per signal: CTAs = n_sig * groups run one per SM (the FFT buffer fills the shared memory), so the group
// count is chosen to fill whole waves of 148 CTAs (342 segments: 1 group = 2.3 waves -> 3 rounds, 3 groups = 6.93 -> 7 rounds).
static int psd_groups(int n_sig, int pairs) {
    const int sm = 148;            // sizing only (also used by the GPU-less workspace query); any SM count gives a valid split
    int best = 1;
    double best_eff = 0.0;
    for (int g = 1; g <= 8 && g <= pairs; ++g) {
        const double waves = (double)n_sig * g / sm;
        const double rounds = (double)((long long)((n_sig * (long long)g + sm - 1) / sm));
        const double eff = waves / rounds - 0.004 * g;      // small penalty: every CTA recomputes the mean and reloads the twiddles
        if (eff > best_eff) { best_eff = eff; best = g; }
    }
    return best;
}

__global__ void twiddle_kernel(double2* w, long long n) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double s, c;
    sincospi(-2.0 * (double)j / (double)n, &s, &c);      // exp(-2 pi i j / n)
    w[j] = make_double2(c, s);
}

__device__ __forceinline__ unsigned bitrev_bits(unsigned v, int bits) { return bits ? (__brev(v) >> (32 - bits)) : 0u; }
__device__ __forceinline__ double2 cmul(const double2 a, const double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}

struct PsdParams {
    const double* x;         // (n_sig, n)
    const double* tapers;    // (K, n)
    const double* weights;   // (K)
    const double2* tw;       // (n) W_n^j
    double2* gbuf;           // global FFT buffers (n_sig*groups, n) when n*16 B does not fit shared memory, else null
    double* partial;         // (n_sig, groups, nb)
    long long n;
    int n1, n2, log2n2;      // n = n1 * n2, n2 = 2^log2n2
    int K, groups, k_lo, nb, n_sig;
    int tw_in_smem;          // 1: the half twiddle table sits behind the FFT buffer in shared memory
    int skew;                // 1: FFT buffer in shared memory with the i + (i >> 3) layout
};

__global__ void __launch_bounds__(kPsdThreads, 1) mt_psd_kernel(const PsdParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int sig = blockIdx.x / P.groups, grp = blockIdx.x % P.groups;
    const long long n = P.n;
    const int n1 = P.n1, n2 = P.n2;
    double* acc = reinterpret_cast<double*>(smem_raw);                       // [nb]
    double2* buf = P.gbuf ? (P.gbuf + (size_t)blockIdx.x * n)
                          : reinterpret_cast<double2*>(smem_raw + (((size_t)P.nb * sizeof(double) + 15) / 16) * 16);
    // Shared-memory layout of the FFT buffer: element i sits at i + (i >> 3) (one 16-byte pad per 128-byte row), so the strided
    // accesses of the late stages (stride 4, 16, ... elements) and the bit-reversed read-out spread over the banks
    // (61 % of the wavefronts were bank-conflict replays without it).  No skew for the global-memory fallback.
    const int skew = P.skew;
    auto sk = [skew](const long long i) -> long long { return skew ? i + (i >> 3) : i; };
    double2* tws = P.tw_in_smem ? buf + (n + (n >> 3) + 1) : nullptr;        // W_n^j, j < n/2 (only with the shared-memory FFT buffer)
    if (tws)
        for (int e = threadIdx.x; e < (int)(n >> 1); e += kPsdThreads) tws[e] = P.tw[e];
    __shared__ double red[kPsdThreads];
    const double* xs = P.x + (size_t)sig * n;
    const int tid = threadIdx.x;

    // mean of the signal (fixed-order tree)
    double a0 = 0.0;
    for (long long t = tid; t < n; t += kPsdThreads) a0 += xs[t];
    red[tid] = a0;
    __syncthreads();
    for (int off = kPsdThreads / 2; off > 0; off >>= 1) {
        if (tid < off) red[tid] += red[tid + off];
        __syncthreads();
    }
    const double mean = red[0] / (double)n;
    for (int b = tid; b < P.nb; b += kPsdThreads) acc[b] = 0.0;

    const int pairs = (P.K + 1) / 2;
    for (int pr = grp; pr < pairs; pr += P.groups) {
        const int ka = 2 * pr, kb = 2 * pr + 1;
        const double* ha = P.tapers + (size_t)ka * n;
        const double* hb = (kb < P.K) ? P.tapers + (size_t)kb * n : nullptr;
        const double wa = P.weights[ka], wb = hb ? P.weights[kb] : 0.0;
        __syncthreads();
        // ---- step 1+2: Y[k1][t2] = W_n^(t2 k1) * sum_t1 z[n2 t1 + t2] W_n1^(t1 k1)
        for (long long e = tid; e < n; e += kPsdThreads) {
            const int k1 = (int)(e / n2), t2 = (int)(e - (long long)k1 * n2);
            double2 y;
            if (n1 == 1) {
                const double v = xs[t2] - mean;
                y = make_double2(v * ha[t2], hb ? v * hb[t2] : 0.0);
            } else {
                double yr = 0.0, yi = 0.0;
                long long tw_idx = 0;                               // (n2 * t1 * k1) mod n
                const long long tw_step = ((long long)n2 * k1) % n;
                for (int t1 = 0; t1 < n1; ++t1) {
                    const long long t = (long long)n2 * t1 + t2;
                    const double v = xs[t] - mean;
                    const double2 z = make_double2(v * ha[t], hb ? v * hb[t] : 0.0);
                    const double2 wv = P.tw[tw_idx];
                    yr += fma(z.x, wv.x, -z.y * wv.y);
                    yi += fma(z.x, wv.y, z.y * wv.x);
                    tw_idx += tw_step;
                    if (tw_idx >= n) tw_idx -= n;
                }
                y = cmul(make_double2(yr, yi), P.tw[((long long)t2 * k1) % n]);
            }
            buf[sk(e)] = y;
        }
        __syncthreads();
        // ---- step 3: n1 independent in-place DIF FFTs of length n2 (32-bit index arithmetic, shifts only).  Two radix-2 stages
        //      are fused into one radix-4 pass (same data placement as the two stages, so the bit-reversed read-out below is
        //      unchanged): half the passes through shared memory and half the barriers; a single radix-2 stage is left over
        //      when log2(n2) is odd.  Twiddles come from the shared-memory half table when it fits (tws), else from L2.
        {
            const int ni = (int)n;
            auto twid = [&](const int idx) -> double2 {
                if (tws) {
                    const int h = ni >> 1;
                    const double2 v = tws[idx >= h ? idx - h : idx];
                    return idx >= h ? make_double2(-v.x, -v.y) : v;             // W^(k + n/2) = -W^k
                }
                return P.tw[idx];
            };
            int s = P.log2n2 - 1;
            for (; s >= 1; s -= 2) {
                const int q4 = 1 << (s - 1);                                    // quarter of the block length L = 2^(s+1)
                const int tw_mul = ni >> (s + 1);                               // W_L^j = W_n^(j * n / L)
                for (int r = tid; r < (ni >> 2); r += kPsdThreads) {
                    const int j = r & (q4 - 1);
                    const int e0 = ((r >> (s - 1)) << (s + 1)) + j;
                    double2* pa = buf + sk(e0);
                    double2* pb = buf + sk(e0 + q4);
                    double2* pc = buf + sk(e0 + 2 * q4);
                    double2* pd = buf + sk(e0 + 3 * q4);
                    const double2 a = *pa, b = *pb, c = *pc, d = *pd;
                    const double2 w1 = twid(j * tw_mul), w2 = twid(2 * j * tw_mul), w3 = twid(3 * j * tw_mul);
                    const double2 apc = make_double2(a.x + c.x, a.y + c.y), amc = make_double2(a.x - c.x, a.y - c.y);
                    const double2 bpd = make_double2(b.x + d.x, b.y + d.y), bmd = make_double2(b.x - d.x, b.y - d.y);
                    *pa = make_double2(apc.x + bpd.x, apc.y + bpd.y);
                    *pb = cmul(make_double2(apc.x - bpd.x, apc.y - bpd.y), w2);
                    *pc = cmul(make_double2(amc.x + bmd.y, amc.y - bmd.x), w1);            // (a - c) - i (b - d)
                    *pd = cmul(make_double2(amc.x - bmd.y, amc.y + bmd.x), w3);            // (a - c) + i (b - d)
                }
                __syncthreads();
            }
            if (s == 0) {                                                       // last radix-2 stage: blocks of 2, twiddle 1
                for (int r = tid; r < (ni >> 1); r += kPsdThreads) {
                    double2* p = buf + sk(2 * r);                               // 2 r and 2 r + 1 share a row of 8: adjacent
                    const double2 u = p[0], v = p[1];
                    p[0] = make_double2(u.x + v.x, u.y + v.y);
                    p[1] = make_double2(u.x - v.x, u.y - v.y);
                }
                __syncthreads();
            }
        }
        // ---- power of both tapers at the requested bins; X[k1 + n1 k2] sits at buf[k1 * n2 + bitrev(k2)]
        for (int b = tid; b < P.nb; b += kPsdThreads) {
            const long long k = P.k_lo + b, km = (n - k) % n;
            const double2 z = buf[sk((k % n1) * (long long)n2 + bitrev_bits((unsigned)(k / n1), P.log2n2))];
            const double2 zm = buf[sk((km % n1) * (long long)n2 + bitrev_bits((unsigned)(km / n1), P.log2n2))];
            const double sr = z.x + zm.x, si = z.y - zm.y;          // Z[k] + conj Z[n-k]  = 2 X_a[k]
            const double dr = z.x - zm.x, di = z.y + zm.y;          // Z[k] - conj Z[n-k]  = 2 i X_b[k]
            const double pa = 0.25 * fma(sr, sr, si * si), pb = 0.25 * fma(dr, dr, di * di);
            acc[b] += fma(wa * wa, pa, wb * wb * pb);
        }
    }
    __syncthreads();
    double* out = P.partial + ((size_t)sig * P.groups + grp) * P.nb;
    for (int b = tid; b < P.nb; b += kPsdThreads) out[b] = acc[b];
}

__global__ void mt_psd_finish_kernel(const double* partial, const double* weights, int K, int groups, int nb, int k_lo, long long n,
                                     double* psd) {
    const int sig = blockIdx.y;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    double wsum = 0.0;
    for (int k = 0; k < K; ++k) wsum = fma(weights[k], weights[k], wsum);
    double acc = 0.0;
    for (int g = 0; g < groups; ++g) acc += partial[((size_t)sig * groups + g) * nb + b];
    const long long k = k_lo + b;
    if (k == 0 || ((n & 1) == 0 && k == n / 2)) acc *= 0.5;          // amplitude / sqrt(2) at DC and Nyquist
    psd[(size_t)sig * nb + b] = acc * 2.0 / wsum;
}

static int launch_mt_psd_smem(const double* x, int n_sig, long long n, const double* tapers, const double* weights, int K, int k_lo,
                         int k_hi, double* psd, void* d_ws, cudaStream_t st) {
    PsdParams P;
    P.x = x;
    P.tapers = tapers;
    P.weights = weights;
    P.n = n;
    long long n2 = 1;
    int l2 = 0;
    while (((n / n2) & 1) == 0) { n2 <<= 1; ++l2; }
    if (n / n2 > 4096) return set_error(HS_ERR_UNSUPPORTED, "hs_mt_psd_f64: odd factor %lld of n=%lld is too large for the direct stage", n / n2, n);
    P.n1 = (int)(n / n2);
    P.n2 = (int)n2;
    P.log2n2 = l2;
    P.K = K;
    P.k_lo = k_lo;
    P.nb = k_hi - k_lo;
    P.n_sig = n_sig;
    const int pairs = (K + 1) / 2;
    P.groups = psd_groups(n_sig, pairs);
    unsigned char* ws = reinterpret_cast<unsigned char*>(d_ws);
    double2* tw = reinterpret_cast<double2*>(ws);
    ws += ((size_t)n * 16 + 255) / 256 * 256;
    P.tw = tw;
    P.partial = reinterpret_cast<double*>(ws);
    // NOTE: sized with (n/2+1) bins per partial spectrum in hs_mt_psd_ws_bytes
    ws += ((size_t)n_sig * P.groups * (n / 2 + 1) * sizeof(double) + 255) / 256 * 256;
    const size_t accb = (((size_t)P.nb * sizeof(double) + 15) / 16) * 16;
    const size_t buf_elems = (size_t)n + (size_t)(n >> 3) + 1;        // skewed layout: one pad element per row of 8
    const size_t smem_plain = accb + (size_t)n * 16, smem_skew = accb + buf_elems * 16;
    size_t smem;
    P.gbuf = nullptr;
    P.tw_in_smem = 0;
    P.skew = 0;
    if (smem_skew + (size_t)(n / 2) * 16 <= 220 * 1024 && (n & 1) == 0) {       // skewed buffer + half twiddle table
        smem = smem_skew + (size_t)(n / 2) * 16;
        P.tw_in_smem = 1;
        P.skew = 1;
    } else if (smem_skew <= kPsdSmemMax) {                                       // skewed buffer, twiddles from L2
        smem = smem_skew;
        P.skew = 1;
    } else if (smem_plain <= kPsdSmemMax) {                                      // plain buffer
        smem = smem_plain;
    } else {                                                                     // FFT buffer in global memory
        P.gbuf = reinterpret_cast<double2*>(ws);
        smem = accb;
        if (smem > 200 * 1024) return set_error(HS_ERR_UNSUPPORTED, "hs_mt_psd_f64: too many bins requested");
    }
    twiddle_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(tw, n);
    int rc = check_launch("twiddle_kernel");
    if (rc) return rc;
    cudaError_t e = cudaFuncSetAttribute(mt_psd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(HS_ERR_CUDA, "mt_psd: %s", cudaGetErrorString(e));
    mt_psd_kernel<<<n_sig * P.groups, kPsdThreads, smem, st>>>(P);
    rc = check_launch("mt_psd_kernel");
    if (rc) return rc;
    dim3 grid((P.nb + 255) / 256, n_sig);
    mt_psd_finish_kernel<<<grid, 256, 0, st>>>(P.partial, weights, K, P.groups, P.nb, k_lo, n, psd);
    return check_launch("mt_psd_finish_kernel");
}


// =====================================================================================
// Path (1): register radix-16 transform, N = 2^LOG2N in {4096, 8192}.
//   T = N / 16 threads.  With R3 = N / 256:
//   stage 1  thread t holds z[t + T j], j < 16:  16-point DFT over j, twiddle W_N^(t k1)        -> S1[k1][t]
//   stage 2  thread (k1, t2) holds S1[k1][t2 + R3 j2]:  16-point DFT, twiddle W_T^(t2 k2)        -> S2[k1*16 + k2][t2]
//   stage 3  R3 = 16: thread (k1, k2) transforms its 16 values;  R3 = 32: two threads per (k1, k2), h = 0 / 1 takes the
//            even / odd outputs of the length-32 transform (one radix-2 DIF step on the 32 values both threads read --
//            a shared-memory broadcast -- then a 16-point DFT)
//   X[k1 + 16 k2 + 256 q3] comes out in thread (k1, k2[, h]) register q3 (q3 = h + 2 q4 for R3 = 32) and is handed to the
//   bin owners through the same buffer.  Every exchange is conflict free: rows padded by one element where a warp walks
//   rows (S2: R3 + 1; read-out: tid + tid / 32).
// =====================================================================================
__device__ __forceinline__ double2 cadd(const double2 a, const double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(const double2 a, const double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 csqr(const double2 a) { return make_double2(fma(a.x, a.x, -a.y * a.y), 2.0 * a.x * a.y); }

// a * exp(-2 pi i j / 16); j is a compile-time constant after unrolling
__device__ __forceinline__ double2 mul_w16(const double2 a, const int j) {
    constexpr double C1 = 0.92387953251128673848, S1 = 0.38268343236508978178, R = 0.70710678118654752440;
    switch (j) {
        case 0: return a;
        case 1: return cmul(a, make_double2(C1, -S1));
        case 2: return make_double2(R * (a.x + a.y), R * (a.y - a.x));
        case 3: return cmul(a, make_double2(S1, -C1));
        case 4: return make_double2(a.y, -a.x);
        case 5: return cmul(a, make_double2(-S1, -C1));
        case 6: return make_double2(R * (a.y - a.x), -R * (a.x + a.y));
        case 7: return cmul(a, make_double2(-C1, -S1));
        default: return make_double2(-a.x, -a.y);      // j = 8 (never reached: the DIF stages use j <= 7)
    }
}

__device__ __forceinline__ constexpr int bitrev4(const int k) { return ((k & 1) << 3) | ((k & 2) << 1) | ((k & 4) >> 1) | ((k & 8) >> 3); }

// in-place 16-point forward DFT, radix-2 DIF: Y[k] ends up in v[bitrev4(k)]
__device__ __forceinline__ void fft16(double2 (&v)[16]) {
#pragma unroll
    for (int span = 8; span >= 1; span >>= 1) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if ((i & span) == 0) {
                const int j = i & (span - 1);
                const double2 a = v[i], b = v[i + span];
                v[i] = cadd(a, b);
                v[i + span] = mul_w16(csub(a, b), j * (8 / span));
            }
        }
    }
}

// v[bitrev4(k)] *= w^k for k = 1..15: powers by squaring / one product each (depth <= 4 products)
__device__ __forceinline__ void twiddle16(double2 (&v)[16], const double2 w1) {
    const double2 w2 = csqr(w1), w4 = csqr(w2), w8 = csqr(w4);
    v[bitrev4(1)] = cmul(v[bitrev4(1)], w1);
    v[bitrev4(2)] = cmul(v[bitrev4(2)], w2);
    v[bitrev4(4)] = cmul(v[bitrev4(4)], w4);
    v[bitrev4(8)] = cmul(v[bitrev4(8)], w8);
    {
        const double2 w3 = cmul(w2, w1);
        v[bitrev4(3)] = cmul(v[bitrev4(3)], w3);
        v[bitrev4(11)] = cmul(v[bitrev4(11)], cmul(w8, w3));
        const double2 w7 = cmul(w4, w3);
        v[bitrev4(7)] = cmul(v[bitrev4(7)], w7);
        v[bitrev4(15)] = cmul(v[bitrev4(15)], cmul(w8, w7));
    }
    {
        const double2 w5 = cmul(w4, w1);
        v[bitrev4(5)] = cmul(v[bitrev4(5)], w5);
        v[bitrev4(13)] = cmul(v[bitrev4(13)], cmul(w8, w5));
        const double2 w6 = cmul(w4, w2);
        v[bitrev4(6)] = cmul(v[bitrev4(6)], w6);
        v[bitrev4(14)] = cmul(v[bitrev4(14)], cmul(w8, w6));
    }
    v[bitrev4(9)] = cmul(v[bitrev4(9)], cmul(w8, w1));
    v[bitrev4(10)] = cmul(v[bitrev4(10)], cmul(w8, w2));
    v[bitrev4(12)] = cmul(v[bitrev4(12)], cmul(w8, w4));
}

// exp(-2 pi i t / 32), t = 0..15
__device__ __forceinline__ double2 w32(const int t) {
    constexpr double c[16] = {1.0, 0.98078528040323044913, 0.92387953251128673848, 0.83146961230254523708, 0.70710678118654752440,
                              0.55557023301960222474, 0.38268343236508978178, 0.19509032201612826785, 0.0, -0.19509032201612826785,
                              -0.38268343236508978178, -0.55557023301960222474, -0.70710678118654752440, -0.83146961230254523708,
                              -0.92387953251128673848, -0.98078528040323044913};
    constexpr double s[16] = {0.0, 0.19509032201612826785, 0.38268343236508978178, 0.55557023301960222474, 0.70710678118654752440,
                              0.83146961230254523708, 0.92387953251128673848, 0.98078528040323044913, 1.0, 0.98078528040323044913,
                              0.92387953251128673848, 0.83146961230254523708, 0.70710678118654752440, 0.55557023301960222474,
                              0.38268343236508978178, 0.19509032201612826785};
    return make_double2(c[t], -s[t]);
}

template <int LOG2N>
struct R16Cfg {
    static constexpr int N = 1 << LOG2N;
    static constexpr int T = N / 16;            // threads
    static constexpr int R3 = N / 256;          // 16 or 32
    static constexpr int LD2 = R3 + 1;          // row stride of S2
    static constexpr int LDO = T + T / 32;      // row stride of the read-out layout
    static constexpr int BUF = (256 * LD2 > 16 * LDO ? 256 * LD2 : 16 * LDO) > N ? (256 * LD2 > 16 * LDO ? 256 * LD2 : 16 * LDO) : N;
};

template <int LOG2N>
__global__ void __launch_bounds__(R16Cfg<LOG2N>::T, 1) mt_psd_r16_kernel(const PsdParams P) {
    using C = R16Cfg<LOG2N>;
    constexpr int N = C::N, T = C::T, R3 = C::R3, LD2 = C::LD2, LDO = C::LDO;
    static_assert(R3 == 16 || R3 == 32, "stage 3 handles 16- and 32-point sub-transforms");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* acc = reinterpret_cast<double*>(smem_raw);                                        // [nb]
    double2* buf = reinterpret_cast<double2*>(smem_raw + (((size_t)P.nb * sizeof(double) + 15) / 16) * 16);
    __shared__ double red[T];
    const int tid = threadIdx.x;
    const int sig = blockIdx.x / P.groups, grp = blockIdx.x % P.groups;
    const double* xs = P.x + (size_t)sig * N;

    // mean of the signal (fixed-order tree)
    double a0 = 0.0;
#pragma unroll
    for (int j = 0; j < 16; ++j) a0 += xs[tid + T * j];
    red[tid] = a0;
    __syncthreads();
    for (int off = T / 2; off > 0; off >>= 1) {
        if (tid < off) red[tid] += red[tid + off];
        __syncthreads();
    }
    const double mean = red[0] / (double)N;
    for (int b = tid; b < P.nb; b += T) acc[b] = 0.0;

    const double2 tw1 = P.tw[tid];                                   // W_N^t
    const int k1s = tid / R3, t2 = tid % R3;                         // stage-2 role
    const double2 tw2 = P.tw[16 * t2];                               // W_T^t2 = W_N^(16 t2)
    const int f3 = (R3 == 32) ? (tid >> 1) : tid, h3 = (R3 == 32) ? (tid & 1) : 0;      // stage-3 role: sub-transform, half
    // which rows hi = k >> 8 of the spectrum are needed (bins k_lo..k_hi-1 and their mirrors N - k)
    unsigned need = 0;
    {
        const int k_hi = P.k_lo + P.nb;      // exclusive
        for (int hi = 0; hi < N / 256; ++hi) {
            const int lo_k = hi * 256, hi_k = lo_k + 255;
            const bool direct = hi_k >= P.k_lo && lo_k < k_hi;
            const bool mirror = hi_k >= N - (k_hi - 1) && lo_k <= N - P.k_lo;      // N - k for k in [k_lo, k_hi)
            if (direct || mirror || hi == 0) need |= 1u << hi;
        }
    }
    const int pairs = (P.K + 1) / 2;
    for (int pr = grp; pr < pairs; pr += P.groups) {
        const int ka = 2 * pr, kb = 2 * pr + 1;
        const double* ha = P.tapers + (size_t)ka * N;
        const double* hb = (kb < P.K) ? P.tapers + (size_t)kb * N : nullptr;
        const double wa = P.weights[ka], wb = hb ? P.weights[kb] : 0.0;
        double2 v[16];
        // ---- stage 1: loads with the tapers multiplied in, 16-point DFT over j, twiddle, -> S1[k1][t]
        if (hb) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const double xv = xs[tid + T * j] - mean;
                v[j] = make_double2(xv * ha[tid + T * j], xv * hb[tid + T * j]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = make_double2((xs[tid + T * j] - mean) * ha[tid + T * j], 0.0);
        }
        fft16(v);
        twiddle16(v, tw1);
#pragma unroll
        for (int k = 0; k < 16; ++k) buf[k * T + tid] = v[bitrev4(k)];
        __syncthreads();
        // ---- stage 2
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = buf[k1s * T + t2 + R3 * j];
        __syncthreads();
        fft16(v);
        twiddle16(v, tw2);
#pragma unroll
        for (int k = 0; k < 16; ++k) buf[(k1s * 16 + k) * LD2 + t2] = v[bitrev4(k)];
        __syncthreads();
        // ---- stage 3
        if (R3 == 32) {
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                const double2 a = buf[f3 * LD2 + t], b = buf[f3 * LD2 + t + 16];
                const double2 s = cadd(a, b), d = cmul(csub(a, b), w32(t));
                v[t] = h3 ? d : s;
            }
        } else {
#pragma unroll
            for (int t = 0; t < 16; ++t) v[t] = buf[f3 * LD2 + t];
        }
        __syncthreads();
        fft16(v);
        // ---- hand the needed rows to the bin owners: X[k1 + 16 k2 + 256 hi], hi = q (R3 = 16) or h + 2 q (R3 = 32)
        {
            const int slot = tid + (tid >> 5);
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int hi = (R3 == 32) ? (h3 + 2 * q) : q;
                if ((need >> hi) & 1u) buf[q * LDO + slot] = v[bitrev4(q)];
            }
        }
        __syncthreads();
        auto fetch = [&](const int k) -> double2 {
            const int k1 = k & 15, k2 = (k >> 4) & 15, hi = k >> 8;
            const int th = (R3 == 32) ? (((k1 << 4) | k2) << 1 | (hi & 1)) : ((k1 << 4) | k2);
            const int q = (R3 == 32) ? (hi >> 1) : hi;
            return buf[q * LDO + th + (th >> 5)];
        };
        for (int b = tid; b < P.nb; b += T) {
            const int k = P.k_lo + b, km = (N - k) & (N - 1);
            const double2 z = fetch(k), zm = fetch(km);
            const double sr = z.x + zm.x, si = z.y - zm.y;          // Z[k] + conj Z[n-k]  = 2 X_a[k]
            const double dr = z.x - zm.x, di = z.y + zm.y;          // Z[k] - conj Z[n-k]  = 2 i X_b[k]
            const double pa = 0.25 * fma(sr, sr, si * si), pb = 0.25 * fma(dr, dr, di * di);
            acc[b] += fma(wa * wa, pa, wb * wb * pb);
        }
        __syncthreads();
    }
    double* out = P.partial + ((size_t)sig * P.groups + grp) * P.nb;
    for (int b = tid; b < P.nb; b += T) out[b] = acc[b];
}

template <int LOG2N>
static int launch_r16(PsdParams P, const double* weights, double* psd, cudaStream_t st) {
    using C = R16Cfg<LOG2N>;
    const size_t smem = (((size_t)P.nb * sizeof(double) + 15) / 16) * 16 + (size_t)C::BUF * sizeof(double2);
    cudaError_t e = cudaFuncSetAttribute(mt_psd_r16_kernel<LOG2N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(HS_ERR_CUDA, "mt_psd_r16: %s", cudaGetErrorString(e));
    mt_psd_r16_kernel<LOG2N><<<P.n_sig * P.groups, C::T, smem, st>>>(P);
    int rc = check_launch("mt_psd_r16_kernel");
    if (rc) return rc;
    dim3 grid((P.nb + 255) / 256, P.n_sig);
    mt_psd_finish_kernel<<<grid, 256, 0, st>>>(P.partial, weights, P.K, P.groups, P.nb, P.k_lo, P.n, psd);
    return check_launch("mt_psd_finish_kernel");
}

// =====================================================================================
// Path (3): any length.  Units (signal, taper pair) are transformed in chunks by the batched global-memory FFT:
//   smooth n (prime factors <= 31):   Z = FFT_n(z)
//   otherwise (Bluestein):            a[t] = z[t] c[t], c[t] = exp(-pi i t^2 / n), zero padded to M = 2^ceil(log2(2n - 1));
//                                     Z[k] = c[k] * IFFT_M(FFT_M(a) * FFT_M(b))[k],  b[j] = conj c[|j|] (circular)
//   (the angle uses t^2 mod 2n in integers, so the chirp is exact to rounding for any n < 2^24).
// The power of the chunk's pairs is accumulated per (signal, bin) in ascending pair order (deterministic).
// =====================================================================================
__global__ void psd_mean_kernel(const double* __restrict__ x, const long long n, double* __restrict__ mean) {
    __shared__ double red[256];
    const double* xs = x + (size_t)blockIdx.x * n;
    double a = 0.0;
    for (long long t = threadIdx.x; t < n; t += 256) a += xs[t];
    red[threadIdx.x] = a;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) mean[blockIdx.x] = red[0] / (double)n;
}

__global__ void chirp_kernel(double2* __restrict__ c, const long long n) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long long r = (t * t) % (2 * n);
    double s, co;
    sincospi(-(double)r / (double)n, &s, &co);
    c[t] = make_double2(co, s);
}

__global__ void bluestein_filter_kernel(const double2* __restrict__ c, const long long n, const long long M, double2* __restrict__ b) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    double2 v = make_double2(0.0, 0.0);
    if (j < n) v = make_double2(c[j].x, -c[j].y);
    else if (M - j < n) v = make_double2(c[M - j].x, -c[M - j].y);
    b[j] = v;
}

struct GenChunk {
    const double* x;
    const double* tapers;
    const double* mean;
    const double2* chirp;       // null: smooth length
    long long n, L;             // signal length, transform length
    int K, s0, p0, ns, np;      // chunk: signals [s0, s0 + ns), pairs [p0, p0 + np)
};

__global__ void psd_gen_load_kernel(const GenChunk G, double2* __restrict__ z) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= G.L) return;
    const int u = blockIdx.y, s = G.s0 + u / G.np, pr = G.p0 + u % G.np;
    double2 v = make_double2(0.0, 0.0);
    if (t < G.n) {
        const double xv = G.x[(size_t)s * G.n + t] - G.mean[s];
        const int ka = 2 * pr, kb = 2 * pr + 1;
        v = make_double2(xv * G.tapers[(size_t)ka * G.n + t], kb < G.K ? xv * G.tapers[(size_t)kb * G.n + t] : 0.0);
        if (G.chirp) v = cmul(v, G.chirp[t]);
    }
    z[(size_t)u * G.L + t] = v;
}

// a <- conj(a * B): the next FORWARD transform then yields conj(IFFT(a B)) * M
__global__ void bluestein_mul_kernel(double2* __restrict__ a, const double2* __restrict__ B, const long long M) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    double2* p = a + (size_t)blockIdx.y * M + j;
    const double2 v = cmul(*p, B[j]);
    *p = make_double2(v.x, -v.y);
}

__global__ void psd_gen_power_kernel(const GenChunk G, const double2* __restrict__ z, const double* __restrict__ weights, const int k_lo,
                                     const int nb, double* __restrict__ acc) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const int sl = blockIdx.y, s = G.s0 + sl;
    const long long k = k_lo + b, km = (G.n - k) % G.n;
    double a = acc[(size_t)s * nb + b];
    const double invM = 1.0 / (double)G.L;
    for (int pl = 0; pl < G.np; ++pl) {
        const int pr = G.p0 + pl;
        const double2* zu = z + (size_t)(sl * G.np + pl) * G.L;
        double2 zk = zu[k], zm = zu[km];
        if (G.chirp) {       // Z[k] = c[k] conj(y[k]) / M
            zk = cmul(G.chirp[k], make_double2(zk.x * invM, -zk.y * invM));
            zm = cmul(G.chirp[km], make_double2(zm.x * invM, -zm.y * invM));
        }
        const double wa = weights[2 * pr], wb = (2 * pr + 1 < G.K) ? weights[2 * pr + 1] : 0.0;
        const double sr = zk.x + zm.x, si = zk.y - zm.y;
        const double dr = zk.x - zm.x, di = zk.y + zm.y;
        const double pa = 0.25 * fma(sr, sr, si * si), pb = 0.25 * fma(dr, dr, di * di);
        a += fma(wa * wa, pa, wb * wb * pb);
    }
    acc[(size_t)s * nb + b] = a;
}

static inline size_t up256(size_t b) { return (b + 255) / 256 * 256; }
static long long bluestein_len(long long n) {
    long long M = 1;
    while (M < 2 * n - 1) M <<= 1;
    return M;
}
constexpr size_t kGenChunkBytes = (size_t)96 << 20;       // both ping-pong buffers of a chunk: L2-sized

struct GenPlan {
    bool bluestein;
    long long L;
    int ns, np;              // signals and pairs per chunk
    size_t bytes;            // workspace of the general path
};
static GenPlan gen_plan(int n_sig, long long n, int K) {
    GenPlan g;
    g.bluestein = !fft_smooth(n);
    g.L = g.bluestein ? bluestein_len(n) : n;
    const int pairs = (K + 1) / 2;
    const size_t unit = 2 * (size_t)g.L * sizeof(double2);
    long long units = (long long)(kGenChunkBytes / unit);
    if (units < 1) units = 1;
    if (units > 65535) units = 65535;
    g.np = (int)(units < pairs ? units : pairs);
    g.ns = (int)(units / g.np);
    if (g.ns < 1) g.ns = 1;
    if (g.ns > n_sig) g.ns = n_sig;
    g.bytes = up256((size_t)n_sig * sizeof(double)) + 2 * up256((size_t)g.ns * g.np * g.L * sizeof(double2));
    if (g.bluestein) g.bytes += up256((size_t)n * sizeof(double2)) + 2 * up256((size_t)g.L * sizeof(double2));
    return g;
}

static int launch_mt_psd_general(const double* x, int n_sig, long long n, const double* tapers, const double* weights, int K, int k_lo,
                                 int k_hi, double* psd, void* d_ws, cudaStream_t st) {
    const GenPlan g = gen_plan(n_sig, n, K);
    const int nb = k_hi - k_lo, pairs = (K + 1) / 2;
    unsigned char* ws = reinterpret_cast<unsigned char*>(d_ws);
    double* mean = reinterpret_cast<double*>(ws);
    ws += up256((size_t)n_sig * sizeof(double));
    double2* bufA = reinterpret_cast<double2*>(ws);
    ws += up256((size_t)g.ns * g.np * g.L * sizeof(double2));
    double2* bufB = reinterpret_cast<double2*>(ws);
    ws += up256((size_t)g.ns * g.np * g.L * sizeof(double2));
    double2* chirp = nullptr;
    double2* Bspec = nullptr;
    int rc;
    if (g.bluestein) {
        chirp = reinterpret_cast<double2*>(ws);
        ws += up256((size_t)n * sizeof(double2));
        double2* b0 = reinterpret_cast<double2*>(ws);
        ws += up256((size_t)g.L * sizeof(double2));
        double2* b1 = reinterpret_cast<double2*>(ws);
        chirp_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(chirp, n);
        if ((rc = check_launch("chirp_kernel"))) return rc;
        bluestein_filter_kernel<<<(unsigned)((g.L + 255) / 256), 256, 0, st>>>(chirp, n, g.L, b0);
        if ((rc = check_launch("bluestein_filter_kernel"))) return rc;
        if ((rc = fft_forward_batched(b0, b1, 1, g.L, st))) return rc;
        Bspec = b0;
    }
    psd_mean_kernel<<<n_sig, 256, 0, st>>>(x, n, mean);
    if ((rc = check_launch("psd_mean_kernel"))) return rc;
    if (cudaMemsetAsync(psd, 0, (size_t)n_sig * nb * sizeof(double), st) != cudaSuccess) return set_error(HS_ERR_CUDA, "mt_psd: memset failed");
    GenChunk G;
    G.x = x;
    G.tapers = tapers;
    G.mean = mean;
    G.chirp = chirp;
    G.n = n;
    G.L = g.L;
    G.K = K;
    for (int s0 = 0; s0 < n_sig; s0 += g.ns) {
        G.s0 = s0;
        G.ns = (n_sig - s0 < g.ns) ? (n_sig - s0) : g.ns;
        for (int p0 = 0; p0 < pairs; p0 += g.np) {
            G.p0 = p0;
            G.np = (pairs - p0 < g.np) ? (pairs - p0) : g.np;
            const int units = G.ns * G.np;
            double2* a = bufA;
            double2* b = bufB;
            dim3 gl((unsigned)((g.L + 255) / 256), units);
            psd_gen_load_kernel<<<gl, 256, 0, st>>>(G, a);
            if ((rc = check_launch("psd_gen_load_kernel"))) return rc;
            if ((rc = fft_forward_batched(a, b, units, g.L, st))) return rc;
            if (g.bluestein) {
                bluestein_mul_kernel<<<gl, 256, 0, st>>>(a, Bspec, g.L);
                if ((rc = check_launch("bluestein_mul_kernel"))) return rc;
                if ((rc = fft_forward_batched(a, b, units, g.L, st))) return rc;
            }
            dim3 gp((nb + 127) / 128, G.ns);
            psd_gen_power_kernel<<<gp, 128, 0, st>>>(G, a, weights, k_lo, nb, psd);
            if ((rc = check_launch("psd_gen_power_kernel"))) return rc;
        }
    }
    dim3 grid((nb + 255) / 256, n_sig);
    mt_psd_finish_kernel<<<grid, 256, 0, st>>>(psd, weights, K, 1, nb, k_lo, n, psd);      // in place: one partial per signal
    return check_launch("mt_psd_finish_kernel");
}

static int g_psd_path = 0;      // 0: automatic

static bool r16_ok(long long n, int nb) { return (n == 4096 || n == 8192) && (size_t)nb * 8 + 16 + (size_t)R16Cfg<13>::BUF * 16 <= 200 * 1024; }
static bool smem_ok(long long n, int nb) {
    long long n2 = 1;
    while (((n / n2) & 1) == 0) n2 <<= 1;
    const size_t accb = (((size_t)nb * sizeof(double) + 15) / 16) * 16;
    return n / n2 <= 64 && accb + ((size_t)n + (size_t)(n >> 3) + 1) * 16 <= kPsdSmemMax;
}

static int launch_mt_psd(const double* x, int n_sig, long long n, const double* tapers, const double* weights, int K, int k_lo, int k_hi,
                         double* psd, void* d_ws, cudaStream_t st) {
    const int nb = k_hi - k_lo;
    int path = g_psd_path;
    if (path == 1 && !r16_ok(n, nb)) return set_error(HS_ERR_UNSUPPORTED, "hs_mt_psd_f64: the register radix-16 kernel takes n = 4096 or 8192");
    if (path == 0) path = r16_ok(n, nb) ? 1 : (smem_ok(n, nb) ? 2 : 3);
    // the general path's buffers start behind the twiddle table + partial spectra of the shared-memory paths
    if (path == 3) {
        unsigned char* ws = reinterpret_cast<unsigned char*>(d_ws);
        return launch_mt_psd_general(x, n_sig, n, tapers, weights, K, k_lo, k_hi, psd, ws, st);
    }
    if (path == 2) return launch_mt_psd_smem(x, n_sig, n, tapers, weights, K, k_lo, k_hi, psd, d_ws, st);
    PsdParams P;
    memset(&P, 0, sizeof(P));
    P.x = x;
    P.tapers = tapers;
    P.weights = weights;
    P.n = n;
    P.K = K;
    P.k_lo = k_lo;
    P.nb = nb;
    P.n_sig = n_sig;
    P.groups = psd_groups(n_sig, (K + 1) / 2);
    unsigned char* ws = reinterpret_cast<unsigned char*>(d_ws);
    double2* tw = reinterpret_cast<double2*>(ws);
    ws += up256((size_t)n * 16);
    P.tw = tw;
    P.partial = reinterpret_cast<double*>(ws);
    twiddle_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(tw, n);
    int rc = check_launch("twiddle_kernel");
    if (rc) return rc;
    return n == 4096 ? launch_r16<12>(P, weights, psd, st) : launch_r16<13>(P, weights, psd, st);
}

}  // namespace hs

using namespace hs;

extern "C" {

int hs_mt_psd_set_path(int path) {
    if (path < 0 || path > 3) return set_error(HS_ERR_INVALID, "hs_mt_psd_set_path: 0 (automatic), 1 (register radix-16), 2 (shared memory), 3 (general)");
    g_psd_path = path;
    return HS_OK;
}

size_t hs_mt_psd_ws_bytes(int n_sig, int64_t n, int K) {
    if (n_sig <= 0 || n <= 0 || K <= 0) return 256;
    const int pairs = (K + 1) / 2;
    int groups = psd_groups(n_sig, pairs);
    size_t b = 0;
    b += up256((size_t)n * 16);                                              // twiddle table W_n^j
    b += up256((size_t)n_sig * groups * (n / 2 + 1) * sizeof(double));       // partial spectra
    if ((size_t)n * 16 + (size_t)(n / 2 + 1) * 8 + 16 > kPsdSmemMax && g_psd_path == 2)
        b += up256((size_t)n_sig * groups * n * 16);                         // forced shared-memory path with its global FFT buffers
    const size_t gen = gen_plan(n_sig, n, K).bytes;
    return (b > gen ? b : gen) + 256;
}

int hs_mt_psd_f64(const double* d_x, int n_sig, int64_t n, const double* d_tapers, const double* d_weights, int K, int k_lo,
                  int k_hi, double* d_psd, void* d_ws, void* stream) {
    if (!d_x || !d_tapers || !d_weights || !d_psd || !d_ws) return set_error(HS_ERR_INVALID, "hs_mt_psd_f64: null pointer");
    if (n < 2 || K < 1 || k_lo < 0 || k_hi > n / 2 + 1 || k_lo > k_hi) return set_error(HS_ERR_INVALID, "hs_mt_psd_f64: bad sizes");
    if (n > (1 << 24)) return set_error(HS_ERR_UNSUPPORTED, "hs_mt_psd_f64: n too large");
    if (n_sig <= 0 || k_lo == k_hi) return HS_OK;
    return launch_mt_psd(d_x, n_sig, (long long)n, d_tapers, d_weights, K, k_lo, k_hi, d_psd, d_ws, (cudaStream_t)stream);
}

}  // extern "C"
