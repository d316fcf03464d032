// MVAR hot path on sm_100a, FP64:  lag covariances (K3) -> LWR Yule-Walker (K4) ->
// A(f)^-1 / DTF / ffDTF (K5).  See DESIGN.md for the data layout and roofline of each.
//
// Replaces, for batches of windows:
//   count_corr              /root/reference  src/mtmvar.py:35-87
//   ar_coeff                src/mtmvar.py:90-123
//   mvar_transfer_function  src/mtmvar.py:126-162
//   dtf_multivariate        src/mtmvar.py:204-234   (|H|^2, un-normalised)
//   full_freq_dtf           src/mtmvar.py:237-284
//   multivariate_spectra    src/mtmvar.py:165-201   (H V H^T, plain transpose)
#include <cstdio>
#include <cstdlib>
#include "hs_tile.cuh"
#include "hs_internal.h"
#include "mvar_launch.h"

namespace hs {

// =====================================================================================
// z table:  z[k][f] = exp(-(k+1) * 2*pi*i * f / fs), same expression order as
// mtmvar.py:153 (argument rounded exactly as NumPy rounds it, then sincos).
// =====================================================================================
__global__ void ztable_kernel(const double* __restrict__ freqs, int F, int p, double fs, double2* __restrict__ z) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p * F) return;
    const int k = idx / F + 1, f = idx % F;
    // numpy: (((-k * 2) * pi) * 1j) * freqs / fs   ->  imaginary part ((-2k*pi) * f) / fs
    const double t = (((double)(-k * 2) * 3.141592653589793) * freqs[f]) / fs;
    double s, c;
    sincos(t, &s, &c);
    z[idx] = make_double2(c, s);
}

// =====================================================================================
// K5  transfer_dtf :  per (window, frequency)  A(f) = I - sum_k A_k z_k(f),  H = A(f)^-1,
//                     dtf = |H|^2, plus per-row partial sums for the ffDTF denominator.
//
// One CTA = NG tile groups (64 threads each) that share one window's AR coefficients in
// shared memory and work on NG consecutive frequency bins at a time; |H|^2 of the NG
// bins is staged in shared memory so the (m, m, F) output (F fastest, as the reference
// lays it out) is written in NG*8-byte contiguous runs.
// Work unit = (window, segment of seg_len bins); units are dealt round-robin to a
// persistent grid of one CTA per SM.
// =====================================================================================
template <int T>
struct K5Smem {
    // doubles per coefficient row, chosen = 2 (mod 16) so that 8 consecutive rows hit 8 distinct 16 B bank groups;
    // rows and columns are zero-padded to the 8T x 8T tile grid so the assembly loop needs no bounds checks
    static __host__ __device__ size_t coef_stride(int m, int p) { const int mp = 8 * T * p; return (size_t)(mp + ((2 - (mp & 15) + 16) & 15)); }
    static __host__ __device__ size_t coef_bytes(int m, int p) { return coef_stride(m, p) * 8 * T * sizeof(double); }
    static __host__ __device__ size_t group_bytes() {
        return sizeof(GJScratch) + sizeof(GJLScratch) + 2 * kPadMax * sizeof(double) + 3 * kPadMax * sizeof(double2) + 16;
    }
    // probe[i][k] = sum_j A_k[i][j] u_j (complex), the per-window part of the a-posteriori check vector A(f) u
    static __host__ __device__ size_t probe_bytes(int p) { return (size_t)kPadMax * p * sizeof(double2); }
    static __host__ __device__ size_t total(int m, int p, int ng) { return coef_bytes(m, p) + probe_bytes(p) + (size_t)ng * group_bytes() + 64; }
};

// probe vector of the a-posteriori check (any fixed vector without structure the elimination could preserve)
__device__ __forceinline__ double2 probe_u(const int j) {
    return make_double2(1.0 + 0.03125 * j, ((j & 1) ? -1.0 : 1.0) * (0.5 + 0.015625 * j));
}

// out_part[warp][i] = sum over this warp's columns of tile[i][j] * x[j]   (x given per column slot b)
template <int T>
__device__ __forceinline__ void tile_matvec_partial(const double (&ar)[T][T], const double (&ai)[T][T], const double2 (&x)[T],
                                                    double2* __restrict__ out_part, const Group& g) {
#pragma unroll
    for (int a = 0; a < T; ++a) {
        double sr = 0.0, si = 0.0;
#pragma unroll
        for (int b = 0; b < T; ++b) {
            sr = fma(ar[a][b], x[b].x, fma(-ai[a][b], x[b].y, sr));
            si = fma(ar[a][b], x[b].y, fma(ai[a][b], x[b].x, si));
        }
        sr += __shfl_xor_sync(0xffffffffu, sr, 8);
        si += __shfl_xor_sync(0xffffffffu, si, 8);
        sr += __shfl_xor_sync(0xffffffffu, sr, 16);
        si += __shfl_xor_sync(0xffffffffu, si, 16);
        if ((g.l64 & 24) == 0) out_part[(g.l64 >> 5) * kPadMax + g.tr + 8 * a] = make_double2(sr, si);
    }
}

// MODE 0: pivoted Gauss-Jordan for every matrix.
// MODE 1: optimistic static-pivot Gauss-Jordan + a-posteriori check ||H (A u) - u|| <= tol; failures are flagged in
//         P.bad / P.bad_count and left to a MODE 2 launch.
// MODE 2: pivoted Gauss-Jordan for the flagged matrices only (row sums go to P.rowpart2).
template <int T, int NG, int MODE>
__global__ void __launch_bounds__(NG * 64, 1) transfer_dtf_kernel(const K5Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int m = P.m, p = P.p, F = P.F;
    // MODE 2 walks the compact list of flagged matrices: one unit = one matrix, handled by group 0 of the CTA
    const int n_bad = (MODE == 2) ? min(*P.bad_count, P.n_win * P.F) : 0;
    if (MODE == 2 && n_bad == 0) return;
    const int cstride = (int)K5Smem<T>::coef_stride(m, p);
    double* coef = reinterpret_cast<double*>(smem_raw);
    const Group g = make_group(P.flip);
    double2* probe = reinterpret_cast<double2*>(smem_raw + K5Smem<T>::coef_bytes(m, p));          // [40][p]
    unsigned char* gbase0 = smem_raw + K5Smem<T>::coef_bytes(m, p) + K5Smem<T>::probe_bytes(p);
    unsigned char* gbase = gbase0 + (size_t)g.gid * K5Smem<T>::group_bytes();
    GJScratch* sh = reinterpret_cast<GJScratch*>(gbase);
    GJLScratch* shl = reinterpret_cast<GJLScratch*>(gbase + sizeof(GJScratch));
    double* rs_grp = reinterpret_cast<double*>(gbase + sizeof(GJScratch) + sizeof(GJLScratch));                  // [2][40]
    double2* vfull = reinterpret_cast<double2*>(gbase + sizeof(GJScratch) + sizeof(GJLScratch) + 2 * kPadMax * sizeof(double));   // [40] A u
    double2* wpart = vfull + kPadMax;                                                            // [2][40] S (A u)
    double* rs_mine = rs_grp + (g.l64 >> 5) * kPadMax;
    const int n_units = (MODE == 2) ? n_bad : P.n_win * P.n_seg;
    const bool vec_ok = ((p & 1) == 0);
    const int off00 = g.tr * cstride + g.tc * p;      // this thread's tile entry (a = 0, b = 0) in the coefficient block
    double* rowpart = (MODE == 2) ? P.rowpart2 : P.rowpart;

    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        int w = unit / P.n_seg, seg = unit % P.n_seg;
        int f_begin = seg * P.seg_len, f_end = min(F, f_begin + P.seg_len);
        if (MODE == 2) {
            const int idx = P.bad_list[unit];
            w = idx / F;
            f_begin = idx - w * F;
            f_end = f_begin + 1;
        }
        __syncthreads();
        {   // AR coefficients of window w -> shared (zero padded to 8T x 8T, row stride cstride)
            const double* Aw = P.A + (size_t)w * m * m * p;
            const int row_len = m * p, pad_len = 8 * T * p;
            for (int i = g.gid * 2 + (g.l64 >> 5); i < 8 * T; i += NG * 2) {       // one warp per row
                for (int c = threadIdx.x & 31; c < pad_len; c += 32)
                    coef[i * cstride + c] = (i < m && c < row_len) ? Aw[(size_t)i * row_len + c] : 0.0;
            }
            for (int e = g.l64; e < 2 * kPadMax; e += 64) rs_grp[e] = 0.0;
        }
        __syncthreads();
        if (MODE == 1) {
            // probe[i][k] = sum_j A_k[i][j] u_j : the window-dependent part of v = A(f) u = u - sum_k z_k(f) probe[.][k]
            for (int e = threadIdx.x; e < kPadMax * p; e += NG * 64) {
                const int i = e / p, k = e - i * p;
                double sr = 0.0, si = 0.0;
                if (i < 8 * T) {
                    const double* cp = coef + i * cstride + k;
                    for (int j = 0; j < m; ++j) {
                        const double2 u = probe_u(j);
                        const double c = cp[j * p];
                        sr = fma(c, u.x, sr);
                        si = fma(c, u.y, si);
                    }
                }
                probe[e] = make_double2(sr, si);
            }
            __syncthreads();
        }

        for (int f = f_begin + g.gid; f < f_end; f += NG) {
            double ar[T][T], ai[T][T];
            // ---- build A(f) = I - sum_k A_k z_k(f)
#pragma unroll
            for (int a = 0; a < T; ++a)
#pragma unroll
                for (int b = 0; b < T; ++b) {
                    ar[a][b] = ((g.tr + 8 * a) == (g.tc + 8 * b)) ? 1.0 : 0.0;
                    ai[a][b] = 0.0;
                }
            // v = A(f) u for the check: thread l64 (< 40) accumulates entry l64 next to the assembly
            const int vi = min(g.l64, kPadMax - 1);
            double2 vacc = (MODE == 1 && g.l64 < m) ? probe_u(g.l64) : make_double2(0.0, 0.0);
            if (vec_ok) {
                for (int k = 0; k < p; k += 2) {
                    const double2 z0 = __ldg(&P.z[(size_t)k * F + f]);
                    const double2 z1 = __ldg(&P.z[(size_t)(k + 1) * F + f]);
                    const double* cp = coef + off00 + k;
#pragma unroll
                    for (int a = 0; a < T; ++a) {
#pragma unroll
                        for (int b = 0; b < T; ++b) {
                            const double2 c = *reinterpret_cast<const double2*>(cp + a * 8 * cstride + b * 8 * p);
                            ar[a][b] = fma(-c.x, z0.x, fma(-c.y, z1.x, ar[a][b]));
                            ai[a][b] = fma(-c.x, z0.y, fma(-c.y, z1.y, ai[a][b]));
                        }
                    }
                    if (MODE == 1) {
                        const double2 q0 = probe[vi * p + k], q1 = probe[vi * p + k + 1];
                        vacc.x = fma(-q0.x, z0.x, fma(q0.y, z0.y, fma(-q1.x, z1.x, fma(q1.y, z1.y, vacc.x))));
                        vacc.y = fma(-q0.x, z0.y, fma(-q0.y, z0.x, fma(-q1.x, z1.y, fma(-q1.y, z1.x, vacc.y))));
                    }
                }
            } else {
                for (int k = 0; k < p; ++k) {
                    const double2 zz = __ldg(&P.z[(size_t)k * F + f]);
                    const double* cp = coef + off00 + k;
#pragma unroll
                    for (int a = 0; a < T; ++a) {
#pragma unroll
                        for (int b = 0; b < T; ++b) {
                            const double c = cp[a * 8 * cstride + b * 8 * p];
                            ar[a][b] = fma(-c, zz.x, ar[a][b]);
                            ai[a][b] = fma(-c, zz.y, ai[a][b]);
                        }
                    }
                    if (MODE == 1) {
                        const double2 q0 = probe[vi * p + k];
                        vacc.x = fma(-q0.x, zz.x, fma(q0.y, zz.y, vacc.x));
                        vacc.y = fma(-q0.x, zz.y, fma(-q0.y, zz.x, vacc.y));
                    }
                }
            }
            const size_t wbase = (size_t)w * m * m;
            if (P.Af && MODE != 2) {
#pragma unroll
                for (int a = 0; a < T; ++a)
#pragma unroll
                    for (int b = 0; b < T; ++b) {
                        const int i = g.tr + 8 * a, j = g.tc + 8 * b;
                        if (i < m && j < m) P.Af[(wbase + (size_t)i * m + j) * F + f] = make_double2(ar[a][b], ai[a][b]);
                    }
            }
            bool good = true;
            double2 sc[T];          // MODE 1: H[i][j] = sc[a] * tile[a][b] (deferred row scaling of gj_inverse_la)
            if (MODE == 1) {
                group_sync(g);                      // previous matrix' check has finished reading vfull / wpart
                if (g.l64 < kPadMax) vfull[g.l64] = vacc;
                if (g.l64 == 0) sh->singular = 0;
                gj_inverse_la<T, true>(ar, ai, m, g, shl);
                double2 x[T];
#pragma unroll
                for (int b = 0; b < T; ++b) x[b] = vfull[g.tc + 8 * b];         // 0 in the padding
#pragma unroll
                for (int a = 0; a < T; ++a) sc[a] = shl->pinv[g.tr + 8 * a];
                tile_matvec_partial<T>(ar, ai, x, wpart, g);
                group_sync(g);
                if (g.l64 < m) {
                    const double2 u = probe_u(g.l64), iv = shl->pinv[g.l64];
                    const double sr = wpart[g.l64].x + wpart[kPadMax + g.l64].x, si = wpart[g.l64].y + wpart[kPadMax + g.l64].y;
                    const double er = fma(iv.x, sr, -iv.y * si) - u.x;
                    const double ei = fma(iv.x, si, iv.y * sr) - u.y;
                    const double err = fma(er, er, ei * ei), ref = fma(u.x, u.x, u.y * u.y);
                    if (!(err <= P.verify_tol2 * ref)) sh->singular = 1;      // also catches NaN / Inf
                }
                group_sync(g);
                good = (sh->singular == 0);
                if (!good && g.l64 == 0) {
                    P.bad[(size_t)w * F + f] = 1;
                    P.bad_list[atomicAdd(P.bad_count, 1)] = w * F + f;
                }
            } else {
                gj_inverse<T, true>(ar, ai, m, g, sh);
                if (g.l64 == 0 && sh->singular) atomicOr(&P.status[w], 1);
            }
            if (good) {
                // ---- |H|^2 straight to global (8 B stores; the 32 B sectors are completed in L2 by the neighbouring
                //      bins, written by the other groups of this CTA within the same ~100 us), plus this bin's
                //      contribution to the ffDTF row sums.
                int cj[T];
#pragma unroll
                for (int b = 0; b < T; ++b) cj[b] = (MODE == 1) ? (g.tc + 8 * b) : sh->colmap[min(g.tc + 8 * b, kPadMax - 1)];
#pragma unroll
                for (int a = 0; a < T; ++a) {
                    const int i = g.tr + 8 * a;
                    const int ri = (MODE == 1) ? i : sh->rowmap[min(i, kPadMax - 1)];
                    const double s2 = (MODE == 1) ? fma(sc[a].x, sc[a].x, sc[a].y * sc[a].y) : 1.0;
                    double rsum = 0.0;
#pragma unroll
                    for (int b = 0; b < T; ++b) {
                        const int j = g.tc + 8 * b;
                        if (i < m && j < m) {
                            double v = fma(ar[a][b], ar[a][b], ai[a][b] * ai[a][b]);
                            if (MODE == 1) v *= s2;
                            rsum += v;
                            const size_t o = (wbase + (size_t)ri * m + cj[b]) * F + f;
                            if (P.dtf) P.dtf[P.dtf_fij ? (((size_t)w * m + ri) * F + f) * m + cj[b] : o] = v;
                            if (P.H) {
                                double2 h = make_double2(ar[a][b], ai[a][b]);
                                if (MODE == 1) h = make_double2(fma(sc[a].x, ar[a][b], -sc[a].y * ai[a][b]), fma(sc[a].x, ai[a][b], sc[a].y * ar[a][b]));
                                P.H[o] = h;
                            }
                        }
                    }
                    // reduce over the 4 column groups of this warp (lane bits 3, 4), fixed order
                    rsum += __shfl_xor_sync(0xffffffffu, rsum, 8);
                    rsum += __shfl_xor_sync(0xffffffffu, rsum, 16);
                    if ((g.l64 & 24) == 0 && i < m) rs_mine[ri] += rsum;      // one writer per (warp, row): deterministic
                }
            }
        }
        // ---- per-unit row sums (fixed summation order -> deterministic)
        __syncthreads();
        if (MODE != 2 && rowpart && threadIdx.x < m) {
            double acc = 0.0;
            for (int q = 0; q < NG; ++q) {
                const double* rg = reinterpret_cast<const double*>(gbase0 + (size_t)q * K5Smem<T>::group_bytes() + sizeof(GJScratch) + sizeof(GJLScratch));
                acc += rg[threadIdx.x] + rg[kPadMax + threadIdx.x];
            }
            rowpart[((size_t)w * P.n_seg + seg) * m + threadIdx.x] = acc;
        }
    }
}

// ffdtf[w][i][j][f] = dtf[w][i][j][f] / sum_seg rowpart[w][seg][i]     (mtmvar.py:281-283)
__global__ void ffdtf_normalize_kernel(double* __restrict__ dtf, const double* __restrict__ rowpart, const double* __restrict__ rowpart2, int m,
                                       int F, int n_seg, double* __restrict__ out) {
    const int w = blockIdx.y;
    const int i = blockIdx.x;
    __shared__ double inv_s;
    if (threadIdx.x == 0) {
        double acc = 0.0;
        for (int s = 0; s < n_seg; ++s) acc += rowpart[((size_t)w * n_seg + s) * m + i];
        if (rowpart2)
            for (int s = 0; s < n_seg; ++s) acc += rowpart2[((size_t)w * n_seg + s) * m + i];
        inv_s = acc;
    }
    __syncthreads();
    const double denom = inv_s;
    const size_t base = ((size_t)w * m + i) * (size_t)m * F;
    const size_t n = (size_t)m * F;
    if ((n & 1) == 0 && ((reinterpret_cast<uintptr_t>(dtf + base) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out + base) & 15) == 0)) {
        const double2* src = reinterpret_cast<const double2*>(dtf + base);
        double2* dst = reinterpret_cast<double2*>(out + base);
        for (size_t e = threadIdx.x; e < n / 2; e += blockDim.x) {
            double2 v = src[e];
            v.x = v.x / denom;
            v.y = v.y / denom;
            dst[e] = v;
        }
    } else {
        for (size_t e = threadIdx.x; e < n; e += blockDim.x) out[base + e] = dtf[base + e] / denom;
    }
}

// Transposing finalize:  stage (n_win, m, F, m)  ->  dtf / ffdtf (n_win, m, m, F)  with
// ffdtf[w][i][j][f] = dtf[w][i][j][f] / sum_{j,f} dtf[w][i][j][f]   (mtmvar.py:281-283).
// One CTA per (window, row i): its input is ONE contiguous run of F * m doubles (K5 stores row i of every bin's matrix there), read
// with 16-byte loads, transposed through shared memory in chunks of kFinChunk bins and written as contiguous runs along f.
constexpr int kFinChunk = 64;
__global__ void __launch_bounds__(256, 4) dtf_finalize_kernel(const double* __restrict__ stage, const double* __restrict__ rowpart,
                                                           const int* __restrict__ bad, int m, int F, int n_seg,
                                                           double* __restrict__ dtf_out, double* __restrict__ ffdtf_out) {
    extern __shared__ double tile[];                 // [m][kFinChunk + 1]
    const int w = blockIdx.y, i = blockIdx.x;
    __shared__ double denom_s;
    __shared__ int any_bad;
    const size_t obase = ((size_t)w * m + i) * (size_t)m * F;
    constexpr int kPer = 5;                           // double2 per thread and chunk (m * kFinChunk <= 2560)
    const bool vec_ok = ((F & 1) == 0) && ((reinterpret_cast<uintptr_t>(dtf_out) & 15) == 0) && ((reinterpret_cast<uintptr_t>(ffdtf_out) & 15) == 0);
    const int nchunk = (F + kFinChunk - 1) / kFinChunk;
    const double* sbase = stage + obase;              // (f, j) of this (w, i) at sbase[f * m + j]
    const bool in_vec = ((m & 1) == 0) && ((reinterpret_cast<uintptr_t>(stage) & 15) == 0);      // a chunk is a whole number of aligned double2
    const int chunk_len = kFinChunk * m;
    const int total = F * m;
    // (f, j) of the pair e = 2 (tid + 256 u) of a chunk, advanced incrementally (no division in the loops)
    const int f_first = (2 * threadIdx.x) / m, j_first = 2 * threadIdx.x - f_first * m;
    const int df = 512 / m, dj = 512 - df * m;
    double2 v[kPer];
    auto fetch = [&](const int f0) {
        const int base = f0 * m;
#pragma unroll
        for (int u = 0; u < kPer; ++u) {
            const int e = 2 * (threadIdx.x + 256 * u);
            if (in_vec) {
                v[u] = (e < chunk_len && base + e < total) ? __ldcs(reinterpret_cast<const double2*>(sbase + base + e)) : make_double2(0.0, 0.0);
            } else {
                v[u].x = (e < chunk_len && base + e < total) ? __ldcs(sbase + base + e) : 0.0;
                v[u].y = (e + 1 < chunk_len && base + e + 1 < total) ? __ldcs(sbase + base + e + 1) : 0.0;
            }
        }
    };
    fetch(0);                                         // in flight while the denominator is put together
    if (threadIdx.x == 0) any_bad = 0;
    __syncthreads();
    if (rowpart && bad) {
        int mine = 0;
        for (int f = threadIdx.x; f < F; f += blockDim.x) mine |= bad[(size_t)w * F + f];
        if (mine) any_bad = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // row sum of |H|^2 over (j, f): per-segment partials of the optimistic pass (which skips the bins it flagged) ...
        double acc = 0.0;
        if (rowpart)
            for (int s = 0; s < n_seg; ++s) acc += rowpart[((size_t)w * n_seg + s) * m + i];
        // ... plus the bins redone with pivoting, summed here in a fixed order (rare)
        if (rowpart && any_bad)
            for (int f = 0; f < F; ++f)
                if (bad[(size_t)w * F + f])
                    for (int j = 0; j < m; ++j) acc += sbase[(size_t)f * m + j];
        denom_s = acc;
    }
    __syncthreads();
    const double denom = denom_s;
    // x / denom as a multiplication by the reciprocal plus one residual correction (the quotient an IEEE division returns in all but
    // the rarest half-ulp ties; 3 FP64 instructions instead of the ~20 of the division sequence, per output element)
    const double rden = 1.0 / denom;
    for (int c = 0; c < nchunk; ++c) {
        const int f0 = c * kFinChunk;
        const int nf = min(kFinChunk, F - f0);
        {
            int f = f_first, j = j_first;
#pragma unroll
            for (int u = 0; u < kPer; ++u) {
                if (f < kFinChunk) {
                    tile[j * (kFinChunk + 1) + f] = v[u].x;
                    if (j + 1 < m) tile[(j + 1) * (kFinChunk + 1) + f] = v[u].y;
                    else if (f + 1 < kFinChunk) tile[f + 1] = v[u].y;                   // (f + 1, j = 0): odd m only
                }
                f += df;
                j += dj;
                if (j >= m) { j -= m; ++f; }
            }
        }
        __syncthreads();
        if (c + 1 < nchunk) fetch(f0 + kFinChunk);      // the next chunk's reads are in flight while this one is written out
        // write-out: one warp per tile row j, every lane two consecutive bins (16-byte stores when the run is aligned)
        {
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            const int f = 2 * lane;
            for (int j = warp; j < m; j += 8) {
                const double* trow = tile + j * (kFinChunk + 1) + f;
                const double x0 = trow[0], x1 = trow[1];
                const size_t o = obase + (size_t)j * F + f0 + f;
                double q0 = x0 * rden, q1 = x1 * rden;
                q0 = fma(fma(-q0, denom, x0), rden, q0);
                q1 = fma(fma(-q1, denom, x1), rden, q1);
                if (vec_ok && f + 1 < nf) {
                    if (dtf_out) __stcs(reinterpret_cast<double2*>(dtf_out + o), make_double2(x0, x1));
                    if (ffdtf_out) __stcs(reinterpret_cast<double2*>(ffdtf_out + o), make_double2(q0, q1));
                } else {
                    if (f < nf) {
                        if (dtf_out) dtf_out[o] = x0;
                        if (ffdtf_out) ffdtf_out[o] = q0;
                    }
                    if (f + 1 < nf) {
                        if (dtf_out) dtf_out[o + 1] = x1;
                        if (ffdtf_out) ffdtf_out[o + 1] = q1;
                    }
                }
            }
        }
        __syncthreads();
    }
}

int launch_dtf_finalize(const double* stage, const double* rowpart, const int* bad, int n_win, int m, int F, int n_seg,
                        double* dtf_out, double* ffdtf_out, cudaStream_t stream) {
    dim3 grid(m, n_win);
    const size_t smem = (size_t)m * (kFinChunk + 1) * sizeof(double);
    dtf_finalize_kernel<<<grid, 256, smem, stream>>>(stage, rowpart, bad, m, F, n_seg, dtf_out, ffdtf_out);
    return check_launch("dtf_finalize_kernel");
}

template <int T, int NG, int MODE>
static int launch_k5_t(const K5Params& P, int sm_count, cudaStream_t stream) {
    const size_t smem = K5Smem<T>::total(P.m, P.p, NG);
    cudaError_t e = cudaFuncSetAttribute(transfer_dtf_kernel<T, NG, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(HS_ERR_CUDA, "transfer_dtf: cannot reserve %zu B shared memory: %s", smem, cudaGetErrorString(e));
    const int n_units = (MODE == 2) ? P.n_win * P.F : P.n_win * P.n_seg;
    const int grid = n_units < sm_count ? n_units : sm_count;
    transfer_dtf_kernel<T, NG, MODE><<<grid, NG * 64, smem, stream>>>(P);
    return check_launch("transfer_dtf_kernel");
}

template <int NG, int MODE>
static int launch_k5_ng(const K5Params& P, int sm, cudaStream_t stream) {
    switch ((P.m + 7) / 8) {
        case 1: return launch_k5_t<1, NG, MODE>(P, sm, stream);
        case 2: return launch_k5_t<2, NG, MODE>(P, sm, stream);
        case 3: return launch_k5_t<3, NG, MODE>(P, sm, stream);
        case 4: return launch_k5_t<4, NG, MODE>(P, sm, stream);
        case 5: return launch_k5_t<5, NG, MODE>(P, sm, stream);
    }
    return set_error(HS_ERR_UNSUPPORTED, "transfer_dtf: no register-tile kernel for m=%d", P.m);
}

// mode 0: pivoted only; mode 1: optimistic + verify; mode 2: pivoted redo of the flagged matrices
int launch_transfer_dtf(const K5Params& P, int ng, int mode, cudaStream_t stream) {
    const int sm = device_sm_count();
#ifdef HS_EXPERIMENT
    if (ng == 8) {
        if (mode == 0) return launch_k5_ng<8, 0>(P, sm, stream);
        if (mode == 1) return launch_k5_ng<8, 1>(P, sm, stream);
        return launch_k5_ng<8, 2>(P, sm, stream);
    }
    if (ng == 7) {
        if (mode == 0) return launch_k5_ng<7, 0>(P, sm, stream);
        if (mode == 1) return launch_k5_ng<7, 1>(P, sm, stream);
        return launch_k5_ng<7, 2>(P, sm, stream);
    }
    if (mode == 1) return launch_k5_ng<6, 1>(P, sm, stream);
#endif
    // product build: the register-tile kernel only runs WITH pivoting (mode 0: every matrix, for shapes the tensor-pipe
    // kernel does not take; mode 2: the matrices the optimistic tensor-pipe pass flagged)
    if (mode == 0 || mode == 1) return launch_k5_ng<6, 0>(P, sm, stream);
    return launch_k5_ng<6, 2>(P, sm, stream);
}

// =====================================================================================
// K3  lag_cov :  R(L)[i][j] = 1/(n*trials) * sum_trials sum_{t=0}^{n-1-L} x_i(t) x_j(t+L),
//                L = 0..p   (count_corr, mtmvar.py:54-59, 72-73, 78-85; biased, no mean removal).
//
// One CTA per window.  Group g of the CTA owns lag (round*NG + g); all groups consume the
// same time-major panel  xs[t][ch]  staged once per chunk of TC samples (+ halo of the
// largest lag in the round), so the window is read from L2/HBM once per lag-round.
// m > 40 is handled by looping over 40 x 40 output blocks.
// =====================================================================================
static_assert(kPadMaxHost == kPadMax, "pad mismatch");
constexpr int kK3Chunk = 64;
constexpr int kK3Ld = 41;       // odd stride: transposed staging stores are conflict-light


template <int T>
__global__ void __launch_bounds__(576, 1) lagcov_kernel(const K3Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NG = blockDim.x >> 6;
    const Group g = make_group();
    const int m = P.m, n = P.n, p = P.p;
    const int w = blockIdx.x;
    const int nblk = (m + kPadMax - 1) / kPadMax;
    double* panelB = reinterpret_cast<double*>(smem_raw);                       // [(TC + NG-1 .. ) x 41]
    const int halo_max = p;                                                     // rows of halo reserved
    double* panelA = panelB + (size_t)(kK3Chunk + halo_max) * kK3Ld;            // [TC x 41], only when bi != bj
    const double scale = 1.0 / ((double)n * (double)P.trials);

    for (int bi = 0; bi < nblk; ++bi)
        for (int bj = 0; bj < nblk; ++bj) {
            const int mi = min(kPadMax, m - bi * kPadMax), mj = min(kPadMax, m - bj * kPadMax);
            for (int l0 = 0; l0 <= p; l0 += NG) {
                const int lag = l0 + g.gid;
                const bool has_lag = lag <= p;
                const int halo = min(l0 + NG - 1, p);          // largest lag of this round: x_j(t + lag) lives `lag` rows below x_i(t)
                double acc[T][T];
#pragma unroll
                for (int a = 0; a < T; ++a)
#pragma unroll
                    for (int b = 0; b < T; ++b) acc[a][b] = 0.0;
                for (int tr = 0; tr < P.trials; ++tr) {
                    const double* xu = P.x + P.offsets[(size_t)w * P.trials + tr];
                    for (int t0 = 0; t0 < n; t0 += kK3Chunk) {
                        __syncthreads();
                        // stage rows [t0, t0+TC+halo) of block bj (and TC rows of block bi if different)
                        const int rowsB = kK3Chunk + halo;
                        for (int e = threadIdx.x; e < kPadMax * rowsB; e += blockDim.x) {
                            const int c = e / rowsB, t = e - c * rowsB;
                            double v = 0.0;
                            if (c < mj && t0 + t < n) v = xu[(size_t)(bj * kPadMax + c) * P.ch_stride + t0 + t];
                            panelB[t * kK3Ld + c] = v;
                        }
                        if (bi != bj) {
                            for (int e = threadIdx.x; e < kPadMax * kK3Chunk; e += blockDim.x) {
                                const int c = e / kK3Chunk, t = e - c * kK3Chunk;
                                double v = 0.0;
                                if (c < mi && t0 + t < n) v = xu[(size_t)(bi * kPadMax + c) * P.ch_stride + t0 + t];
                                panelA[t * kK3Ld + c] = v;
                            }
                        }
                        __syncthreads();
                        if (has_lag) {
                            const double* pa = (bi != bj) ? panelA : panelB;
                            tile_mac<T, false>(acc, pa, kK3Ld, panelB + (size_t)lag * kK3Ld, kK3Ld, kK3Chunk, g);
                        }
                    }
                }
                if (has_lag) {
                    double* Rw = P.R + (((size_t)w * (p + 1) + lag) * m + (size_t)bi * kPadMax) * m + (size_t)bj * kPadMax;
#pragma unroll
                    for (int a = 0; a < T; ++a) {
                        const int i = g.tr + 8 * a;
#pragma unroll
                        for (int b = 0; b < T; ++b) {
                            const int j = g.tc + 8 * b;
                            if (i < mi && j < mj) Rw[(size_t)i * m + j] = acc[a][b] * scale;
                        }
                    }
                }
            }
        }
}

// -------------------------------------------------------------------------------------
// K3 on the FP64 tensor pipe (m <= 40, 5 (p+1) <= 96 output column tiles, window fits in shared memory).
//   R(L)[i][j] = scale * sum_t x_i(t) x_j(t + L)  is, for all lags at once, the product  X (40 x n) . [X_0 | X_1 | ... | X_p]^T
//   where X_L is X shifted by L samples: one channel-major copy of the window in shared memory (zero padded by p + 3
//   samples, so the upper summation limit n - 1 - L needs no test) serves every operand.
//   mma.m8n8k4:  A fragment = x[8 ta + g4][t0 + t4],  B fragment = x[8 tb + g4][t0 + t4 + L]  -- both plain 8-byte loads, conflict
//   free because the row stride is = 4 (mod 16) doubles.  One warp owns three of the 5 (p+1) column tiles (lag, tb) for all
//   row tiles: 8 fragment loads per 15 DMMAs.  The DFMA kernel above issues 25 three-operand DFMAs (2.47 cycles each) + 10 LDS
//   per k step and lag; this form needs 1/8 of the issue slots and leaves the pipe to the DMMAs.
// -------------------------------------------------------------------------------------

__device__ __forceinline__ void dmma884_k3(double& c0, double& c1, const double a, const double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---- bulk asynchronous copies (TMA unit, cp.async.bulk + mbarrier): one elected warp issues a copy per channel row, the
//      barrier's transaction count tells every thread when the bytes have landed; no registers and no LSU instructions are
//      spent on the staging.  16-byte aligned source rows only (window starts on an even sample), else plain loads.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, const int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, const unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, const unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_%=;\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, const unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int KC>      // column tiles per warp: 3 (up to 48 column tiles with 16 warps), 6 (up to 96)
__global__ void __launch_bounds__(512, 1) lagcov_mma_kernel(const K3Params P, const int ld, const int n_full, const int n_parts, const int wpp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long stage_bar;
    double* xs = reinterpret_cast<double*>(smem_raw);                 // [40][ld], channel-major, zero padded
    const int m = P.m, n = P.n, p = P.p;
    // CTAs 0 .. n_full - 1 take a whole window each (full waves of the grid).  The windows of the last, partial wave are split over
    // n_parts CTAs each: part q runs the column tiles of warps q wpp .. (q + 1) wpp - 1 of a whole-window CTA -- the same tiles by
    // the same instructions (bit-identical results), but with the SM's tensor pipe to itself the tail takes 1 / (warps per
    // sub-partition) of a full wave instead of a full wave for a handful of windows.
    const bool split = (int)blockIdx.x >= n_full;
    const int w = split ? n_full + ((int)blockIdx.x - n_full) / n_parts : (int)blockIdx.x;
    const int part = split ? ((int)blockIdx.x - n_full) % n_parts : 0;
    const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int warp_cta = threadIdx.x >> 5;
    const int warp = split ? part * wpp + warp_cta : warp_cta;        // the whole-window warp whose tiles this warp computes
    const bool warp_on = !split || (warp_cta < wpp && warp < nwarps);
    const int g4 = lane >> 2, t4 = lane & 3;
    const int TA = (m + 7) >> 3;                                        // row / column tiles actually populated
    const int n_cols = TA * (p + 1);
    const double scale = 1.0 / ((double)n * (double)P.trials);
    // zero the padding once: columns n .. ld-1 of every row, and rows m .. 8 TA - 1 completely
    for (int e = threadIdx.x; e < kPadMax * ld; e += blockDim.x) {
        const int r = e / ld, c = e - r * ld;
        if (r >= m || c >= n) xs[e] = 0.0;
    }
    // this warp's column tiles
    int lag[KC], tb[KC];
    bool okc[KC];
#pragma unroll
    for (int u = 0; u < KC; ++u) {
        const int c = warp * KC + u;
        okc[u] = warp_on && c < n_cols;
        lag[u] = okc[u] ? c / TA : 0;
        tb[u] = okc[u] ? c - lag[u] * TA : 0;
    }
    double acc[kTileMax][KC][2];
#pragma unroll
    for (int ta = 0; ta < kTileMax; ++ta)
#pragma unroll
        for (int u = 0; u < KC; ++u) acc[ta][u][0] = acc[ta][u][1] = 0.0;

    if (threadIdx.x == 0) mbar_init(&stage_bar, 1);
    unsigned stage_phase = 0;
    const bool rows_aligned = ((n & 1) == 0) && ((P.ch_stride & 1) == 0) && ((reinterpret_cast<uintptr_t>(P.x) & 15) == 0);
    for (int tr = 0; tr < P.trials; ++tr) {
        const long long off = P.offsets[(size_t)w * P.trials + tr];
        const double* xu = P.x + off;
        __syncthreads();                                               // previous trial's products are done with xs (and the barrier is initialised)
        if (rows_aligned && (off & 1) == 0) {
            // TMA path: warp 0 issues one bulk copy per channel row (n * 8 bytes each), all threads wait on the transaction barrier
            if (warp_cta == 0) {
                fence_proxy_async();                                   // the generic-proxy reads of the previous trial precede these async writes
                if (lane == 0) mbar_expect_tx(&stage_bar, (unsigned)(m * n * sizeof(double)));
                __syncwarp();
                for (int r = lane; r < m; r += 32) bulk_g2s(xs + (size_t)r * ld, xu + (size_t)r * P.ch_stride, (unsigned)(n * sizeof(double)), &stage_bar);
            }
            mbar_wait(&stage_bar, stage_phase);
            stage_phase ^= 1;
        } else
        for (int r = warp_cta; r < m; r += nwarps) {                   // one warp per channel row: coalesced
            const double* src = xu + (size_t)r * P.ch_stride;
            double* dst = xs + (size_t)r * ld;
            for (int t0 = lane; t0 < n; t0 += 8 * 32) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = (t0 + 32 * u < n) ? src[t0 + 32 * u] : 0.0;
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (t0 + 32 * u < n) dst[t0 + 32 * u] = v[u];
            }
        }
        __syncthreads();
        if (okc[0]) {
            const double* pa = xs + (size_t)g4 * ld + t4;
            int ob[KC];                                               // element offsets of the B fragments (invalid tiles read tile 0: harmless)
#pragma unroll
            for (int u = 0; u < KC; ++u) ob[u] = (8 * tb[u] + g4) * ld + t4 + lag[u];
            const int row8 = 8 * ld;
#pragma unroll(KC == 3 ? 4 : 1)
            for (int t0 = 0; t0 < n; t0 += 4) {
                double a[kTileMax];
#pragma unroll
                for (int ta = 0; ta < kTileMax; ++ta) a[ta] = (ta < TA) ? pa[ta * row8 + t0] : 0.0;
                double b[KC];
#pragma unroll
                for (int u = 0; u < KC; ++u) b[u] = xs[ob[u] + t0];
#pragma unroll
                for (int ta = 0; ta < kTileMax; ++ta) {
                    if (ta < TA) {
#pragma unroll
                        for (int u = 0; u < KC; ++u) dmma884_k3(acc[ta][u][0], acc[ta][u][1], a[ta], b[u]);
                    }
                }
            }
        }
    }
    // R[w][lag][i][j],  i = 8 ta + g4,  j = 8 tb + 2 t4 + {0, 1}
#pragma unroll
    for (int u = 0; u < KC; ++u) {
        if (!okc[u]) continue;
        double* Rl = P.R + ((size_t)w * (p + 1) + lag[u]) * m * m;
        const int j = 8 * tb[u] + 2 * t4;
#pragma unroll
        for (int ta = 0; ta < kTileMax; ++ta) {
            const int i = 8 * ta + g4;
            if (ta < TA && i < m) {
                if (j < m) Rl[(size_t)i * m + j] = acc[ta][u][0] * scale;
                if (j + 1 < m) Rl[(size_t)i * m + j + 1] = acc[ta][u][1] * scale;
            }
        }
    }
}


// -------------------------------------------------------------------------------------
// K3 for many channels / many epochs (cfg5: 2 x 64 ch, 100 epochs per window, p = 15: the dense lag-covariance contraction,
// 26.5 GFLOP per window): a tiled FP64 tensor-core GEMM.
//   R(L)[i][j] = scale * sum_trials sum_t x_i(t) x_j(t + L):  per (window, lag, 128 x 128 block of (i, j)) one CTA; K dimension =
//   epochs x samples, walked in chunks of 64 samples.  A chunk (128 channel rows x 80 samples: 64 + a 16-sample halo for the lag)
//   is brought in by ONE bulk copy per row (cp.async.bulk, transaction barrier) into one of two buffers while the 16 warps
//   (4 x 4 grid, 32 x 32 outputs = 16 accumulator tiles each) multiply the other: A = X[i][t], B = X[j][t + L] are fragments of the
//   same staged rows (row stride = 4 mod 16 doubles: conflict free).  Samples beyond the epoch's end are zero, which implements the
//   shrinking sum range of count_corr (mtmvar.py:57-59) without branches.
// -------------------------------------------------------------------------------------
constexpr int kG3Block = 128;               // output block edge
constexpr int kG3Halo = 16;                 // halo (>= max lag, multiple of 2)
constexpr int kG3MaxLag = kG3Halo;

// CH = samples per K chunk, ROWS = staged rows per buffer: (64, 128) when the channel count fits one block (only diagonal blocks:
// the i and j operands are the same rows), (32, 256) otherwise; row stride CH + 20 = 4 (mod 16) doubles in both cases.
template <int kG3Chunk, int kG3Rows>
__global__ void __launch_bounds__(512, 1) lagcov_gemm_kernel(const K3Params P) {
    constexpr int kG3Ld = kG3Chunk + kG3Halo + 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long full_bar[2];
    double* buf[2] = {reinterpret_cast<double*>(smem_raw), reinterpret_cast<double*>(smem_raw) + (size_t)kG3Rows * kG3Ld};
    // buf[b]: rows 0..127 = the i block, rows 128..255 = the j block (kG3Rows == 128: diagonal blocks only, one set of rows)
    const int m = P.m, n = P.n, p = P.p;
    const int nbk = (m + kG3Block - 1) / kG3Block;
    const int w = blockIdx.x / (p + 1), L = blockIdx.x % (p + 1);
    const int bi = blockIdx.y / nbk, bj = blockIdx.y % nbk;
    const bool diag = (bi == bj);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g4 = lane >> 2, t4 = lane & 3;
    const int wr = warp >> 2, wc = warp & 3;                       // 32 x 32 sub-block of the warp
    const int rows_i = min(kG3Block, m - bi * kG3Block), rows_j = min(kG3Block, m - bj * kG3Block);
    const int chunks_per_trial = (n + kG3Chunk - 1) / kG3Chunk;
    const int n_chunks = chunks_per_trial * P.trials;
    // rows beyond m and the 4 pad columns stay zero for the whole kernel
    for (int e = threadIdx.x; e < 2 * kG3Rows * kG3Ld; e += blockDim.x) reinterpret_cast<double*>(smem_raw)[e] = 0.0;
    if (threadIdx.x == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
    }
    __syncthreads();
    auto issue = [&](const int c) {          // warp 0: bulk copies of chunk c into buffer c & 1
        const int tr = c / chunks_per_trial, t0 = (c - tr * chunks_per_trial) * kG3Chunk;
        const int valid = min(kG3Chunk + kG3Halo, n - t0);       // samples of this chunk that exist (even: n and t0 are even)
        const double* xu = P.x + P.offsets[(size_t)w * P.trials + tr] + t0;
        double* dst = buf[c & 1];
        const int n_rows = rows_i + (diag ? 0 : rows_j);
        fence_proxy_async();
        if (lane == 0) mbar_expect_tx(&full_bar[c & 1], (unsigned)((size_t)n_rows * valid * sizeof(double)));
        __syncwarp();
        for (int r = lane; r < rows_i; r += 32)
            bulk_g2s(dst + (size_t)r * kG3Ld, xu + (size_t)(bi * kG3Block + r) * P.ch_stride, (unsigned)(valid * sizeof(double)), &full_bar[c & 1]);
        if (!diag)
            for (int r = lane; r < rows_j; r += 32)
                bulk_g2s(dst + (size_t)(kG3Block + r) * kG3Ld, xu + (size_t)(bj * kG3Block + r) * P.ch_stride, (unsigned)(valid * sizeof(double)),
                         &full_bar[c & 1]);
    };
    // 16-byte aligned rows for every epoch of this window?  (uniform per CTA)
    bool aligned = true;
    for (int tr = 0; tr < P.trials; ++tr) aligned = aligned && ((P.offsets[(size_t)w * P.trials + tr] & 1) == 0);
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    auto multiply = [&](const double* xb) {
        const double* pa = xb + (size_t)(32 * wr + g4) * kG3Ld + t4;
        const double* pb = xb + (size_t)((diag ? 0 : kG3Block) + 32 * wc + g4) * kG3Ld + t4 + L;
#pragma unroll 4
        for (int t = 0; t < kG3Chunk; t += 4) {
            double a[4], b[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                a[q] = pa[(size_t)q * 8 * kG3Ld + t];
                b[q] = pb[(size_t)q * 8 * kG3Ld + t];
            }
#pragma unroll
            for (int qa = 0; qa < 4; ++qa)
#pragma unroll
                for (int qb = 0; qb < 4; ++qb) dmma884_k3(acc[qa][qb][0], acc[qa][qb][1], a[qa], b[qb]);
        }
    };
    if (aligned) {
        if (warp == 0) {
            issue(0);
            if (n_chunks > 1) issue(1);
        }
        for (int c = 0; c < n_chunks; ++c) {
            mbar_wait(&full_bar[c & 1], (unsigned)((c >> 1) & 1));
            const int t0 = (c % chunks_per_trial) * kG3Chunk;
            const int valid = min(kG3Chunk + kG3Halo, n - t0);
            double* xb = buf[c & 1];
            if (valid < kG3Chunk + kG3Halo) {
                // last chunk of an epoch: the columns behind the end still hold the previous contents of this buffer -> zero them
                __syncthreads();
                const int nz = kG3Chunk + kG3Halo - valid;
                for (int e = threadIdx.x; e < kG3Rows * nz; e += blockDim.x) xb[(size_t)(e / nz) * kG3Ld + valid + e % nz] = 0.0;
                __syncthreads();
            }
            multiply(xb);
            __syncthreads();                                           // everyone is done with this buffer
            if (warp == 0 && c + 2 < n_chunks) issue(c + 2);
        }
    } else {
        // odd epoch offset: rows are only 8-byte aligned -> plain loads, one buffer, no overlap
        double* xb = buf[0];
        for (int c = 0; c < n_chunks; ++c) {
            const int tr = c / chunks_per_trial, t0 = (c - tr * chunks_per_trial) * kG3Chunk;
            const int valid = min(kG3Chunk + kG3Halo, n - t0);
            const double* xu = P.x + P.offsets[(size_t)w * P.trials + tr] + t0;
            __syncthreads();
            for (int e = threadIdx.x; e < kG3Rows * (kG3Chunk + kG3Halo); e += blockDim.x) {
                const int r = e / (kG3Chunk + kG3Halo), col = e - r * (kG3Chunk + kG3Halo);
                const bool second = r >= kG3Block;
                if (second && diag) continue;
                const int rr = second ? r - kG3Block : r;
                const int ch = (second ? bj : bi) * kG3Block + rr;
                xb[(size_t)r * kG3Ld + col] = (ch < m && col < valid) ? xu[(size_t)ch * P.ch_stride + col] : 0.0;
            }
            __syncthreads();
            multiply(xb);
        }
    }
    const double scale = 1.0 / ((double)n * (double)P.trials);
    double* Rl = P.R + ((size_t)w * (p + 1) + L) * m * m;
#pragma unroll
    for (int qa = 0; qa < 4; ++qa) {
        const int i = bi * kG3Block + 32 * wr + 8 * qa + g4;
#pragma unroll
        for (int qb = 0; qb < 4; ++qb) {
            const int j = bj * kG3Block + 32 * wc + 8 * qb + 2 * t4;
            if (i < m) {
                if (j < m) Rl[(size_t)i * m + j] = acc[qa][qb][0] * scale;
                if (j + 1 < m) Rl[(size_t)i * m + j + 1] = acc[qa][qb][1] * scale;
            }
        }
    }
}

static bool lagcov_gemm_ok(const K3Params& P) {
    // bulk copies need 16-byte aligned rows: even sample counts / strides here, even epoch offsets checked per window on the device
    // (a window with an odd offset is staged with plain loads instead)
    return P.m > kPadMax && P.p <= kG3MaxLag && (P.n & 1) == 0 && (P.ch_stride & 1) == 0 && (reinterpret_cast<uintptr_t>(P.x) & 15) == 0;
}

static int lagcov_mma_ld(int n, int p) {
    int ld = n + p + 4;                        // fragments read up to column (n - 4) + 3 + p
    ld += (4 - ld % 16 + 16) % 16;             // = 4 (mod 16) doubles: the 32 lanes of a fragment load hit 32 distinct banks
    return ld;
}

int launch_lagcov(const K3Params& P, cudaStream_t stream) {
    static const bool legacy = exp_env_int("HS_K3_LEGACY", 0) == 1;
    if (!legacy && lagcov_gemm_ok(P)) {
        const int nbk = (P.m + kG3Block - 1) / kG3Block;
        const bool one = nbk == 1;
        const size_t smem = one ? (size_t)2 * 128 * (64 + kG3Halo + 4) * sizeof(double) : (size_t)2 * 256 * (32 + kG3Halo + 4) * sizeof(double);
        auto kern = one ? lagcov_gemm_kernel<64, 128> : lagcov_gemm_kernel<32, 256>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(HS_ERR_CUDA, "lagcov: %s", cudaGetErrorString(e));
        dim3 grid(P.n_win * (P.p + 1), nbk * nbk);
        kern<<<grid, 512, smem, stream>>>(P);
        return check_launch("lagcov_gemm_kernel");
    }
    if (!legacy && P.m <= kPadMax) {
        const int ld = lagcov_mma_ld(P.n, P.p);
        const size_t smem_mma = (size_t)kPadMax * ld * sizeof(double);
        const int n_cols = ((P.m + 7) / 8) * (P.p + 1);
        if (smem_mma <= 220 * 1024 && n_cols <= 96) {
            const int kc = n_cols <= 48 ? 3 : 6;
            const int nwarps = (n_cols + kc - 1) / kc;
            auto kern = kc == 3 ? lagcov_mma_kernel<3> : lagcov_mma_kernel<6>;
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mma);
            if (e != cudaSuccess) return set_error(HS_ERR_CUDA, "lagcov: %s", cudaGetErrorString(e));
            // last, partial wave: its windows are split over several CTAs (see the kernel)
            int per_sm = 1;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nwarps * 32, smem_mma) != cudaSuccess || per_sm < 1) per_sm = 1;
            const int slots = device_sm_count() * per_sm;
            const int rem = P.n_win % slots;
            int n_parts = rem ? slots / rem : 1;
            if (n_parts > nwarps) n_parts = nwarps;
            const int wpp = (nwarps + n_parts - 1) / n_parts;
            n_parts = (nwarps + wpp - 1) / wpp;
            const int n_full = n_parts > 1 ? P.n_win - rem : P.n_win;
            const int grid = n_full + (P.n_win - n_full) * n_parts;
            kern<<<grid, nwarps * 32, smem_mma, stream>>>(P, ld, n_full, n_parts, wpp);
            return check_launch("lagcov_mma_kernel");
        }
    }
    int ng = P.p + 1;
    if (ng > 9) ng = 9;
    const size_t smem = ((size_t)(kK3Chunk + P.p) + kK3Chunk) * kK3Ld * sizeof(double);
    if (smem > 200 * 1024) return set_error(HS_ERR_UNSUPPORTED, "lagcov: model order %d too large for the staged panel", P.p);
    cudaError_t e = cudaFuncSetAttribute(lagcov_kernel<kTileMax>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(HS_ERR_CUDA, "lagcov: %s", cudaGetErrorString(e));
    lagcov_kernel<kTileMax><<<P.n_win, ng * 64, smem, stream>>>(P);
    return check_launch("lagcov_kernel");
}

// Block-Toeplitz Yule-Walker system exactly as count_corr returns it (mtmvar.py:65-76):
// G[a][b] = R(a-b) (a>b), R(b-a)^T (a<b), R(0) (a=b);  rhs = [R(1); ...; R(p)].
__global__ void toeplitz_assemble_kernel(const double* __restrict__ R, int m, int p, double* __restrict__ G, double* __restrict__ rhs) {
    const int w = blockIdx.y;
    const size_t mp = (size_t)m * p;
    const double* Rw = R + (size_t)w * (p + 1) * m * m;
    const size_t total = mp * mp + mp * m;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        if (e < mp * mp) {
            const size_t row = e / mp, col = e - row * mp;
            const int a = (int)(row / m), i = (int)(row % m), b = (int)(col / m), j = (int)(col % m);
            double v;
            if (a >= b) v = Rw[((size_t)(a - b) * m + i) * m + j];
            else v = Rw[((size_t)(b - a) * m + j) * m + i];
            G[(size_t)w * mp * mp + e] = v;
        } else {
            const size_t q = e - mp * mp;
            const size_t row = q / m, j = q - row * m;
            const int a = (int)(row / m), i = (int)(row % m);
            rhs[(size_t)w * mp * m + q] = Rw[((size_t)(a + 1) * m + i) * m + j];
        }
    }
}

int launch_toeplitz(const double* R, int n_win, int m, int p, double* G, double* rhs, cudaStream_t stream) {
    dim3 grid(64, n_win);
    toeplitz_assemble_kernel<<<grid, 256, 0, stream>>>(R, m, p, G, rhs);
    return check_launch("toeplitz_assemble_kernel");
}

// =====================================================================================
// K4  lwr_solve :  Levinson-Wiggins-Robinson (Whittle) recursion for the block-Toeplitz
//     Yule-Walker system that ar_coeff solves with a dense LU (mtmvar.py:116-122).
//     With Gamma(l) = R(l)^T:
//        D    = Gamma(k+1) - sum_{j=1..k} A_j Gamma(k+1-j)
//        A_{k+1} = D Vb^-1          B_{k+1} = D^T Vf^-1
//        A_j <- A_j - A_{k+1} B_{k+1-j}      B_j <- B_j - B_{k+1} A_{k+1-j}
//        Vf  <- Vf - A_{k+1} D^T    Vb  <- Vb - B_{k+1} D
//     A[i][j][k] = A_{k+1}[i][j],  V = Vf(p)  (= R0 - X rhs of mtmvar.py:119).
//     One CTA (4 tile groups) per window, persistent over windows; A_j/B_j live in an
//     L2-resident global scratch (double buffered per order), operands are staged as
//     k-major 40 x 40 panels in shared memory.
// =====================================================================================

#ifdef HS_EXPERIMENT
constexpr int kK4Groups = 4;
constexpr int kPanel = kPadMax * kPadMax;

template <int T>
__global__ void __launch_bounds__(kK4Groups * 64, 1) lwr_kernel(const K4Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* panels = reinterpret_cast<double*>(smem_raw);              // [NG][2][kPanel]
    double* D = panels + (size_t)kK4Groups * 2 * kPanel;                // D[q][c]  = Delta[q][c]
    double* DT = D + kPanel;                                            // DT[q][c] = Delta[c][q]
    double* KFT = DT + kPanel;                                          // KFT[q][i] = Kf[i][q]
    double* KBT = KFT + kPanel;
    GJScratch* gjs = reinterpret_cast<GJScratch*>(KBT + kPanel);
    const Group g = make_group();
    double* pA = panels + (size_t)g.gid * 2 * kPanel;
    double* pB = pA + kPanel;
    GJScratch* sh = gjs + g.gid;
    const int m = P.m, p = P.p;
    const size_t mm = (size_t)m * m;
    double* ws = P.ws + (size_t)blockIdx.x * 4 * p * mm;               // [par][A|B][p][mm]
    auto stA = [&](int par, int j) { return ws + ((size_t)(par * 2 + 0) * p + j) * mm; };   // j = 0-based index of A_{j+1}
    auto stB = [&](int par, int j) { return ws + ((size_t)(par * 2 + 1) * p + j) * mm; };

    // zero the shared panels once (padding stays zero / finite)
    for (int e = threadIdx.x; e < (kK4Groups * 2 + 4) * kPanel; e += blockDim.x) panels[e] = 0.0;

    for (int w = blockIdx.x; w < P.n_win; w += gridDim.x) {
        const double* Rw = P.R + (size_t)w * (p + 1) * mm;
        __syncthreads();
        double Vt[T][T];          // group 0: Vf, group 1: Vb   (both start as Gamma(0) = R(0)^T)
        double dummy[T][T];
        if (g.gid < 2) {
#pragma unroll
            for (int a = 0; a < T; ++a)
#pragma unroll
                for (int b = 0; b < T; ++b) {
                    const int i = g.tr + 8 * a, j = g.tc + 8 * b;
                    Vt[a][b] = (i < m && j < m) ? Rw[(size_t)j * m + i] : 0.0;
                }
        }
        for (int kk = 0; kk < p; ++kk) {
            const int cur = kk & 1, nxt = cur ^ 1;
            const bool last = (kk == p - 1);
            // ---- phase 1: partial Delta_g = - sum_{j in group} A_j Gamma(kk+1-j)
            {
                double acc[T][T];
#pragma unroll
                for (int a = 0; a < T; ++a)
#pragma unroll
                    for (int b = 0; b < T; ++b) acc[a][b] = 0.0;
                for (int j = g.gid; j < kk; j += kK4Groups) {        // A_{j+1} with Gamma(kk - j)
                    group_sync(g);
                    load_panel<true>(pA, stA(cur, j), m, m, g);                       // pA[q][i] = A[i][q]
                    load_panel<true>(pB, Rw + (size_t)(kk - j) * mm, m, m, g);        // pB[q][c] = R(l)[c][q] = Gamma(l)[q][c]
                    group_sync(g);
                    tile_mac<T, true>(acc, pA, kPadMax, pB, kPadMax, m, g);
                }
                group_sync(g);
                store_tile<T>(pA, kPadMax, acc, kPadMax, g);           // partial, row-major (full padded tile)
            }
            __syncthreads();
            // ---- phase 2: Delta, inverses
            if (g.gid < 2) {
                double dl[T][T];
#pragma unroll
                for (int a = 0; a < T; ++a)
#pragma unroll
                    for (int b = 0; b < T; ++b) {
                        const int i = g.tr + 8 * a, j = g.tc + 8 * b;
                        double v = (i < m && j < m) ? Rw[((size_t)(kk + 1) * m + j) * m + i] : 0.0;   // Gamma(kk+1)[i][j]
                        for (int q = 0; q < kK4Groups; ++q) v += panels[(size_t)q * 2 * kPanel + i * kPadMax + j];
                        dl[a][b] = v;
                    }
                if (g.gid == 0) {
                    store_tile<T>(D, kPadMax, dl, m, g);
                    store_tile_t<T>(DT, kPadMax, dl, m, g);
                }
                // invert own V (copy; padding -> identity) and publish it in own pB (row-major = k-major right operand)
                if (!(last && g.gid == 0)) {
                    double iv[T][T];
#pragma unroll
                    for (int a = 0; a < T; ++a)
#pragma unroll
                        for (int b = 0; b < T; ++b) {
                            const int i = g.tr + 8 * a, j = g.tc + 8 * b;
                            iv[a][b] = (i < m && j < m) ? Vt[a][b] : ((i == j) ? 1.0 : 0.0);
                        }
                    // residual covariances are symmetric positive definite: unpivoted elimination is stable
                    gj_inverse_static<T, false>(iv, dummy, m, g, sh);
                    bool finite = true;
#pragma unroll
                    for (int a = 0; a < T; ++a) {
                        const int i = g.tr + 8 * a;
#pragma unroll
                        for (int b = 0; b < T; ++b) {
                            const int j = g.tc + 8 * b;
                            if (i < m && j < m) {
                                pB[i * kPadMax + j] = iv[a][b];
                                finite = finite && (fabs(iv[a][b]) <= 1.79e308);
                            }
                        }
                    }
                    if (!finite) atomicOr(&P.status[w], 2);      // singular (or not positive definite) residual covariance
                }
            }
            __syncthreads();
            // ---- phase 3: reflection coefficients and residual covariances
            if (g.gid == 0) {
                // Kf = Delta * Vb^-1   (Vb^-1 is in group 1's pB)
                double kf[T][T];
#pragma unroll
                for (int a = 0; a < T; ++a)
#pragma unroll
                    for (int b = 0; b < T; ++b) kf[a][b] = 0.0;
                tile_mac<T, false>(kf, DT, kPadMax, panels + (size_t)(1 * 2 + 1) * kPanel, kPadMax, m, g);
                store_tile_t<T>(KFT, kPadMax, kf, m, g);
                store_tile<T>(stA(nxt, kk), m, kf, m, g);
                group_sync(g);
                tile_mac<T, true>(Vt, KFT, kPadMax, DT, kPadMax, m, g);      // Vf -= Kf * Delta^T
                if (P.Vall) store_tile<T>(P.Vall + ((size_t)w * p + kk) * mm, m, Vt, m, g);
            } else if (g.gid == 1 && !last) {
                // Kb = Delta^T * Vf^-1  (Vf^-1 is in group 0's pB)
                double kb[T][T];
#pragma unroll
                for (int a = 0; a < T; ++a)
#pragma unroll
                    for (int b = 0; b < T; ++b) kb[a][b] = 0.0;
                tile_mac<T, false>(kb, D, kPadMax, panels + (size_t)(0 * 2 + 1) * kPanel, kPadMax, m, g);
                store_tile_t<T>(KBT, kPadMax, kb, m, g);
                store_tile<T>(stB(nxt, kk), m, kb, m, g);
                group_sync(g);
                tile_mac<T, true>(Vt, KBT, kPadMax, D, kPadMax, m, g);       // Vb -= Kb * Delta
            }
            __syncthreads();
            // ---- phase 4: order update of A_j (and B_j unless this is the last order)
            const int n_items = last ? kk : 2 * kk;
            for (int it = g.gid; it < n_items; it += kK4Groups) {
                const bool isA = it < kk;
                const int j = isA ? it : it - kk;             // 0-based: A_{j+1} pairs with B_{kk-j}
                double acc[T][T];
                const double* own = isA ? stA(cur, j) : stB(cur, j);
                const double* other = isA ? stB(cur, kk - 1 - j) : stA(cur, kk - 1 - j);
                group_sync(g);
                load_panel<false>(pB, other, m, m, g);
                load_tile<T>(acc, own, m, m, g);
                group_sync(g);
                tile_mac<T, true>(acc, isA ? KFT : KBT, kPadMax, pB, kPadMax, m, g);
                store_tile<T>(isA ? stA(nxt, j) : stB(nxt, j), m, acc, m, g);
            }
            __syncthreads();
        }
        // ---- outputs: A[w][i][j][k] = A_{k+1}[i][j]
        {
            const int fin = p & 1;
            double* Aw = P.A + (size_t)w * mm * p;
            for (size_t e = threadIdx.x; e < mm * p; e += blockDim.x) {
                const size_t ij = e / p;
                const int k = (int)(e - ij * p);
                Aw[e] = stA(fin, k)[ij];
            }
            if (g.gid == 0) store_tile<T>(P.V + (size_t)w * mm, m, Vt, m, g);
        }
    }
}

#endif  // HS_EXPERIMENT (legacy 4-group K4 kernel)
// -------------------------------------------------------------------------------------
// K4, one-group variant (the one that runs): ONE tile group (64 threads) per window and FIVE
// CTAs per SM.  The recursion of a window is a sequential chain of ~107 small GEMMs and 15
// SPD inverses; the 4-group kernel above spreads that chain over 256 threads and leaves half of
// them idle in the inverse / reflection phases, with one window per SM in flight (599 windows on
// 148 SMs = 5 rounds).  Here every window runs the whole chain on its own 64 threads, five
// windows share an SM (740 slots >= 599 windows: one round) and hide each other's latencies.
// Three shared-memory panels with an ODD row stride (41 doubles: row and column fragment reads are
// both bank-conflict free) rotate through the operand roles, so no transposed copies (D^T, Kf^T,
// Kb^T) are ever materialised:
//      phase 1   P0 = A_j, P1 = R(kk-j)            Delta -= A_j Gamma(kk-j)        -> P2 = Delta
//      phase 3   P1 = Vb^-1:  Kf = Delta Vb^-1 -> P0;   P1 = Vf^-1:  Kb = Delta^T Vf^-1 (regs)
//                Vf -= Kf Delta^T;   P1 = Kb;   Vb -= Kb Delta
//      phase 4   P2 = B_{kk-1-j} / A_{kk-1-j}      A_j -= Kf B_{kk-1-j},  B_j -= Kb A_{kk-1-j}
// -------------------------------------------------------------------------------------
constexpr int kK4Ld = kPadMax + 1;              // 41
constexpr int kK4Panel = kPadMax * kK4Ld;       // doubles per panel
constexpr int kK4PerSM = 5;

// acc[a][b] +=/-= sum_k  L[i = tr+8a][k] * Rt[k][j = tc+8b];  TRA: Pa holds L row-major (else k-major), TRB: Pb holds Rt^T
template <int T, bool SUB, bool TRA, bool TRB>
__device__ __forceinline__ void tile_mac_x(double (&acc)[T][T], const double* __restrict__ Pa, const double* __restrict__ Pb,
                                           const int depth, const Group& g) {
    const double* pa = TRA ? Pa + g.tr * kK4Ld : Pa + g.tr;
    const double* pb = TRB ? Pb + g.tc * kK4Ld : Pb + g.tc;
#pragma unroll 4
    for (int k = 0; k < depth; ++k) {
        double av[T], bv[T];
#pragma unroll
        for (int a = 0; a < T; ++a) av[a] = TRA ? pa[8 * a * kK4Ld] : pa[8 * a];
#pragma unroll
        for (int b = 0; b < T; ++b) bv[b] = TRB ? pb[8 * b * kK4Ld] : pb[8 * b];
#pragma unroll
        for (int a = 0; a < T; ++a)
#pragma unroll
            for (int b = 0; b < T; ++b) acc[a][b] = fma(SUB ? -av[a] : av[a], bv[b], acc[a][b]);
        pa += TRA ? 1 : kK4Ld;
        pb += TRB ? 1 : kK4Ld;
    }
}

// A row-major m x m matrix in global memory -> registers (coalesced: element e = l64 + 64 u, all kK4Pre loads in flight at
// once) -> shared panel dst[r*41 + c].  Fetch and commit are separate so that the loads of BOTH operands of a product are
// issued before the first one is waited for.
constexpr int kK4Pre = (kPadMax * kPadMax + 63) / 64;      // 25
__device__ __forceinline__ void fetch_rowmajor(double (&v)[kK4Pre], const double* __restrict__ src, const int total, const int l64) {
#pragma unroll
    for (int u = 0; u < kK4Pre; ++u) v[u] = (l64 + 64 * u < total) ? src[l64 + 64 * u] : 0.0;
}
__device__ __forceinline__ void commit_rowmajor(double* __restrict__ dst, const double (&v)[kK4Pre], const int m, const int l64) {
    const int total = m * m;
    int r = l64 / m, c = l64 - r * m;
    const int dr = 64 / m, dc = 64 - dr * m;
#pragma unroll
    for (int u = 0; u < kK4Pre; ++u) {
        if (l64 + 64 * u < total) dst[r * kK4Ld + c] = v[u];
        r += dr;
        c += dc;
        if (c >= m) { c -= m; ++r; }
    }
}

// Variants measured and dropped (599 windows): operand panels software-pipelined through registers with V_f / V_b in
// the scratch (1.68 ms: ptxas sinks the loads to their use under the 168-register cap of 5 CTAs/SM; forcing their issue
// with 200 registers costs the fifth CTA: 2.2 ms).  This form: 1.43 ms (4-group kernel: 1.71 ms).
template <int T>
__global__ void __launch_bounds__(64, kK4PerSM) lwr1_kernel(const K4Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* P0 = reinterpret_cast<double*>(smem_raw);
    double* P1 = P0 + kK4Panel;
    double* P2 = P1 + kK4Panel;
    GJScratch* sh = reinterpret_cast<GJScratch*>(P2 + kK4Panel);
    Group g = make_group();
    g.gid = 0;
    g.bar = 1;
    const int m = P.m, p = P.p;
    const int mmi = m * m;
    const size_t mm = (size_t)mmi;
    double* ws = P.ws + (size_t)blockIdx.x * (4 * p + 2) * mm;         // [par][A|B][p][mm], Vf, Vb
    auto stA = [&](int par, int j) { return ws + ((size_t)(par * 2 + 0) * p + j) * mm; };   // j = 0-based index of A_{j+1}
    auto stB = [&](int par, int j) { return ws + ((size_t)(par * 2 + 1) * p + j) * mm; };
    // The residual covariances live in the scratch between the orders (2 tile loads + 2 stores per order): keeping them
    // in registers would pin 100 of the 168 registers through phases 1 and 4, where they buy 50 loads in flight instead.
    double* gVf = ws + (size_t)4 * p * mm;
    double* gVb = gVf + mm;

    for (int e = threadIdx.x; e < 3 * kK4Panel; e += 64) P0[e] = 0.0;      // padding rows / columns stay zero for good

    for (int w = blockIdx.x; w < P.n_win; w += gridDim.x) {
        const double* Rw = P.R + (size_t)w * (p + 1) * mm;
        double dummy[T][T];
        {
            double v0[T][T];
#pragma unroll
            for (int a = 0; a < T; ++a)
#pragma unroll
                for (int b = 0; b < T; ++b) {
                    const int i = g.tr + 8 * a, j = g.tc + 8 * b;
                    v0[a][b] = (i < m && j < m) ? Rw[(size_t)j * m + i] : 0.0;      // Gamma(0) = R(0)^T
                }
            store_tile<T>(gVf, m, v0, m, g);
            store_tile<T>(gVb, m, v0, m, g);
        }
        for (int kk = 0; kk < p; ++kk) {
            const int cur = kk & 1, nxt = cur ^ 1;
            const bool last = (kk == p - 1);
            // ---- phase 1: Delta = Gamma(kk+1) - sum_j A_{j+1} Gamma(kk-j)
            {
                double acc[T][T];
#pragma unroll
                for (int a = 0; a < T; ++a)
#pragma unroll
                    for (int b = 0; b < T; ++b) {
                        const int i = g.tr + 8 * a, j = g.tc + 8 * b;
                        acc[a][b] = (i < m && j < m) ? Rw[((size_t)(kk + 1) * m + j) * m + i] : 0.0;
                    }
                for (int j = 0; j < kk; ++j) {
                    double va[kK4Pre], vr[kK4Pre];
                    fetch_rowmajor(va, stA(cur, j), mmi, g.l64);                       // A_{j+1}[i][q]
                    fetch_rowmajor(vr, Rw + (size_t)(kk - j) * mm, mmi, g.l64);        // R(l)[c][q] = Gamma(l)[q][c]
                    __syncthreads();                                                    // the previous product is done with P0 / P1
                    commit_rowmajor(P0, va, m, g.l64);
                    commit_rowmajor(P1, vr, m, g.l64);
                    __syncthreads();
                    tile_mac_x<T, true, true, true>(acc, P0, P1, m, g);
                }
                __syncthreads();
                store_tile<T>(P2, kK4Ld, acc, m, g);                                    // Delta, row-major
            }
            // ---- phases 2+3: Kf = Delta Vb^-1, Kb = Delta^T Vf^-1, residual covariances
            {
                double iv[T][T];
                bool finite = true;
                load_tile<T>(iv, gVb, m, m, g);
#pragma unroll
                for (int a = 0; a < T; ++a)
#pragma unroll
                    for (int b = 0; b < T; ++b) {
                        const int i = g.tr + 8 * a, j = g.tc + 8 * b;
                        if (!(i < m && j < m)) iv[a][b] = (i == j) ? 1.0 : 0.0;
                    }
                // residual covariances are symmetric positive definite: unpivoted elimination is stable
                gj_inverse_static<T, false>(iv, dummy, m, g, sh);
#pragma unroll
                for (int a = 0; a < T; ++a)
#pragma unroll
                    for (int b = 0; b < T; ++b) {
                        const int i = g.tr + 8 * a, j = g.tc + 8 * b;
                        if (i < m && j < m) {
                            P1[i * kK4Ld + j] = iv[a][b];
                            finite = finite && (fabs(iv[a][b]) <= 1.79e308);
                        }
                    }
                __syncthreads();
#pragma unroll
                for (int a = 0; a < T; ++a)
#pragma unroll
                    for (int b = 0; b < T; ++b) iv[a][b] = 0.0;
                tile_mac_x<T, false, true, false>(iv, P2, P1, m, g);                    // Kf = sum_q Delta[i][q] Vbinv[q][j]
                store_tile<T>(P0, kK4Ld, iv, m, g);                                     // (phase 1 is done with P0)
                store_tile<T>(stA(nxt, kk), m, iv, m, g);
                if (!last) {
                    load_tile<T>(iv, gVf, m, m, g);
#pragma unroll
                    for (int a = 0; a < T; ++a)
#pragma unroll
                        for (int b = 0; b < T; ++b) {
                            const int i = g.tr + 8 * a, j = g.tc + 8 * b;
                            if (!(i < m && j < m)) iv[a][b] = (i == j) ? 1.0 : 0.0;
                        }
                    gj_inverse_static<T, false>(iv, dummy, m, g, sh);
                    __syncthreads();                                                     // every thread is done reading Vb^-1
#pragma unroll
                    for (int a = 0; a < T; ++a)
#pragma unroll
                        for (int b = 0; b < T; ++b) {
                            const int i = g.tr + 8 * a, j = g.tc + 8 * b;
                            if (i < m && j < m) {
                                P1[i * kK4Ld + j] = iv[a][b];
                                finite = finite && (fabs(iv[a][b]) <= 1.79e308);
                            }
                        }
                    __syncthreads();
#pragma unroll
                    for (int a = 0; a < T; ++a)
#pragma unroll
                        for (int b = 0; b < T; ++b) iv[a][b] = 0.0;
                    tile_mac_x<T, false, false, false>(iv, P2, P1, m, g);               // Kb = sum_q Delta[q][i] Vfinv[q][j]
                    store_tile<T>(stB(nxt, kk), m, iv, m, g);
                    __syncthreads();                                                     // every thread is done reading Vf^-1
                    store_tile<T>(P1, kK4Ld, iv, m, g);
                    load_tile<T>(iv, gVb, m, m, g);
                    __syncthreads();
                    tile_mac_x<T, true, true, false>(iv, P1, P2, m, g);                 // Vb -= Kb Delta
                    store_tile<T>(gVb, m, iv, m, g);
                } else {
                    __syncthreads();
                }
                load_tile<T>(iv, gVf, m, m, g);
                tile_mac_x<T, true, true, true>(iv, P0, P2, m, g);                      // Vf -= Kf Delta^T   (P0 written before the barriers above)
                store_tile<T>(gVf, m, iv, m, g);
                if (!finite) atomicOr(&P.status[w], 2);      // singular (or not positive definite) residual covariance
                if (P.Vall) store_tile<T>(P.Vall + ((size_t)w * p + kk) * mm, m, iv, m, g);
                if (last) store_tile<T>(P.V + (size_t)w * mm, m, iv, m, g);
            }
            // ---- phase 4: order update  A_j -= Kf B_{kk-1-j}  (and  B_j -= Kb A_{kk-1-j}  unless this is the last order)
            for (int it = 0; it < (last ? kk : 2 * kk); ++it) {
                const bool isA = it < kk;
                const int j = isA ? it : it - kk;
                const double* own = isA ? stA(cur, j) : stB(cur, j);
                const double* other = isA ? stB(cur, kk - 1 - j) : stA(cur, kk - 1 - j);
                double acc[T][T];
                double vo[kK4Pre];
                fetch_rowmajor(vo, other, mmi, g.l64);
                load_tile<T>(acc, own, m, m, g);
                __syncthreads();                                                         // P2 (Delta / previous operand) is free
                commit_rowmajor(P2, vo, m, g.l64);
                __syncthreads();
                if (isA) tile_mac_x<T, true, true, false>(acc, P0, P2, m, g);
                else tile_mac_x<T, true, true, false>(acc, P1, P2, m, g);
                store_tile<T>(isA ? stA(nxt, j) : stB(nxt, j), m, acc, m, g);
            }
            __syncthreads();
        }
        // ---- outputs: A[w][i][j][k] = A_{k+1}[i][j]
        {
            const int fin = p & 1;
            double* Aw = P.A + (size_t)w * mm * p;
            const int total = mmi * p;
            for (int e0 = threadIdx.x; e0 < total; e0 += 8 * 64) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e = e0 + 64 * u;
                    const int ij = e / p, k = e - ij * p;
                    v[u] = (e < total) ? stA(fin, k)[ij] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (e0 + 64 * u < total) Aw[e0 + 64 * u] = v[u];
            }
        }
        __syncthreads();
    }
}


// -------------------------------------------------------------------------------------
// K4 on the FP64 tensor pipe (the one that runs): the recursion and the panel roles of lwr1_kernel, with
//   * every product as mma.sync.m8n8k4.f64: lwr1_kernel issues 1000 DFMAs + 400 LDS per thread and product; a DMMA does the work of 8
//     DFMAs per lane from two fragment loads, so a product is 130 DMMAs + 80 LDS per lane.  The 25 accumulator tiles of a 40 x 40
//     product are split 13 / 12 between the two warps of the window (tile order row-major: warp 0 rows 0, 1 and three tiles of row 2;
//     warp 1 the rest): per k-step 3 A fragments + 5 B fragments for 13 DMMAs;
//   * panels (P0 / P1 / P2 rotate through A_j, Gamma, Delta, V^-1, Kf, Kb) with a row stride of 44 doubles (= 12 mod 16): a fragment is
//     read along a row or along a column of the panel -- (ld g4 + t4) or (ld t4 + g4) mod 16 distinct over a half-warp -- so no
//     transposed copies exist, as before;
//   * the data movement the profiler showed to dominate lwr1_kernel taken out of the warps: scratch matrices (A_j, B_j, V_f, V_b) are
//     zero-padded 40 x 40, go global -> panel by cp.async and global <-> accumulator as unconditional 16-byte accesses; the order update
//     is done in place (A_ja and B_jb, ja + jb = kk - 1, read only each other's old values); the final A_k go straight to the output;
//   * accumulators in the m8n8k4 layout (lane (g4, t4): row 8a + g4, columns 8b + 2 t4, +1); the two SPD inverses of an order run side
//     by side, one warp each, on the whole matrix in that warp's registers (k4_spd_inverse below).
// 599 cfg2 windows: 0.68 ms (lwr1_kernel: 1.08 ms); dram 1.2 GB per launch.
// -------------------------------------------------------------------------------------
constexpr int kK4Ld2 = kPadMax + 4;             // 44
constexpr int kK4Panel2 = kPadMax * kK4Ld2;

template <int W>
struct K4Split {
    static constexpr int first = W ? 13 : 0, count = W ? 12 : 13, row0 = W ? 2 : 0;
};

// acc (tiles of warp W) +=/-= sum_k L[i][k] Rt[k][j];  TRA: Pa holds L row-major (else k-major);  TRB: Pb holds Rt^T
template <int W, bool SUB, bool TRA, bool TRB>
__device__ __forceinline__ void k4_mma(double (&acc)[13][2], const double* __restrict__ Pa, const double* __restrict__ Pb, const int ksteps,
                                       const int g4, const int t4) {
    using S = K4Split<W>;
    const double* pa = TRA ? Pa + g4 * kK4Ld2 + t4 : Pa + t4 * kK4Ld2 + g4;
    const double* pb = TRB ? Pb + g4 * kK4Ld2 + t4 : Pb + t4 * kK4Ld2 + g4;
#pragma unroll 2
    for (int s = 0; s < ksteps; ++s) {
        double av[3], bv[5];
#pragma unroll
        for (int r = 0; r < 3; ++r) av[r] = TRA ? pa[8 * (S::row0 + r) * kK4Ld2] : pa[8 * (S::row0 + r)];
#pragma unroll
        for (int b = 0; b < 5; ++b) bv[b] = TRB ? pb[8 * b * kK4Ld2] : pb[8 * b];
#pragma unroll
        for (int r = 0; r < 3; ++r) av[r] = SUB ? -av[r] : av[r];
#pragma unroll
        for (int n = 0; n < S::count; ++n) dmma884_k3(acc[n][0], acc[n][1], av[(S::first + n) / 5 - S::row0], bv[(S::first + n) % 5]);
        pa += TRA ? 4 : 4 * kK4Ld2;
        pb += TRB ? 4 : 4 * kK4Ld2;
    }
}

template <int W, bool TRANS>      // acc <- src (row-major m x m, leading dimension ld; TRANS: acc[i][j] = src[j][i]), zero outside m x m
__device__ __forceinline__ void k4_load(double (&acc)[13][2], const double* __restrict__ src, const int ld, const int m, const int g4, const int t4) {
    using S = K4Split<W>;
#pragma unroll
    for (int n = 0; n < S::count; ++n) {
        const int i = 8 * ((S::first + n) / 5) + g4, j = 8 * ((S::first + n) % 5) + 2 * t4;
#pragma unroll
        for (int c = 0; c < 2; ++c) acc[n][c] = (i < m && j + c < m) ? (TRANS ? src[(size_t)(j + c) * ld + i] : src[(size_t)i * ld + j + c]) : 0.0;
    }
}

template <int W>                  // bounded store to a packed m x m array (the outputs V, Vall)
__device__ __forceinline__ void k4_store(double* __restrict__ dst, const int ld, const double (&acc)[13][2], const int m, const int g4, const int t4) {
    using S = K4Split<W>;
#pragma unroll
    for (int n = 0; n < S::count; ++n) {
        const int i = 8 * ((S::first + n) / 5) + g4, j = 8 * ((S::first + n) % 5) + 2 * t4;
#pragma unroll
        for (int c = 0; c < 2; ++c)
            if (i < m && j + c < m) dst[(size_t)i * ld + j + c] = acc[n][c];
    }
}

// Scratch matrices and panels are 40 x 40 with zero padding (every producer writes whole tiles, the padding of an accumulator is
// exactly zero), so they move as unconditional 16-byte accesses: LD = 40 for the global scratch, 44 for a shared-memory panel.
template <int W, int LD>
__device__ __forceinline__ void k4_load_pad(double (&acc)[13][2], const double* __restrict__ src, const int g4, const int t4) {
    using S = K4Split<W>;
#pragma unroll
    for (int n = 0; n < S::count; ++n) {
        const double2 v = *reinterpret_cast<const double2*>(src + (8 * ((S::first + n) / 5) + g4) * LD + 8 * ((S::first + n) % 5) + 2 * t4);
        acc[n][0] = v.x;
        acc[n][1] = v.y;
    }
}
template <int W, int LD>
__device__ __forceinline__ void k4_store_pad(double* __restrict__ dst, const double (&acc)[13][2], const int g4, const int t4) {
    using S = K4Split<W>;
#pragma unroll
    for (int n = 0; n < S::count; ++n)
        *reinterpret_cast<double2*>(dst + (8 * ((S::first + n) / 5) + g4) * LD + 8 * ((S::first + n) % 5) + 2 * t4) = make_double2(acc[n][0], acc[n][1]);
}
// A[w][i][j][k] = A_{k+1}[i][j]: the final coefficient matrices go straight to the output (lag fastest, stride p)
template <int W>
__device__ __forceinline__ void k4_store_lag(double* __restrict__ Aw, const int p, const int k, const double (&acc)[13][2], const int m, const int g4, const int t4) {
    using S = K4Split<W>;
#pragma unroll
    for (int n = 0; n < S::count; ++n) {
        const int i = 8 * ((S::first + n) / 5) + g4, j = 8 * ((S::first + n) % 5) + 2 * t4;
#pragma unroll
        for (int c = 0; c < 2; ++c)
            if (i < m && j + c < m) Aw[((size_t)i * m + j + c) * p + k] = acc[n][c];
    }
}

constexpr int kK4Sd = kPadMax * kPadMax;        // doubles per scratch matrix

// global -> panel without a detour through registers: cp.async (L2 only, so stores of other threads of the CTA made visible by a
// barrier are seen), CH doubles per copy.  src rows are `row_len` doubles long and contiguous (row_len = 40: a scratch matrix;
// row_len = m: a packed R(l) from K3), panel rows are kK4Ld2 apart.  Caller: k4_async_wait(); __syncthreads().
template <int CH>
__device__ __forceinline__ void k4_async_rows(double* __restrict__ panel, const double* __restrict__ src, const int row_len, const int n_rows,
                                              const int l64) {
    const int rc = row_len / CH;                 // copies per row
    const int total = rc * n_rows;
    int r = l64 / rc, c = l64 - r * rc;
    const int dr = 64 / rc, dc = 64 - dr * rc;
    const unsigned base = (unsigned)__cvta_generic_to_shared(panel);
    for (int e = l64; e < total; e += 64) {
        const unsigned dst = base + (unsigned)((r * kK4Ld2 + c * CH) * 8);
        const double* sp = src + (size_t)e * CH;
        if (CH == 2) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(sp) : "memory");
        else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(sp) : "memory");
        r += dr;
        c += dc;
        if (c >= rc) { c -= rc; ++r; }
    }
}
__device__ __forceinline__ void k4_async_wait() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }

// ---- SPD inverse of a residual covariance by ONE warp: the 40 x 40 matrix lives in the warp's registers in the m8n8k4 accumulator
//      layout (50 doubles per lane), in-place Gauss-Jordan with 1 x 1 pivots and no pivoting (residual covariances are symmetric
//      positive definite, so this is backward stable -- same arithmetic as gj_inverse_static, which it replaces).  Per step the
//      owners publish the pivot row and column to shared memory (double buffered by step parity: one __syncwarp per step, no CTA
//      barrier), every lane takes its 5 multipliers and 10 row entries back and does its 50 DFMAs.  The two inverses of an order
//      (V_b^-1 by warp 0, V_f^-1 by warp 1) run side by side.  Scratch: the start of the panel that receives the result.
//      Measured and dropped: block Gauss-Jordan with 4 x 4 pivot blocks on the tensor pipe (explicit D^-1 by cofactors, L = C D^-1 and the
//      rank-4 update as DMMAs, like transfer_mma.cu): 0.63 instead of 0.70 ms for the whole stage, but the explicit inverse of a pivot block puts
//      cond(D) into the backward error -- A came out 30 x further from the reference (6e-9 .. 1e-8 instead of 3e-10 on the cfg2
//      goldens, 8e-7 on the cond 1.7e9 stress fixture).  K5's complex A(f) blocks are well conditioned and verified; these are not.
struct K4InvBuf {
    double row[2][kPadMax];                     // [step parity][j]  pivot row
    double col[2][kPadMax];                     // [step parity][i]  pivot column
};
static_assert(sizeof(K4InvBuf) <= kK4Panel2 * sizeof(double), "inverse scratch must fit the destination panel");

template <int b>      // the pivots k = 8 b .. 8 b + 7
__device__ __forceinline__ void k4_gj_steps(double (&c)[5][5][2], const int m, K4InvBuf* __restrict__ ib, int& step, const int g4, const int t4) {
#pragma unroll 1
    for (int kc = 0; kc < 8; ++kc) {
        const int k = 8 * b + kc;
        if (k >= m) break;
        const int par = step & 1;
        ++step;
        const bool odd = kc & 1;
        const bool own_row = (g4 == kc), own_col = (t4 == (kc >> 1));
        double* row = ib->row[par];
        double* col = ib->col[par];
        if (own_row) {
#pragma unroll
            for (int tb = 0; tb < 5; ++tb) *reinterpret_cast<double2*>(&row[8 * tb + 2 * t4]) = make_double2(c[b][tb][0], c[b][tb][1]);
        }
        if (own_col) {
#pragma unroll
            for (int ta = 0; ta < 5; ++ta) col[8 * ta + g4] = odd ? c[ta][b][1] : c[ta][b][0];
        }
        __syncwarp();
        const double ip = rcp_newton(row[k]);
        double ci[5], cfix[5];
        double2 r[5];
#pragma unroll
        for (int ta = 0; ta < 5; ++ta) cfix[ta] = ci[ta] = col[8 * ta + g4] * ip;
#pragma unroll
        for (int tb = 0; tb < 5; ++tb) r[tb] = *reinterpret_cast<const double2*>(&row[8 * tb + 2 * t4]);
        if (own_col) {                               // the rank-1 term must not touch column k
            if (odd) r[b].y = 0.0;
            else r[b].x = 0.0;
        }
        if (own_row) ci[b] = 0.0;                    // ... nor row k
#pragma unroll
        for (int ta = 0; ta < 5; ++ta)
#pragma unroll
            for (int tb = 0; tb < 5; ++tb) {
                c[ta][tb][0] = fma(-ci[ta], r[tb].x, c[ta][tb][0]);
                c[ta][tb][1] = fma(-ci[ta], r[tb].y, c[ta][tb][1]);
            }
        if (own_row) {                               // row k := row k / p
#pragma unroll
            for (int tb = 0; tb < 5; ++tb) {
                c[b][tb][0] = r[tb].x * ip;
                c[b][tb][1] = r[tb].y * ip;
            }
        }
        if (own_col) {                               // column k := -a_ik / p, 1 / p on the diagonal
#pragma unroll
            for (int ta = 0; ta < 5; ++ta) {
                const double v = (ta == b && own_row) ? ip : -cfix[ta];
                if (odd) c[ta][b][1] = v;
                else c[ta][b][0] = v;
            }
        }
    }
}

template <int b>
__device__ __forceinline__ void k4_inv_all(double (&c)[5][5][2], const int m, K4InvBuf* ib, int& step, const int g4, const int t4) {
    if constexpr (b < 5) {
        k4_gj_steps<b>(c, m, ib, step, g4, t4);
        k4_inv_all<b + 1>(c, m, ib, step, g4, t4);
    }
}

// panel (row stride kK4Ld2) <- inverse of the padded scratch matrix src; returns false if the result is not finite
__device__ __forceinline__ bool k4_spd_inverse(const double* __restrict__ src, double* __restrict__ panel, const int m, const int lane,
                                               const int g4, const int t4) {
    double c[5][5][2];
#pragma unroll
    for (int ta = 0; ta < 5; ++ta)
#pragma unroll
        for (int tb = 0; tb < 5; ++tb) {
            const int i = 8 * ta + g4, j = 8 * tb + 2 * t4;
            const double2 v = *reinterpret_cast<const double2*>(src + i * kPadMax + j);
            c[ta][tb][0] = (i >= m && i == j) ? 1.0 : v.x;          // identity on the padding diagonal
            c[ta][tb][1] = (i >= m && i == j + 1) ? 1.0 : v.y;
        }
    int step = 0;
    k4_inv_all<0>(c, m, reinterpret_cast<K4InvBuf*>(panel), step, g4, t4);
    __syncwarp();                                  // the scratch is dead: the result may overwrite it
    bool finite = true;
#pragma unroll
    for (int ta = 0; ta < 5; ++ta)
#pragma unroll
        for (int tb = 0; tb < 5; ++tb) {
            const int i = 8 * ta + g4, j = 8 * tb + 2 * t4;
            const double x0 = (i < m && j < m) ? c[ta][tb][0] : 0.0, x1 = (i < m && j + 1 < m) ? c[ta][tb][1] : 0.0;
            finite = finite && (fabs(x0) <= 1.79e308) && (fabs(x1) <= 1.79e308);
            *reinterpret_cast<double2*>(panel + i * kK4Ld2 + j) = make_double2(x0, x1);
        }
    return finite;
}

// one warp's share of the recursion of the windows w = blockIdx.x, + gridDim.x, ...
template <int W>
__device__ __forceinline__ void lwr2_body(const K4Params& P, double* P0, double* P1, double* P2) {
    const int lane = threadIdx.x & 31, g4 = lane >> 2, t4 = lane & 3, l64 = threadIdx.x & 63;
    const int m = P.m, p = P.p;
    const size_t mm = (size_t)m * m;
    const int ksteps = (m + 3) >> 2;
    const bool r16 = ((m & 1) == 0) && ((reinterpret_cast<size_t>(P.R) & 15) == 0);       // rows of R(l) are whole 16-byte chunks
    double* ws = P.ws + (size_t)blockIdx.x * (2 * p + 2) * kK4Sd;       // [A|B][p] padded matrices (updated in place), Vf, Vb
    auto stA = [&](int j) { return ws + (size_t)j * kK4Sd; };            // j = 0-based index of A_{j+1}
    auto stB = [&](int j) { return ws + ((size_t)p + j) * kK4Sd; };
    double* gVf = ws + (size_t)2 * p * kK4Sd;
    double* gVb = gVf + kK4Sd;

    for (int w = blockIdx.x; w < P.n_win; w += gridDim.x) {
        const double* Rw = P.R + (size_t)w * (p + 1) * mm;
        double* Aw = P.A + (size_t)w * mm * p;
        double acc[13][2];
        k4_load<W, true>(acc, Rw, m, m, g4, t4);                                       // Gamma(0) = R(0)^T
        k4_store_pad<W, kPadMax>(gVf, acc, g4, t4);
        k4_store_pad<W, kPadMax>(gVb, acc, g4, t4);
        __syncthreads();
        for (int kk = 0; kk < p; ++kk) {
            const bool last = (kk == p - 1);
            // ---- phase 1: Delta = Gamma(kk+1) - sum_j A_{j+1} Gamma(kk-j)
            k4_load<W, true>(acc, Rw + (size_t)(kk + 1) * mm, m, m, g4, t4);
            for (int j = 0; j < kk; ++j) {
                __syncthreads();                                                        // the previous product is done with P0 / P1
                k4_async_rows<2>(P0, stA(j), kPadMax, m, l64);                    // A_{j+1}[i][q]
                if (r16) k4_async_rows<2>(P1, Rw + (size_t)(kk - j) * mm, m, m, l64);        // R(l)[c][q] = Gamma(l)[q][c]
                else k4_async_rows<1>(P1, Rw + (size_t)(kk - j) * mm, m, m, l64);
                k4_async_wait();
                __syncthreads();
                k4_mma<W, true, true, true>(acc, P0, P1, ksteps, g4, t4);
            }
            __syncthreads();
            k4_store_pad<W, kK4Ld2>(P2, acc, g4, t4);                                   // Delta, row-major
            // ---- phases 2+3: Vb^-1 -> P1 (warp 0) next to Vf^-1 -> P0 (warp 1);  Kf = Delta Vb^-1, Kb = Delta^T Vf^-1;  residual covariances
            {
                bool finite = true;
                if (W == 0) finite = k4_spd_inverse(gVb, P1, m, lane, g4, t4);
                else if (!last) finite = k4_spd_inverse(gVf, P0, m, lane, g4, t4);
                __syncthreads();
                double acc2[13][2];
#pragma unroll
                for (int n = 0; n < 13; ++n) acc[n][0] = acc[n][1] = acc2[n][0] = acc2[n][1] = 0.0;
                k4_mma<W, false, true, false>(acc, P2, P1, ksteps, g4, t4);             // Kf = sum_q Delta[i][q] Vbinv[q][j]
                if (!last) k4_mma<W, false, false, false>(acc2, P2, P0, ksteps, g4, t4);      // Kb = sum_q Delta[q][i] Vfinv[q][j]
                __syncthreads();                                                         // every thread is done reading the inverses
                k4_store_pad<W, kK4Ld2>(P0, acc, g4, t4);
                if (last) k4_store_lag<W>(Aw, p, kk, acc, m, g4, t4);                   // A_p = Kf
                else k4_store_pad<W, kPadMax>(stA(kk), acc, g4, t4);
                if (!last) {
                    k4_store_pad<W, kK4Ld2>(P1, acc2, g4, t4);
                    k4_store_pad<W, kPadMax>(stB(kk), acc2, g4, t4);
                    k4_load_pad<W, kPadMax>(acc2, gVb, g4, t4);
                }
                k4_load_pad<W, kPadMax>(acc, gVf, g4, t4);
                __syncthreads();
                if (!last) {
                    k4_mma<W, true, true, false>(acc2, P1, P2, ksteps, g4, t4);         // Vb -= Kb Delta
                    k4_store_pad<W, kPadMax>(gVb, acc2, g4, t4);
                }
                k4_mma<W, true, true, true>(acc, P0, P2, ksteps, g4, t4);               // Vf -= Kf Delta^T
                k4_store_pad<W, kPadMax>(gVf, acc, g4, t4);
                if (!finite) atomicOr(&P.status[w], 2);      // singular (or not positive definite) residual covariance
                if (P.Vall) k4_store<W>(P.Vall + ((size_t)w * p + kk) * mm, m, acc, m, g4, t4);
                if (last) k4_store<W>(P.V + (size_t)w * mm, m, acc, m, g4, t4);
            }
            // ---- phase 4: order update  A_j -= Kf B_{kk-1-j},  B_j -= Kb A_{kk-1-j}, IN PLACE: the updates of A_ja and B_jb (ja + jb = kk - 1)
            //      read only each other's old values, so both results are formed before either is stored (no second copy of the A / B
            //      sets: the scratch of all 599 windows is 138 MB instead of 261 MB and mostly stays in L2)
            if (last) {
                for (int j = 0; j < kk; ++j) {
                    __syncthreads();                                                     // P2 (Delta / previous operand) is free
                    k4_async_rows<2>(P2, stB(kk - 1 - j), kPadMax, m, l64);
                    k4_load_pad<W, kPadMax>(acc, stA(j), g4, t4);
                    k4_async_wait();
                    __syncthreads();
                    k4_mma<W, true, true, false>(acc, P0, P2, ksteps, g4, t4);
                    k4_store_lag<W>(Aw, p, j, acc, m, g4, t4);                          // final A_{j+1}
                }
            } else {
                for (int it = 0; it < kk; ++it) {
                    const int ja = it, jb = kk - 1 - it;
                    double acc2[13][2];
                    __syncthreads();                                                     // P2 is free
                    k4_async_rows<2>(P2, stB(jb), kPadMax, m, l64);
                    k4_load_pad<W, kPadMax>(acc, stA(ja), g4, t4);
                    k4_async_wait();
                    __syncthreads();
                    k4_mma<W, true, true, false>(acc, P0, P2, ksteps, g4, t4);          // A_ja - Kf B_jb   (kept in registers)
                    __syncthreads();                                                     // every thread is done with B_jb in P2
                    k4_async_rows<2>(P2, stA(ja), kPadMax, m, l64);                     // the OLD A_ja
                    k4_load_pad<W, kPadMax>(acc2, stB(jb), g4, t4);
                    k4_async_wait();
                    __syncthreads();
                    k4_mma<W, true, true, false>(acc2, P1, P2, ksteps, g4, t4);         // B_jb - Kb A_ja
                    k4_store_pad<W, kPadMax>(stA(ja), acc, g4, t4);
                    k4_store_pad<W, kPadMax>(stB(jb), acc2, g4, t4);
                }
            }
            __syncthreads();
        }
    }
}

template <int T>
__global__ void __launch_bounds__(64, kK4PerSM) lwr2_kernel(const K4Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* P0 = reinterpret_cast<double*>(smem_raw);
    double* P1 = P0 + kK4Panel2;
    double* P2 = P1 + kK4Panel2;
    for (int e = threadIdx.x; e < 3 * kK4Panel2; e += 64) P0[e] = 0.0;      // padding rows / columns stay zero for good
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) lwr2_body<0>(P, P0, P1, P2);
    else lwr2_body<1>(P, P0, P1, P2);
}

size_t lwr_ws_doubles(int grid, int m, int p) {       // lwr2_kernel keeps its scratch matrices padded to 40 x 40 (m <= 40 on this path)
    const size_t mm = (size_t)(m > kPadMax ? m : kPadMax) * (m > kPadMax ? m : kPadMax);
#ifdef HS_EXPERIMENT
    return (size_t)grid * (4 * p + 2) * mm;       // lwr1_kernel keeps two copies of the A / B sets
#else
    return (size_t)grid * (2 * p + 2) * mm;
#endif
}
int lwr_grid(int n_win) { const int slots = device_sm_count() * kK4PerSM; return n_win < slots ? n_win : slots; }

int launch_lwr(const K4Params& P, int grid, cudaStream_t stream) {
#ifdef HS_EXPERIMENT
    if (exp_env_int("HS_K4_LEGACY", 0) == 1) {
        const int sm = device_sm_count();
        const size_t smem = (size_t)(kK4Groups * 2 + 4) * kPanel * sizeof(double) + kK4Groups * sizeof(GJScratch);
        cudaError_t e = cudaFuncSetAttribute(lwr_kernel<kTileMax>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(HS_ERR_CUDA, "lwr: %s", cudaGetErrorString(e));
        lwr_kernel<kTileMax><<<grid < sm ? grid : sm, kK4Groups * 64, smem, stream>>>(P);
        return check_launch("lwr_kernel");
    }
#endif
#ifdef HS_EXPERIMENT
    if (exp_env_int("HS_K4_LEGACY", 0) == 2) {      // register-tiled DFMA products
        const size_t smem1 = (size_t)3 * kK4Panel * sizeof(double) + sizeof(GJScratch);
        cudaError_t e1 = cudaFuncSetAttribute(lwr1_kernel<kTileMax>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
        if (e1 != cudaSuccess) return set_error(HS_ERR_CUDA, "lwr: %s", cudaGetErrorString(e1));
        cudaFuncSetAttribute(lwr1_kernel<kTileMax>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        lwr1_kernel<kTileMax><<<grid, 64, smem1, stream>>>(P);
        return check_launch("lwr1_kernel");
    }
#endif
    const size_t smem = (size_t)3 * kK4Panel2 * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(lwr2_kernel<kTileMax>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(HS_ERR_CUDA, "lwr: %s", cudaGetErrorString(e));
    // five 43 KB CTAs per SM need the large shared-memory carve-out (the default split may leave room for two only)
    cudaFuncSetAttribute(lwr2_kernel<kTileMax>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
#ifdef HS_EXPERIMENT
    if (exp_env_int("HS_DEBUG", 0)) {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, lwr2_kernel<kTileMax>, 64, smem);
        fprintf(stderr, "[hs] lwr2_kernel: %d CTAs/SM, smem %zu, grid %d\n", nb, smem, grid);
    }
#endif
    lwr2_kernel<kTileMax><<<grid, 64, smem, stream>>>(P);
    return check_launch("lwr2_kernel");
}


int launch_ztable(const double* freqs, int F, int p, double fs, void* z, cudaStream_t stream) {
    const int n = p * F;
    ztable_kernel<<<(n + 255) / 256, 256, 0, stream>>>(freqs, F, p, fs, reinterpret_cast<double2*>(z));
    return check_launch("ztable_kernel");
}

int launch_ffdtf_normalize(double* dtf, const double* rowpart, const double* rowpart2, int n_win, int m, int F, int n_seg, double* out,
                           cudaStream_t stream) {
    dim3 grid(m, n_win);
    ffdtf_normalize_kernel<<<grid, 256, 0, stream>>>(dtf, rowpart, rowpart2, m, F, n_seg, out);
    return check_launch("ffdtf_normalize_kernel");
}

// =====================================================================================
// S(f) = H(f) * (V * H(f)^T), plain transpose (quirk of mtmvar.py:199).  One CTA per
// (window, bin); generic in m (shared-memory resident, m <= 64) -- not on the metric path.
// =====================================================================================
__global__ void spectra_kernel(const double2* __restrict__ H, const double* __restrict__ V, int m, int F, double2* __restrict__ S) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* h = reinterpret_cast<double2*>(smem_raw);          // [m][m]
    double2* t = h + (size_t)m * m;                             // T = V * H^T
    const int w = blockIdx.y, f = blockIdx.x;
    const size_t base = (size_t)w * m * m;
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) h[e] = H[(base + e) * F + f];
    __syncthreads();
    const double* Vw = V + base;
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
        const int i = e / m, j = e - i * m;      // T[i][j] = sum_k V[i][k] H[j][k]
        double sr = 0.0, si = 0.0;
        for (int k = 0; k < m; ++k) {
            const double v = Vw[i * m + k];
            const double2 x = h[j * m + k];
            sr = fma(v, x.x, sr);
            si = fma(v, x.y, si);
        }
        t[e] = make_double2(sr, si);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
        const int i = e / m, j = e - i * m;      // S[i][j] = sum_k H[i][k] T[k][j]
        double sr = 0.0, si = 0.0;
        for (int k = 0; k < m; ++k) {
            const double2 a = h[i * m + k], b = t[k * m + j];
            sr = fma(a.x, b.x, fma(-a.y, b.y, sr));
            si = fma(a.x, b.y, fma(a.y, b.x, si));
        }
        S[(base + e) * F + f] = make_double2(sr, si);
    }
}

// Partial coherence from the spectral matrix (partial_coherence, mtmvar.py:287-338) and dDTF = ffDTF * |kappa|
// (direct_dtf, :341-385).  The reference takes the determinant of every (m-1) x (m-1) minor of S(f) -- m^2 LU
// factorisations per bin; the minors are the entries of the adjugate:
//      minor_ij = (-1)^(i+j) det(S) (S^-1)[j][i]
// so ONE pivoted Gauss-Jordan inverse per bin (which also yields det S as the product of the pivots) gives all of them;
//      kappa_ij = minor_ij / sqrt(minor_ii minor_jj)   (principal complex square root, 0 where the denominator is 0),
//      kappa_ii = 1.
// One tile group (64 threads) per (window, bin), four groups per CTA; S is (n_win, m, m, F) complex128.
struct PcohScratch {
    GJScratch gj;
    double2 diag[kPadMax];
};

__device__ __forceinline__ double2 cmul2(const double2 a, const double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
// principal square root, as numpy / C99 csqrt
__device__ __forceinline__ double2 csqrt2(const double2 z) {
    if (z.x == 0.0 && z.y == 0.0) return make_double2(0.0, z.y);
    const double r = hypot(z.x, z.y);
    if (z.x >= 0.0) {
        const double t = sqrt(0.5 * (r + z.x));
        return make_double2(t, z.y / (2.0 * t));
    }
    const double t = sqrt(0.5 * (r - z.x));
    return make_double2(fabs(z.y) / (2.0 * t), copysign(t, z.y));
}
// a / b, Smith's algorithm (what C / numpy complex division does)
__device__ __forceinline__ double2 cdiv2(const double2 a, const double2 b) {
    if (fabs(b.x) >= fabs(b.y)) {
        const double r = b.y / b.x, d = 1.0 / (b.x + b.y * r);
        return make_double2((a.x + a.y * r) * d, (a.y - a.x * r) * d);
    }
    const double r = b.x / b.y, d = 1.0 / (b.x * r + b.y);
    return make_double2((a.x * r + a.y) * d, (a.y * r - a.x) * d);
}

template <int T>
__global__ void __launch_bounds__(256) pcoh_kernel(const double2* __restrict__ S, const int n_mat, const int m, const int F,
                                                   double2* __restrict__ kappa, const double* __restrict__ ffdtf,
                                                   double* __restrict__ ddtf, int* __restrict__ status) {
    __shared__ PcohScratch scr[4];
    const Group g = make_group();
    PcohScratch* sc = scr + g.gid;
    const int q = blockIdx.x * 4 + g.gid;
    const bool active = q < n_mat;
    const int w = active ? q / F : 0, f = active ? q - w * F : 0;
    const size_t sF = (size_t)F;
    const double2* Sw = S + (size_t)w * m * m * sF + f;
    // Diagonal pre-scaling S' = D S D with D_ii = 2^-round(log2 sqrt|S_ii|): the partial coherence is invariant under it (every minor
    // picks up the same factors in numerator and denominator), powers of two make it exact, and det S' stays representable where
    // det S of unscaled data (volts: |S_ii| ~ 1e-12, m = 38) would underflow.
    int ex_r[T], ex_c[T];
#pragma unroll
    for (int a = 0; a < T; ++a) {
        const int i = g.tr + 8 * a, j = g.tc + 8 * a;
        ex_r[a] = ex_c[a] = 0;
        if (active && i < m) {
            const double2 d = Sw[((size_t)i * m + i) * sF];
            const double mag = hypot(d.x, d.y);
            if (mag > 0.0 && mag < 1e300) ex_r[a] = -(ilogb(mag) / 2);
        }
        if (active && j < m) {
            const double2 d = Sw[((size_t)j * m + j) * sF];
            const double mag = hypot(d.x, d.y);
            if (mag > 0.0 && mag < 1e300) ex_c[a] = -(ilogb(mag) / 2);
        }
    }
    double ar[T][T], ai[T][T];
#pragma unroll
    for (int a = 0; a < T; ++a)
#pragma unroll
        for (int b = 0; b < T; ++b) {
            const int i = g.tr + 8 * a, j = g.tc + 8 * b;
            double2 v = make_double2((i == j) ? 1.0 : 0.0, 0.0);
            if (active && i < m && j < m) {
                v = Sw[((size_t)i * m + j) * sF];
                v.x = scalbn(v.x, ex_r[a] + ex_c[b]);
                v.y = scalbn(v.y, ex_r[a] + ex_c[b]);
            }
            ar[a][b] = v.x;
            ai[a][b] = v.y;
        }
    double2 pprod = make_double2(1.0, 0.0);
    gj_inverse<T, true, true>(ar, ai, m, g, &sc->gj, &pprod);
    // sign of the row permutation k -> colmap[k] (every thread walks the cycles: m <= 40 steps)
    double sign = 1.0;
    {
        unsigned long long seen = 0ull;
        for (int k = 0; k < m; ++k) {
            if ((seen >> k) & 1ull) continue;
            int len = 0, c = k;
            while (!((seen >> c) & 1ull)) {
                seen |= 1ull << c;
                c = sc->gj.colmap[c];
                ++len;
            }
            if (!(len & 1)) sign = -sign;
        }
    }
    const double2 det = cdiv2(make_double2(sign, 0.0), pprod);
    if (active && g.l64 == 0 && sc->gj.singular) atomicOr(&status[w], 4);
    // minors: storage entry (i', j') holds inverse[r][c], r = rowmap[i'], c = colmap[j']  ->  minor[c][r]
    int rr[T], cc[T];
#pragma unroll
    for (int a = 0; a < T; ++a) rr[a] = (g.tr + 8 * a < m) ? sc->gj.rowmap[g.tr + 8 * a] : -1;
#pragma unroll
    for (int b = 0; b < T; ++b) cc[b] = (g.tc + 8 * b < m) ? sc->gj.colmap[g.tc + 8 * b] : -2;
#pragma unroll
    for (int a = 0; a < T; ++a)
#pragma unroll
        for (int b = 0; b < T; ++b) {
            double2 v = cmul2(det, make_double2(ar[a][b], ai[a][b]));
            if ((rr[a] + cc[b]) & 1) v = make_double2(-v.x, -v.y);
            ar[a][b] = v.x;
            ai[a][b] = v.y;
            if (rr[a] == cc[b] && rr[a] >= 0) sc->diag[rr[a]] = v;
        }
    group_sync(g);
    if (!active) return;
#pragma unroll
    for (int a = 0; a < T; ++a)
#pragma unroll
        for (int b = 0; b < T; ++b) {
            const int r = rr[a], c = cc[b];
            if (r < 0 || c < 0) continue;
            double2 k;
            if (r == c) {
                k = make_double2(1.0, 0.0);
            } else {
                const double2 den = csqrt2(cmul2(sc->diag[c], sc->diag[r]));
                k = (den.x != 0.0 || den.y != 0.0) ? cdiv2(make_double2(ar[a][b], ai[a][b]), den) : make_double2(0.0, 0.0);
            }
            const size_t o = (((size_t)w * m + c) * m + r) * sF + f;
            if (kappa) kappa[o] = k;
            if (ddtf) ddtf[o] = ffdtf[o] * hypot(k.x, k.y);
        }
}

// Partial coherence for m > 40: the same identity (minor_cr = (-1)^(r+c) det S (S^-1)_rc), one CTA per (window, bin), IN PLACE in the
// kappa output -- the (m, m) slice of bin f (element stride F) is the work matrix, so no scratch is needed:
//   1. slice <- D S D (diagonal pre-scaling by powers of two, as above);
//   2. Gauss-Jordan with partial pivoting and PHYSICAL row swaps; only the PHASE of the determinant is tracked (product of the pivots'
//      unit phases times the swap parity): kappa is invariant under det -> rho det (rho > 0), and |det| of a 128 x 128 matrix leaves
//      the double range;
//   3. the swaps applied to the columns in reverse order: slice = (D S D)^-1 =: X;
//   4. kappa[c][r] = (-1)^(r+c) e^(i phi) X_rc / csqrt(e^(2 i phi) X_cc X_rr), pairwise in place (the diagonal of X is kept in shared
//      memory), 1 on the diagonal.
// Strided and unblocked: a completeness path (the reference's m^2 minor determinants per bin are O(m^5)), not a tuned one.
__global__ void __launch_bounds__(256) pcoh_generic_kernel(const double2* __restrict__ S, const int m, const int F, double2* __restrict__ kappa,
                                                           const double* __restrict__ ffdtf, double* __restrict__ ddtf, int* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char pg_smem[];
    double2* col = reinterpret_cast<double2*>(pg_smem);      // m: pivot column / diagonal of X
    double2* row = col + m;                                  // m: scaled pivot row
    int* ex = reinterpret_cast<int*>(row + m);               // m: scaling exponents
    int* perm = ex + m;                                      // m: row swapped with k at step k
    __shared__ double s_red[256];
    __shared__ int s_idx[256];
    __shared__ double2 s_phase;
    __shared__ int s_sing;
    const int f = blockIdx.x, w = blockIdx.y;
    const size_t sF = (size_t)F;
    const double2* Sw = S + (size_t)w * m * m * sF + f;
    double2* a = kappa + (size_t)w * m * m * sF + f;         // a[e * sF], e = i * m + j
    for (int i = threadIdx.x; i < m; i += 256) {
        const double2 d = Sw[((size_t)i * m + i) * sF];
        const double mag = hypot(d.x, d.y);
        ex[i] = (mag > 0.0 && mag < 1e300) ? -(ilogb(mag) / 2) : 0;
    }
    if (threadIdx.x == 0) { s_phase = make_double2(1.0, 0.0); s_sing = 0; }
    __syncthreads();
    for (int e = threadIdx.x; e < m * m; e += 256) {
        const int i = e / m, j = e - i * m;
        double2 v = Sw[(size_t)e * sF];
        v.x = scalbn(v.x, ex[i] + ex[j]);
        v.y = scalbn(v.y, ex[i] + ex[j]);
        a[(size_t)e * sF] = v;
    }
    __syncthreads();
    for (int k = 0; k < m; ++k) {
        // pivot: largest |a_ik|, i >= k (ties: smallest i)
        double best = -1.0;
        int bi = 1 << 20;
        for (int i = k + threadIdx.x; i < m; i += 256) {
            const double2 v = a[((size_t)i * m + k) * sF];
            const double mag = fma(v.x, v.x, v.y * v.y);
            if (mag > best) { best = mag; bi = i; }
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
        const int r = s_idx[0];
        const bool ok = s_red[0] > 0.0 && s_red[0] < 1.79e308;
        if (!ok) {                                            // singular (or non-finite) S(f): flag the window, leave the bin
            if (threadIdx.x == 0) { s_sing = 1; atomicOr(&status[w], 4); }
            __syncthreads();
            break;
        }
        if (threadIdx.x == 0) perm[k] = r;
        if (r != k)                                           // physical row swap
            for (int j = threadIdx.x; j < m; j += 256) {
                const double2 t0 = a[((size_t)k * m + j) * sF];
                a[((size_t)k * m + j) * sF] = a[((size_t)r * m + j) * sF];
                a[((size_t)r * m + j) * sF] = t0;
            }
        __syncthreads();
        const double2 pv = a[((size_t)k * m + k) * sF];
        const double d = 1.0 / fma(pv.x, pv.x, pv.y * pv.y);
        const double2 iv = make_double2(pv.x * d, -pv.y * d);
        if (threadIdx.x == 0) {                               // phase of det: times pv / |pv|, times -1 for a swap
            const double inv_abs = rsqrt(fma(pv.x, pv.x, pv.y * pv.y));
            double2 ph = cmul2(s_phase, make_double2(pv.x * inv_abs, pv.y * inv_abs));
            const double nrm = rsqrt(fma(ph.x, ph.x, ph.y * ph.y));
            const double sg = (r != k) ? -nrm : nrm;
            s_phase = make_double2(ph.x * sg, ph.y * sg);
        }
        for (int i = threadIdx.x; i < m; i += 256) col[i] = (i == k) ? make_double2(0.0, 0.0) : a[((size_t)i * m + k) * sF];
        for (int j = threadIdx.x; j < m; j += 256) {
            const double2 xv = (j == k) ? make_double2(1.0, 0.0) : a[((size_t)k * m + j) * sF];
            row[j] = cmul2(xv, iv);
        }
        __syncthreads();
        for (int e = threadIdx.x; e < m * m; e += 256) {
            const int i = e / m, j = e - i * m;
            double2 xv;
            if (i == k) {
                xv = row[j];
            } else {
                xv = (j == k) ? make_double2(0.0, 0.0) : a[(size_t)e * sF];
                const double2 c = col[i], rv = row[j];
                xv.x = fma(-c.x, rv.x, fma(c.y, rv.y, xv.x));
                xv.y = fma(-c.x, rv.y, fma(-c.y, rv.x, xv.y));
            }
            a[(size_t)e * sF] = xv;
        }
        __syncthreads();
    }
    if (s_sing) {                                             // defined output for a flagged bin: identity
        for (int e = threadIdx.x; e < m * m; e += 256) {
            const int i = e / m, j = e - i * m;
            a[(size_t)e * sF] = make_double2((i == j) ? 1.0 : 0.0, 0.0);
            if (ddtf) ddtf[(size_t)w * m * m * sF + (size_t)e * sF + f] = (i == j) ? ffdtf[(size_t)w * m * m * sF + (size_t)e * sF + f] : 0.0;
        }
        return;
    }
    // (P S')^-1 -> S'^-1: the row swaps as column swaps, last one first
    for (int k = m - 1; k >= 0; --k) {
        const int r = perm[k];
        if (r != k)
            for (int i = threadIdx.x; i < m; i += 256) {
                const double2 t0 = a[((size_t)i * m + k) * sF];
                a[((size_t)i * m + k) * sF] = a[((size_t)i * m + r) * sF];
                a[((size_t)i * m + r) * sF] = t0;
            }
        __syncthreads();
    }
    const double2 ph = s_phase;
    const double2 ph2 = cmul2(ph, ph);
    for (int i = threadIdx.x; i < m; i += 256) col[i] = a[((size_t)i * m + i) * sF];       // X_ii
    __syncthreads();
    const double* ffw = ffdtf ? ffdtf + (size_t)w * m * m * sF + f : nullptr;
    double* ddw = ddtf ? ddtf + (size_t)w * m * m * sF + f : nullptr;
    for (int e = threadIdx.x; e < m * m; e += 256) {
        const int r = e / m, c = e - r * m;
        if (r > c) continue;
        if (r == c) {
            a[(size_t)e * sF] = make_double2(1.0, 0.0);
            if (ddw) ddw[(size_t)e * sF] = ffw[(size_t)e * sF];
            continue;
        }
        // kappa[c][r] from X_rc, kappa[r][c] from X_cr: the pair is read before either is written
        const size_t e_rc = (size_t)r * m + c, e_cr = (size_t)c * m + r;
        const double2 x_rc = a[e_rc * sF], x_cr = a[e_cr * sF];
        const double2 den = csqrt2(cmul2(ph2, cmul2(col[c], col[r])));
        const bool live = den.x != 0.0 || den.y != 0.0;
        const double sg = ((r + c) & 1) ? -1.0 : 1.0;
        double2 n_rc = cmul2(ph, x_rc), n_cr = cmul2(ph, x_cr);
        n_rc = make_double2(sg * n_rc.x, sg * n_rc.y);
        n_cr = make_double2(sg * n_cr.x, sg * n_cr.y);
        const double2 k_cr = live ? cdiv2(n_rc, den) : make_double2(0.0, 0.0);      // kappa[c][r] <- minor_cr ~ X_rc
        const double2 k_rc = live ? cdiv2(n_cr, den) : make_double2(0.0, 0.0);      // kappa[r][c] <- minor_rc ~ X_cr
        a[e_cr * sF] = k_cr;
        a[e_rc * sF] = k_rc;
        if (ddw) {
            ddw[e_cr * sF] = ffw[e_cr * sF] * hypot(k_cr.x, k_cr.y);
            ddw[e_rc * sF] = ffw[e_rc * sF] * hypot(k_rc.x, k_rc.y);
        }
    }
}

int launch_pcoh(const void* S, int n_win, int m, int F, void* kappa, const double* ffdtf, double* ddtf, int* status, cudaStream_t stream) {
    const long long n_mat = (long long)n_win * F;
    const int grid = (int)((n_mat + 3) / 4);
    const double2* Sp = reinterpret_cast<const double2*>(S);
    double2* kp = reinterpret_cast<double2*>(kappa);
    switch ((m + 7) / 8) {
        case 1: pcoh_kernel<1><<<grid, 256, 0, stream>>>(Sp, (int)n_mat, m, F, kp, ffdtf, ddtf, status); break;
        case 2: pcoh_kernel<2><<<grid, 256, 0, stream>>>(Sp, (int)n_mat, m, F, kp, ffdtf, ddtf, status); break;
        case 3: pcoh_kernel<3><<<grid, 256, 0, stream>>>(Sp, (int)n_mat, m, F, kp, ffdtf, ddtf, status); break;
        case 4: pcoh_kernel<4><<<grid, 256, 0, stream>>>(Sp, (int)n_mat, m, F, kp, ffdtf, ddtf, status); break;
        case 5: pcoh_kernel<5><<<grid, 256, 0, stream>>>(Sp, (int)n_mat, m, F, kp, ffdtf, ddtf, status); break;
        default: {
            // m > 40: in place in the kappa output (no workspace in this entry point's signature)
            if (!kappa) return set_error(HS_ERR_INVALID, "partial coherence: m = %d > %d needs the kappa output as its work matrix", m, kPadMax);
            const size_t smem = (size_t)2 * m * sizeof(double2) + (size_t)2 * m * sizeof(int);
            if (smem > 48 * 1024) return set_error(HS_ERR_UNSUPPORTED, "partial coherence: m = %d too large", m);
            pcoh_generic_kernel<<<dim3(F, n_win), 256, smem, stream>>>(Sp, m, F, kp, ffdtf, ddtf, status);
            return check_launch("pcoh_generic_kernel");
        }
    }
    return check_launch("pcoh_kernel");
}

int launch_spectra(const void* H, const double* V, int n_win, int m, int F, void* S, cudaStream_t stream) {
    const size_t smem = (size_t)2 * m * m * sizeof(double2);
    if (smem > 200 * 1024) return set_error(HS_ERR_UNSUPPORTED, "spectra: m=%d too large", m);
    cudaError_t e = cudaFuncSetAttribute(spectra_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(HS_ERR_CUDA, "spectra: %s", cudaGetErrorString(e));
    dim3 grid(F, n_win);
    spectra_kernel<<<grid, 256, smem, stream>>>(reinterpret_cast<const double2*>(H), V, m, F, reinterpret_cast<double2*>(S));
    return check_launch("spectra_kernel");
}


}  // namespace hs
