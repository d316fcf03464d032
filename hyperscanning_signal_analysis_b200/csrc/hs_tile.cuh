// Register-tile primitives shared by the MVAR kernels (sm_100a, FP64).
//
// A "tile group" is 64 threads (2 warps) arranged as an 8 x 8 thread grid.  Thread
// (tr, tc) owns the cyclic sub-matrix  rows i = tr + 8a,  cols j = tc + 8b,
// a, b < T, of a (8T x 8T)-padded matrix, held entirely in registers.  T = 5 covers
// the 2 x 19 channel dyad (m = 38).  All m x m dense work on the hot path (lag
// covariances, the LWR recursion, A(f)^-1) is built from two primitives on that
// layout:
//
//   tile_mac    C(regs) +=/-= Pa^T-panel * Pb-panel from shared memory (k-major
//               panels, so every LDS is an 8-lane contiguous 64 B read that is
//               broadcast to the other 3 quarter-warps: no bank conflicts)
//   gj_inverse  in-place Gauss-Jordan inversion with implicit row pivoting; the
//               pivot column/row are exchanged through ~1.5 KB of shared memory,
//               everything else is rank-1 updates on registers (DFMA-bound).
//
// The column owners of a given column (fixed tc, tr = 0..7) are 8 consecutive
// lanes of ONE warp, so the pivot search is three xor-shuffles.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hs {

constexpr int kGroupThreads = 64;
constexpr int kTileMax = 5;                 // T <= 5  ->  m <= 40
constexpr int kPadMax = 8 * kTileMax;       // 40

struct Group {
    int tr, tc;      // thread coordinates in the 8 x 8 grid
    int gid;         // group index inside the CTA
    int l64;         // thread index inside the group
    int bar;         // named barrier id (1..15)
};

__device__ __forceinline__ Group make_group() {
    Group g;
    g.l64 = threadIdx.x & 63;
    g.gid = threadIdx.x >> 6;
    const int w = g.l64 >> 5, lane = g.l64 & 31;
    g.tr = lane & 7;
    g.tc = (lane >> 3) + 4 * w;
    g.bar = 1 + g.gid;
    return g;
}

__device__ __forceinline__ void group_sync(const Group& g) {
    asm volatile("bar.sync %0, 64;" ::"r"(g.bar) : "memory");
}

// Scratch for gj_inverse, one per group.
struct __align__(16) GJScratch {
    double2 c[2][kPadMax];   // pivot-column multipliers (double buffered)
    double2 r[kPadMax];      // scaled pivot row
    double2 inv;             // 1 / pivot
    int piv;                 // pivot row of the current step
    int singular;            // sticky flag
    int rowmap[kPadMax];     // storage row i  -> row of the inverse
    int colmap[kPadMax];     // storage col j  -> column of the inverse
};

// ---------------------------------------------------------------------------------
// In-place Gauss-Jordan inverse of the m x m matrix held in the group's register
// tiles (padding rows/cols must hold the identity).  Implicit row pivoting: at the
// step that eliminates column k the unused row r_k with the largest |a_ik| is the
// pivot; rows are never moved.  On exit the registers hold S with
//        inverse[rowmap[i]][colmap[j]] = S[i][j],   rowmap[r_k] = k, colmap[k] = r_k.
// (np.linalg.inv, which the reference calls per frequency bin at mtmvar.py:159,
// is LU with partial pivoting; any pivoted elimination agrees to rounding.)
// ---------------------------------------------------------------------------------
template <int T, bool CPLX>
__device__ __forceinline__ void gj_inverse(double (&ar)[T][T], double (&ai)[T][T], const int m,
                                           const Group& g, GJScratch* sh) {
    unsigned used = 0;
    int step = 0;
    if (g.l64 == 0) sh->singular = 0;
    group_sync(g);   // previous user of the scratch (rowmap/colmap readers) is done
#pragma unroll
    for (int b = 0; b < T; ++b) {
        for (int kc = 0; kc < 8; ++kc) {
            const int k = kc + 8 * b;
            if (k >= m) break;
            const int par = step & 1;
            if (g.tc == kc) {
                // ---- pivot search over the unused rows of column k (8 lanes of one warp)
                const unsigned mask = 0xFFu << (8 * (kc & 3));   // the 8 owner lanes of this warp
                double best = -1.0;
                int bi = -1, fb = 1 << 20;
#pragma unroll
                for (int a = 0; a < T; ++a) {
                    const int i = g.tr + 8 * a;
                    const bool ok = (i < m) && !((used >> a) & 1u);
                    double mag = ar[a][b] * ar[a][b];
                    if (CPLX) mag = fma(ai[a][b], ai[a][b], mag);
                    if (ok && i < fb) fb = i;
                    if (ok && mag > best) { best = mag; bi = i; }
                }
#pragma unroll
                for (int off = 1; off < 8; off <<= 1) {
                    const double ob = __shfl_xor_sync(mask, best, off, 8);
                    const int oi = __shfl_xor_sync(mask, bi, off, 8);
                    const int of = __shfl_xor_sync(mask, fb, off, 8);
                    if (oi >= 0 && (ob > best || (ob == best && oi < bi) || bi < 0)) { best = ob; bi = oi; }
                    fb = min(fb, of);
                }
                const bool bad = (bi < 0) || !(best > 0.0);
                const int r = (bi < 0) ? fb : bi;
                // ---- publish multipliers, reset own column to e_r
#pragma unroll
                for (int a = 0; a < T; ++a) {
                    const int i = g.tr + 8 * a;
                    double2 v = make_double2(ar[a][b], CPLX ? ai[a][b] : 0.0);
                    if (i == r) {
                        // this lane holds the pivot: 1/p = conj(p)/|p|^2
                        double2 iv;
                        if (CPLX) {
                            const double d = 1.0 / fma(v.x, v.x, v.y * v.y);
                            iv = make_double2(v.x * d, -v.y * d);
                        } else {
                            iv = make_double2(1.0 / v.x, 0.0);
                        }
                        sh->inv = iv;
                        sh->piv = r;
                        sh->rowmap[r] = k;
                        sh->colmap[k] = r;
                        if (bad) sh->singular = 1;
                        v = make_double2(0.0, 0.0);
                        ar[a][b] = 1.0;
                    } else {
                        ar[a][b] = 0.0;
                    }
                    if (CPLX) ai[a][b] = 0.0;
                    sh->c[par][i] = v;
                }
            }
            group_sync(g);
            const int r = sh->piv;
            if (g.tr == (r & 7)) {
                const double2 iv = sh->inv;
                const int ra = r >> 3;
                used |= 1u << ra;
#pragma unroll
                for (int a = 0; a < T; ++a) {
                    if (a == ra) {
#pragma unroll
                        for (int bb = 0; bb < T; ++bb) {
                            const double xr = ar[a][bb], xi = CPLX ? ai[a][bb] : 0.0;
                            double yr, yi = 0.0;
                            if (CPLX) {
                                yr = fma(xr, iv.x, -xi * iv.y);
                                yi = fma(xr, iv.y, xi * iv.x);
                                ai[a][bb] = yi;
                            } else {
                                yr = xr * iv.x;
                            }
                            ar[a][bb] = yr;
                            sh->r[g.tc + 8 * bb] = make_double2(yr, yi);
                        }
                    }
                }
            }
            group_sync(g);
            // ---- rank-1 update of every tile (row r has multiplier 0)
            double cr[T], ci[T];
#pragma unroll
            for (int a = 0; a < T; ++a) {
                const double2 v = sh->c[par][g.tr + 8 * a];
                cr[a] = v.x;
                ci[a] = v.y;
            }
#pragma unroll
            for (int bb = 0; bb < T; ++bb) {
                const double2 rv = sh->r[g.tc + 8 * bb];
#pragma unroll
                for (int a = 0; a < T; ++a) {
                    if (CPLX) {
                        ar[a][bb] = fma(-cr[a], rv.x, fma(ci[a], rv.y, ar[a][bb]));
                        ai[a][bb] = fma(-cr[a], rv.y, fma(-ci[a], rv.x, ai[a][bb]));
                    } else {
                        ar[a][bb] = fma(-cr[a], rv.x, ar[a][bb]);
                    }
                }
            }
            ++step;
        }
    }
    group_sync(g);   // rowmap / colmap / singular complete and visible
}

// ---------------------------------------------------------------------------------
// acc[a][b] += sign * sum_{k<depth} Pa[k*lda + tr+8a] * Pb[k*ldb + tc+8b]
// Pa / Pb are k-major panels in shared memory (lda, ldb >= 8T).
// ---------------------------------------------------------------------------------
template <int T, bool SUB>
__device__ __forceinline__ void tile_mac(double (&acc)[T][T], const double* __restrict__ Pa, const int lda,
                                         const double* __restrict__ Pb, const int ldb, const int depth,
                                         const Group& g) {
    const double* pa = Pa + g.tr;
    const double* pb = Pb + g.tc;
#pragma unroll 2
    for (int k = 0; k < depth; ++k) {
        double av[T], bv[T];
#pragma unroll
        for (int a = 0; a < T; ++a) av[a] = pa[8 * a];
#pragma unroll
        for (int b = 0; b < T; ++b) bv[b] = pb[8 * b];
#pragma unroll
        for (int a = 0; a < T; ++a)
#pragma unroll
            for (int b = 0; b < T; ++b) acc[a][b] = fma(SUB ? -av[a] : av[a], bv[b], acc[a][b]);
        pa += lda;
        pb += ldb;
    }
}

// ---------------------------------------------------------------------------------
// Panel helpers (64 threads of one group).  Panels are kPadMax x kPadMax doubles.
// ---------------------------------------------------------------------------------
// dst[r*40 + c] = src[r][c] (or src[c][r] if TR), zero padded; src is m x m, row stride ld.
template <bool TR>
__device__ __forceinline__ void load_panel(double* __restrict__ dst, const double* __restrict__ src, const int ld,
                                           const int m, const Group& g) {
    for (int e = g.l64; e < kPadMax * kPadMax; e += kGroupThreads) {
        const int r = e / kPadMax, c = e - r * kPadMax;
        double v = 0.0;
        if (r < m && c < m) v = TR ? src[(size_t)c * ld + r] : src[(size_t)r * ld + c];
        dst[e] = v;
    }
}

// row-major store of the register tile: dst[i*ld + j] (i, j < m)
template <int T>
__device__ __forceinline__ void store_tile(double* __restrict__ dst, const int ld, const double (&acc)[T][T], const int m,
                                           const Group& g) {
#pragma unroll
    for (int a = 0; a < T; ++a) {
        const int i = g.tr + 8 * a;
#pragma unroll
        for (int b = 0; b < T; ++b) {
            const int j = g.tc + 8 * b;
            if (i < m && j < m) dst[(size_t)i * ld + j] = acc[a][b];
        }
    }
}
// transposed store: dst[j*ld + i]
template <int T>
__device__ __forceinline__ void store_tile_t(double* __restrict__ dst, const int ld, const double (&acc)[T][T], const int m,
                                             const Group& g) {
#pragma unroll
    for (int a = 0; a < T; ++a) {
        const int i = g.tr + 8 * a;
#pragma unroll
        for (int b = 0; b < T; ++b) {
            const int j = g.tc + 8 * b;
            if (i < m && j < m) dst[(size_t)j * ld + i] = acc[a][b];
        }
    }
}
template <int T>
__device__ __forceinline__ void load_tile(double (&acc)[T][T], const double* __restrict__ src, const int ld, const int m,
                                          const Group& g) {
#pragma unroll
    for (int a = 0; a < T; ++a) {
        const int i = g.tr + 8 * a;
#pragma unroll
        for (int b = 0; b < T; ++b) {
            const int j = g.tc + 8 * b;
            acc[a][b] = (i < m && j < m) ? src[(size_t)i * ld + j] : 0.0;
        }
    }
}

}  // namespace hs
