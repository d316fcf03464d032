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
//               pivot column goes through ~1.3 KB of shared memory (one 64-thread
//               barrier per step), the scaled pivot row is broadcast with width-8
//               shuffles (threads that share a column set are 8 consecutive lanes),
//               everything else is rank-1 updates on registers (DFMA-bound).
//
// The column owners of a given column (fixed tc, tr = 0..7) are 8 consecutive
// lanes of ONE warp, so the pivot search is three xor-shuffles; consecutive columns
// belong to alternating warps so the two warps share the pivot work evenly.
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
    int wpar;        // parity of the columns this thread's warp owns
};

__device__ __forceinline__ Group make_group(const int flip = 0) {
    Group g;
    g.l64 = threadIdx.x & 63;
    g.gid = threadIdx.x >> 6;
    const int lane = g.l64 & 31;
    // Column parity owned by this warp.  Groups 2q and 2q+1 of a CTA land on different SMSP pairs, groups g and
    // g+2 on the same pair: flipping the parity for every other group of a pair keeps both SMSPs of the pair busy
    // with a mix of column-owner work and plain updates at every elimination step.
    const int w = (g.l64 >> 5) ^ (((g.gid >> 1) & 1) & flip);
    g.tr = lane & 7;
    g.tc = ((lane >> 3) << 1) | w;     // consecutive columns alternate between the two warps of the group
    g.wpar = w;
    g.bar = 1 + g.gid;
    return g;
}

__device__ __forceinline__ void group_sync(const Group& g) {
    asm volatile("bar.sync %0, 64;" ::"r"(g.bar) : "memory");
}

// Scratch for gj_inverse, one per group.
struct __align__(16) GJScratch {
    double2 c[2][kPadMax];   // pivot-column multipliers (double buffered by step parity)
    double2 inv[2];          // 1 / pivot            (double buffered)
    int piv[2];              // pivot row of the step (double buffered)
    int singular;            // sticky flag
    int pad_;
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
template <int T, bool CPLX, bool PIVOT = true>
__device__ __forceinline__ void gj_inverse(double (&ar)[T][T], double (&ai)[T][T], const int m,
                                           const Group& g, GJScratch* sh, double2* pinv_prod = nullptr) {
    constexpr int kNone = 1 << 20;
    unsigned used = 0;
    int step = 0;
    const int wsel = g.wpar;              // parity of the columns this warp owns
    if (g.l64 == 0) sh->singular = 0;
    group_sync(g);   // previous user of the scratch (rowmap/colmap readers) is done
#pragma unroll
    for (int b = 0; b < T; ++b) {
#pragma unroll 1
        for (int kc = 0; kc < 8; ++kc) {
            const int k = kc + 8 * b;
            if (k >= m) break;
            const int par = step & 1;
            if ((kc & 1) == wsel) {
                // ---- pivot search, whole warp in lock step (warp-uniform branch, full-mask shuffles):
                //      every 8-lane group scans ITS column tc + 8b; only the group with tc == kc publishes.
                double best = -1.0;
                int bi = kNone;
                if (!PIVOT) { best = 1.0; bi = k; }      // experiment: static diagonal pivot
#pragma unroll
                for (int a = 0; a < (PIVOT ? T : 0); ++a) {
                    const int i = g.tr + 8 * a;
                    const bool ok = (i < m) && !((used >> a) & 1u);
                    double mag = ar[a][b] * ar[a][b];
                    if (CPLX) mag = fma(ai[a][b], ai[a][b], mag);
                    const bool gt = mag > best;
                    if (ok && (gt || bi == kNone)) {
                        bi = i;
                        if (gt) best = mag;
                    }
                }
#pragma unroll
                for (int off = 1; off < (PIVOT ? 8 : 0); off <<= 1) {
                    const double ob = __shfl_xor_sync(0xffffffffu, best, off);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                    if (ob > best || (ob == best && oi < bi)) {
                        best = ob;
                        bi = oi;
                    }
                }
                const int r = bi;                 // agreed by the 8 lanes of the group (max |a|, lowest row on ties)
                if (g.tc == kc) {
                    // publish multipliers, reset own column to zero (the pivot lane fixes up row r below)
#pragma unroll
                    for (int a = 0; a < T; ++a) {
                        sh->c[par][g.tr + 8 * a] = make_double2(ar[a][b], CPLX ? ai[a][b] : 0.0);
                        ar[a][b] = 0.0;
                        if (CPLX) ai[a][b] = 0.0;
                    }
                    if (g.tr == (r & 7)) {
                        // this lane holds the pivot: 1/p = conj(p)/|p|^2
                        const double2 pv = sh->c[par][r];
                        double2 iv;
                        if (CPLX) {
                            const double d = 1.0 / fma(pv.x, pv.x, pv.y * pv.y);
                            iv = make_double2(pv.x * d, -pv.y * d);
                        } else {
                            iv = make_double2(1.0 / pv.x, 0.0);
                        }
                        sh->c[par][r] = make_double2(0.0, 0.0);     // row r is not eliminated
                        sh->inv[par] = iv;
                        sh->piv[par] = r;
                        sh->rowmap[r] = k;
                        sh->colmap[k] = r;
                        if (!(best > 0.0)) sh->singular = 1;
                    }
                }
            }
            group_sync(g);
            // ---- pivot row: its owners scale it in registers, then it is broadcast inside each 8-lane
            //      (same tc) group.  ra (which of my T row slots) is uniform over the group, so the
            //      switch is warp-uniform and every register index below is static.
            const int r = sh->piv[par];
            const int rl = r & 7, ra = r >> 3;
            const bool row_owner = (g.tr == rl);
            const double2 iv = sh->inv[par];
            if (pinv_prod) {      // product of the pivot reciprocals: 1 / (det * sign of the row permutation)
                const double2 q = *pinv_prod;
                *pinv_prod = make_double2(fma(q.x, iv.x, -q.y * iv.y), fma(q.x, iv.y, q.y * iv.x));
            }
            if (row_owner) used |= 1u << ra;
            double rr[T], ri[T];
#define HS_GJ_ROW_CASE(A)                                                                         \
    case A:                                                                                       \
        if constexpr (A < T) {                                                                    \
            if (row_owner) {                                                                      \
                if (g.tc == kc) {                                                                 \
                    ar[A][b] = 1.0; /* own column was zeroed: becomes 1 * inv */                  \
                    if (CPLX) ai[A][b] = 0.0;                                                     \
                }                                                                                 \
                _Pragma("unroll") for (int bb = 0; bb < T; ++bb) {                                \
                    const double xr = ar[A][bb], xi = CPLX ? ai[A][bb] : 0.0;                     \
                    if (CPLX) {                                                                   \
                        ar[A][bb] = fma(xr, iv.x, -xi * iv.y);                                    \
                        ai[A][bb] = fma(xr, iv.y, xi * iv.x);                                     \
                    } else {                                                                      \
                        ar[A][bb] = xr * iv.x;                                                    \
                    }                                                                             \
                }                                                                                 \
            }                                                                                     \
            _Pragma("unroll") for (int bb = 0; bb < T; ++bb) {                                    \
                rr[bb] = __shfl_sync(0xffffffffu, ar[A][bb], rl, 8);                              \
                ri[bb] = CPLX ? __shfl_sync(0xffffffffu, ai[A][bb], rl, 8) : 0.0;                 \
            }                                                                                     \
        }                                                                                         \
        break;
            switch (ra) {
                HS_GJ_ROW_CASE(0)
                HS_GJ_ROW_CASE(1)
                HS_GJ_ROW_CASE(2)
                HS_GJ_ROW_CASE(3)
                HS_GJ_ROW_CASE(4)
                default:
#pragma unroll
                    for (int bb = 0; bb < T; ++bb) rr[bb] = ri[bb] = 0.0;
                    break;
            }
#undef HS_GJ_ROW_CASE
            // ---- rank-1 update of every tile (row r has multiplier 0)
#pragma unroll
            for (int a = 0; a < T; ++a) {
                const double2 cv = sh->c[par][g.tr + 8 * a];
#pragma unroll
                for (int bb = 0; bb < T; ++bb) {
                    if (CPLX) {
                        ar[a][bb] = fma(-cv.x, rr[bb], fma(cv.y, ri[bb], ar[a][bb]));
                        ai[a][bb] = fma(-cv.x, ri[bb], fma(-cv.y, rr[bb], ai[a][bb]));
                    } else {
                        ar[a][bb] = fma(-cv.x, rr[bb], ar[a][bb]);
                    }
                }
            }
            ++step;
        }
    }
    group_sync(g);   // rowmap / colmap / singular complete and visible
}

// ---------------------------------------------------------------------------------
// Optimistic Gauss-Jordan: static diagonal pivots (pivot row of column k is row k).
// Nothing about the step is data dependent, so there is no search, no row bookkeeping and no
// permutation; the (unscaled) pivot row is broadcast with width-8 shuffles BEFORE the step's only
// barrier, overlapping the column owners' reciprocal, and the multipliers are published already
// scaled (c~ = c / p) so the rank-1 update needs nothing else.  The result is H itself
// (rowmap = colmap = identity).  It is only as stable as unpivoted elimination: callers either know
// the matrix is SPD (LWR residual covariances) or verify the result and fall back to gj_inverse.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ double rcp_newton(const double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));     // ~20 bits, branch free
    y = fma(y, fma(-x, y, 1.0), y);
    y = fma(y, fma(-x, y, 1.0), y);
    y = fma(y, fma(-x, y, 1.0), y);
    return y;
}

template <int T, bool CPLX>
__device__ __forceinline__ void gj_inverse_static(double (&ar)[T][T], double (&ai)[T][T], const int m,
                                                  const Group& g, GJScratch* sh) {
    int step = 0;
    group_sync(g);   // previous user of the scratch is done
#pragma unroll
    for (int b = 0; b < T; ++b) {
#pragma unroll 1
        for (int kc = 0; kc < 8; ++kc) {
            const int k = kc + 8 * b;
            if (k >= m) break;
            const int par = step & 1;
            const bool col_owner = (g.tc == kc);
            // ---- unscaled pivot row k (row slot b of the lanes with tr == kc), broadcast per 8-lane group
            double rr[T], ri[T];
#pragma unroll
            for (int bb = 0; bb < T; ++bb) {
                rr[bb] = __shfl_sync(0xffffffffu, ar[b][bb], kc, 8);
                ri[bb] = CPLX ? __shfl_sync(0xffffffffu, ai[b][bb], kc, 8) : 0.0;
            }
            if ((kc & 1) == g.wpar) {
                // this warp owns column k: rr[b] / ri[b] of the lanes with tc == kc is the pivot a_kk
                double ivr, ivi = 0.0;
                if (CPLX) {
                    const double d = rcp_newton(fma(rr[b], rr[b], ri[b] * ri[b]));
                    ivr = rr[b] * d;
                    ivi = -ri[b] * d;
                } else {
                    ivr = rcp_newton(rr[b]);
                }
                if (col_owner) {
#pragma unroll
                    for (int a = 0; a < T; ++a) {
                        double cr, ci = 0.0;
                        if (CPLX) {
                            cr = fma(ar[a][b], ivr, -ai[a][b] * ivi);
                            ci = fma(ar[a][b], ivi, ai[a][b] * ivr);
                        } else {
                            cr = ar[a][b] * ivr;
                        }
                        const bool piv = (a == b) && (g.tr == kc);
                        sh->c[par][g.tr + 8 * a] = piv ? make_double2(0.0, 0.0) : make_double2(cr, ci);
                        ar[a][b] = piv ? ivr : -cr;          // column k of the result: -c~ and 1/p
                        if (CPLX) ai[a][b] = piv ? ivi : -ci;
                    }
                    if (g.tr == kc) sh->inv[par] = make_double2(ivr, ivi);
                    rr[b] = 0.0;                              // the update below must not touch column k
                    ri[b] = 0.0;
                }
            }
            group_sync(g);
            // ---- row k := row k / p (its owners; column k of it was set above)
            if (g.tr == kc) {
                const double2 iv = sh->inv[par];
#pragma unroll
                for (int bb = 0; bb < T; ++bb) {
                    if (!(col_owner && bb == b)) {
                        const double xr = ar[b][bb], xi = CPLX ? ai[b][bb] : 0.0;
                        if (CPLX) {
                            ar[b][bb] = fma(xr, iv.x, -xi * iv.y);
                            ai[b][bb] = fma(xr, iv.y, xi * iv.x);
                        } else {
                            ar[b][bb] = xr * iv.x;
                        }
                    }
                }
            }
            // ---- rank-1 update with the scaled multipliers (0 for row k) and the unscaled row
#pragma unroll
            for (int a = 0; a < T; ++a) {
                const double2 cv = sh->c[par][g.tr + 8 * a];
#pragma unroll
                for (int bb = 0; bb < T; ++bb) {
                    if (CPLX) {
                        ar[a][bb] = fma(-cv.x, rr[bb], fma(cv.y, ri[bb], ar[a][bb]));
                        ai[a][bb] = fma(-cv.x, ri[bb], fma(-cv.y, rr[bb], ai[a][bb]));
                    } else {
                        ar[a][bb] = fma(-cv.x, rr[bb], ar[a][bb]);
                    }
                }
            }
            ++step;
        }
    }
}

// ---------------------------------------------------------------------------------
// Look-ahead Gauss-Jordan (static diagonal pivots, deferred row scaling).
//
// Same optimistic elimination as gj_inverse_static, restructured so that nothing but the rank-1
// update sits between two barriers:
//   * the pivot row and the SCALED multipliers of step k+1 are published to shared memory DURING
//     step k: the entries of column k+1 / row k+1 are updated first, their owners publish, and the
//     other 16 of the 25 tile entries are updated afterwards, so the reciprocal / scaling / store
//     latency of the next step overlaps this step's bulk DFMAs;
//   * both operands of the update come from shared memory as broadcast LDS.128 (10 per step and
//     thread) -- no shuffles in the steady state apart from the pivot broadcast of the owner warp;
//   * the pivot row is never scaled: row k keeps the factor p_k it had when it was eliminated and
//     1/p_k is recorded in sh->pinv[k].  The elimination is linear in each row, so the true inverse
//     is  H[i][j] = pinv[i] * S[i][j]  with S the register tiles on exit; |H|^2 needs one real scale
//     per row.  Column k of S is produced by the update itself: its owners zero it when they
//     publish, the published pivot-row entry for that column is 1 and the multiplier of row k is 0.
// One 64-thread barrier per step.  Stability is that of unpivoted elimination: callers know the
// matrix is SPD or verify the result a posteriori and fall back to gj_inverse.
// ---------------------------------------------------------------------------------
struct __align__(16) GJLScratch {
    double2 row[2][kPadMax];   // pivot row of the step (unscaled), double buffered by step parity
    double2 col[2][kPadMax];   // scaled multipliers a_ik / p_k (0 for i == k)
    double2 pinv[kPadMax];     // 1 / p_k
};

// publish pivot row / multipliers of step k1 = kc1 + 8*B1 into buffer `par1`
template <int T, bool CPLX, int B1>
__device__ __forceinline__ void gjl_publish(double (&ar)[T][T], double (&ai)[T][T], const int kc1, const int par1,
                                            const Group& g, GJLScratch* sh) {
    const bool col_lane = (g.tc == kc1);
    if (g.tr == kc1) {
#pragma unroll
        for (int bb = 0; bb < T; ++bb) {
            const bool piv = col_lane && (bb == B1);
            sh->row[par1][g.tc + 8 * bb] = piv ? make_double2(1.0, 0.0) : make_double2(ar[B1][bb], CPLX ? ai[B1][bb] : 0.0);
        }
    }
    if ((kc1 & 1) == g.wpar) {      // warp-uniform: this warp holds column k1
        const double pr = __shfl_sync(0xffffffffu, ar[B1][B1], kc1, 8);
        double ivr, ivi = 0.0;
        if (CPLX) {
            const double pi = __shfl_sync(0xffffffffu, ai[B1][B1], kc1, 8);
            const double d = rcp_newton(fma(pr, pr, pi * pi));
            ivr = pr * d;
            ivi = -pi * d;
        } else {
            ivr = rcp_newton(pr);
        }
#pragma unroll
        for (int a = 0; a < T; ++a) {
            double cr, ci = 0.0;
            if (CPLX) {
                cr = fma(ar[a][B1], ivr, -ai[a][B1] * ivi);
                ci = fma(ar[a][B1], ivi, ai[a][B1] * ivr);
            } else {
                cr = ar[a][B1] * ivr;
            }
            if (col_lane) {
                const bool piv = (a == B1) && (g.tr == kc1);
                sh->col[par1][g.tr + 8 * a] = piv ? make_double2(0.0, 0.0) : make_double2(cr, ci);
                ar[a][B1] = piv ? 1.0 : 0.0;
                if (CPLX) ai[a][B1] = 0.0;
            }
        }
        if (col_lane && g.tr == kc1) sh->pinv[kc1 + 8 * B1] = make_double2(ivr, ivi);
    }
}

template <int T, bool CPLX>
__device__ __forceinline__ void gjl_update(double (&ar)[T][T], double (&ai)[T][T], const double2 (&cc)[T], const double2 (&rr)[T],
                                           const int a, const int bb) {
    if (CPLX) {
        ar[a][bb] = fma(-cc[a].x, rr[bb].x, fma(cc[a].y, rr[bb].y, ar[a][bb]));
        ai[a][bb] = fma(-cc[a].x, rr[bb].y, fma(-cc[a].y, rr[bb].x, ai[a][bb]));
    } else {
        ar[a][bb] = fma(-cc[a].x, rr[bb].x, ar[a][bb]);
    }
}

// one elimination step k = kc + 8*B; LOOK: also publish step k+1 = kc1 + 8*B1
template <int T, bool CPLX, int B, int B1, bool LOOK>
__device__ __forceinline__ void gjl_step(double (&ar)[T][T], double (&ai)[T][T], const int kc, const int kc1, const int par,
                                         const Group& g, GJLScratch* sh) {
    group_sync(g);
    double2 cc[T], rr[T];
#pragma unroll
    for (int a = 0; a < T; ++a) cc[a] = sh->col[par][g.tr + 8 * a];
#pragma unroll
    for (int bb = 0; bb < T; ++bb) rr[bb] = sh->row[par][g.tc + 8 * bb];
    if (LOOK) {
#pragma unroll
        for (int a = 0; a < T; ++a) gjl_update<T, CPLX>(ar, ai, cc, rr, a, B1);
#pragma unroll
        for (int bb = 0; bb < T; ++bb)
            if (bb != B1) gjl_update<T, CPLX>(ar, ai, cc, rr, B1, bb);
        gjl_publish<T, CPLX, B1>(ar, ai, kc1, par ^ 1, g, sh);
    }
#pragma unroll
    for (int a = 0; a < T; ++a)
#pragma unroll
        for (int bb = 0; bb < T; ++bb)
            if (!LOOK || (a != B1 && bb != B1)) gjl_update<T, CPLX>(ar, ai, cc, rr, a, bb);
}

template <int T, bool CPLX, int B>
__device__ __forceinline__ void gjl_block(double (&ar)[T][T], double (&ai)[T][T], const int m, const Group& g, GJLScratch* sh) {
    if constexpr (B < T) {
        const int left = m - 8 * B;                  // steps of this block: kc = 0 .. min(8, left) - 1
        if (left > 0) {
            const int nk = left < 8 ? left : 8;
#pragma unroll 1
            for (int kc = 0; kc < nk - 1; ++kc) gjl_step<T, CPLX, B, B, true>(ar, ai, kc, kc + 1, kc & 1, g, sh);
            if (left > 8) {
                if constexpr (B + 1 < T) gjl_step<T, CPLX, B, B + 1, true>(ar, ai, 7, 0, 1, g, sh);
            } else {
                gjl_step<T, CPLX, B, B, false>(ar, ai, nk - 1, 0, (nk - 1) & 1, g, sh);      // last step of the matrix
            }
            gjl_block<T, CPLX, B + 1>(ar, ai, m, g, sh);
        }
    }
}

// On exit: inverse[i][j] = sh->pinv[i] * tile[i][j]; the last barrier executed is the one that
// opens step m-1, so every pinv entry is visible to the whole group.
template <int T, bool CPLX>
__device__ __forceinline__ void gj_inverse_la(double (&ar)[T][T], double (&ai)[T][T], const int m, const Group& g,
                                              GJLScratch* sh) {
    group_sync(g);   // previous readers of the scratch (pinv, row/col of the previous matrix) are done
    gjl_publish<T, CPLX, 0>(ar, ai, 0, 0, g, sh);
    gjl_block<T, CPLX, 0>(ar, ai, m, g, sh);
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
