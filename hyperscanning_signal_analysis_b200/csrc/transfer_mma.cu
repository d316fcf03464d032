// K5 on the FP64 tensor pipe:  A(f) = I - sum_k A_k z_k(f),  H = A(f)^-1,  |H|^2  (m <= 40).
//
// Replaces, per (window, frequency bin), the body of the reference's loop in
// mvar_transfer_function (/root/reference src/mtmvar.py:155-159: A(f) assembly + np.linalg.inv) and the
// |H|^2 of dtf_multivariate / full_freq_dtf (:232, :278).
//
// Why DMMA: a rank-1 Gauss-Jordan update on register tiles is one DFMA per matrix entry with three
// distinct 64-bit register operands, which this chip issues at 2.46 cycles/warp-instruction instead of 2
// (tools/lat_probe.cu), plus ~1 non-FMA instruction per FMA.  mma.sync.m8n8k4.f64 does 8 FMAs per lane and
// instruction from 4 register operands and reaches the FP64 pipe's peak with the matrix resident in
// registers (tools/dmma_probe.cu: 100 % of peak with fragment LDS + barriers).  So the elimination is a
// BLOCK Gauss-Jordan with 4 x 4 pivot blocks, and everything except the 4 x 4 inverse is DMMA.
//
// Layout: one matrix = one GROUP of two warps; warp 0 holds Re S, warp 1 holds Im S, both as T x T
// tiles of 8 x 8 in the m8n8k4 accumulator layout (lane l: row l>>2, columns 2(l&3), 2(l&3)+1), 50 doubles
// per lane for T = 5.  Re/Im split => both warps do identical work.
//
// Block step for the pivot block K = {k0..k0+3}, D = S[K,K], C = S[:,K], R = S[K,:]:
//   P = D^-1                                (16 lanes, in-place Gauss-Jordan with shuffles)
//   L = C P          for rows outside K     (DMMA: A = C as extracted, B = -P with Re/Im interleaved by column,
//   L = I - P        for the rows K          so the accumulator IS the (-L_re, -L_im) A-fragment of the next DMMAs)
//   S[:,K] := identity pattern (0, and I on D);  U := R with the same pattern
//   S <- S - L U                            (2 DMMAs per real tile: Re S += (-Lr) Ur + (+Li) Ui, Im S += (-Lr) Ui + (-Li) Ur)
// which yields  S[K,:] = P U  (scaled pivot rows incl. P itself in the columns K),  S[i,K] = -L  and the rank-4
// update everywhere else: the in-place block Gauss-Jordan inverse.  After ceil(m/4) steps S = H.
// Data flow per step: both warps copy their part of the NEXT panel (columns as A fragments, rows as B
// fragments) to shared memory right after the update, one 64-thread barrier, then each warp inverts D for itself.
//
// Stability is that of elimination without pivoting across blocks; the kernel verifies every matrix
// (|| H (A u) - u || on a probe vector) and flags failures for the pivoted register-tile kernel
// (transfer_dtf_kernel MODE 2).
#include <cstdlib>
#include "hs_tile.cuh"
#include "hs_internal.h"
#include "mvar_launch.h"

namespace hs {

namespace {

constexpr int kMP = 40;                     // padded matrix dimension (5 tiles of 8)
constexpr int kRowD = 2 * kMP + 2;          // doubles per coefficient row of one lag pair: 40 x (2 lags) + 16 B skew
constexpr int kPlaneD = kMP * kRowD + 4;    // doubles per lag-pair plane (+32 B skew between planes)
constexpr int kUS = 52;                     // row stride of the B-fragment buffer (= 4 mod 16 doubles: conflict-free fragment loads)

__device__ __forceinline__ void dmma884(double& c0, double& c1, const double a, const double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ double flip_sign(const double v, const int mask_hi) {
    return __hiloint2double(__double2hiint(v) ^ mask_hi, __double2loint(v));
}

__device__ __forceinline__ double rcp_newton2(const double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));     // ~20 bits
    y = fma(y, fma(-x, y, 1.0), y);                             // ~40
    y = fma(y, fma(-x, y, 1.0), y);                             // full
    return y;
}

struct __align__(16) MmaGroupSmem {
    union {
        struct {
            double Craw[2][2][kMP * 4];    // [buf][part][row*4 + k]   panel columns  (A-fragment order)
            double Rraw[2][2][4 * kUS];    // [buf][part][k*52 + col]  panel rows with the identity pattern (B-fragment order)
            double2 P[2][16];              // [warp][i*4 + j]          D^-1, private copy per warp
            float Df[2][32];               // [warp][part*16 + i*4 + j] pivot block rounded to FP32 (seed of mma_inverse4_newton)
        } p;
        double2 X[25][32];                 // Re/Im exchange of the epilogue (after the elimination)
    } u;
    double2 vfull[kMP];                    // A(f) u
    double wpart[2][2][kMP];               // [part][vr|vi][row]  partial S v
    double rs[2][kMP];                     // ffDTF row sums of this unit, per warp
    int flag;
    int pad_[3];
    // warp-specialised kernel only (transfer_ws_kernel): the helper warp's mailbox
    double2 Ph[16];                        // D^-1 of the current block step, written by the helper warp
    double Dblk[2][16];                    // [part][i*4 + j]  the NEXT pivot block, published early by the two main warps
    unsigned long long bar_d, bar_p;       // mbarriers: "pivot block published" (2 arrivals), "inverse ready" (1 arrival)
};

struct MmaCtx {
    int part, lane, g4, t4, bar, dbg;
    MmaGroupSmem* gs;
    int* lock;          // tensor-pipe turn lock of this warp's SM sub-partition, or null (transfer_mma_kernel)
};

// Experiment (HS_EXPERIMENT builds, hs_transfer_set_kernel(4)): one warp at a time per SM sub-partition streams its block step's
// DMMAs.  The idea: three warps sharing a sub-partition's FP64 tensor pipe might fall into lock-step (all streaming DMMAs at a third
// of the rate, then all in their scalar phases with the pipe idle); a single warp can feed the pipe alone (tools/dmma_rate.cu: 16.5
// cycles per DMMA with one warp, 16.0 with two or more).  Measured: 5.85 ms against 5.29 ms without the lock -- the hand-over of the
// lock and the exposed latencies at the start of every hold cost more than the interleaving wins.
__device__ __forceinline__ void pipe_lock(const MmaCtx& x) {
#ifdef HS_EXPERIMENT
    if (x.lock) {
        if (x.lane == 0) {
            while (atomicCAS(x.lock, 0, 1) != 0) { }
        }
        __syncwarp();
    }
#endif
}
__device__ __forceinline__ void pipe_unlock(const MmaCtx& x) {
#ifdef HS_EXPERIMENT
    if (x.lock) {
        __syncwarp();
        if (x.lane == 0) atomicExch(x.lock, 0);
    }
#endif
}

__device__ __forceinline__ void mma_group_sync(const MmaCtx& x) { asm volatile("bar.sync %0, 64;" ::"r"(x.bar) : "memory"); }

// ---- copy the panel (tile t, half h: columns / rows 8t+4h .. 8t+4h+3) of the own part to shared memory, then put
//      the identity pattern into the panel columns of the register tiles
template <int T, int t>
__device__ __forceinline__ void mma_extract(double (&c)[T][T][2], const int h, const MmaCtx& x) {
    MmaGroupSmem* gs = x.gs;
    const bool col_lane = (x.t4 >> 1) == h;
    if (col_lane) {
#pragma unroll
        for (int ta = 0; ta < T; ++ta)
            *reinterpret_cast<double2*>(&gs->u.p.Craw[h][x.part][(8 * ta + x.g4) * 4 + 2 * (x.t4 & 1)]) = make_double2(c[ta][t][0], c[ta][t][1]);
    }
    // identity pattern of the panel columns
    const double d0 = (x.part == 0 && x.g4 == 2 * x.t4) ? 1.0 : 0.0, d1 = (x.part == 0 && x.g4 == 2 * x.t4 + 1) ? 1.0 : 0.0;
    if (col_lane) {
#pragma unroll
        for (int ta = 0; ta < T; ++ta) {
            c[ta][t][0] = (ta == t) ? d0 : 0.0;
            c[ta][t][1] = (ta == t) ? d1 : 0.0;
        }
    }
    if ((x.lane >> 4) == h) {
#pragma unroll
        for (int tb = 0; tb < T; ++tb)
            *reinterpret_cast<double2*>(&gs->u.p.Rraw[h][x.part][(x.g4 & 3) * kUS + 8 * tb + 2 * x.t4]) = make_double2(c[t][tb][0], c[t][tb][1]);
    }
}

// ---- P = D^-1 of the panel in buffer `buf`: lanes 0..15 (mirrored in 16..31) hold D[i][j], in-place Gauss-Jordan with shuffles
__device__ __forceinline__ void mma_inverse4(const MmaCtx& x, const int K0, const int buf) {
    MmaGroupSmem* gs = x.gs;
    const int i = (x.lane >> 2) & 3, j = x.lane & 3;
    double dr = gs->u.p.Craw[buf][0][(K0 + i) * 4 + j];
    double di = gs->u.p.Craw[buf][1][(K0 + i) * 4 + j];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const double pr = __shfl_sync(0xffffffffu, dr, s * 5, 16), pi = __shfl_sync(0xffffffffu, di, s * 5, 16);
        const double rjr = __shfl_sync(0xffffffffu, dr, s * 4 + j, 16), rji = __shfl_sync(0xffffffffu, di, s * 4 + j, 16);
        const double cir = __shfl_sync(0xffffffffu, dr, i * 4 + s, 16), cii = __shfl_sync(0xffffffffu, di, i * 4 + s, 16);
        const bool prow = (i == s), pcol = (j == s);
        const double sr = prow ? dr : cir, si = prow ? di : cii;       // pivot row: scale the own entry; other rows: multiplier
        // t = s / p = y q,  u = d - t r = d - y (q r)  with  y = 1 / |p|^2,  q = s conj(p):  q and q r are computed next to the
        // reciprocal chain, so only ONE dependent FMA follows it
        const double y = rcp_newton2(fma(pr, pr, pi * pi));
        const double qr = fma(sr, pr, si * pi), qi = fma(si, pr, -sr * pi);
        const double wr = fma(qr, rjr, -qi * rji), wi = fma(qr, rji, qi * rjr);
        const double tr = qr * y, ti = qi * y;
        const double ivr = pr * y, ivi = -pi * y;
        const double ur = fma(-y, wr, dr), ui = fma(-y, wi, di);
        dr = prow ? (pcol ? ivr : tr) : (pcol ? -tr : ur);
        di = prow ? (pcol ? ivi : ti) : (pcol ? -ti : ui);
    }
    if (x.lane < 16) gs->u.p.P[x.part][x.lane] = make_double2(dr, di);
    __syncwarp();
}

// ---- P = D^-1 by cofactors: lane (i, j) (all 32 lanes, 16..31 mirror 0..15) takes the 3 x 3 minor that deletes row j and column i
//      straight from the panel in shared memory (9 complex loads, any lane pattern inside one 128-byte block is conflict free),
//      the determinant follows from one round of shuffles (det = sum_c D[0][c] adj[c][0]), then ONE reciprocal.
//      Dependent depth ~20 FP64 instructions and 1 shuffle round, against 4 x (shuffle round + reciprocal chain) for the
//      in-place elimination above; no pivot order inside the block, so small leading minors of D are harmless.
template <bool SCALED>      // SCALED: P = adj / det in shared memory;  otherwise P = adj and the caller applies 1 / det (returned in det)
__device__ __forceinline__ void mma_inverse4_adj(const MmaCtx& x, const int K0, const int buf, double2& det) {
    MmaGroupSmem* gs = x.gs;
    const int i = (x.lane >> 2) & 3, j = x.lane & 3;
    const double* Dr = gs->u.p.Craw[buf][0] + K0 * 4;
    const double* Di = gs->u.p.Craw[buf][1] + K0 * 4;
    const int r0 = (0 >= j) ? 1 : 0, r1 = (1 >= j) ? 2 : 1, r2 = (2 >= j) ? 3 : 2;
    const int c0 = (0 >= i) ? 1 : 0, c1 = (1 >= i) ? 2 : 1, c2 = (2 >= i) ? 3 : 2;
#define HS_LD(R, C, vr, vi) const double vr = Dr[(R) * 4 + (C)], vi = Di[(R) * 4 + (C)]
    HS_LD(r0, c0, a00r, a00i); HS_LD(r0, c1, a01r, a01i); HS_LD(r0, c2, a02r, a02i);
    HS_LD(r1, c0, a10r, a10i); HS_LD(r1, c1, a11r, a11i); HS_LD(r1, c2, a12r, a12i);
    HS_LD(r2, c0, a20r, a20i); HS_LD(r2, c1, a21r, a21i); HS_LD(r2, c2, a22r, a22i);
    HS_LD(0, 0, d0r, d0i); HS_LD(0, 1, d1r, d1i); HS_LD(0, 2, d2r, d2i); HS_LD(0, 3, d3r, d3i);
#undef HS_LD
    // 2 x 2 minors of rows r1, r2:  m0 = a11 a22 - a12 a21,  m1 = a10 a22 - a12 a20,  m2 = a10 a21 - a11 a20
#define HS_DET2(pr, pi, qr, qi, sr, si, tr, ti, outr, outi)                                        \
    const double outr = fma(pr, qr, -pi * qi) - fma(sr, tr, -si * ti);                              \
    const double outi = fma(pr, qi, pi * qr) - fma(sr, ti, si * tr)
    HS_DET2(a11r, a11i, a22r, a22i, a12r, a12i, a21r, a21i, m0r, m0i);
    HS_DET2(a10r, a10i, a22r, a22i, a12r, a12i, a20r, a20i, m1r, m1i);
    HS_DET2(a10r, a10i, a21r, a21i, a11r, a11i, a20r, a20i, m2r, m2i);
#undef HS_DET2
    // cofactor = a00 m0 - a01 m1 + a02 m2, signed
    double cr = fma(a00r, m0r, -a00i * m0i) - fma(a01r, m1r, -a01i * m1i) + fma(a02r, m2r, -a02i * m2i);
    double ci = fma(a00r, m0i, a00i * m0r) - fma(a01r, m1i, a01i * m1r) + fma(a02r, m2i, a02i * m2r);
    if ((i + j) & 1) { cr = -cr; ci = -ci; }
    // det = sum_c D[0][c] adj[c][0];  adj[c][0] lives in lane 4c
    const double b0r = __shfl_sync(0xffffffffu, cr, 0, 16), b0i = __shfl_sync(0xffffffffu, ci, 0, 16);
    const double b1r = __shfl_sync(0xffffffffu, cr, 4, 16), b1i = __shfl_sync(0xffffffffu, ci, 4, 16);
    const double b2r = __shfl_sync(0xffffffffu, cr, 8, 16), b2i = __shfl_sync(0xffffffffu, ci, 8, 16);
    const double b3r = __shfl_sync(0xffffffffu, cr, 12, 16), b3i = __shfl_sync(0xffffffffu, ci, 12, 16);
    const double detr = (fma(d0r, b0r, -d0i * b0i) + fma(d1r, b1r, -d1i * b1i)) + (fma(d2r, b2r, -d2i * b2i) + fma(d3r, b3r, -d3i * b3i));
    const double deti = (fma(d0r, b0i, d0i * b0r) + fma(d1r, b1i, d1i * b1r)) + (fma(d2r, b2i, d2i * b2r) + fma(d3r, b3i, d3i * b3r));
    det = make_double2(detr, deti);
    if (SCALED) {
        const double y = rcp_newton2(fma(detr, detr, deti * deti));
        const double ivr = detr * y, ivi = -deti * y;                      // 1 / det
        const double pr = fma(cr, ivr, -ci * ivi), pi = fma(cr, ivi, ci * ivr);
        if (x.lane < 16) gs->u.p.P[x.part][x.lane] = make_double2(pr, pi);
    } else {
        if (x.lane < 16) gs->u.p.P[x.part][x.lane] = make_double2(cr, ci);
    }
    __syncwarp();
}

// ---- P = D^-1 by cofactors with the Re / Im parts of every entry split over the two half-warps: lane (s, i, j) produces the
//      component s (0: Re, 1: Im) of P[i][j].  A scalar FP64 instruction costs the warp a turn of the FP64 pipe no matter how many
//      lanes are live, and while the other warps of the sub-partition stream 16-cycle DMMAs such a turn comes up about once per
//      DMMA: what the inverse costs is its NUMBER of FP64 instructions, and mirrored lanes are wasted turns.  With
//      (q1, q2) = (Re q, -Im q) in the Re lanes and (Im q, Re q) in the Im lanes, the lane's component of a complex product p q
//      is p.re q1 + p.im q2: 2 instructions instead of 4, the same code in both halves.  Operands that come from the panel are
//      loaded in that form straight from the Re / Im planes; computed operands take their other component from lane ^ 16.
//      det = (D adj)[j][j] = sum_i D[j][i] adj[i][j]: one term per lane, summed over i by two butterfly rounds.
//      31 FP64 instructions + 14 shuffles per block step against 84 + 16 for mma_inverse4_adj.
__device__ __forceinline__ void mma_inverse4_split(const MmaCtx& x, const double* __restrict__ Dr, const double* __restrict__ Di) {
    MmaGroupSmem* gs = x.gs;
    const int s = x.lane >> 4, i = (x.lane >> 2) & 3, j = x.lane & 3;
    const double* D1 = s ? Di : Dr;                    // plane of the component this lane produces
    const double* D2 = s ? Dr : Di;
    const int neg = s ? 0 : (int)0x80000000;           // the Re lanes take -Im of a second operand
    const int r0 = (0 >= j) ? 1 : 0, r1 = (1 >= j) ? 2 : 1, r2 = (2 >= j) ? 3 : 2;
    const int c0 = (0 >= i) ? 1 : 0, c1 = (1 >= i) ? 2 : 1, c2 = (2 >= i) ? 3 : 2;
#define HS_LD1(R, C, vr, vi) const double vr = Dr[(R) * 4 + (C)], vi = Di[(R) * 4 + (C)]                              /* first operand: (re, im) */
#define HS_LD2(R, C, v1, v2) const double v1 = D1[(R) * 4 + (C)], v2 = flip_sign(D2[(R) * 4 + (C)], neg)             /* second operand form */
    HS_LD1(r1, c0, a10r, a10i); HS_LD1(r1, c1, a11r, a11i); HS_LD1(r1, c2, a12r, a12i);
    HS_LD2(r2, c0, a20p, a20q); HS_LD2(r2, c1, a21p, a21q); HS_LD2(r2, c2, a22p, a22q);
    HS_LD1(r0, c0, a00r, a00i); HS_LD1(r0, c1, a01r, a01i); HS_LD1(r0, c2, a02r, a02i);
    HS_LD1(j, i, djr, dji);
#undef HS_LD1
#undef HS_LD2
    // own component of the 2 x 2 minors of rows r1, r2
    const double m0 = fma(-a12i, a21q, fma(-a12r, a21p, fma(a11i, a22q, a11r * a22p)));
    const double m1 = fma(-a12i, a20q, fma(-a12r, a20p, fma(a10i, a22q, a10r * a22p)));
    const double m2 = fma(-a11i, a20q, fma(-a11r, a20p, fma(a10i, a21q, a10r * a21p)));
    const double m0o = flip_sign(__shfl_xor_sync(0xffffffffu, m0, 16), neg);
    const double m1o = flip_sign(__shfl_xor_sync(0xffffffffu, m1, 16), neg);
    const double m2o = flip_sign(__shfl_xor_sync(0xffffffffu, m2, 16), neg);
    // own component of the signed cofactor  a00 m0 - a01 m1 + a02 m2
    double cf = fma(a02i, m2o, fma(a02r, m2, fma(-a01i, m1o, fma(-a01r, m1, fma(a00i, m0o, a00r * m0)))));
    cf = flip_sign(cf, ((i + j) & 1) ? (int)0x80000000 : 0);
    const double cfo = __shfl_xor_sync(0xffffffffu, cf, 16);           // the other component of the own cofactor
    // det = sum_i D[j][i] adj[i][j]
    double dt = fma(dji, flip_sign(cfo, neg), djr * cf);
    dt += __shfl_xor_sync(0xffffffffu, dt, 4);
    dt += __shfl_xor_sync(0xffffffffu, dt, 8);
    const double dto = __shfl_xor_sync(0xffffffffu, dt, 16);
    const double y = rcp_newton2(fma(dt, dt, dto * dto));
    // P = adj conj(det) / |det|^2:  Re = cr dr + ci di,  Im = ci dr - cr di
    const double u1 = s ? dto : dt, u2 = s ? -dt : dto;
    const double pv = fma(cfo, u2, cf * u1) * y;
    reinterpret_cast<double*>(gs->u.p.P[x.part])[2 * (x.lane & 15) + s] = pv;
    __syncwarp();
}

// ---- P = D^-1 with (almost) no scalar FP64 work: FP32 seed + two Newton-Schulz steps on the tensor pipe.
//      While the other warps of the sub-partition stream DMMAs a dependent scalar FP64 instruction waits for the FP64 pipe
//      about as long as two DMMAs take (ncu: math_pipe_throttle on every DFMA of the cofactor chain, 44 % of a block step), so the
//      31-instruction chain of mma_inverse4_split is the most expensive part of the step.  Here the chain runs on the FP32 pipe,
//      which nothing else uses: the pivot block is rounded to FP32 (one entry component per lane), inverted by the same split
//      cofactor formula in FP32 (X0: relative error ~1e-7 cond(D)), and refined in FP64 by
//          R = I - X D,   X <- X + R X          (error -> error^2 per step; two steps: ~1e-7^4 cond^4, i.e. rounding level)
//      as four DMMAs per step: X lives in the accumulator layout with Re / Im interleaved by column, which IS the pair of A
//      fragments (Re X, Im X) of the next product, exactly like the L panel below; the B-side copy of X (lane (g4, t4) <- X[t4][g4 >> 1])
//      is one shuffle.  Rows 4..7 of the 8-row fragments mirror rows 0..3.  A block too ill-conditioned for the FP32 seed
//      (cond(D) >~ 1e5) does not converge; the a-posteriori check then flags the matrix for the pivoted kernel like any other failure.
//      Returns pk = P[g4 & 3][t4] (the lane's own accumulator) and pv = P[t4][g4 >> 1] without a round trip through shared memory.
//      Measured (599 cfg2 windows): 4.85 ms against 4.81 ms for mma_inverse4_split, 36 instead of 49 matrices flagged -- inside the block
//      steps the FP64 pipe is already ~85 % busy (ncu), so trading 31 scalar instructions for 8 DMMAs buys nothing; selected by
//      HS_K5_ADJ=4 in HS_EXPERIMENT builds.
__device__ __forceinline__ float flip_sign_f(const float v, const int mask) { return __int_as_float(__float_as_int(v) ^ mask); }

__device__ __forceinline__ void mma_inverse4_newton(const MmaCtx& x, const int K0, const int buf, double2& pv, double2& pk) {
    MmaGroupSmem* gs = x.gs;
    const int s = x.lane >> 4, i = (x.lane >> 2) & 3, j = x.lane & 3;
    const double* Dr = gs->u.p.Craw[buf][0] + K0 * 4;
    const double* Di = gs->u.p.Craw[buf][1] + K0 * 4;
    float* Df = gs->u.p.Df[x.part];
    Df[x.lane] = (float)(s ? Di : Dr)[x.lane & 15];
    // D as the B operand of X D:  column n = 2 j' + c of the product is the Re (c = 0) / Im (c = 1) part of complex column j'
    const bool odd = x.g4 & 1;
    const int bsel = x.t4 * 4 + (x.g4 >> 1);
    const double dBr = Dr[bsel], dBi = Di[bsel];
    const double dB1 = odd ? dBi : dBr;                                                     // multiplies Re X
    const double dB2 = flip_sign(odd ? dBr : dBi, odd ? 0 : (int)0x80000000);                // multiplies Im X
    __syncwarp();
    // ---- FP32 seed: component s of adj(D)[i][j] / det D, same scheme as mma_inverse4_split
    float x0;
    {
        const float* F1 = Df + 16 * s;
        const float* F2 = Df + 16 * (s ^ 1);
        const int neg = s ? 0 : (int)0x80000000;
        const int r0 = (0 >= j) ? 1 : 0, r1 = (1 >= j) ? 2 : 1, r2 = (2 >= j) ? 3 : 2;
        const int c0 = (0 >= i) ? 1 : 0, c1 = (1 >= i) ? 2 : 1, c2 = (2 >= i) ? 3 : 2;
#define HS_LD1(R, C, vr, vi) const float vr = Df[(R) * 4 + (C)], vi = Df[16 + (R) * 4 + (C)]
#define HS_LD2(R, C, v1, v2) const float v1 = F1[(R) * 4 + (C)], v2 = flip_sign_f(F2[(R) * 4 + (C)], neg)
        HS_LD1(r1, c0, a10r, a10i); HS_LD1(r1, c1, a11r, a11i); HS_LD1(r1, c2, a12r, a12i);
        HS_LD2(r2, c0, a20p, a20q); HS_LD2(r2, c1, a21p, a21q); HS_LD2(r2, c2, a22p, a22q);
        HS_LD1(r0, c0, a00r, a00i); HS_LD1(r0, c1, a01r, a01i); HS_LD1(r0, c2, a02r, a02i);
        HS_LD1(j, i, djr, dji);
#undef HS_LD1
#undef HS_LD2
        const float m0 = fmaf(-a12i, a21q, fmaf(-a12r, a21p, fmaf(a11i, a22q, a11r * a22p)));
        const float m1 = fmaf(-a12i, a20q, fmaf(-a12r, a20p, fmaf(a10i, a22q, a10r * a22p)));
        const float m2 = fmaf(-a11i, a20q, fmaf(-a11r, a20p, fmaf(a10i, a21q, a10r * a21p)));
        const float m0o = flip_sign_f(__shfl_xor_sync(0xffffffffu, m0, 16), neg);
        const float m1o = flip_sign_f(__shfl_xor_sync(0xffffffffu, m1, 16), neg);
        const float m2o = flip_sign_f(__shfl_xor_sync(0xffffffffu, m2, 16), neg);
        float cf = fmaf(a02i, m2o, fmaf(a02r, m2, fmaf(-a01i, m1o, fmaf(-a01r, m1, fmaf(a00i, m0o, a00r * m0)))));
        cf = flip_sign_f(cf, ((i + j) & 1) ? (int)0x80000000 : 0);
        const float cfo = __shfl_xor_sync(0xffffffffu, cf, 16);
        float dt = fmaf(dji, flip_sign_f(cfo, neg), djr * cf);
        dt += __shfl_xor_sync(0xffffffffu, dt, 4);
        dt += __shfl_xor_sync(0xffffffffu, dt, 8);
        const float dto = __shfl_xor_sync(0xffffffffu, dt, 16);
        const float y = __fdividef(1.0f, fmaf(dt, dt, dto * dto));
        const float u1 = s ? dto : dt, u2 = s ? -dt : dto;
        x0 = fmaf(cfo, u2, cf * u1) * y;
    }
    // ---- X0 into the accumulator layout: lane (g4, t4) <- X0[g4 & 3][t4] (Re, Im), from the lanes (s, i, j) = (0 | 1, g4 & 3, t4)
    const int own = (x.g4 & 3) * 4 + x.t4;
    double ar = (double)__shfl_sync(0xffffffffu, x0, own);
    double ai = (double)__shfl_sync(0xffffffffu, x0, own + 16);
    const double idr = ((x.g4 & 3) == x.t4) ? 1.0 : 0.0;
    double b1, b2;      // X as the B operand of R X
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const double tr = __shfl_sync(0xffffffffu, ar, bsel), ti = __shfl_sync(0xffffffffu, ai, bsel);
        b1 = odd ? ti : tr;
        b2 = flip_sign(odd ? tr : ti, odd ? 0 : (int)0x80000000);
        double q0 = idr, q1 = 0.0;
        dmma884(q0, q1, ar, -dB1);              // R = I - X D
        dmma884(q0, q1, ai, -dB2);
        dmma884(ar, ai, q0, b1);                // X += R X
        dmma884(ar, ai, q1, b2);
    }
    pk = make_double2(ar, ai);
    pv = make_double2(__shfl_sync(0xffffffffu, ar, bsel), __shfl_sync(0xffffffffu, ai, bsel));
}

// ---- the two block steps of tile t (h = 0, 1): hand-over of the panel, 4 x 4 inverse, L panel by DMMA, rank-4 update
template <int T, int t, int ADJ>
__device__ __forceinline__ void mma_tile_steps(double (&c)[T][T][2], const int m, const MmaCtx& x) {
    MmaGroupSmem* gs = x.gs;
    const int fo = x.g4 * 4 + x.t4;           // A-fragment offset inside a tile row
    const int sgn = x.part ? 0 : (int)0x80000000;       // Re warp needs +Li = -(NL_i)
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
        const int K0 = 8 * t + 4 * h;
        if (K0 >= m) break;
        mma_extract<T, t>(c, h, x);
        mma_group_sync(x);
        double2 det = make_double2(1.0, 0.0);      // ADJ == 2: P holds adj(D), L comes out multiplied by det and U is divided by it below
        double2 pv, pk;       // P[t4][g4 >> 1] (B fragments of -P) and P[g4 & 3][t4] (rows K of the panel)
#ifdef HS_EXPERIMENT
        if (x.dbg & 1) {      // experiment: no 4 x 4 inverse (timing only, results are wrong)
            if (x.lane < 16) gs->u.p.P[x.part][x.lane] = make_double2((x.lane % 5 == 0) ? 1.0 : 0.0, 0.0);
            __syncwarp();
            pv = gs->u.p.P[x.part][x.t4 * 4 + (x.g4 >> 1)];
            pk = gs->u.p.P[x.part][(x.g4 & 3) * 4 + x.t4];
        } else
#endif
        if (ADJ == 4) {
            mma_inverse4_newton(x, K0, h, pv, pk);
        } else {
            if (ADJ == 3) mma_inverse4_split(x, gs->u.p.Craw[h][0] + K0 * 4, gs->u.p.Craw[h][1] + K0 * 4);
            else if (ADJ == 2) mma_inverse4_adj<false>(x, K0, h, det);
            else if (ADJ == 1) mma_inverse4_adj<true>(x, K0, h, det);
            else mma_inverse4(x, K0, h);
            pv = gs->u.p.P[x.part][x.t4 * 4 + (x.g4 >> 1)];
            pk = gs->u.p.P[x.part][(x.g4 & 3) * 4 + x.t4];
        }
        // B fragments of -P with Re/Im interleaved by output column:  n = 2j -> Re, n = 2j+1 -> Im
        double a0[T], a1[T];
        pipe_lock(x);
        {
            const bool odd = x.g4 & 1;
            const double bB1 = odd ? -pv.y : -pv.x;        // multiplies Re C
            const double bB2 = odd ? -pv.x : pv.y;         // multiplies Im C
            const double* Cr = gs->u.p.Craw[h][0] + fo;
            const double* Ci = gs->u.p.Craw[h][1] + fo;
#pragma unroll
            for (int ta = 0; ta < T; ++ta) {
                double l0 = 0.0, l1 = 0.0;
                dmma884(l0, l1, Cr[32 * ta], bB1);
                dmma884(l0, l1, Ci[32 * ta], bB2);
                a0[ta] = l0;
                a1[ta] = l1;
            }
            // rows K of the panel: -L = P - I
            if ((x.g4 >> 2) == h) {
                const bool dg = (x.g4 & 3) == x.t4;
                a0[t] = pk.x - (dg ? (ADJ == 2 ? det.x : 1.0) : 0.0);
                a1[t] = pk.y - ((ADJ == 2 && dg) ? det.y : 0.0);
            }
#pragma unroll
            for (int ta = 0; ta < T; ++ta) a1[ta] = flip_sign(a1[ta], sgn);
        }
        {
            const double* Ua = gs->u.p.Rraw[h][x.part] + x.t4 * kUS + x.g4;
            const double* Ub = gs->u.p.Rraw[h][x.part ^ 1] + x.t4 * kUS + x.g4;
            double b0[T], b1[T];
#pragma unroll
            for (int tb = 0; tb < T; ++tb) {
                b0[tb] = Ua[8 * tb];
                b1[tb] = Ub[8 * tb];
            }
            if (ADJ == 2) {
                // U / det: the reciprocal chain runs next to the L-panel DMMAs above instead of in front of them.
                // (b0, b1) = (Re U, Im U) in the Re warp and (Im U, Re U) in the Im warp.
                const double y = rcp_newton2(fma(det.x, det.x, det.y * det.y));
                const double yr = det.x * y, yi = -det.y * y;
#pragma unroll
                for (int tb = 0; tb < T; ++tb) {
                    const double ur = x.part ? b1[tb] : b0[tb], ui = x.part ? b0[tb] : b1[tb];
                    const double sr = fma(ur, yr, -ui * yi), si = fma(ur, yi, ui * yr);
                    b0[tb] = x.part ? si : sr;
                    b1[tb] = x.part ? sr : si;
                }
            }
#pragma unroll
            for (int ta = 0; ta < T; ++ta) {
#pragma unroll
                for (int tb = 0; tb < T; ++tb) {
                    dmma884(c[ta][tb][0], c[ta][tb][1], a0[ta], b0[tb]);
                    dmma884(c[ta][tb][0], c[ta][tb][1], a1[ta], b1[tb]);
                }
            }
        }
        pipe_unlock(x);
    }
}

// ---- ADJ == 5: the same two block steps with the pivot-block inverse taken ONE STEP AHEAD inside the warp.  The diagonal tile that
//      holds the next pivot block is updated first; its 4 x 4 block (own component) goes to gs->Dblk, one group barrier later both
//      warps hold both components and run the cofactor inverse for the NEXT step -- 31 FP64 instructions + 14 shuffles that depend on
//      nothing else in flight -- while the other 48 update DMMAs of THIS step are issued around them.  The chain's turns at the FP64 pipe
//      (50 - 100 cycles each behind the other warps' DMMAs) no longer sit between a step's hand-over and its first DMMA.  Same
//      operations on the same values as ADJ == 3: bit-identical results.
//      Written plainly (ADJ == 5, TIED = false) this does not happen: ptxas believes a DFMA to be short, keeps the whole chain in front
//      of the 48 independent DMMAs (~200 chain instructions, then 46 DMMAs in a row in the SASS) and the variant is slower (6.11 vs
//      4.91 ms with 6 groups and the spills of the longer live ranges, 5.16 vs 5.09 ms with 4 groups and none).  ADJ == 6 (TIED = true)
//      therefore cuts the chain into five pieces and gives each piece a lane / address offset computed from the last accumulator of
//      the DMMA batch (one column of the update, its two B fragments loaded in front of it: 16 registers less than keeping all ten)
//      that precedes it in the source: a value that is always 0 but that the compiler cannot fold (tie_zero), i.e. a true data
//      dependency.  The SASS then alternates chain instructions and DMMAs.  Measured: in the HS_EXPERIMENT build 4.98 -> 4.82 ms with
//      6 groups and 5.09 -> 4.79 ms with 4; compiled as the product kernel 4.91 -> 4.91 / 4.94 ms (two block steps per tile unrolled:
//      9000 instead of 7100 instructions, 240 instead of 76 bytes of spill stores).  With the chain hidden the kernel is no faster
//      than with the chain exposed: the FP64 pipe's turns, not the chain's latency, are what a block step costs.  Both variants stay
//      in HS_EXPERIMENT builds (HS_K5_ADJ = 5 / 6); the product kernel is ADJ == 3.
template <int T, int tn>
__device__ __forceinline__ void mma_publish_next(const double (&c)[T][T][2], const int hn, const MmaCtx& x) {
    if ((x.g4 >> 2) == hn && (x.t4 >> 1) == hn)
        *reinterpret_cast<double2*>(&x.gs->Dblk[x.part][(x.g4 & 3) * 4 + 2 * (x.t4 & 1)]) = make_double2(c[tn][tn][0], c[tn][tn][1]);
}

// An integer that is always 0 but that the compiler cannot prove to be (squares are 0 or 1 mod 4): a data dependency on `v` without a
// change of value.  Used by the TIED look-ahead to pin pieces of the inverse chain BEHIND batches of update DMMAs in the SASS.
__device__ __forceinline__ int tie_zero(const double v) {
    const int hi = __double2hiint(v);
    return (hi * hi) & 2;
}

template <int T, int t, bool TIED>
__device__ __forceinline__ void mma_tile_steps_la(double (&c)[T][T][2], const int m, const MmaCtx& x) {
    MmaGroupSmem* gs = x.gs;
    const int fo = x.g4 * 4 + x.t4;
    const int sgn = x.part ? 0 : (int)0x80000000;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int K0 = 8 * t + 4 * h;
        if (K0 >= m) break;
        mma_extract<T, t>(c, h, x);
        mma_group_sync(x);
        if (t == 0 && h == 0) mma_inverse4_split(x, gs->u.p.Craw[0][0], gs->u.p.Craw[0][1]);      // nothing to look ahead from
        const double2 pv = gs->u.p.P[x.part][x.t4 * 4 + (x.g4 >> 1)];
        const double2 pk = gs->u.p.P[x.part][(x.g4 & 3) * 4 + x.t4];
        double a0[T], a1[T];
        {
            const bool odd = x.g4 & 1;
            const double bB1 = odd ? -pv.y : -pv.x;
            const double bB2 = odd ? -pv.x : pv.y;
            const double* Cr = gs->u.p.Craw[h][0] + fo;
            const double* Ci = gs->u.p.Craw[h][1] + fo;
#pragma unroll
            for (int ta = 0; ta < T; ++ta) {
                double l0 = 0.0, l1 = 0.0;
                dmma884(l0, l1, Cr[32 * ta], bB1);
                dmma884(l0, l1, Ci[32 * ta], bB2);
                a0[ta] = l0;
                a1[ta] = l1;
            }
            if ((x.g4 >> 2) == h) {
                const bool dg = (x.g4 & 3) == x.t4;
                a0[t] = pk.x - (dg ? 1.0 : 0.0);
                a1[t] = pk.y;
            }
#pragma unroll
            for (int ta = 0; ta < T; ++ta) a1[ta] = flip_sign(a1[ta], sgn);
        }
        const double* Ua = gs->u.p.Rraw[h][x.part] + x.t4 * kUS + x.g4;
        const double* Ub = gs->u.p.Rraw[h][x.part ^ 1] + x.t4 * kUS + x.g4;
        // the diagonal tile that holds the NEXT pivot block first, then that block's inverse next to the rest of the update
        constexpr bool kLastTile = (t + 1 >= T);
        const int tn = (h == 0 || kLastTile) ? t : t + 1;
        // T = ceil(m / 8): every block of the tiles before the last one has a successor (known at compile time: one code path)
        const bool has_next = !kLastTile || ((K0 + 4 < m) && h == 0);
        double b0[T], b1[T];
        if (!TIED || !has_next) {       // the tied path loads a column's two fragments in front of that column's batch: 16 registers less
#pragma unroll
            for (int tb = 0; tb < T; ++tb) {
                b0[tb] = Ua[8 * tb];
                b1[tb] = Ub[8 * tb];
            }
        }
        if (has_next) {
            if (h == 0) {
                const double u0 = TIED ? Ua[8 * t] : b0[t], u1 = TIED ? Ub[8 * t] : b1[t];
                dmma884(c[t][t][0], c[t][t][1], a0[t], u0);
                dmma884(c[t][t][0], c[t][t][1], a1[t], u1);
                mma_publish_next<T, t>(c, 1, x);
            } else if constexpr (!kLastTile) {
                const double u0 = TIED ? Ua[8 * (t + 1)] : b0[t + 1], u1 = TIED ? Ub[8 * (t + 1)] : b1[t + 1];
                dmma884(c[t + 1][t + 1][0], c[t + 1][t + 1][1], a0[t + 1], u0);
                dmma884(c[t + 1][t + 1][0], c[t + 1][t + 1][1], a1[t + 1], u1);
                mma_publish_next<T, t + 1>(c, 0, x);
            }
            mma_group_sync(x);                     // both components of the next block are in Dblk (and this warp is past its pv / pk loads)
            if (!TIED) mma_inverse4_split(x, gs->Dblk[0], gs->Dblk[1]);
        }
        if (!TIED || !has_next) {
#pragma unroll
            for (int ta = 0; ta < T; ++ta) {
#pragma unroll
                for (int tb = 0; tb < T; ++tb) {
                    if (has_next && ta == tb && ta == tn) continue;      // done above
                    dmma884(c[ta][tb][0], c[ta][tb][1], a0[ta], b0[tb]);
                    dmma884(c[ta][tb][0], c[ta][tb][1], a1[ta], b1[tb]);
                }
            }
        } else {
            // the inverse chain of mma_inverse4_split in five pieces, a batch of update DMMAs in front of each of the last four; piece k
            // takes a (zero) lane / address offset from the last accumulator of batch k - 1, so it cannot be scheduled before that batch
            const int s2 = x.lane >> 4, ii = (x.lane >> 2) & 3, jj = x.lane & 3;
            const double* Dr = gs->Dblk[0];
            const double* Di = gs->Dblk[1];
            const double* D1 = s2 ? Di : Dr;
            const double* D2 = s2 ? Dr : Di;
            const int neg = s2 ? 0 : (int)0x80000000;
            const int r0 = (0 >= jj) ? 1 : 0, r1 = (1 >= jj) ? 2 : 1, r2 = (2 >= jj) ? 3 : 2;
            const int c0 = (0 >= ii) ? 1 : 0, c1 = (1 >= ii) ? 2 : 1, c2 = (2 >= ii) ? 3 : 2;
            double last = 0.0;
            auto batch = [&](const int tb) {              // column tb of the update (without the tile done above)
                if (tb < T) {
                    const double u0 = Ua[8 * tb], u1 = Ub[8 * tb];
#pragma unroll
                    for (int ta = 0; ta < T; ++ta) {
                        if (ta == tb && ta == tn) continue;
                        dmma884(c[ta][tb][0], c[ta][tb][1], a0[ta], u0);
                        dmma884(c[ta][tb][0], c[ta][tb][1], a1[ta], u1);
                        last = c[ta][tb][0];
                    }
                }
            };
#define HS_LD1(R, C, vr, vi) const double vr = Dr[(R) * 4 + (C)], vi = Di[(R) * 4 + (C)]
#define HS_LD2(R, C, v1, v2) const double v1 = D1[(R) * 4 + (C)], v2 = flip_sign(D2[(R) * 4 + (C)], neg)
            // piece 0: operands and the 2 x 2 minors
            HS_LD1(r1, c0, a10r, a10i); HS_LD1(r1, c1, a11r, a11i); HS_LD1(r1, c2, a12r, a12i);
            HS_LD2(r2, c0, a20p, a20q); HS_LD2(r2, c1, a21p, a21q); HS_LD2(r2, c2, a22p, a22q);
            const double m0 = fma(-a12i, a21q, fma(-a12r, a21p, fma(a11i, a22q, a11r * a22p)));
            const double m1 = fma(-a12i, a20q, fma(-a12r, a20p, fma(a10i, a22q, a10r * a22p)));
            const double m2 = fma(-a11i, a20q, fma(-a11r, a20p, fma(a10i, a21q, a10r * a21p)));
            batch(0);
            // piece 1: the other component of the minors, the signed cofactor (its operands are loaded here, behind the tie: short live ranges)
            int z = tie_zero(last);
            Dr += z;
            Di += z;
            HS_LD1(r0, c0, a00r, a00i); HS_LD1(r0, c1, a01r, a01i); HS_LD1(r0, c2, a02r, a02i);
            const double m0o = flip_sign(__shfl_xor_sync(0xffffffffu, m0, 16 + z), neg);
            const double m1o = flip_sign(__shfl_xor_sync(0xffffffffu, m1, 16 + z), neg);
            const double m2o = flip_sign(__shfl_xor_sync(0xffffffffu, m2, 16 + z), neg);
            double cf = fma(a02i, m2o, fma(a02r, m2, fma(-a01i, m1o, fma(-a01r, m1, fma(a00i, m0o, a00r * m0)))));
            cf = flip_sign(cf, ((ii + jj) & 1) ? (int)0x80000000 : 0);
            batch(1);
            // piece 2: determinant
            z = tie_zero(last);
            Dr += z;
            Di += z;
            HS_LD1(jj, ii, djr, dji);
#undef HS_LD1
#undef HS_LD2
            const double cfo = __shfl_xor_sync(0xffffffffu, cf, 16 + z);
            double dt = fma(dji, flip_sign(cfo, neg), djr * cf);
            dt += __shfl_xor_sync(0xffffffffu, dt, 4);
            dt += __shfl_xor_sync(0xffffffffu, dt, 8);
            batch(2);
            // piece 3: 1 / |det|^2
            z = tie_zero(last);
            const double dto = __shfl_xor_sync(0xffffffffu, dt, 16 + z);
            const double y = rcp_newton2(fma(dt, dt, dto * dto));
            batch(3);
            // piece 4: P = adj conj(det) / |det|^2
            z = tie_zero(last);
            const double u1 = s2 ? dto : dt, u2 = s2 ? -dt : dto;
            const double pvn = fma(cfo, u2, cf * u1) * y;
            reinterpret_cast<double*>(gs->u.p.P[x.part])[2 * (x.lane & 15) + s2 + z] = pvn;
            __syncwarp();
#pragma unroll
            for (int tb = 4; tb < T; ++tb) batch(tb);
        }
    }
}

template <int T, int t, int ADJ>
__device__ __forceinline__ void mma_all_tiles(double (&c)[T][T][2], const int m, const MmaCtx& x) {
    if constexpr (t < T) {
        if constexpr (ADJ == 5) mma_tile_steps_la<T, t, false>(c, m, x);
        else if constexpr (ADJ == 6) mma_tile_steps_la<T, t, true>(c, m, x);
        else mma_tile_steps<T, t, ADJ>(c, m, x);
        mma_all_tiles<T, t + 1, ADJ>(c, m, x);
    }
}

// in-place inverse of the matrix held by the two warps of the group
template <int T, int ADJ>
__device__ __forceinline__ void mma_gauss_jordan(double (&c)[T][T][2], const int m, const MmaCtx& x) {
    mma_group_sync(x);                 // previous users of the panel buffers (assembly exchange) are done
    mma_all_tiles<T, 0, ADJ>(c, m, x);
}


// ---- A(f) = I - sum_k A_k z_k(f).  Each warp assembles ONE of the two entries every lane owns per tile
//      (Re warp: column 2*t4, Im warp: column 2*t4 + 1) as a complex number, keeps the part it owns and hands the
//      other part to the partner warp through gs->u.X: every coefficient is read from shared memory once per
//      matrix, and both warps run the same code (the part only selects operands).
template <int T, int NP>      // NP > 0: number of lag pairs known at compile time (z kept in registers, tile-major order)
__device__ __forceinline__ void mma_assemble(double (&c)[T][T][2], const double* __restrict__ coef, const double2* __restrict__ zf,
                                             const int n_planes, const MmaCtx& x) {
    const double* cbase = coef + x.g4 * kRowD + 4 * x.t4 + 2 * x.part;
    const double dg = (x.g4 == 2 * x.t4 + x.part) ? 1.0 : 0.0;      // identity (real part) of the entry this warp assembled
    double* X = reinterpret_cast<double*>(x.gs->u.X) + x.part * (25 * 32) + x.lane;          // outbox of this warp
    if constexpr (NP > 0) {
        double2 z[2 * NP];
#pragma unroll
        for (int k = 0; k < 2 * NP; ++k) z[k] = zf[k];       // shared memory, zero padded to an even number of lags
#pragma unroll
        for (int ta = 0; ta < T; ++ta) {
#pragma unroll
            for (int tb = 0; tb < T; ++tb) {
                double vr = (ta == tb) ? dg : 0.0, vi = 0.0;
#pragma unroll
                for (int kp = 0; kp < NP; ++kp) {
                    const double2 q = *reinterpret_cast<const double2*>(cbase + kp * kPlaneD + ta * 8 * kRowD + tb * 16);
                    vr = fma(-q.x, z[2 * kp].x, fma(-q.y, z[2 * kp + 1].x, vr));
                    vi = fma(-q.x, z[2 * kp].y, fma(-q.y, z[2 * kp + 1].y, vi));
                }
                X[(ta * T + tb) * 32] = x.part ? vr : vi;      // the part the partner owns
                if (x.part) c[ta][tb][1] = vi;
                else c[ta][tb][0] = vr;
            }
        }
    } else {
#pragma unroll
        for (int ta = 0; ta < T; ++ta)
#pragma unroll
            for (int tb = 0; tb < T; ++tb) c[ta][tb][0] = c[ta][tb][1] = 0.0;      // [0]: Re, [1]: Im of the assembled entry, for now
        for (int kp = 0; kp < n_planes; ++kp) {
            const double2 z0 = zf[2 * kp], z1 = zf[2 * kp + 1];
            const double* cp = cbase + kp * kPlaneD;
#pragma unroll
            for (int ta = 0; ta < T; ++ta) {
#pragma unroll
                for (int tb = 0; tb < T; ++tb) {
                    const double2 q = *reinterpret_cast<const double2*>(cp + ta * 8 * kRowD + tb * 16);
                    c[ta][tb][0] = fma(-q.x, z0.x, fma(-q.y, z1.x, c[ta][tb][0]));
                    c[ta][tb][1] = fma(-q.x, z0.y, fma(-q.y, z1.y, c[ta][tb][1]));
                }
            }
        }
#pragma unroll
        for (int ta = 0; ta < T; ++ta) {
#pragma unroll
            for (int tb = 0; tb < T; ++tb) {
                const double vr = c[ta][tb][0] + ((ta == tb) ? dg : 0.0), vi = c[ta][tb][1];
                X[(ta * T + tb) * 32] = x.part ? vr : vi;
                if (x.part) c[ta][tb][1] = vi;
                else c[ta][tb][0] = vr;
            }
        }
    }
    mma_group_sync(x);
    const double* Xo = reinterpret_cast<const double*>(x.gs->u.X) + (x.part ^ 1) * (25 * 32) + x.lane;      // the partner's outbox
#pragma unroll
    for (int ta = 0; ta < T; ++ta) {
#pragma unroll
        for (int tb = 0; tb < T; ++tb) {
            const double got = Xo[(ta * T + tb) * 32];
            if (x.part) c[ta][tb][0] = got;
            else c[ta][tb][1] = got;
        }
    }
}

// Outputs in the reference's (w, i, j, f) layout (A(f), H, or |H|^2 without the staging buffer): not the metric path.
// The register tiles must never have their address taken (the compiler would mirror them in local memory), so both
// warps dump their part to gs->u.X, one of the two entries per lane and tile at a time, and a compact loop with one
// thread per matrix row does the addressing.  FINAL: H / |H|^2 / row sums; otherwise A(f).
template <int T, bool FINAL>
__device__ __forceinline__ void mma_store_generic(const double (&c)[T][T][2], const K5Params& P, const int w, const int f, const MmaCtx& x) {
    const int m = P.m;
    const size_t sF = (size_t)P.F;
    const int l64 = x.part * 32 + x.lane;
    double* Xd = reinterpret_cast<double*>(x.gs->u.X);
    double rsum = 0.0;
#pragma unroll 1
    for (int e = 0; e < 2; ++e) {
        mma_group_sync(x);
#pragma unroll
        for (int q = 0; q < T * T; ++q) Xd[x.part * (25 * 32) + q * 32 + x.lane] = e ? c[q / T][q % T][1] : c[q / T][q % T][0];
        mma_group_sync(x);
        if (l64 < m) {
            const int i = l64, ta = i >> 3, g4 = i & 7;
            for (int j = e; j < m; j += 2) {
                const int tb = j >> 3, t4 = (j & 7) >> 1;
                const int off = (ta * T + tb) * 32 + g4 * 4 + t4;
                const double re = Xd[off], im = Xd[25 * 32 + off];
                const size_t o = (((size_t)w * m + i) * m + j) * sF + f;
                if (FINAL) {
                    const double v = fma(re, re, im * im);
                    rsum += v;
                    if (P.dtf) P.dtf[P.dtf_fij ? (((size_t)w * m + i) * sF + f) * m + j : o] = v;
                    if (P.H) P.H[o] = make_double2(re, im);
                } else {
                    P.Af[o] = make_double2(re, im);
                }
            }
        }
    }
    if (FINAL && l64 < m) x.gs->rs[0][l64] += rsum;
    mma_group_sync(x);
}

__device__ __forceinline__ double2 probe_u2(const int j) {
    return make_double2(1.0 + 0.03125 * j, ((j & 1) ? -1.0 : 1.0) * (0.5 + 0.015625 * j));
}


#ifdef HS_EXPERIMENT
// =====================================================================================
// Warp-specialised variant (transfer_ws_kernel).  In transfer_mma_kernel every dependent scalar FP64 instruction of the 4 x 4
// pivot-block inverse queues behind the other warps' 16-cycle DMMAs, and the two warps of a group cannot issue a DMMA until
// the chain (64 instructions, both warps redundantly) is through: 24 % of the kernel.  Here a group is THREE warps: the Re and
// Im warps only stream DMMAs; right after the L panel they update the diagonal tile that holds the NEXT pivot block, publish
// those 16 entries (bar_d) and go on with the other 48 update DMMAs, while the group's HELPER warp inverts the block (cofactors,
// one reciprocal) and hands P back (bar_p) before the main warps need it.  Register budgets follow the roles (setmaxnreg: the
// helper warpgroup gives registers to the main warpgroups), so the main warps hold the 50 accumulators + fragments without spills.
// =====================================================================================
__device__ __forceinline__ unsigned ws_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ws_mbar_init(unsigned long long* bar, const int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ws_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ws_mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ws_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool ws_mbar_test(unsigned long long* bar, const unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(ws_smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void ws_mbar_wait(unsigned long long* bar, const unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WS_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WS_WAIT_%=;\n"
        "}\n" ::"r"(ws_smem_u32(bar)), "r"(parity) : "memory");
}

// helper warp: P = D^-1 by cofactors from the published block (same arithmetic as mma_inverse4_adj<true>)
__device__ __forceinline__ void ws_inverse4(MmaGroupSmem* gs, const int lane) {
    const int i = (lane >> 2) & 3, j = lane & 3;
    const double* Dr = gs->Dblk[0];
    const double* Di = gs->Dblk[1];
    const int r0 = (0 >= j) ? 1 : 0, r1 = (1 >= j) ? 2 : 1, r2 = (2 >= j) ? 3 : 2;
    const int c0 = (0 >= i) ? 1 : 0, c1 = (1 >= i) ? 2 : 1, c2 = (2 >= i) ? 3 : 2;
#define HS_LD(R, C, vr, vi) const double vr = Dr[(R) * 4 + (C)], vi = Di[(R) * 4 + (C)]
    HS_LD(r0, c0, a00r, a00i); HS_LD(r0, c1, a01r, a01i); HS_LD(r0, c2, a02r, a02i);
    HS_LD(r1, c0, a10r, a10i); HS_LD(r1, c1, a11r, a11i); HS_LD(r1, c2, a12r, a12i);
    HS_LD(r2, c0, a20r, a20i); HS_LD(r2, c1, a21r, a21i); HS_LD(r2, c2, a22r, a22i);
    HS_LD(0, 0, d0r, d0i); HS_LD(0, 1, d1r, d1i); HS_LD(0, 2, d2r, d2i); HS_LD(0, 3, d3r, d3i);
#undef HS_LD
#define HS_DET2(pr, pi, qr, qi, sr, si, tr, ti, outr, outi)                                        \
    const double outr = fma(pr, qr, -pi * qi) - fma(sr, tr, -si * ti);                              \
    const double outi = fma(pr, qi, pi * qr) - fma(sr, ti, si * tr)
    HS_DET2(a11r, a11i, a22r, a22i, a12r, a12i, a21r, a21i, m0r, m0i);
    HS_DET2(a10r, a10i, a22r, a22i, a12r, a12i, a20r, a20i, m1r, m1i);
    HS_DET2(a10r, a10i, a21r, a21i, a11r, a11i, a20r, a20i, m2r, m2i);
#undef HS_DET2
    double cr = fma(a00r, m0r, -a00i * m0i) - fma(a01r, m1r, -a01i * m1i) + fma(a02r, m2r, -a02i * m2i);
    double ci = fma(a00r, m0i, a00i * m0r) - fma(a01r, m1i, a01i * m1r) + fma(a02r, m2i, a02i * m2r);
    if ((i + j) & 1) { cr = -cr; ci = -ci; }
    const double b0r = __shfl_sync(0xffffffffu, cr, 0, 16), b0i = __shfl_sync(0xffffffffu, ci, 0, 16);
    const double b1r = __shfl_sync(0xffffffffu, cr, 4, 16), b1i = __shfl_sync(0xffffffffu, ci, 4, 16);
    const double b2r = __shfl_sync(0xffffffffu, cr, 8, 16), b2i = __shfl_sync(0xffffffffu, ci, 8, 16);
    const double b3r = __shfl_sync(0xffffffffu, cr, 12, 16), b3i = __shfl_sync(0xffffffffu, ci, 12, 16);
    const double detr = (fma(d0r, b0r, -d0i * b0i) + fma(d1r, b1r, -d1i * b1i)) + (fma(d2r, b2r, -d2i * b2i) + fma(d3r, b3r, -d3i * b3i));
    const double deti = (fma(d0r, b0i, d0i * b0r) + fma(d1r, b1i, d1i * b1r)) + (fma(d2r, b2i, d2i * b2r) + fma(d3r, b3i, d3i * b3r));
    const double y = rcp_newton2(fma(detr, detr, deti * deti));
    const double ivr = detr * y, ivi = -deti * y;
    const double pr = fma(cr, ivr, -ci * ivi), pi = fma(cr, ivi, ci * ivr);
    if (lane < 16) gs->Ph[lane] = make_double2(pr, pi);
}

// main warps: publish the 4 x 4 block (half hn of tile (tn, tn)) for the helper
template <int T, int tn>
__device__ __forceinline__ void ws_publish(const double (&c)[T][T][2], const int hn, const MmaCtx& x) {
    if ((x.g4 >> 2) == hn && (x.t4 >> 1) == hn)
        *reinterpret_cast<double2*>(&x.gs->Dblk[x.part][(x.g4 & 3) * 4 + 2 * (x.t4 & 1)]) = make_double2(c[tn][tn][0], c[tn][tn][1]);
    __syncwarp();
    if (x.lane == 0) ws_mbar_arrive(&x.gs->bar_d);
}

template <int T, int t>
__device__ __forceinline__ void ws_tile_steps(double (&c)[T][T][2], const int m, const MmaCtx& x, unsigned& ph) {
    MmaGroupSmem* gs = x.gs;
    const int fo = x.g4 * 4 + x.t4;
    const int sgn = x.part ? 0 : (int)0x80000000;
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
        const int K0 = 8 * t + 4 * h;
        if (K0 >= m) break;
        mma_extract<T, t>(c, h, x);
        mma_group_sync(x);
        ws_mbar_wait(&gs->bar_p, ph & 1u);                 // P = D^-1 of this step, from the helper warp
        ++ph;
        double a0[T], a1[T];
        {
            const double2 pv = gs->Ph[x.t4 * 4 + (x.g4 >> 1)];
            const bool odd = x.g4 & 1;
            const double bB1 = odd ? -pv.y : -pv.x;
            const double bB2 = odd ? -pv.x : pv.y;
            const double* Cr = gs->u.p.Craw[h][0] + fo;
            const double* Ci = gs->u.p.Craw[h][1] + fo;
#pragma unroll
            for (int ta = 0; ta < T; ++ta) {
                double l0 = 0.0, l1 = 0.0;
                dmma884(l0, l1, Cr[32 * ta], bB1);
                dmma884(l0, l1, Ci[32 * ta], bB2);
                a0[ta] = l0;
                a1[ta] = l1;
            }
            const double2 pk = gs->Ph[(x.g4 & 3) * 4 + x.t4];
            if ((x.g4 >> 2) == h) {
                const bool dg = (x.g4 & 3) == x.t4;
                a0[t] = pk.x - (dg ? 1.0 : 0.0);
                a1[t] = pk.y;
            }
#pragma unroll
            for (int ta = 0; ta < T; ++ta) a1[ta] = flip_sign(a1[ta], sgn);
        }
        const double* Ua = gs->u.p.Rraw[h][x.part] + x.t4 * kUS + x.g4;
        const double* Ub = gs->u.p.Rraw[h][x.part ^ 1] + x.t4 * kUS + x.g4;
        double b0[T], b1[T];
#pragma unroll
        for (int tb = 0; tb < T; ++tb) {
            b0[tb] = Ua[8 * tb];
            b1[tb] = Ub[8 * tb];
        }
        // the diagonal tile that holds the NEXT pivot block first, so the helper can start on its inverse
        const bool has_next = (K0 + 4 < m);
        if (h == 0) {
            dmma884(c[t][t][0], c[t][t][1], a0[t], b0[t]);
            dmma884(c[t][t][0], c[t][t][1], a1[t], b1[t]);
            if (has_next) ws_publish<T, t>(c, 1, x);
        } else if constexpr (t + 1 < T) {
            dmma884(c[t + 1][t + 1][0], c[t + 1][t + 1][1], a0[t + 1], b0[t + 1]);
            dmma884(c[t + 1][t + 1][0], c[t + 1][t + 1][1], a1[t + 1], b1[t + 1]);
            if (has_next) ws_publish<T, t + 1>(c, 0, x);
        }
#pragma unroll
        for (int ta = 0; ta < T; ++ta) {
#pragma unroll
            for (int tb = 0; tb < T; ++tb) {
                if (ta == tb && ((h == 0 && ta == t) || (h == 1 && ta == t + 1))) continue;      // done above
                dmma884(c[ta][tb][0], c[ta][tb][1], a0[ta], b0[tb]);
                dmma884(c[ta][tb][0], c[ta][tb][1], a1[ta], b1[tb]);
            }
        }
    }
}

template <int T, int t>
__device__ __forceinline__ void ws_all_tiles(double (&c)[T][T][2], const int m, const MmaCtx& x, unsigned& ph) {
    if constexpr (t < T) {
        ws_tile_steps<T, t>(c, m, x, ph);
        ws_all_tiles<T, t + 1>(c, m, x, ph);
    }
}

#endif  // HS_EXPERIMENT (warp-specialised device functions)

template <int T>
struct MmaSmem {
    static __host__ __device__ int planes(int p) { return (p + 1) / 2; }
    static __host__ __device__ size_t coef_bytes(int p) { return (size_t)planes(p) * kPlaneD * sizeof(double); }
    static __host__ __device__ size_t probe_bytes(int p) { return (size_t)kMP * p * sizeof(double2); }
    static __host__ __device__ size_t z_bytes(int p, int seg_len) { return (size_t)seg_len * 2 * planes(p) * sizeof(double2); }    // z of the unit's bins
    static __host__ __device__ size_t chk_bytes() { return (size_t)2 * kMP * sizeof(double2); }      // [part][row] {expected pad-column value, limit}
    static __host__ __device__ size_t total(int p, int ng, int seg_len) {
        return coef_bytes(p) + probe_bytes(p) + z_bytes(p, seg_len) + chk_bytes() + (size_t)ng * sizeof(MmaGroupSmem) + 64;
    }
};

template <int T, int NG, int ADJ>
__global__ void __launch_bounds__(NG * 64, 1) transfer_mma_kernel(const K5Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int m = P.m, p = P.p, F = P.F;
    double* coef = reinterpret_cast<double*>(smem_raw);
    double2* probe = reinterpret_cast<double2*>(smem_raw + MmaSmem<T>::coef_bytes(p));
    double2* zs = reinterpret_cast<double2*>(smem_raw + MmaSmem<T>::coef_bytes(p) + MmaSmem<T>::probe_bytes(p));      // [bin of the unit][2 * n_planes]
    double2* chk = reinterpret_cast<double2*>(smem_raw + MmaSmem<T>::coef_bytes(p) + MmaSmem<T>::probe_bytes(p) + MmaSmem<T>::z_bytes(p, P.seg_len));
    MmaGroupSmem* groups = reinterpret_cast<MmaGroupSmem*>(smem_raw + MmaSmem<T>::coef_bytes(p) + MmaSmem<T>::probe_bytes(p) + MmaSmem<T>::z_bytes(p, P.seg_len) +
                                                           MmaSmem<T>::chk_bytes());
    const int warp = threadIdx.x >> 5;
    MmaCtx x;
    x.lane = threadIdx.x & 31;
    x.part = warp & 1;
    x.g4 = x.lane >> 2;
    x.t4 = x.lane & 3;
    const int gid = warp >> 1;
    x.bar = 1 + gid;
    x.dbg = (P.flip >> 27) & 1;
    x.gs = groups + gid;
#ifdef HS_EXPERIMENT
    __shared__ int pipe_locks[4];
    if (threadIdx.x < 4) pipe_locks[threadIdx.x] = 0;
    x.lock = P.pipe_turns ? &pipe_locks[warp & 3] : nullptr;       // warp w runs on sub-partition w % 4
#else
    x.lock = nullptr;
#endif
    MmaGroupSmem* gs = x.gs;
    const int l64 = x.part * 32 + x.lane;
    const int n_planes = MmaSmem<T>::planes(p);

    // zero the coefficient planes once: padding rows / columns / the odd lag stay zero for every unit
    for (int e = threadIdx.x; e < n_planes * kPlaneD; e += NG * 64) coef[e] = 0.0;
    // A-posteriori check through the padding: when the last column of the padded matrix is free (m < 8 T) it carries v = A(f) u
    // through the elimination like a right-hand side ([[A, v], [0, 1]]^-1 = [[H, -H v], [0, 1]]), so H v comes out of the update DMMAs
    // and only has to be compared with u: -u if the pad column lies inside the last eliminated block (8 T - 4 < m), else +u.
    const bool padchk = (m < 8 * T);
    if (threadIdx.x < 2 * kMP) {
        const int part = threadIdx.x / kMP, row = threadIdx.x - part * kMP;
        const double2 u = probe_u2(row);
        const double want = part ? u.y : u.x;
        chk[threadIdx.x] = make_double2((8 * T - 4 < m) ? -want : want, 0.5 * P.verify_tol2 * fma(u.x, u.x, u.y * u.y));
    }

    // Balanced static partition: CTA c owns the matrices q = w * F + f in [c * per_cta, (c + 1) * per_cta): every CTA gets the
    // same count (+-1 round of NG), and a window's coefficients are loaded once per PIECE (= the part of a window inside the range).
    const long long q_total = (long long)P.n_win * F;
    long long q = (long long)blockIdx.x * P.per_cta;
    const long long q_end = min(q_total, q + (long long)P.per_cta);
    while (q < q_end) {
        const int w = (int)(q / F);
        const int f_begin = (int)(q - (long long)w * F);
        const int f_end = (int)min((long long)F, f_begin + (q_end - q));
        q += f_end - f_begin;
        __syncthreads();
        {   // AR coefficients of window w -> shared, layout [lag pair][row][col][lag & 1]
            const double* Aw = P.A + (size_t)w * m * m * p;
            const int row_len = m * p;
            for (int i = warp; i < m; i += NG * 2) {       // one warp per row, 5 independent loads in flight per lane
                const double* Ar = Aw + (size_t)i * row_len;
                for (int c0 = x.lane; c0 < row_len; c0 += 5 * 32) {
                    double v[5];
#pragma unroll
                    for (int u = 0; u < 5; ++u) v[u] = (c0 + 32 * u < row_len) ? Ar[c0 + 32 * u] : 0.0;
#pragma unroll
                    for (int u = 0; u < 5; ++u) {
                        const int cidx = c0 + 32 * u;
                        if (cidx < row_len) {
                            const int j = cidx / p, k = cidx - j * p;
                            coef[(k >> 1) * kPlaneD + i * kRowD + 2 * j + (k & 1)] = v[u];
                        }
                    }
                }
            }
            for (int e = l64; e < 2 * kMP; e += 64) (&gs->rs[0][0])[e] = 0.0;
        }
        __syncthreads();
        // probe[i][k] = sum_j A_k[i][j] u_j : the window-dependent part of v = A(f) u = u - sum_k z_k(f) probe[.][k]
        for (int e = threadIdx.x; e < kMP * p; e += NG * 64) {
            const int i = e / p, k = e - i * p;
            double sr = 0.0, si = 0.0;
            const double* cp = coef + (k >> 1) * kPlaneD + i * kRowD + (k & 1);
            for (int j = 0; j < m; ++j) {
                const double2 u = probe_u2(j);
                const double cv = cp[2 * j];
                sr = fma(cv, u.x, sr);
                si = fma(cv, u.y, si);
            }
            probe[e] = make_double2(sr, si);
        }
        __syncthreads();

#ifdef HS_EXPERIMENT
        if ((P.flip & 0xfffff) > 0) {
            // phase skew between the groups that share an SMSP pair (experiment): all groups run identical work, so after
            // every CTA barrier they would otherwise hit the FP64 pipe in the same phases
            const long long until = clock64() + (long long)(gid >> 1) * (P.flip & 0xfffff);
            while (clock64() < until) { }
        }
#endif
        // z_k(f) of the group's current bin: a private 2 * n_planes slot in shared memory, filled from the (L2-resident) table by the
        // lanes l64 < 2 * n_planes, which fetch the NEXT bin's values one matrix ahead
        double2* zg = zs + gid * 2 * n_planes;
        auto z_fetch = [&](const int f) {
            return (l64 < p && f < f_end) ? P.z[(size_t)l64 * F + f] : make_double2(0.0, 0.0);
        };
        double2 z_next = z_fetch(f_begin + gid);
        for (int f = f_begin + gid; f < f_end; f += NG) {
            double c[T][T][2];
            mma_group_sync(x);                      // previous matrix' check / epilogue is done with zg, vfull, wpart, X
            if (l64 < 2 * n_planes) zg[l64] = z_next;
            z_next = z_fetch(f + NG);
            mma_group_sync(x);
            // ---- v = A(f) u for the check (thread l64 < 40 owns entry l64)
            const int vi = min(l64, kMP - 1);
            double2 vacc = (l64 < m) ? probe_u2(l64) : make_double2(0.0, 0.0);
            for (int k = 0; k < p; ++k) {
                const double2 zz = zg[k];
                const double2 q0 = probe[vi * p + k];
                vacc.x = fma(-q0.x, zz.x, fma(q0.y, zz.y, vacc.x));
                vacc.y = fma(-q0.x, zz.y, fma(-q0.y, zz.x, vacc.y));
            }
            if (l64 < kMP) gs->vfull[l64] = vacc;
            if (l64 == 0) gs->flag = 0;
            // ---- A(f) = I - sum_k A_k z_k(f)
#ifdef HS_EXPERIMENT
            if (P.flip & (1 << 30)) {       // experiment: skip the assembly (diagonally dominant dummy matrix)
#pragma unroll
                for (int ta = 0; ta < T; ++ta)
#pragma unroll
                    for (int tb = 0; tb < T; ++tb) {
                        c[ta][tb][0] = ((ta == tb && x.g4 == 2 * x.t4 && x.part == 0) ? 4.0 : 0.0) + 1e-3 * (x.lane + ta - tb + f);
                        c[ta][tb][1] = ((ta == tb && x.g4 == 2 * x.t4 + 1 && x.part == 0) ? 4.0 : 0.0) + 1e-3 * (x.lane - ta + tb);
                    }
            } else
#endif
            if (n_planes == 4) mma_assemble<T, 4>(c, coef, zg, n_planes, x);
            else mma_assemble<T, 0>(c, coef, zg, n_planes, x);
            if (P.Af) mma_store_generic<T, false>(c, P, w, f, x);
            if (padchk && x.t4 == 3) {       // v into the pad column (entry [1] of the last column tile in the lanes t4 == 3)
#pragma unroll
                for (int ta = 0; ta < T; ++ta) {
                    const int row = 8 * ta + x.g4;
                    if (row < m) {
                        const double2 vv = gs->vfull[row];
                        c[ta][T - 1][1] = x.part ? vv.y : vv.x;
                    }
                }
            }
            // ---- blocked Gauss-Jordan on the tensor pipe
            mma_gauss_jordan<T, ADJ>(c, m, x);
            // ---- a-posteriori check:  S v == u ?   (v = A(f) u)
            if (padchk) {
                if (x.t4 == 3) {
                    bool bad = false;
#pragma unroll
                    for (int ta = 0; ta < T; ++ta) {
                        const int row = 8 * ta + x.g4;
                        const double2 ck = chk[x.part * kMP + row];
                        const double e = c[ta][T - 1][1] - ck.x;
                        if (row < m && !(e * e <= ck.y)) bad = true;      // also catches NaN / Inf
                    }
                    if (bad) gs->flag = 1;
                }
                mma_group_sync(x);             // the partner is past its last fragment loads: the panel buffers may become the exchange buffer
            } else {
                double sr[T], si[T];
#pragma unroll
                for (int ta = 0; ta < T; ++ta) sr[ta] = si[ta] = 0.0;
#pragma unroll
                for (int tb = 0; tb < T; ++tb) {
                    const double2 v0 = gs->vfull[8 * tb + 2 * x.t4], v1 = gs->vfull[8 * tb + 2 * x.t4 + 1];
#pragma unroll
                    for (int ta = 0; ta < T; ++ta) {
                        sr[ta] = fma(c[ta][tb][0], v0.x, fma(c[ta][tb][1], v1.x, sr[ta]));
                        si[ta] = fma(c[ta][tb][0], v0.y, fma(c[ta][tb][1], v1.y, si[ta]));
                    }
                }
#pragma unroll
                for (int ta = 0; ta < T; ++ta) {
                    sr[ta] += __shfl_xor_sync(0xffffffffu, sr[ta], 1);
                    si[ta] += __shfl_xor_sync(0xffffffffu, si[ta], 1);
                    sr[ta] += __shfl_xor_sync(0xffffffffu, sr[ta], 2);
                    si[ta] += __shfl_xor_sync(0xffffffffu, si[ta], 2);
                    if (x.t4 == 0) {
                        gs->wpart[x.part][0][8 * ta + x.g4] = sr[ta];
                        gs->wpart[x.part][1][8 * ta + x.g4] = si[ta];
                    }
                }
                mma_group_sync(x);
                if (l64 < m) {
                    // S v = (Sr vr - Si vi) + i (Sr vi + Si vr)
                    const double wr = gs->wpart[0][0][l64] - gs->wpart[1][1][l64];
                    const double wi = gs->wpart[0][1][l64] + gs->wpart[1][0][l64];
                    const double2 u = probe_u2(l64);
                    const double er = wr - u.x, ei = wi - u.y;
                    const double err = fma(er, er, ei * ei), ref = fma(u.x, u.x, u.y * u.y);
                    if (!(err <= P.verify_tol2 * ref)) gs->flag = 1;      // also catches NaN / Inf
                }
            }
            // ---- Re/Im exchange for |H|^2 (the panel buffers are free: the barrier above is past every fragment load).
            //      The Re warp finishes the entries in even columns, the Im warp those in odd columns.
            {
                double* X = reinterpret_cast<double*>(gs->u.X) + x.part * (25 * 32) + x.lane;      // outbox of this warp
#pragma unroll
                for (int q = 0; q < T * T; ++q) X[q * 32] = x.part ? c[q / T][q % T][0] : c[q / T][q % T][1];
            }
            mma_group_sync(x);
#ifdef HS_EXPERIMENT
            if (P.flip & (1 << 29)) continue;       // experiment: skip the epilogue
#endif
            const bool good = (gs->flag == 0);
            if (!good) {
                if (l64 == 0) {
                    P.bad[(size_t)w * F + f] = 1;
                    P.bad_list[atomicAdd(P.bad_count, 1)] = w * F + f;
                }
                continue;
            }
            if (P.H || (P.dtf && !P.dtf_fij)) {
                mma_store_generic<T, true>(c, P, w, f, x);        // outputs in the reference layout: not the metric path
            } else {
                const double* X = reinterpret_cast<const double*>(gs->u.X) + (x.part ^ 1) * (25 * 32) + x.lane;      // the partner's outbox
                const int j0 = 2 * x.t4 + x.part;
                double* dst = P.dtf ? P.dtf + (((size_t)w * m + x.g4) * F + f) * m + j0 : nullptr;      // staging (w, i, f, j)
                const size_t row_step = (size_t)8 * F * m;
#pragma unroll
                for (int ta = 0; ta < T; ++ta) {
                    double rsum = 0.0;
                    const bool row_ok = (8 * ta + x.g4 < m);
#pragma unroll
                    for (int tb = 0; tb < T; ++tb) {
                        const double mine = x.part ? c[ta][tb][1] : c[ta][tb][0];
                        const double got = X[(ta * T + tb) * 32];
                        const double v = fma(mine, mine, got * got);
                        if (row_ok && (8 * tb + j0 < m)) {
                            rsum += v;
                            if (dst) dst[ta * row_step + 8 * tb] = v;
                        }
                    }
                    rsum += __shfl_xor_sync(0xffffffffu, rsum, 1);
                    rsum += __shfl_xor_sync(0xffffffffu, rsum, 2);
                    if (x.t4 == 0 && row_ok) gs->rs[x.part][8 * ta + x.g4] += rsum;      // one writer per (warp, row): deterministic
                }
            }
        }
        // ---- row sums of the piece (fixed summation order -> deterministic); slot = CTA index - first CTA of the window
        __syncthreads();
        if (P.rowpart && threadIdx.x < m) {
            double acc = 0.0;
            for (int g2 = 0; g2 < NG; ++g2) acc += groups[g2].rs[0][threadIdx.x] + groups[g2].rs[1][threadIdx.x];
            const int slot = (int)blockIdx.x - (int)(((long long)w * F) / P.per_cta);
            P.rowpart[((size_t)w * P.n_seg + slot) * m + threadIdx.x] = acc;
        }
    }
}

#ifdef HS_EXPERIMENT
// Role layout: warps 0 .. 2 NG - 1 = main warps (whole warpgroups of 4), then ONE warpgroup of helpers.
//   NG = 4: 4 helpers, one per group;            registers 256 x 200 + 128 x 96 = 63 488
//   NG = 6: 3 helpers, each serving two groups;  registers 384 x 152 + 128 x 56 = 65 536   (the fourth warp of the helper warpgroup idles)
template <int NG> struct WsCfg;
template <> struct WsCfg<4> { static constexpr int kMainRegs = 200, kHelperRegs = 96, kPerHelper = 1; };
template <> struct WsCfg<6> { static constexpr int kMainRegs = 152, kHelperRegs = 56, kPerHelper = 2; };

template <int T, int NG>
__global__ void __launch_bounds__(NG * 64 + 128, 1) transfer_ws_kernel(const K5Params P) {
    constexpr int kThreads = NG * 64 + 128;
    constexpr int kPerHelper = WsCfg<NG>::kPerHelper;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int m = P.m, p = P.p, F = P.F;
    double* coef = reinterpret_cast<double*>(smem_raw);
    double2* probe = reinterpret_cast<double2*>(smem_raw + MmaSmem<T>::coef_bytes(p));
    double2* zs = reinterpret_cast<double2*>(smem_raw + MmaSmem<T>::coef_bytes(p) + MmaSmem<T>::probe_bytes(p));      // [bin of the unit][2 * n_planes]
    MmaGroupSmem* groups = reinterpret_cast<MmaGroupSmem*>(smem_raw + MmaSmem<T>::coef_bytes(p) + MmaSmem<T>::probe_bytes(p) + MmaSmem<T>::z_bytes(p, P.seg_len));
    const int warp = threadIdx.x >> 5;
    const bool helper = warp >= 2 * NG;
    MmaCtx x;
    x.lane = threadIdx.x & 31;
    x.part = warp & 1;
    x.g4 = x.lane >> 2;
    x.t4 = x.lane & 3;
    const int gid = helper ? min((warp - 2 * NG) * kPerHelper, NG - 1) : warp >> 1;      // helper: its first group
    const int n_served = helper ? max(0, min(kPerHelper, NG - (warp - 2 * NG) * kPerHelper)) : 0;
    x.bar = 1 + gid;
    x.dbg = 0;
    x.lock = nullptr;
    x.gs = groups + gid;
    MmaGroupSmem* gs = x.gs;
    const int l64 = x.part * 32 + x.lane;
    const int n_planes = MmaSmem<T>::planes(p);
    const int n_steps = (m + 3) >> 2;
    constexpr int NW = kThreads / 32;          // warps per CTA
    if (!helper && l64 == 0) {
        ws_mbar_init(&gs->bar_d, 2);
        ws_mbar_init(&gs->bar_p, 1);
    }
    unsigned ph = 0;                           // block steps done by this group since the kernel started (barrier phase)
    unsigned ph2[kPerHelper];                  // the same per served group (helpers serving several groups)
#pragma unroll
    for (int u = 0; u < kPerHelper; ++u) ph2[u] = 0;
    // zero the coefficient planes once: padding rows / columns / the odd lag stay zero for every unit
    for (int e = threadIdx.x; e < n_planes * kPlaneD; e += kThreads) coef[e] = 0.0;

    // Balanced static partition: CTA c owns the matrices q = w * F + f in [c * per_cta, (c + 1) * per_cta): every CTA gets the
    // same count (+-1 round of NG), and a window's coefficients are loaded once per PIECE (= the part of a window inside the range).
    const long long q_total = (long long)P.n_win * F;
    long long q = (long long)blockIdx.x * P.per_cta;
    const long long q_end = min(q_total, q + (long long)P.per_cta);
    // The two roles run SEPARATE copies of the piece loop (same barriers, same order): ptxas gives each role its own register
    // budget only when the roles never re-join after setmaxnreg -- a shared loop body would be compiled for the smaller one.
    auto prepare_piece = [&](const int w) {
    __syncthreads();
    {   // AR coefficients of window w -> shared, layout [lag pair][row][col][lag & 1]
        const double* Aw = P.A + (size_t)w * m * m * p;
        const int row_len = m * p;
        for (int i = warp; i < m; i += NW) {       // one warp per row, 5 independent loads in flight per lane
            const double* Ar = Aw + (size_t)i * row_len;
            for (int c0 = x.lane; c0 < row_len; c0 += 5 * 32) {
                double v[5];
#pragma unroll
                for (int u = 0; u < 5; ++u) v[u] = (c0 + 32 * u < row_len) ? Ar[c0 + 32 * u] : 0.0;
#pragma unroll
                for (int u = 0; u < 5; ++u) {
                    const int cidx = c0 + 32 * u;
                    if (cidx < row_len) {
                        const int j = cidx / p, k = cidx - j * p;
                        coef[(k >> 1) * kPlaneD + i * kRowD + 2 * j + (k & 1)] = v[u];
                    }
                }
            }
        }
        if (!helper)
            for (int e = l64; e < 2 * kMP; e += 64) (&gs->rs[0][0])[e] = 0.0;
    }
    __syncthreads();
    // probe[i][k] = sum_j A_k[i][j] u_j : the window-dependent part of v = A(f) u = u - sum_k z_k(f) probe[.][k]
    for (int e = threadIdx.x; e < kMP * p; e += kThreads) {
        const int i = e / p, k = e - i * p;
        double sr = 0.0, si = 0.0;
        const double* cp = coef + (k >> 1) * kPlaneD + i * kRowD + (k & 1);
        for (int j = 0; j < m; ++j) {
            const double2 u = probe_u2(j);
            const double cv = cp[2 * j];
            sr = fma(cv, u.x, sr);
            si = fma(cv, u.y, si);
        }
        probe[e] = make_double2(sr, si);
    }
    __syncthreads();

    };
    if (helper) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WsCfg<NG>::kHelperRegs));
        while (q < q_end) {
            const int w = (int)(q / F);
            const int f_begin = (int)(q - (long long)w * F);
            const int f_end = (int)min((long long)F, f_begin + (q_end - q));
            q += f_end - f_begin;
            prepare_piece(w);
            // ---- helper warp: one pivot-block inverse per block step of every matrix of the group(s) it serves; with two
            //      groups it polls both mailboxes and serves whichever published first
            if constexpr (kPerHelper == 1) {
                for (int f = f_begin + gid; f < f_end; f += NG) {
                    for (int st = 0; st < n_steps; ++st) {
                        ws_mbar_wait(&gs->bar_d, ph & 1u);
                        ++ph;
                        ws_inverse4(gs, x.lane);
                        __syncwarp();
                        if (x.lane == 0) ws_mbar_arrive(&gs->bar_p);
                    }
                }
            } else {
                int left[kPerHelper];
                int total = 0;
#pragma unroll
                for (int u = 0; u < kPerHelper; ++u) {
                    const int first = f_begin + gid + u;
                    left[u] = (u < n_served && first < f_end) ? ((f_end - first + NG - 1) / NG) * n_steps : 0;
                    total += left[u];
                }
                while (total > 0) {
#pragma unroll
                    for (int u = 0; u < kPerHelper; ++u) {
                        if (left[u] > 0 && ws_mbar_test(&gs[u].bar_d, (ph2[u] & 1u))) {
                            ++ph2[u];
                            --left[u];
                            --total;
                            ws_inverse4(gs + u, x.lane);
                            __syncwarp();
                            if (x.lane == 0) ws_mbar_arrive(&gs[u].bar_p);
                        }
                    }
                }
            }
            __syncthreads();       // the main warps' row-sum write-out
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(WsCfg<NG>::kMainRegs));
        while (q < q_end) {
            const int w = (int)(q / F);
            const int f_begin = (int)(q - (long long)w * F);
            const int f_end = (int)min((long long)F, f_begin + (q_end - q));
            q += f_end - f_begin;
            prepare_piece(w);
        // z_k(f) of the group's current bin: a private 2 * n_planes slot in shared memory, filled from the (L2-resident) table by the
        // lanes l64 < 2 * n_planes, which fetch the NEXT bin's values one matrix ahead
        double2* zg = zs + gid * 2 * n_planes;
        auto z_fetch = [&](const int f) {
            return (l64 < p && f < f_end) ? P.z[(size_t)l64 * F + f] : make_double2(0.0, 0.0);
        };
        double2 z_next = z_fetch(f_begin + gid);
        for (int f = f_begin + gid; f < f_end; f += NG) {
            double c[T][T][2];
            mma_group_sync(x);                      // previous matrix' check / epilogue is done with zg, vfull, wpart, X
            if (l64 < 2 * n_planes) zg[l64] = z_next;
            z_next = z_fetch(f + NG);
            mma_group_sync(x);
            // ---- v = A(f) u for the check (thread l64 < 40 owns entry l64)
            const int vi = min(l64, kMP - 1);
            double2 vacc = (l64 < m) ? probe_u2(l64) : make_double2(0.0, 0.0);
            for (int k = 0; k < p; ++k) {
                const double2 zz = zg[k];
                const double2 q0 = probe[vi * p + k];
                vacc.x = fma(-q0.x, zz.x, fma(q0.y, zz.y, vacc.x));
                vacc.y = fma(-q0.x, zz.y, fma(-q0.y, zz.x, vacc.y));
            }
            if (l64 < kMP) gs->vfull[l64] = vacc;
            if (l64 == 0) gs->flag = 0;
            // ---- A(f) = I - sum_k A_k z_k(f)
            if (n_planes == 4) mma_assemble<T, 4>(c, coef, zg, n_planes, x);
            else mma_assemble<T, 0>(c, coef, zg, n_planes, x);
            if (P.Af) mma_store_generic<T, false>(c, P, w, f, x);
            // ---- blocked Gauss-Jordan on the tensor pipe
            mma_group_sync(x);                 // previous users of the panel buffers (assembly exchange) are done
            ws_publish<T, 0>(c, 0, x);         // pivot block of step 0
            ws_all_tiles<T, 0>(c, m, x, ph);
            // ---- a-posteriori check:  S v == u ?   (v = A(f) u)
            {
                double sr[T], si[T];
#pragma unroll
                for (int ta = 0; ta < T; ++ta) sr[ta] = si[ta] = 0.0;
#pragma unroll
                for (int tb = 0; tb < T; ++tb) {
                    const double2 v0 = gs->vfull[8 * tb + 2 * x.t4], v1 = gs->vfull[8 * tb + 2 * x.t4 + 1];
#pragma unroll
                    for (int ta = 0; ta < T; ++ta) {
                        sr[ta] = fma(c[ta][tb][0], v0.x, fma(c[ta][tb][1], v1.x, sr[ta]));
                        si[ta] = fma(c[ta][tb][0], v0.y, fma(c[ta][tb][1], v1.y, si[ta]));
                    }
                }
#pragma unroll
                for (int ta = 0; ta < T; ++ta) {
                    sr[ta] += __shfl_xor_sync(0xffffffffu, sr[ta], 1);
                    si[ta] += __shfl_xor_sync(0xffffffffu, si[ta], 1);
                    sr[ta] += __shfl_xor_sync(0xffffffffu, sr[ta], 2);
                    si[ta] += __shfl_xor_sync(0xffffffffu, si[ta], 2);
                    if (x.t4 == 0) {
                        gs->wpart[x.part][0][8 * ta + x.g4] = sr[ta];
                        gs->wpart[x.part][1][8 * ta + x.g4] = si[ta];
                    }
                }
            }
            mma_group_sync(x);
            if (l64 < m) {
                // S v = (Sr vr - Si vi) + i (Sr vi + Si vr)
                const double wr = gs->wpart[0][0][l64] - gs->wpart[1][1][l64];
                const double wi = gs->wpart[0][1][l64] + gs->wpart[1][0][l64];
                const double2 u = probe_u2(l64);
                const double er = wr - u.x, ei = wi - u.y;
                const double err = fma(er, er, ei * ei), ref = fma(u.x, u.x, u.y * u.y);
                if (!(err <= P.verify_tol2 * ref)) gs->flag = 1;      // also catches NaN / Inf
            }
            // ---- Re/Im exchange for |H|^2 (the panel buffers are free: the barrier above is past every fragment load).
            //      The Re warp finishes the entries in even columns, the Im warp those in odd columns.
            {
                double* X = reinterpret_cast<double*>(gs->u.X) + x.part * (25 * 32) + x.lane;      // outbox of this warp
#pragma unroll
                for (int q = 0; q < T * T; ++q) X[q * 32] = x.part ? c[q / T][q % T][0] : c[q / T][q % T][1];
            }
            mma_group_sync(x);
            const bool good = (gs->flag == 0);
            if (!good) {
                if (l64 == 0) {
                    P.bad[(size_t)w * F + f] = 1;
                    P.bad_list[atomicAdd(P.bad_count, 1)] = w * F + f;
                }
                continue;
            }
            if (P.H || (P.dtf && !P.dtf_fij)) {
                mma_store_generic<T, true>(c, P, w, f, x);        // outputs in the reference layout: not the metric path
            } else {
                const double* X = reinterpret_cast<const double*>(gs->u.X) + (x.part ^ 1) * (25 * 32) + x.lane;      // the partner's outbox
                const int j0 = 2 * x.t4 + x.part;
                double* dst = P.dtf ? P.dtf + (((size_t)w * m + x.g4) * F + f) * m + j0 : nullptr;      // staging (w, i, f, j)
                const size_t row_step = (size_t)8 * F * m;
#pragma unroll
                for (int ta = 0; ta < T; ++ta) {
                    double rsum = 0.0;
                    const bool row_ok = (8 * ta + x.g4 < m);
#pragma unroll
                    for (int tb = 0; tb < T; ++tb) {
                        const double mine = x.part ? c[ta][tb][1] : c[ta][tb][0];
                        const double got = X[(ta * T + tb) * 32];
                        const double v = fma(mine, mine, got * got);
                        if (row_ok && (8 * tb + j0 < m)) {
                            rsum += v;
                            if (dst) dst[ta * row_step + 8 * tb] = v;
                        }
                    }
                    rsum += __shfl_xor_sync(0xffffffffu, rsum, 1);
                    rsum += __shfl_xor_sync(0xffffffffu, rsum, 2);
                    if (x.t4 == 0 && row_ok) gs->rs[x.part][8 * ta + x.g4] += rsum;      // one writer per (warp, row): deterministic
                }
            }
        }
        // ---- row sums of the piece (fixed summation order -> deterministic); slot = CTA index - first CTA of the window
        __syncthreads();
        if (P.rowpart && threadIdx.x < m) {
            double acc = 0.0;
            for (int g2 = 0; g2 < NG; ++g2) acc += groups[g2].rs[0][threadIdx.x] + groups[g2].rs[1][threadIdx.x];
            const int slot = (int)blockIdx.x - (int)(((long long)w * F) / P.per_cta);
            P.rowpart[((size_t)w * P.n_seg + slot) * m + threadIdx.x] = acc;
        }
        }
    }
}

#endif  // HS_EXPERIMENT (transfer_ws_kernel)

template <int T, int NG>
int launch_mma_t(const K5Params& P, int sm_count, cudaStream_t stream) {
    const size_t smem = MmaSmem<T>::total(P.p, NG, P.seg_len);
    if (smem > 227 * 1024) return set_error(HS_ERR_UNSUPPORTED, "transfer_mma: model order %d needs %zu B shared memory", P.p, smem);
    // 4 x 4 pivot-block inverse: 3 (default) cofactors with Re / Im split over the half-warps; 4 FP32 seed + Newton-Schulz on the tensor pipe
    // (measured the same: 4.85 vs 4.81 ms -- 8 more DMMAs per block step, see mma_inverse4_newton); 1 cofactors, every lane a whole entry; 2 cofactors with the division by det deferred to the U panel (measured
    // slower: 5.49 vs 5.24 ms, the 20 extra FP64 instructions per lane and step sit in front of the 50 update DMMAs); 0 in-place elimination
#ifdef HS_EXPERIMENT
    static const int adj = exp_env_int("HS_K5_ADJ", 3);
    constexpr bool kMain = (T == 5 && NG == 6);
    constexpr bool kMain5 = (T == 5);       // the look-ahead variant: measured at 4 and 6 groups
    auto kern = (adj == 6 && kMain5) ? transfer_mma_kernel<T, NG, kMain5 ? 6 : 3> : (adj == 5 && kMain5) ? transfer_mma_kernel<T, NG, kMain5 ? 5 : 3> : (adj == 4) ? transfer_mma_kernel<T, NG, 4> : (adj == 3) ? transfer_mma_kernel<T, NG, 3> : (adj == 2 && kMain) ? transfer_mma_kernel<T, NG, kMain ? 2 : 0>
              : (adj == 1 && kMain) ? transfer_mma_kernel<T, NG, kMain ? 1 : 0> : transfer_mma_kernel<T, NG, 0>;
#else
    auto kern = transfer_mma_kernel<T, NG, 3>;       // cofactor pivot-block inverse, Re / Im split over the half-warps, for every tile count
#endif
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(HS_ERR_CUDA, "transfer_mma: cannot reserve %zu B shared memory: %s", smem, cudaGetErrorString(e));
    const long long q_total = (long long)P.n_win * P.F;
    const int grid = (int)((q_total + P.per_cta - 1) / P.per_cta);
    kern<<<grid, NG * 64, smem, stream>>>(P);
    return check_launch("transfer_mma_kernel");
}

#ifdef HS_EXPERIMENT
template <int T, int NG>
int launch_ws_t(const K5Params& P, cudaStream_t stream) {
    const size_t smem = MmaSmem<T>::total(P.p, NG, P.seg_len);
    if (smem > 227 * 1024) return set_error(HS_ERR_UNSUPPORTED, "transfer_ws: model order %d needs %zu B shared memory", P.p, smem);
    auto kern = transfer_ws_kernel<T, NG>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(HS_ERR_CUDA, "transfer_ws: cannot reserve %zu B shared memory: %s", smem, cudaGetErrorString(e));
    const long long q_total = (long long)P.n_win * P.F;
    const int grid = (int)((q_total + P.per_cta - 1) / P.per_cta);
    kern<<<grid, NG * 64 + 128, smem, stream>>>(P);
    return check_launch("transfer_ws_kernel");
}

#endif

}  // namespace

// Warp-specialised optimistic pass (groups of Re / Im warps + helper warps that invert the pivot blocks one step ahead); same
// contract as launch_transfer_mma.  Measured on 599 cfg2 windows: 5.48 ms (4 groups, one helper each) and 5.91 ms (6 groups, three
// helpers) against 5.31 ms for transfer_mma_kernel, so it is compiled only into HS_EXPERIMENT builds.
int launch_transfer_ws(const K5Params& P, int ng, cudaStream_t stream) {
#ifndef HS_EXPERIMENT
    (void)P; (void)ng; (void)stream;
    return set_error(HS_ERR_UNSUPPORTED, "transfer_ws_kernel is compiled into HS_EXPERIMENT builds only");
#else
    if (ng == 6) {
        switch ((P.m + 7) / 8) {
            case 1: return launch_ws_t<1, 6>(P, stream);
            case 2: return launch_ws_t<2, 6>(P, stream);
            case 3: return launch_ws_t<3, 6>(P, stream);
            case 4: return launch_ws_t<4, 6>(P, stream);
            case 5: return launch_ws_t<5, 6>(P, stream);
        }
    }
    switch ((P.m + 7) / 8) {
        case 1: return launch_ws_t<1, 4>(P, stream);
        case 2: return launch_ws_t<2, 4>(P, stream);
        case 3: return launch_ws_t<3, 4>(P, stream);
        case 4: return launch_ws_t<4, 4>(P, stream);
        case 5: return launch_ws_t<5, 4>(P, stream);
    }
    return set_error(HS_ERR_UNSUPPORTED, "transfer_ws: no kernel for m=%d", P.m);
#endif
}

void transfer_mma_partition(int n_win, int F, int* per_cta, int* slots, int ng) {
    const long long total = (long long)n_win * F;
    int sm = compute_sm_count();
    if (sm < 1) sm = 148;
    long long per = (total + sm - 1) / sm;
    per = (per + ng - 1) / ng * ng;              // whole rounds of the ng groups
    if (per < ng) per = ng;
    *per_cta = (int)per;
    *slots = (int)((F + per - 1) / per) + 1;
}

bool transfer_mma_fits(int p, int ng, int seg_len) { return ng >= 4 && ng <= 6 && MmaSmem<5>::total(p, ng, seg_len) <= 227 * 1024; }

// optimistic (unpivoted, verified) pass on the tensor pipe; the caller follows up with the pivoted redo
int launch_transfer_mma(const K5Params& P, int ng, cudaStream_t stream) {
    const int sm = device_sm_count();
#ifdef HS_EXPERIMENT
    if (ng == 4 && (P.m + 7) / 8 == 5) return launch_mma_t<5, 4>(P, sm, stream);      // experiments: fewer matrices in flight per SM
    if (ng == 5 && (P.m + 7) / 8 == 5) return launch_mma_t<5, 5>(P, sm, stream);
#endif
    if (ng != 6) return set_error(HS_ERR_INVALID, "transfer_mma: %d groups per CTA not built", ng);
    switch ((P.m + 7) / 8) {
        case 1: return launch_mma_t<1, 6>(P, sm, stream);
        case 2: return launch_mma_t<2, 6>(P, sm, stream);
        case 3: return launch_mma_t<3, 6>(P, sm, stream);
        case 4: return launch_mma_t<4, 6>(P, sm, stream);
        case 5: return launch_mma_t<5, 6>(P, sm, stream);
    }
    return set_error(HS_ERR_UNSUPPORTED, "transfer_mma: no kernel for m=%d", P.m);
}

}  // namespace hs
