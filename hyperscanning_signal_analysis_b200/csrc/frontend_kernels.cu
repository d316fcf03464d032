// Front end on sm_100a, FP64:  zero-phase IIR (K1), FIR decimation (K2), multitaper PSD (K6).
//
// K1 replaces the per-channel Python loop of dataloader._apply_filters (src/dataloader.py:786-803,
// IIR branch :789-792: DC removal, then scipy.signal.filtfilt notch -> low -> high) and the filter block
// of mne_bridge.load_eeg_signals (src/mne_bridge.py:161-184, order-4 Butterworth, axis 0).
// K2 replaces scipy.signal.decimate(x, q, ftype='fir', zero_phase=True) (src/data_structures.py:792).
//
// filtfilt semantics reproduced exactly: odd extension by e = 3*ntaps samples on both
// sides, DF2T recursion started from zi * ext[0], second pass over the reversed forward output started
// from zi * f[last], middle n samples returned.
//
// The recursion is a linear recurrence, so it is solved as a chunked scan:
//   pass A  every thread runs the recursion over its chunk of C samples from a ZERO state and keeps only
//           the end state (the chunk's zero-state response);
//   pass B  one warp per signal composes the chunk end states with the state-transition power M^C
//           (s_{c+1} = M^C s_c + e_c) -- the only sequential part, n/C steps of a d x d mat-vec;
//   pass C  every thread re-runs its chunk from its true start state and writes the output.
// Forward and backward sweeps use the same three kernels (the backward sweep walks the index backwards).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <vector>

#include "hs_internal.h"

namespace hs {

constexpr int kMaxOrder = 4;      // DF2T state dimension d = ntaps - 1 <= 4 (the reference uses biquads and Butterworth-4)
constexpr int kChunk = 512;       // samples per thread-chunk

struct IirCoef {
    double b[kMaxOrder + 1];
    double a[kMaxOrder + 1];
    double zi[kMaxOrder];
    double pw[6][kMaxOrder * kMaxOrder];  // (M^C)^(2^k), k = 0..5, row-major d x d
    int d;                                // state dimension
    int e;                                // pad length 3 * ntaps
    int warm;                             // warm-up chunks replayed by the apply pass (order > 2 only)
};

struct IirPass {
    const double* in;      // source signal base
    double* out;           // destination base
    long long in_sig_stride, in_t_stride;     // element strides of the source
    long long out_sig_stride, out_t_stride;
    long long n_in;        // valid samples of the source (n for the forward sweep, L for the backward sweep)
    long long L;           // sweep length n + 2e
    int e;                 // extension length (forward sweep synthesises the odd extension on the fly)
    int forward;           // 1: source = x with virtual odd extension, writes f[0..L); 0: source = f walked backwards, writes y
    int n_sig;
    long long n_chunks;
    const double* mean;    // per-signal DC to subtract while reading (forward sweep of the first filter) or null
    double* states;        // (n_sig, n_chunks, d) chunk end states (pass A) -> chunk start states (pass B)
};

// value of the (virtually extended, optionally DC-removed) source at sweep position u
__device__ __forceinline__ double sweep_read(const IirPass& P, const double* base, const double dc, const long long u) {
    if (P.forward) {
        const long long n = P.n_in, e = P.e;
        long long t = u - e;
        if (t < 0) {                      // 2*x[0] - x[e-u]
            return 2.0 * (base[0] - dc) - (base[(e - u) * P.in_t_stride] - dc);
        } else if (t >= n) {              // 2*x[n-1] - x[n-2-(t-n)]
            return 2.0 * (base[(n - 1) * P.in_t_stride] - dc) - (base[(2 * n - 2 - t) * P.in_t_stride] - dc);
        }
        return base[t * P.in_t_stride] - dc;
    }
    return base[(P.L - 1 - u) * P.in_t_stride];      // reversed forward output
}

template <int D>
__device__ __forceinline__ double df2t_step(const IirCoef& c, double (&z)[D], const double u) {
    const double y = fma(c.b[0], u, z[0]);
#pragma unroll
    for (int k = 0; k < D - 1; ++k) z[k] = fma(-c.a[k + 1], y, fma(c.b[k + 1], u, z[k + 1]));
    z[D - 1] = fma(-c.a[D], y, c.b[D] * u);
    return y;
}

template <int D>
__global__ void __launch_bounds__(128) iir_local_kernel(const IirPass P, const IirCoef c) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)P.n_sig * P.n_chunks) return;
    const int s = (int)(gid / P.n_chunks);
    const long long ch = gid - (long long)s * P.n_chunks;
    const double* base = P.in + (long long)s * P.in_sig_stride;
    const double dc = P.mean ? P.mean[s] : 0.0;
    double z[D];
#pragma unroll
    for (int k = 0; k < D; ++k) z[k] = 0.0;
    const long long u0 = ch * kChunk, u1 = min(P.L, u0 + kChunk);
    for (long long u = u0; u < u1; ++u) df2t_step<D>(c, z, sweep_read(P, base, dc, u));
    double* st = P.states + ((long long)s * P.n_chunks + ch) * D;
#pragma unroll
    for (int k = 0; k < D; ++k) st[k] = z[k];
}

// One warp per signal: start state of every chunk.  With P = M^C:  s_0 = zi * ext[0],  s_{c+1} = P s_c + e_c.
// 32 chunks per iteration: an inclusive warp scan of the affine maps with the precomputed powers P^(2^k).
template <int D>
__device__ __forceinline__ void matvec_acc(const double* __restrict__ Pm, const double (&v)[D], double (&acc)[D]) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
        double a = acc[i];
#pragma unroll
        for (int j = 0; j < D; ++j) a = fma(Pm[i * D + j], v[j], a);
        acc[i] = a;
    }
}

template <int D>
__global__ void __launch_bounds__(128) iir_carry_kernel(const IirPass P, const IirCoef c) {
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (s >= P.n_sig) return;
    const double* base = P.in + (long long)s * P.in_sig_stride;
    const double dc = P.mean ? P.mean[s] : 0.0;
    const double x0 = sweep_read(P, base, dc, 0);
    double carry[D];
#pragma unroll
    for (int k = 0; k < D; ++k) carry[k] = c.zi[k] * x0;
    double* st = P.states + (long long)s * P.n_chunks * D;
    for (long long c0 = 0; c0 < P.n_chunks; c0 += 32) {
        const long long ch = c0 + lane;
        double v[D];
#pragma unroll
        for (int k = 0; k < D; ++k) v[k] = (ch < P.n_chunks) ? st[ch * D + k] : 0.0;
        if (lane == 0) matvec_acc<D>(c.pw[0], carry, v);            // e_0 + P s_c0
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            double o[D];
#pragma unroll
            for (int q = 0; q < D; ++q) o[q] = __shfl_up_sync(0xffffffffu, v[q], 1 << k);
            if (lane >= (1 << k)) matvec_acc<D>(c.pw[k], o, v);
        }
        // v_l = s_{c0+l+1}; the start state of chunk c0+l is s_{c0+l}
        double prev[D];
#pragma unroll
        for (int q = 0; q < D; ++q) {
            prev[q] = __shfl_up_sync(0xffffffffu, v[q], 1);
            if (lane == 0) prev[q] = carry[q];
            carry[q] = __shfl_sync(0xffffffffu, v[q], 31);
        }
        if (ch < P.n_chunks) {
#pragma unroll
            for (int k = 0; k < D; ++k) st[ch * D + k] = prev[k];
        }
    }
}

template <int D>
__global__ void __launch_bounds__(128) iir_apply_kernel(const IirPass P, const IirCoef c) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)P.n_sig * P.n_chunks) return;
    const int s = (int)(gid / P.n_chunks);
    const long long ch = gid - (long long)s * P.n_chunks;
    const double* base = P.in + (long long)s * P.in_sig_stride;
    double* ob = P.out + (long long)s * P.out_sig_stride;
    const double dc = P.mean ? P.mean[s] : 0.0;
    // Start `warm` chunks early from that chunk's scanned state and replay: the filter is stable, so the error of
    // the scanned state (superposition noise, ~1e-9 for clustered poles) decays geometrically and the arithmetic
    // that follows is the sequential recursion SciPy runs.  warm == 0 for biquads (their scan is accurate to 1e-13).
    const long long ch0 = max(0LL, ch - (long long)c.warm);
    const double* st = P.states + ((long long)s * P.n_chunks + ch0) * D;
    double z[D];
#pragma unroll
    for (int k = 0; k < D; ++k) z[k] = st[k];
    for (long long u = ch0 * kChunk; u < ch * kChunk; ++u) df2t_step<D>(c, z, sweep_read(P, base, dc, u));
    const long long u0 = ch * kChunk, u1 = min(P.L, u0 + kChunk);
    if (P.forward) {
        for (long long u = u0; u < u1; ++u) ob[u * P.out_t_stride] = df2t_step<D>(c, z, sweep_read(P, base, dc, u));
    } else {
        // backward sweep position u corresponds to forward index L-1-u; keep only the middle n samples
        const long long n = P.L - 2 * (long long)P.e;
        for (long long u = u0; u < u1; ++u) {
            const double y = df2t_step<D>(c, z, sweep_read(P, base, dc, u));
            const long long t = P.L - 1 - u - P.e;
            if (t >= 0 && t < n) ob[t * P.out_t_stride] = y;
        }
    }
}

// ------------------------------------------------------------------ K1, tiled path (state dimension <= 2, unit time stride)
// The thread-per-chunk kernels above stream every chunk through its own thread, i.e. with 4 KB between the addresses of
// neighbouring lanes.  For the biquads of the loader (notch / low-pass / high-pass, dataloader.py:687-708) the sweep is
// re-tiled so that global memory only sees coalesced traffic:
//   * a CTA owns a TILE of 8192 consecutive sweep positions, staged in shared memory TRANSPOSED (sample i of thread t at
//     [i][t], row stride 257) by coalesced loads; thread t then walks "its" 32 samples conflict-free;
//   * tile_local : zero-state response of every 32-sample piece, folded into the tile's zero-state end state
//                  sum_t P^(255-t) z_t  with the table P^j = M^(32 j)  (fixed-order block reduction);
//   * the existing one-warp-per-signal carry kernel turns the tile end states into tile start states (powers of M^8192);
//   * tile_apply : stages the tile again, repeats the zero-state pieces, scans them inside the CTA (warp scan with
//                  P^(2^k), then the 8 warp totals), adds P^t s_in, re-runs the 32 samples from the true state, and writes
//                  the tile back coalesced.
// Per sweep: 2 coalesced reads + 1 coalesced write of the signal (the thread-per-chunk path does the same amount of
// traffic, uncoalesced).
constexpr int kTileThreads = 128;      // (512 threads x 16 samples, two CTAs per SM, measured slower: 5.6 vs 4.4 ms for the cfg4 cascade)
constexpr int kTilePer = 32;
constexpr int kTilePerLog2 = 5;
constexpr int kTile = kTileThreads * kTilePer;      // 8192 sweep positions per CTA
constexpr int kTileLd = kTileThreads + 1;
constexpr int kPpowEntries = kTileThreads + 1;      // P^0 .. P^256, 4 doubles each (row-major 2 x 2, zero padded for D = 1)

template <int D>
__device__ __forceinline__ void mv2(const double* __restrict__ Pm, const double (&v)[D], double (&out)[D]) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
        double a = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) a = fma(Pm[i * 2 + j], v[j], a);
        out[i] = a;
    }
}

// barrier of the 256 compute threads: the whole CTA in the two-kernel path, a named barrier in the fused kernel (whose CTA has a
// ninth warp that only does the look-back)
template <bool NAMED>
__device__ __forceinline__ void tile_sync() {
    if (NAMED) asm volatile("bar.sync 1, %0;" ::"n"(kTileThreads) : "memory");
    else __syncthreads();
}

template <int D, bool NAMED = false>
__device__ __forceinline__ void tile_stage_and_local(const IirPass& P, const IirCoef& c, double* sm, const int s, const long long tile,
                                                     double (&z)[D]) {
    const double* base = P.in + (long long)s * P.in_sig_stride;
    const double dc = P.mean ? P.mean[s] : 0.0;
    const long long u0 = tile * kTile;
    // interior tile of a contiguous source: branch-free coalesced loads, 16 in flight per thread
    const bool interior = P.in_t_stride == 1 && (P.forward ? (u0 >= P.e && u0 + kTile <= P.e + P.n_in) : (u0 + kTile <= P.L));
    const double* src = P.forward ? base + (u0 - P.e) : base + (P.L - u0 - kTile);      // forward: x[u - e]; backward: f[L-1-u], walked downwards
    if (interior) {
        double v[kTilePer];            // the whole tile's loads in flight: 32 per thread
#pragma unroll
        for (int k = 0; k < kTilePer; ++k) v[k] = __ldcs(src + k * kTileThreads + threadIdx.x);
#pragma unroll
        for (int k = 0; k < kTilePer; ++k) {
            const int j = k * kTileThreads + threadIdx.x;               // offset inside the staged span
            const int i0 = P.forward ? j : kTile - 1 - j;
            sm[(i0 & (kTilePer - 1)) * kTileLd + (i0 >> kTilePerLog2)] = v[k] - dc;
        }
    } else {
        for (int idx = threadIdx.x; idx < kTile; idx += kTileThreads) {
            const long long u = u0 + idx;
            sm[(idx & (kTilePer - 1)) * kTileLd + (idx >> kTilePerLog2)] = (u < P.L) ? sweep_read(P, base, dc, u) : 0.0;
        }
    }
    tile_sync<NAMED>();
#pragma unroll
    for (int k = 0; k < D; ++k) z[k] = 0.0;
#pragma unroll 8
    for (int i = 0; i < kTilePer; ++i) df2t_step<D>(c, z, sm[i * kTileLd + threadIdx.x]);
}

template <int D>
__global__ void __launch_bounds__(kTileThreads) iir_tile_local_kernel(const IirPass P, const IirCoef c, const double* __restrict__ ppow) {
    extern __shared__ double tile_sm[];
    __shared__ double red[kTileThreads / 32][2];
    const int s = blockIdx.y;
    const long long tile = blockIdx.x;
    double z[D];
    tile_stage_and_local<D>(P, c, tile_sm, s, tile, z);
    double w[D];
    mv2<D>(ppow + (size_t)(kTileThreads - 1 - threadIdx.x) * 4, z, w);
#pragma unroll
    for (int k = 0; k < D; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) w[k] += __shfl_xor_sync(0xffffffffu, w[k], off);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = w[k];
    }
    __syncthreads();
    if (threadIdx.x < D) {
        double acc = 0.0;
        for (int q = 0; q < kTileThreads / 32; ++q) acc += red[q][threadIdx.x];
        P.states[((long long)s * P.n_chunks + tile) * D + threadIdx.x] = acc;
    }
}

template <int D>
__global__ void __launch_bounds__(kTileThreads) iir_tile_apply_kernel(const IirPass P, const IirCoef c, const double* __restrict__ ppow) {
    extern __shared__ double tile_sm[];
    __shared__ double wtot[kTileThreads / 32][2];
    const int s = blockIdx.y;
    const long long tile = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double z[D];
    tile_stage_and_local<D>(P, c, tile_sm, s, tile, z);
    // inclusive scan of the zero-state pieces inside the warp:  v_l = sum_{l' <= l} P^(l-l') z_l'
    double v[D];
#pragma unroll
    for (int k = 0; k < D; ++k) v[k] = z[k];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        double o[D], t[D];
#pragma unroll
        for (int q = 0; q < D; ++q) o[q] = __shfl_up_sync(0xffffffffu, v[q], 1 << k);
        mv2<D>(ppow + (size_t)(1 << k) * 4, o, t);
        if (lane >= (1 << k)) {
#pragma unroll
            for (int q = 0; q < D; ++q) v[q] += t[q];
        }
    }
    if (lane == 31) {
#pragma unroll
        for (int q = 0; q < D; ++q) wtot[warp][q] = v[q];
    }
    __syncthreads();
    // state contributed by the warps before this one
    double W[D];
#pragma unroll
    for (int q = 0; q < D; ++q) W[q] = 0.0;
    for (int j = 0; j < warp; ++j) {
        double t[D];
        mv2<D>(ppow + (size_t)32 * 4, W, t);
#pragma unroll
        for (int q = 0; q < D; ++q) W[q] = t[q] + wtot[j][q];
    }
    // true start state of this thread's piece: P^t s_in + (pieces before it)
    double sin_[D], st[D], e[D];
#pragma unroll
    for (int q = 0; q < D; ++q) sin_[q] = P.states[((long long)s * P.n_chunks + tile) * D + q];
    mv2<D>(ppow + (size_t)threadIdx.x * 4, sin_, st);
    mv2<D>(ppow + (size_t)lane * 4, W, e);
#pragma unroll
    for (int q = 0; q < D; ++q) {
        const double prev = __shfl_up_sync(0xffffffffu, v[q], 1);
        st[q] += e[q] + (lane ? prev : 0.0);
    }
#pragma unroll 8
    for (int i = 0; i < kTilePer; ++i) {
        double* slot = tile_sm + i * kTileLd + threadIdx.x;
        *slot = df2t_step<D>(c, st, *slot);
    }
    __syncthreads();
    double* ob = P.out + (long long)s * P.out_sig_stride;
    const long long u0 = tile * kTile;
    const long long n = P.L - 2 * (long long)P.e;
    for (int idx = threadIdx.x; idx < kTile; idx += kTileThreads) {
        const long long u = u0 + idx;
        if (u >= P.L) break;
        const double y = tile_sm[(idx & (kTilePer - 1)) * kTileLd + (idx >> kTilePerLog2)];
        if (P.forward) {
            ob[u] = y;
        } else {
            const long long t = P.L - 1 - u - P.e;
            if (t >= 0 && t < n) ob[t] = y;
        }
    }
}

// ------------------------------------------------------------------ K1, single-pass tiled sweep (decoupled look-back)
// tile_local + carry + tile_apply read the signal twice per sweep.  Here ONE kernel does the sweep: a CTA (tiles of a signal
// in sweep order) stages its tile once, computes the tile's zero-state end state e (its "aggregate"), publishes it, and then
// obtains its true start state by LOOKING BACK over the predecessors' aggregates:
//      s_in(t) = e(t-1) + Q e(t-2) + Q^2 e(t-3) + ... ,   Q = M^kTile,
// stopping at the beginning of the signal (s_0 = zi * x_0) or when Q^k has decayed below 1e-30 (|Q| ~ 2e-8 for the loader's
// 1 Hz high-pass at 1024 Hz: four records, fetched by four lanes at once).  Only aggregates are read -- never a predecessor's
// finished state, which would shorten the walk but make the summation order depend on timing: the result is bit-reproducible.
// Aggregates depend on nothing but their own tile, so the wait is deadlock free (CTAs are dispatched in index order: every
// predecessor is running or done).  Filters that decay too slowly for this (|Q^8| >= 1e-3) take the three-kernel path.
// Per sweep: 1 coalesced read + 1 coalesced write of the signal.
// A record is ONE 16-byte word {state 0, state 1} that carries its own validity: every sweep of a call has its own records,
// preset to all-ones (a NaN no arithmetic produces; a published value that happens to be that pattern is replaced by the
// canonical NaN), written by one 16-byte store and polled by one 16-byte load.  No flag word, no release / acquire pair:
// a look-back costs one L2 round trip instead of two dependent ones.
struct IirLookback {
    double2* rec;        // (n_sig, n_tiles) of this sweep: zero-state end state of every tile (its aggregate)
    int sweep;           // sweep number within the call
};

constexpr unsigned long long kRecEmpty = ~0ull;
__device__ __forceinline__ double2 ld_rec(const double2* p) {
    double2 v;
    asm volatile("ld.relaxed.gpu.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_rec(double2* p, double a, double b) {
    if ((unsigned long long)__double_as_longlong(a) == kRecEmpty) a = __longlong_as_double(0x7ff8000000000000LL);
    if ((unsigned long long)__double_as_longlong(b) == kRecEmpty) b = __longlong_as_double(0x7ff8000000000000LL);
    asm volatile("st.relaxed.gpu.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ bool rec_valid(const double2 v) {
    return (unsigned long long)__double_as_longlong(v.x) != kRecEmpty && (unsigned long long)__double_as_longlong(v.y) != kRecEmpty;
}
// record j for the look-back: 1 = published, 0 = not yet
template <int D>
__device__ __forceinline__ int rec_fetch(const double2* rec, const long long j, double (&a)[D]) {
    const double2 r = ld_rec(rec + j);
    a[0] = r.x;
    if (D > 1) a[D - 1] = r.y;
    return rec_valid(r) ? 1 : 0;
}

constexpr int kFusedThreads = kTileThreads + 32;      // 8 compute warps + the look-back warp

// Start state of (signal s, tile) from the predecessors' aggregates (one warp; see the header comment).  acc = s_in.
template <int D>
__device__ __forceinline__ void lookback_start_state(const IirPass& P, const IirCoef& c, const IirCoef& ct, const IirLookback& S, const int s,
                                                     const long long tile, const long long rec0, const int lane, double (&acc)[D]) {
    constexpr int f_agg = 1;
    // ---- look-back warp: the start state of this tile from the predecessors' records.  It needs nothing of this tile, so it
    //      runs WHILE the compute warps stage and scan their samples.
    const double* base = P.in + (long long)s * P.in_sig_stride;
    const double dc = P.mean ? P.mean[s] : 0.0;
    auto times_Q = [&](double (&Pw)[D][D]) -> double {      // Pw <- Pw Q; returns max |entry|
        double nx[D][D], mx = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int q = 0; q < D; ++q) {
                double v = 0.0;
#pragma unroll
                for (int r = 0; r < D; ++r) v = fma(Pw[i][r], ct.pw[0][r * D + q], v);
                nx[i][q] = v;
                mx = fmax(mx, fabs(v));
            }
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int q = 0; q < D; ++q) Pw[i][q] = nx[i][q];
        return mx;
    };
    double Pw[D][D];
    // 1. how many terms the sum has: until |Q^k| < 1e-30 (what lies further back changes the state by less than 1e-14 of an
    //    ulp) or the beginning of the signal, whose term is the start state s_0 = zi * x_0 itself.  Depends on Q only.
    long long n_terms = 0;
    {
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) Pw[i][j] = (i == j) ? 1.0 : 0.0;
        double mx;
        do {
            ++n_terms;
            mx = times_Q(Pw);
        } while (mx >= 1e-30 && n_terms < tile + 1);
    }
    // 2. the terms, 32 at a time: every lane waits for ITS record (all of them are needed, and the waits overlap: one L2 round
    //    trip after the last of them is published), then they are summed in order, nearest first
#pragma unroll
    for (int i = 0; i < D; ++i) {
        acc[i] = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) Pw[i][j] = (i == j) ? 1.0 : 0.0;
    }
    for (long long k0 = 0; k0 < n_terms; k0 += 32) {
        const long long k = k0 + lane, j = tile - 1 - k;
        double a[D];
#pragma unroll
        for (int q = 0; q < D; ++q) a[q] = 0.0;
        if (k < n_terms) {
            if (j >= 0) {
                while (rec_fetch<D>(S.rec, rec0 + j, a) < f_agg) {}
            } else {            // beginning of the signal: s_0 = zi * x_0 (the first sample of the extended sweep)
                const double x0 = sweep_read(P, base, dc, 0);
#pragma unroll
                for (int q = 0; q < D; ++q) a[q] = c.zi[q] * x0;
            }
        }
        __syncwarp();
        const int cnt = (int)((n_terms - k0 < 32) ? (n_terms - k0) : 32);
        for (int l = 0; l < cnt; ++l) {
            double al[D];
#pragma unroll
            for (int q = 0; q < D; ++q) al[q] = __shfl_sync(0xffffffffu, a[q], l);
#pragma unroll
            for (int i = 0; i < D; ++i) {
                double v = acc[i];
#pragma unroll
                for (int q = 0; q < D; ++q) v = fma(Pw[i][q], al[q], v);
                acc[i] = v;
            }
            times_Q(Pw);
        }
    }}

template <int D>
__global__ void __launch_bounds__(kFusedThreads, 768 / kTileThreads) iir_tile_fused_kernel(const IirPass P, const IirCoef c, const IirCoef ct, const double* __restrict__ ppow,
                                                                        const IirLookback S) {
    extern __shared__ double tile_sm[];
    __shared__ double red[kTileThreads / 32][2];
    __shared__ double wtot[kTileThreads / 32][2];
    __shared__ double sin_sh[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // CTAs are dispatched in index order, so every predecessor of a running CTA (a lower index) is running or done (the assumption
    // cub::DeviceScan's decoupled look-back makes).  Signal-minor order: CTA b = tile (b / n_sig) of signal (b % n_sig), so the
    // tiles a look-back reads belong to CTAs n_sig, 2 n_sig, ... positions earlier in the dispatch order -- they have a head start
    // instead of having been launched in the same instant, and their aggregates are usually there when the look-back asks.
    const unsigned tk = blockIdx.x;
    const long long n_tiles = P.n_chunks;
    const int s = (int)(tk % (unsigned)P.n_sig);
    const long long tile = tk / (unsigned)P.n_sig;
    const long long rec0 = (long long)s * n_tiles;
    if (warp == kTileThreads / 32) {
        double acc[D];
        lookback_start_state<D>(P, c, ct, S, s, tile, rec0, lane, acc);
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < D; ++i) sin_sh[i] = acc[i];
        }
        asm volatile("bar.sync 2, %0;" ::"n"(kFusedThreads) : "memory");      // rendezvous: sin_sh is in place
        return;
    }
    double z[D];
    tile_stage_and_local<D, true>(P, c, tile_sm, s, tile, z);
    // ---- aggregate of the tile: e = sum_t P^(255-t) z_t (fixed-order block reduction), published at once
    {
        double w[D];
        mv2<D>(ppow + (size_t)(kTileThreads - 1 - threadIdx.x) * 4, z, w);
#pragma unroll
        for (int k = 0; k < D; ++k) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) w[k] += __shfl_xor_sync(0xffffffffu, w[k], off);
            if (lane == 0) red[warp][k] = w[k];
        }
    }
    tile_sync<true>();
    if (threadIdx.x == 0) {
        double e[2] = {0.0, 0.0};
#pragma unroll
        for (int k = 0; k < D; ++k) {
            double acc = 0.0;
            for (int q = 0; q < kTileThreads / 32; ++q) acc += red[q][k];
            e[k] = acc;
        }
        st_rec(S.rec + rec0 + tile, e[0], e[1]);
    }
    // ---- inclusive scan of the zero-state pieces inside the warp:  v_l = sum_{l' <= l} P^(l-l') z_l'
    double v[D];
#pragma unroll
    for (int k = 0; k < D; ++k) v[k] = z[k];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        double o[D], t[D];
#pragma unroll
        for (int q = 0; q < D; ++q) o[q] = __shfl_up_sync(0xffffffffu, v[q], 1 << k);
        mv2<D>(ppow + (size_t)(1 << k) * 4, o, t);
        if (lane >= (1 << k)) {
#pragma unroll
            for (int q = 0; q < D; ++q) v[q] += t[q];
        }
    }
    if (lane == 31) {
#pragma unroll
        for (int q = 0; q < D; ++q) wtot[warp][q] = v[q];
    }
    asm volatile("bar.sync 2, %0;" ::"n"(kFusedThreads) : "memory");      // rendezvous with the look-back warp: sin_sh is known
    double W[D];
#pragma unroll
    for (int q = 0; q < D; ++q) W[q] = 0.0;
    for (int j = 0; j < warp; ++j) {
        double t[D];
        mv2<D>(ppow + (size_t)32 * 4, W, t);
#pragma unroll
        for (int q = 0; q < D; ++q) W[q] = t[q] + wtot[j][q];
    }
    double sin_[D], st[D], e2[D];
#pragma unroll
    for (int q = 0; q < D; ++q) sin_[q] = sin_sh[q];
    mv2<D>(ppow + (size_t)threadIdx.x * 4, sin_, st);
    mv2<D>(ppow + (size_t)lane * 4, W, e2);
#pragma unroll
    for (int q = 0; q < D; ++q) {
        const double prev = __shfl_up_sync(0xffffffffu, v[q], 1);
        st[q] += e2[q] + (lane ? prev : 0.0);
    }
#pragma unroll 8
    for (int i = 0; i < kTilePer; ++i) {
        double* slot = tile_sm + i * kTileLd + threadIdx.x;
        *slot = df2t_step<D>(c, st, *slot);
    }
    tile_sync<true>();
    double* ob = P.out + (long long)s * P.out_sig_stride;
    const long long u0 = tile * kTile;
    const long long n = P.L - 2 * (long long)P.e;
    if (u0 + kTile <= P.L && (P.forward || (P.L - u0 - kTile - P.e >= 0 && P.L - 1 - u0 - P.e < n))) {
        // interior tile: branch-free coalesced stores, 8 values per thread in flight
#pragma unroll 4
        for (int it = 0; it < kTilePer; ++it) {
            const int idx = it * kTileThreads + threadIdx.x;
            const double y = tile_sm[(idx & (kTilePer - 1)) * kTileLd + (idx >> kTilePerLog2)];
            if (P.forward) ob[u0 + idx] = y;
            else ob[P.L - 1 - (u0 + idx) - P.e] = y;
        }
    } else {
        for (int idx = threadIdx.x; idx < kTile; idx += kTileThreads) {
            const long long u = u0 + idx;
            if (u >= P.L) break;
            const double y = tile_sm[(idx & (kTilePer - 1)) * kTileLd + (idx >> kTilePerLog2)];
            if (P.forward) {
                ob[u] = y;
            } else {
                const long long t = P.L - 1 - u - P.e;
                if (t >= 0 && t < n) ob[t] = y;
            }
        }
    }
}

// ------------------------------------------------------------------ K1, single-pass sweep with TWO tiles per CTA in a software pipeline
// Same arithmetic as iir_tile_fused_kernel (bit-identical results), different schedule: a CTA owns two consecutive tiles A, B of a
// signal.  Both tiles are brought in by cp.async (8-byte copies straight to their transposed slots: no registers, no waiting thread),
// then  scan A | publish A | scan B | publish B | apply A | store A | apply B | store B  -- B's loads land under A's scan, A's look-back
// (and the stores of A) under B's phases.  The DC removal of the first sweep moves from the staging to the two reads of a sample.
// Measured on the cfg4 cascade (parity tests green, bit-reproducible): 3.77 ms with 2 x 4096 samples per CTA (three CTAs per SM) and
// 4.08 ms with 2 x 2048 (six CTAs per SM) against 3.50 ms for one 4096-sample tile per CTA, six CTAs per SM: six independent CTAs
// overlap their phases better than three pipelined ones, and smaller tiles pay for more look-back terms and barriers per byte.
// HS_EXPERIMENT builds only (HS_IIR_TWO_TILES=1).
#ifdef HS_EXPERIMENT

__device__ __forceinline__ void cp_async8(double* dst_smem, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}

template <int D>
__global__ void __launch_bounds__(kFusedThreads, 384 / kTileThreads) iir_tile_fused2_kernel(const IirPass P, const IirCoef c, const IirCoef ct,
                                                                                            const double* __restrict__ ppow, const IirLookback S) {
    extern __shared__ double tile_sm2[];                    // [2][kTilePer * kTileLd]
    __shared__ double red[kTileThreads / 32][2];
    __shared__ double wtot[2][kTileThreads / 32][2];
    __shared__ double sin_sh[2][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned tk = blockIdx.x;                          // signal-minor order of the tile PAIRS (see iir_tile_fused_kernel)
    const long long n_tiles = P.n_chunks;
    const int s = (int)(tk % (unsigned)P.n_sig);
    const long long tile0 = 2 * (long long)(tk / (unsigned)P.n_sig);
    const int n_here = (tile0 + 1 < n_tiles) ? 2 : 1;
    const long long rec0 = (long long)s * n_tiles;
    if (warp == kTileThreads / 32) {
        // look-back warp: the start states of A and B (B's first term is A's aggregate, published by this CTA's compute warps)
        for (int q = 0; q < n_here; ++q) {
            double acc[D];
            lookback_start_state<D>(P, c, ct, S, s, tile0 + q, rec0, lane, acc);
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < D; ++i) sin_sh[q][i] = acc[i];
            }
            __syncwarp();
            if (q == 0) asm volatile("bar.arrive 2, %0;" ::"n"(kFusedThreads) : "memory");
            else asm volatile("bar.arrive 3, %0;" ::"n"(kFusedThreads) : "memory");
        }
        return;
    }
    const double* base = P.in + (long long)s * P.in_sig_stride;
    const double dc = P.mean ? P.mean[s] : 0.0;
    double dcq[2] = {0.0, 0.0};                              // what the scans still have to subtract from a staged sample
    // ---- stage both tiles
    for (int q = 0; q < n_here; ++q) {
        double* sm = tile_sm2 + (size_t)q * kTilePer * kTileLd;
        const long long u0 = (tile0 + q) * kTile;
        const bool interior = P.in_t_stride == 1 && (P.forward ? (u0 >= P.e && u0 + kTile <= P.e + P.n_in) : (u0 + kTile <= P.L));
        if (interior) {
            const double* src = P.forward ? base + (u0 - P.e) : base + (P.L - u0 - kTile);
#pragma unroll 8
            for (int k = 0; k < kTilePer; ++k) {
                const int j = k * kTileThreads + threadIdx.x;
                const int i0 = P.forward ? j : kTile - 1 - j;
                cp_async8(sm + (i0 & (kTilePer - 1)) * kTileLd + (i0 >> kTilePerLog2), src + j);
            }
            dcq[q] = dc;
        } else {
            for (int idx = threadIdx.x; idx < kTile; idx += kTileThreads) {
                const long long u = u0 + idx;
                sm[(idx & (kTilePer - 1)) * kTileLd + (idx >> kTilePerLog2)] = (u < P.L) ? sweep_read(P, base, dc, u) : 0.0;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // ---- zero-state pieces, aggregates, warp scans
    double v[2][D];
    for (int q = 0; q < n_here; ++q) {
        double* sm = tile_sm2 + (size_t)q * kTilePer * kTileLd;
        if (q == 0 && n_here == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        tile_sync<true>();
        double z[D];
#pragma unroll
        for (int k = 0; k < D; ++k) z[k] = 0.0;
        const double dq = dcq[q];
#pragma unroll 8
        for (int i = 0; i < kTilePer; ++i) df2t_step<D>(c, z, sm[i * kTileLd + threadIdx.x] - dq);
        {
            double w[D];
            mv2<D>(ppow + (size_t)(kTileThreads - 1 - threadIdx.x) * 4, z, w);
#pragma unroll
            for (int k = 0; k < D; ++k) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) w[k] += __shfl_xor_sync(0xffffffffu, w[k], off);
                if (lane == 0) red[warp][k] = w[k];
            }
        }
        tile_sync<true>();
        if (threadIdx.x == 0) {
            double e[2] = {0.0, 0.0};
#pragma unroll
            for (int k = 0; k < D; ++k) {
                double acc = 0.0;
                for (int r = 0; r < kTileThreads / 32; ++r) acc += red[r][k];
                e[k] = acc;
            }
            st_rec(S.rec + rec0 + tile0 + q, e[0], e[1]);
        }
#pragma unroll
        for (int k = 0; k < D; ++k) v[q][k] = z[k];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            double o[D], t[D];
#pragma unroll
            for (int r = 0; r < D; ++r) o[r] = __shfl_up_sync(0xffffffffu, v[q][r], 1 << k);
            mv2<D>(ppow + (size_t)(1 << k) * 4, o, t);
            if (lane >= (1 << k)) {
#pragma unroll
                for (int r = 0; r < D; ++r) v[q][r] += t[r];
            }
        }
        if (lane == 31) {
#pragma unroll
            for (int r = 0; r < D; ++r) wtot[q][warp][r] = v[q][r];
        }
    }
    // ---- true states, outputs
    double* ob = P.out + (long long)s * P.out_sig_stride;
    const long long n = P.L - 2 * (long long)P.e;
    for (int q = 0; q < n_here; ++q) {
        double* sm = tile_sm2 + (size_t)q * kTilePer * kTileLd;
        if (q == 0) asm volatile("bar.sync 2, %0;" ::"n"(kFusedThreads) : "memory");      // sin_sh[q] is known (and wtot[q] complete)
        else asm volatile("bar.sync 3, %0;" ::"n"(kFusedThreads) : "memory");
        double W[D];
#pragma unroll
        for (int r = 0; r < D; ++r) W[r] = 0.0;
        for (int j = 0; j < warp; ++j) {
            double t[D];
            mv2<D>(ppow + (size_t)32 * 4, W, t);
#pragma unroll
            for (int r = 0; r < D; ++r) W[r] = t[r] + wtot[q][j][r];
        }
        double sin_[D], st[D], e2[D];
#pragma unroll
        for (int r = 0; r < D; ++r) sin_[r] = sin_sh[q][r];
        mv2<D>(ppow + (size_t)threadIdx.x * 4, sin_, st);
        mv2<D>(ppow + (size_t)lane * 4, W, e2);
#pragma unroll
        for (int r = 0; r < D; ++r) {
            const double prev = __shfl_up_sync(0xffffffffu, v[q][r], 1);
            st[r] += e2[r] + (lane ? prev : 0.0);
        }
        const double dq = dcq[q];
#pragma unroll 8
        for (int i = 0; i < kTilePer; ++i) {
            double* slot = sm + i * kTileLd + threadIdx.x;
            *slot = df2t_step<D>(c, st, *slot - dq);
        }
        tile_sync<true>();
        const long long u0 = (tile0 + q) * kTile;
        if (u0 + kTile <= P.L && (P.forward || (P.L - u0 - kTile - P.e >= 0 && P.L - 1 - u0 - P.e < n))) {
#pragma unroll 4
            for (int it = 0; it < kTilePer; ++it) {
                const int idx = it * kTileThreads + threadIdx.x;
                const double y = sm[(idx & (kTilePer - 1)) * kTileLd + (idx >> kTilePerLog2)];
                if (P.forward) ob[u0 + idx] = y;
                else ob[P.L - 1 - (u0 + idx) - P.e] = y;
            }
        } else {
            for (int idx = threadIdx.x; idx < kTile; idx += kTileThreads) {
                const long long u = u0 + idx;
                if (u >= P.L) break;
                const double y = sm[(idx & (kTilePer - 1)) * kTileLd + (idx >> kTilePerLog2)];
                if (P.forward) {
                    ob[u] = y;
                } else {
                    const long long t = P.L - 1 - u - P.e;
                    if (t >= 0 && t < n) ob[t] = y;
                }
            }
        }
    }
}

#endif  // HS_EXPERIMENT (two-tile pipelined sweep)

// per-signal mean in two deterministic stages: kMeanParts CTAs per signal reduce contiguous slices (fixed tree order),
// one more CTA per signal adds the partials in index order.
constexpr int kMeanParts = 64;
__global__ void __launch_bounds__(256) mean_partial_kernel(const double* __restrict__ x, long long n, long long sig_stride, long long t_stride,
                                                           double* __restrict__ partial) {
    __shared__ double sh[256];
    const double* base = x + (long long)blockIdx.y * sig_stride;
    const long long per = (n + kMeanParts - 1) / kMeanParts;
    const long long t0 = (long long)blockIdx.x * per, t1 = min(n, t0 + per);
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    long long t = t0 + threadIdx.x;
    for (; t + 3 * 256 < t1; t += 4 * 256) {
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] += base[(t + k * 256) * t_stride];
    }
    for (; t < t1; t += 256) acc[0] += base[t * t_stride];
    sh[threadIdx.x] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[(long long)blockIdx.y * kMeanParts + blockIdx.x] = sh[0];
}

__global__ void mean_finish_kernel(const double* __restrict__ partial, long long n, int n_sig, double* __restrict__ mean) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sig) return;
    double acc = 0.0;
    for (int k = 0; k < kMeanParts; ++k) acc += partial[(long long)s * kMeanParts + k];
    mean[s] = acc / (double)n;
}

__global__ void sub_mean_kernel(double* x, long long n, long long sig_stride, long long t_stride, const double* mean, int n_sig) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * n_sig) return;
    const int s = (int)(i / n);
    const long long t = i - (long long)s * n;
    x[(long long)s * sig_stride + t * t_stride] -= mean[s];
}

// ------------------------------------------------------------------ host-side filter preparation
static bool solve_small(int d, std::vector<double>& Amat, std::vector<double>& rhs) {
    for (int col = 0; col < d; ++col) {
        int piv = col;
        for (int r = col + 1; r < d; ++r)
            if (fabs(Amat[r * d + col]) > fabs(Amat[piv * d + col])) piv = r;
        if (Amat[piv * d + col] == 0.0) return false;
        if (piv != col) {
            for (int k = 0; k < d; ++k) std::swap(Amat[piv * d + k], Amat[col * d + k]);
            std::swap(rhs[piv], rhs[col]);
        }
        for (int r = col + 1; r < d; ++r) {
            const double f = Amat[r * d + col] / Amat[col * d + col];
            for (int k = col; k < d; ++k) Amat[r * d + k] -= f * Amat[col * d + k];
            rhs[r] -= f * rhs[col];
        }
    }
    for (int r = d - 1; r >= 0; --r) {
        double acc = rhs[r];
        for (int k = r + 1; k < d; ++k) acc -= Amat[r * d + k] * rhs[k];
        rhs[r] = acc / Amat[r * d + r];
    }
    return true;
}

static int prepare_filter(const double* b, const double* a, int ntaps, IirCoef* c) {
    const int d = ntaps - 1;
    if (d < 1 || d > kMaxOrder) return set_error(HS_ERR_UNSUPPORTED, "filtfilt: filter order %d not in 1..%d", d, kMaxOrder);
    if (a[0] == 0.0) return set_error(HS_ERR_INVALID, "filtfilt: a[0] == 0");
    memset(c, 0, sizeof(*c));
    c->d = d;
    c->e = 3 * ntaps;
    for (int k = 0; k < ntaps; ++k) {
        c->b[k] = b[k] / a[0];
        c->a[k] = a[k] / a[0];
    }
    // lfilter_zi: (I - companion(a)^T) zi = b[1:] - a[1:] * b[0]
    std::vector<double> Am(d * d, 0.0), rhs(d);
    for (int i = 0; i < d; ++i) {
        for (int j = 0; j < d; ++j) {
            double comp_t = 0.0;                 // companion(a)^T[i][j] = companion[j][i]
            if (j == 0) comp_t = -c->a[i + 1];   // companion[0][i] = -a[i+1]
            else if (i == j - 1) comp_t = 1.0;   // companion[j][j-1] = 1
            Am[i * d + j] = (i == j ? 1.0 : 0.0) - comp_t;
        }
        rhs[i] = c->b[i + 1] - c->a[i + 1] * c->b[0];
    }
    if (!solve_small(d, Am, rhs)) return set_error(HS_ERR_INVALID, "filtfilt: singular lfilter_zi system (filter has a pole at z = 1)");
    for (int i = 0; i < d; ++i) c->zi[i] = rhs[i];
    // (M^C)^(2^k) = M^(C 2^k): column j = state after C 2^k homogeneous DF2T steps from e_j
    // (z'_k = z_{k+1} - a_{k+1} z_0).  Running the recurrence itself (long double) keeps the natural rounding of the
    // recursion; repeated squaring of this non-normal matrix loses ~cond(V)^2 eps (1e-3 for Butterworth-4 at 1 Hz).
    long double zst[kMaxOrder][kMaxOrder];      // zst[j] = current image of e_j
    for (int j = 0; j < d; ++j)
        for (int k = 0; k < d; ++k) zst[j][k] = (j == k) ? 1.0L : 0.0L;
    long long done = 0;
    c->warm = 0;
    bool warm_found = (d <= 2);
    for (int kpow = 0; kpow < 6; ++kpow) {
        const long long target = (long long)kChunk << kpow;
        for (; done < target; ++done) {
            for (int j = 0; j < d; ++j) {
                const long double y = zst[j][0];
                for (int k = 0; k < d - 1; ++k) zst[j][k] = zst[j][k + 1] - (long double)c->a[k + 1] * y;
                zst[j][d - 1] = -(long double)c->a[d] * y;
            }
            if (!warm_found && (done + 1) % kChunk == 0) {
                long double mx = 0.0L;
                for (int j = 0; j < d; ++j)
                    for (int k = 0; k < d; ++k) mx = fmaxl(mx, fabsl(zst[j][k]));
                if (mx < 1e-6L) {
                    c->warm = (int)((done + 1) / kChunk);
                    warm_found = true;
                }
            }
        }
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) c->pw[kpow][i * d + j] = (double)zst[j][i];
    }
    if (!warm_found) c->warm = 32;      // very slow poles: replay at most 32 chunks (documented accuracy limit)
    return HS_OK;
}

// Tables of the tiled path: ppow[j] = M^(32 j), j = 0..256, by running the homogeneous recurrence in long double;
// ct->pw[k] = (M^8192)^(2^k) by squaring M^8192 in long double (for a stable biquad these are tiny numbers).
static void prepare_tile_tables(const IirCoef& c, std::vector<double>& ppow, IirCoef* ct) {
    const int d = c.d;
    ppow.assign((size_t)kPpowEntries * 4, 0.0);
    long double zst[2][2] = {{1.0L, 0.0L}, {0.0L, 1.0L}};        // zst[j] = image of e_j
    for (int j = 0; j <= kTileThreads; ++j) {
        for (int i = 0; i < d; ++i)
            for (int q = 0; q < d; ++q) ppow[(size_t)j * 4 + i * 2 + q] = (double)zst[q][i];
        if (j == kTileThreads) break;
        for (int step = 0; step < kTilePer; ++step) {
            for (int q = 0; q < d; ++q) {
                const long double y = zst[q][0];
                for (int k = 0; k < d - 1; ++k) zst[q][k] = zst[q][k + 1] - (long double)c.a[k + 1] * y;
                zst[q][d - 1] = -(long double)c.a[d] * y;
            }
        }
    }
    *ct = c;
    long double Q[2][2] = {{0, 0}, {0, 0}};
    for (int i = 0; i < d; ++i)
        for (int q = 0; q < d; ++q) Q[i][q] = zst[q][i];
    for (int kpow = 0; kpow < 6; ++kpow) {
        for (int i = 0; i < d; ++i)
            for (int q = 0; q < d; ++q) ct->pw[kpow][i * d + q] = (double)Q[i][q];
        long double R[2][2] = {{0, 0}, {0, 0}};
        for (int i = 0; i < d; ++i)
            for (int q = 0; q < d; ++q)
                for (int r = 0; r < d; ++r) R[i][q] += Q[i][r] * Q[r][q];
        for (int i = 0; i < d; ++i)
            for (int q = 0; q < d; ++q) Q[i][q] = R[i][q];
    }
}

template <int D>
static int run_sweep_fused(IirPass P, const IirCoef& c, const IirCoef& ct, const double* d_ppow, const IirLookback& S, cudaStream_t st) {
    const long long n_tiles = (P.L + kTile - 1) / kTile;
    P.n_chunks = n_tiles;
    const size_t smem = (size_t)kTilePer * kTileLd * sizeof(double);
    static bool attr_done[3] = {false, false, false};
    if (!attr_done[D]) {
        if (cudaFuncSetAttribute(iir_tile_fused_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return set_error(HS_ERR_CUDA, "filtfilt: cannot reserve %zu B shared memory", smem);
        cudaFuncSetAttribute(iir_tile_fused_kernel<D>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr_done[D] = true;
    }
    const long long ctas = n_tiles * P.n_sig;
    if (ctas > 0x7fffffffLL) return set_error(HS_ERR_UNSUPPORTED, "filtfilt: too many tiles");
#ifdef HS_EXPERIMENT
    static const int two_tiles = exp_env_int("HS_IIR_TWO_TILES", 0);
    if (two_tiles) {
        static bool attr2_done[3] = {false, false, false};
        if (!attr2_done[D]) {
            if (cudaFuncSetAttribute(iir_tile_fused2_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * smem)) != cudaSuccess)
                return set_error(HS_ERR_CUDA, "filtfilt: cannot reserve %zu B shared memory", 2 * smem);
            cudaFuncSetAttribute(iir_tile_fused2_kernel<D>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            attr2_done[D] = true;
        }
        const long long pairs = (n_tiles + 1) / 2 * P.n_sig;
        iir_tile_fused2_kernel<D><<<(unsigned)pairs, kFusedThreads, 2 * smem, st>>>(P, c, ct, d_ppow, S);
        return check_launch("iir_tile_fused2_kernel");
    }
#endif
    iir_tile_fused_kernel<D><<<(unsigned)ctas, kFusedThreads, smem, st>>>(P, c, ct, d_ppow, S);
    return check_launch("iir_tile_fused_kernel");
}

template <int D>
static int run_sweep_tiled(IirPass P, const IirCoef& c, const IirCoef& ct, const double* d_ppow, cudaStream_t st) {
    const long long n_tiles = (P.L + kTile - 1) / kTile;
    P.n_chunks = n_tiles;
    const size_t smem = (size_t)kTilePer * kTileLd * sizeof(double);
    static bool attr_done[3] = {false, false, false};
    if (!attr_done[D]) {
        if (cudaFuncSetAttribute(iir_tile_local_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(iir_tile_apply_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return set_error(HS_ERR_CUDA, "filtfilt: cannot reserve %zu B shared memory", smem);
        attr_done[D] = true;
    }
    dim3 grid((unsigned)n_tiles, (unsigned)P.n_sig);
    iir_tile_local_kernel<D><<<grid, kTileThreads, smem, st>>>(P, c, d_ppow);
    int rc = check_launch("iir_tile_local_kernel");
    if (rc) return rc;
    iir_carry_kernel<D><<<(P.n_sig + 3) / 4, 128, 0, st>>>(P, ct);
    rc = check_launch("iir_carry_kernel");
    if (rc) return rc;
    iir_tile_apply_kernel<D><<<grid, kTileThreads, smem, st>>>(P, c, d_ppow);
    return check_launch("iir_tile_apply_kernel");
}

template <int D>
static int run_sweep(const IirPass& P, const IirCoef& c, cudaStream_t st) {
    const long long n_thr = (long long)P.n_sig * P.n_chunks;
    const int blocks = (int)((n_thr + 127) / 128);
    iir_local_kernel<D><<<blocks, 128, 0, st>>>(P, c);
    int rc = check_launch("iir_local_kernel");
    if (rc) return rc;
    iir_carry_kernel<D><<<(P.n_sig + 3) / 4, 128, 0, st>>>(P, c);
    rc = check_launch("iir_carry_kernel");
    if (rc) return rc;
    iir_apply_kernel<D><<<blocks, 128, 0, st>>>(P, c);
    return check_launch("iir_apply_kernel");
}

static int run_sweep_d(int d, const IirPass& P, const IirCoef& c, cudaStream_t st) {
    switch (d) {
        case 1: return run_sweep<1>(P, c, st);
        case 2: return run_sweep<2>(P, c, st);
        case 3: return run_sweep<3>(P, c, st);
        case 4: return run_sweep<4>(P, c, st);
    }
    return set_error(HS_ERR_UNSUPPORTED, "filtfilt: order %d", d);
}

// ------------------------------------------------------------------ K2
// y[k] = sum_j b[j] x[q k + half - j], zero outside [0, n)   (scipy.signal.decimate(..., ftype='fir', zero_phase=True),
// data_structures.py:792).  HBM-bound: every input sample is read from global memory exactly once.
// One CTA = kDecOut consecutive outputs of one signal.  The input span (q*kDecOut + ntaps - 1 samples) is staged in shared
// memory de-interleaved into its q polyphase components, so that for a fixed tap every thread reads the same component at
// consecutive positions.  With jj = ntaps-1-j:  y[o] = sum_r sum_u b[ntaps-1-(q u + r)] * S_r[o + u]:  a thread owns 4
// consecutive outputs and slides a 4-register window along S_r (1 shared load per 4 FMAs).  Positions are skewed by
// pos + (pos >> kDecSkewShift) so the 32 B-strided window loads of a half-warp hit distinct banks.
constexpr int kDecThreads = 128;           // 512 outputs per CTA: 42 KB of staged input (q = 8), 5 CTAs/SM (256 threads: 2 CTAs/SM, 0.118 vs 0.107 ms)
constexpr int kDecPer = 4;                 // (8 per thread with 128-thread CTAs measured 45 % slower: too few loads in flight)
constexpr int kDecSkewShift = 2;           // pos + (pos >> 2): lanes 4 positions apart land in distinct banks
constexpr int kDecOut = kDecThreads * kDecPer;

__host__ __device__ inline int dec_phase_len(int q, int ntaps) { return kDecOut + (ntaps - 1 + q - 1) / q + 1; }
__host__ __device__ inline int dec_phase_stride(int q, int ntaps) {
    const int pl = dec_phase_len(q, ntaps);
    int st = pl + (pl >> kDecSkewShift) + 1;
    st += (2 - (st & 15) + 16) & 15;          // = 2 (mod 16) doubles: the q components start 16 B apart modulo 128 B
    return st;
}

template <int Q>      // Q > 0: compile-time decimation factor (divisions become shifts for powers of two); 0: runtime q
__global__ void __launch_bounds__(kDecThreads) fir_decimate_kernel(const double* __restrict__ x, long long n, long long sig_stride, int q_rt,
                                                                   long long off, const double* __restrict__ b, int ntaps,
                                                                   double* __restrict__ y, long long n_out, long long y_stride) {
    extern __shared__ double dec_smem[];
    const int q = Q > 0 ? Q : q_rt;
    const int pst = dec_phase_stride(q, ntaps);
    double* S = dec_smem;                       // [q][pst]
    double* taps = dec_smem + (size_t)q * pst;  // taps[jj] = b[ntaps-1-jj]
    const long long k0 = (long long)blockIdx.x * kDecOut;
    const double* xs = x + (long long)blockIdx.y * sig_stride;
    const long long lo = (long long)q * k0 + off - (ntaps - 1);      // input index of staged element 0 (y[k] = sum_j b[j] x[q k + off - j])
    const int span = q * kDecOut + ntaps - 1;
    if (lo >= 0 && lo + span <= n) {          // interior span: branch-free coalesced loads, 8 in flight per thread
        const double* src = xs + lo;
        int i = threadIdx.x;
        for (; i + 7 * kDecThreads < span; i += 8 * kDecThreads) {
            double v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = src[i + k * kDecThreads];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int ii = i + k * kDecThreads, r = ii % q, pos = ii / q;
                S[(size_t)r * pst + pos + (pos >> kDecSkewShift)] = v[k];
            }
        }
        for (; i < span; i += kDecThreads) {
            const int r = i % q, pos = i / q;
            S[(size_t)r * pst + pos + (pos >> kDecSkewShift)] = src[i];
        }
    } else {
        for (int i = threadIdx.x; i < span; i += kDecThreads) {
            const long long g = lo + i;
            const int r = i % q, pos = i / q;
            S[(size_t)r * pst + pos + (pos >> kDecSkewShift)] = (g >= 0 && g < n) ? xs[g] : 0.0;
        }
    }
    for (int j = threadIdx.x; j < ntaps; j += kDecThreads) taps[j] = b[ntaps - 1 - j];
    __syncthreads();
    const int o0 = threadIdx.x * kDecPer;
    double acc[kDecPer];
#pragma unroll
    for (int i = 0; i < kDecPer; ++i) acc[i] = 0.0;
    for (int r = 0; r < q; ++r) {
        const double* Sr = S + (size_t)r * pst;
        const int nu = (ntaps - r + q - 1) / q;            // taps jj = q u + r < ntaps
        double w[kDecPer];
#pragma unroll
        for (int i = 0; i < kDecPer - 1; ++i) {
            const int pos = o0 + i;
            w[i + 1] = Sr[pos + (pos >> kDecSkewShift)];
        }
        for (int u = 0; u < nu; ++u) {
#pragma unroll
            for (int i = 0; i < kDecPer - 1; ++i) w[i] = w[i + 1];
            const int pos = o0 + kDecPer - 1 + u;
            w[kDecPer - 1] = Sr[pos + (pos >> kDecSkewShift)];
            const double t = taps[q * u + r];
#pragma unroll
            for (int i = 0; i < kDecPer; ++i) acc[i] = fma(t, w[i], acc[i]);
        }
    }
    double* ys = y + (long long)blockIdx.y * y_stride + k0 + o0;
#pragma unroll
    for (int i = 0; i < kDecPer; ++i)
        if (k0 + o0 + i < n_out) ys[i] = acc[i];
}

}  // namespace hs

using namespace hs;

constexpr int kMaxTiledFilters = 16;

struct PreparedFilter {
    int ntaps;
    std::vector<double> b, a, ppow;
    hs::IirCoef c, ct;
};
static std::mutex g_prep_mutex;
static std::deque<PreparedFilter> g_prep_cache;

// look-back records: one set per sweep of a call (2 sweeps per tiled filter), 16 bytes per (signal, tile)
static size_t lookback_bytes(int n_sig, long long tiles) {
    const size_t recs = (size_t)n_sig * (size_t)tiles;
    return (size_t)2 * kMaxTiledFilters * recs * sizeof(double2) + 256;
}


// ------------------------------------------------------------------ K2 on the FP64 tensor pipe
// FIR filtering / decimation as a Toeplitz product in the natural sample index: for 8 consecutive outputs o0 .. o0 + 7 of 8 SIGNALS,
//     C[s][n] = sum_e S_s[q o0 + e] B[e][n],      B[e][n] = taps[e - q n]  (0 outside the filter),   e = 0 .. 7 q + ntaps - 1,
// i.e. ceil((7 q + ntaps) / 4) k-steps of mma.sync.m8n8k4.f64 per 8 x 8 outputs (q = 8, 161 taps: 55 DMMAs, 74 % of their FMAs on
// non-zero band entries; q = 1, 201 taps: 97 %).  B is a fixed banded Toeplitz matrix: its fragments are constants per lane, held in
// registers for the whole kernel (the two halves of the k range go to two warps, whose partial sums are added through shared
// memory); an A fragment is one shared load of CONSECUTIVE samples (lane (g4, t4): signal g4, sample q o0 + 4 kk + t4), so the input
// spans are staged exactly as they lie in memory -- 16-byte cp.async copies straight from global to shared memory, double buffered
// (the next tile's copies run under this tile's products) -- with a signal stride = 4 (mod 16) doubles for conflict-free fragment loads.  A DMMA (256 FMAs) costs one
// LDS instead of the 64 LDS + address computations the register-window kernel spends on as many FMAs: that kernel is bound by
// instruction issue (ncu: LSU 69 %, issue 57 %, FP64 pipe at a third of its peak), not by memory.
// One CTA = 8 signals x kMmaTilesPerCta tiles of kMmaOut outputs, 8 warps.  Used when 7 q + ntaps <= 224 and there are >= 5 signals.
constexpr int kMmaOut = 64;                                 // outputs per tile
constexpr int kMmaTilesPerCta = 16;                         // consecutive output tiles per CTA: the B fragments are built once per CTA
__host__ __device__ inline int mma_dec_ksteps(int q, int ntaps) { return (7 * q + ntaps + 3) / 4; }
__host__ __device__ inline int mma_dec_kh(int q, int ntaps) { const int ks = mma_dec_ksteps(q, ntaps); return ks <= 14 ? 7 : (ks <= 28 ? 14 : 28); }
__host__ __device__ inline int mma_dec_sig_stride(int q, int ntaps) {
    const int span = q * kMmaOut + ntaps - 1, reach = q * (kMmaOut - 8) + 8 * mma_dec_kh(q, ntaps);      // the k range is padded to 2 KH steps
    int st = span > reach ? span : reach;
    st += (4 - (st & 15) + 16) & 15;                        // = 4 (mod 16)
    return st;
}

__device__ __forceinline__ void dmma884_fe(double& c0, double& c1, const double a, const double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// Stage the input spans of the 8 signals of one tile exactly as they lie in memory, S[s][i] = x_s[lo + i] (zero outside the signal),
// with cp.async straight from global to shared memory: 16 bytes per copy on interior tiles with aligned spans (all of a thread's ~11
// copies in flight at once, committed as one group), 8-byte copies with bounds checks and zero fill on edge tiles.
// (Measured: one cp.async.bulk per signal and tile with a transaction mbarrier -- no copy instructions in the warps at all -- is slower
//  here, 0.58 against 0.31 ms for cfg4: with two buffers the copy of the next tile can only start one product phase (~0.9 us) ahead,
//  which a 5 KB bulk copy does not cover, and a third buffer costs the second CTA per SM.)
__device__ __forceinline__ void mma_dec_stage(double* __restrict__ buf, const double* __restrict__ x, const long long n,
                                              const long long sig_stride, const int n_sig, const int s0, const long long lo, const int span,
                                              const int ss, const bool aligned) {
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(buf);
    const bool whole = aligned && lo >= 0 && lo + span <= n;          // the span lies inside the signal and starts on a 16-byte boundary
    const int n_live = min(8, n_sig - s0);                             // a missing signal keeps the zeros the buffers start with
    if (whole) {
        const int chunks = span >> 1;
        const unsigned ss8 = 8u * (unsigned)ss;
        for (int c = threadIdx.x; c < chunks; c += 256) {              // two rounds for q = 8, 161 taps
            const double* src = x + (long long)s0 * sig_stride + lo + 2 * c;
            unsigned dst = sbase + 16u * c;
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                if (s < n_live) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                src += sig_stride;
                dst += ss8;
            }
        }
    } else {
#pragma unroll 1
        for (int s = 0; s < n_live; ++s) {
            const double* xs = x + (long long)(s0 + s) * sig_stride + lo;
            const unsigned ds = sbase + 8u * (unsigned)(s * ss);
            for (int i = threadIdx.x; i < span; i += 256) {
                if (lo + i >= 0 && lo + i < n) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(ds + 8u * i), "l"(xs + i) : "memory");
                else buf[s * ss + i] = 0.0;
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int KH>      // k-steps per warp half: 2 KH >= ceil((7 q + ntaps) / 4)
__global__ void __launch_bounds__(256, 2) fir_mma_kernel(const double* __restrict__ x, long long n, long long sig_stride, int n_sig, int q,
                                                         long long off, const double* __restrict__ b, int ntaps,
                                                         double* __restrict__ y, long long n_out, long long y_stride) {
    extern __shared__ __align__(16) double dec_smem[];
    const int ss = mma_dec_sig_stride(q, ntaps);
    double2* red = reinterpret_cast<double2*>(dec_smem + 2 * 8 * ss);  // [4 warps][2 tiles][32 lanes] partial sums of the upper k half
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g4 = lane >> 2, t4 = lane & 3;
    // warp (pw, nw): the k-steps [pw KH, (pw + 1) KH) of the 8-output tiles 2 nw, 2 nw + 1
    const int pw = warp >> 2, nw = warp & 3;
    const int s0 = blockIdx.y * 8;
    const int span = q * kMmaOut + ntaps - 1;
    const long long lo0 = off - (ntaps - 1);
    // bulk copies need 16-byte aligned sources and sizes: even span, even first index (q k0 + lo0 for every tile), even signal stride
    const bool aligned = ((span & 1) == 0) && ((lo0 & 1) == 0) && (((q * kMmaOut) & 1) == 0) && ((sig_stride & 1) == 0) &&
                         ((reinterpret_cast<size_t>(x) & 15) == 0);
    // the samples behind a span that the padded k range reaches meet zero taps, but must be finite: zero both buffers once
    for (int i = threadIdx.x; i < 2 * 8 * ss; i += 256) dec_smem[i] = 0.0;
    // ---- B fragments: lane (n = g4, e = 4 kk + t4) <- taps[e - q n] = b[ntaps - 1 - (e - q n)]
    double bf[KH];
#pragma unroll
    for (int k = 0; k < KH; ++k) {
        const int d = 4 * (pw * KH + k) + t4 - q * g4;
        bf[k] = (d >= 0 && d < ntaps) ? b[ntaps - 1 - d] : 0.0;
    }
    __syncthreads();
    const long long tile0 = (long long)blockIdx.x * kMmaTilesPerCta;
    if (tile0 * kMmaOut < n_out) mma_dec_stage(dec_smem, x, n, sig_stride, n_sig, s0, (long long)q * tile0 * kMmaOut + lo0, span, ss, aligned);
    for (int tile = 0; tile < kMmaTilesPerCta; ++tile) {
        const long long k0 = (tile0 + tile) * kMmaOut;
        if (k0 >= n_out) break;
        // the next tile's copy runs while this tile is multiplied (double buffer; the barrier after the products of tile - 1 freed it)
        const bool more = tile + 1 < kMmaTilesPerCta && k0 + kMmaOut < n_out;
        if (more) mma_dec_stage(dec_smem + ((tile + 1) & 1) * 8 * ss, x, n, sig_stride, n_sig, s0, (long long)q * (k0 + kMmaOut) + lo0, span, ss, aligned);
        if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const double* Sa = dec_smem + (tile & 1) * 8 * ss + g4 * ss + q * (nw * 16) + 4 * pw * KH + t4;
        double c[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int k = 0; k < KH; ++k) {
#pragma unroll
            for (int h = 0; h < 2; ++h) dmma884_fe(c[h][0], c[h][1], Sa[8 * q * h + 4 * k], bf[k]);
        }
        if (pw == 1) {
#pragma unroll
            for (int h = 0; h < 2; ++h) red[(nw * 2 + h) * 32 + lane] = make_double2(c[h][0], c[h][1]);
        }
        __syncthreads();
        if (pw == 0 && s0 + g4 < n_sig) {
            double* ys = y + (long long)(s0 + g4) * y_stride;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double2 o = red[(nw * 2 + h) * 32 + lane];
                const long long k = k0 + (2 * nw + h) * 8 + 2 * t4;
                if (k < n_out) ys[k] = c[h][0] + o.x;
                if (k + 1 < n_out) ys[k + 1] = c[h][1] + o.y;
            }
        }
        // (no barrier here: the next tile's copies go to the OTHER buffer, and the barrier that follows them comes before the next
        //  write of red and before anything overwrites this tile's buffer)
    }
}

extern "C" {

size_t hs_filtfilt_ws_bytes(int n_sig, int64_t n) {
    if (n_sig <= 0 || n <= 0) return 256;
    const long long L = n + 2 * 3 * (kMaxOrder + 1);
    const long long n_chunks = (L + kChunk - 1) / kChunk;
    size_t b = 0;
    b += ((size_t)n_sig * L * sizeof(double) + 255) / 256 * 256;                        // forward output f
    b += ((size_t)n_sig * n_chunks * kMaxOrder * sizeof(double) + 255) / 256 * 256;     // chunk states
    b += (((size_t)n_sig + 31) / 32 * 32 + (size_t)n_sig * kMeanParts) * sizeof(double) / 256 * 256 + 256;   // means + partial sums
    b += (size_t)kMaxTiledFilters * kPpowEntries * 4 * sizeof(double);                  // power tables of the tiled path
    b += lookback_bytes(n_sig, (L + kTile - 1) / kTile);                                // look-back records of the single-pass sweep
    return b + 256;
}

// causal_out == nullptr: zero-phase cascade in place (filtfilt).  Otherwise ONE causal filter (scipy.signal.lfilter, zero
// initial state): a single forward sweep without extension, x -> causal_out (row stride causal_stride).
static int iir_run(double* d_x, int n_sig, int64_t n, int64_t sig_stride, int64_t t_stride, const double* h_b, const double* h_a,
                   int n_filt, int ntaps, int remove_dc, void* d_ws, void* stream, double* causal_out, int64_t causal_stride) {
    const bool causal = causal_out != nullptr;
    if (!d_x || !d_ws) return set_error(HS_ERR_INVALID, "hs_iir_filtfilt_f64: null pointer");
    if (n_sig < 0 || n < 0 || n_filt < 0 || (n_filt > 0 && (!h_b || !h_a || ntaps < 2)))
        return set_error(HS_ERR_INVALID, "hs_iir_filtfilt_f64: bad arguments");
    if (n_sig == 0 || n == 0) return HS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (!causal && n_filt > 0 && n <= 3 * (int64_t)ntaps)
        return set_error(HS_ERR_INVALID, "The length of the input vector x must be greater than padlen, which is %d.", 3 * ntaps);
    unsigned char* ws = reinterpret_cast<unsigned char*>(d_ws);
    const long long Lmax = n + 2 * 3 * (kMaxOrder + 1);
    const long long chunks_max = (Lmax + kChunk - 1) / kChunk;
    double* f = reinterpret_cast<double*>(ws);
    ws += ((size_t)n_sig * Lmax * sizeof(double) + 255) / 256 * 256;
    double* states = reinterpret_cast<double*>(ws);
    ws += ((size_t)n_sig * chunks_max * kMaxOrder * sizeof(double) + 255) / 256 * 256;
    double* mean = reinterpret_cast<double*>(ws);
    ws += (((size_t)n_sig + 31) / 32 * 32 + (size_t)n_sig * kMeanParts) * sizeof(double) / 256 * 256 + 256;
    double* ppow_dev = reinterpret_cast<double*>(ws);
    static int tiled_mode = -1;      // HS_IIR_TILED=0 forces the thread-per-chunk kernels
    if (tiled_mode < 0) tiled_mode = exp_env_int("HS_IIR_TILED", 1) ? 1 : 0;
    static int fused_mode = -1;      // HS_IIR_FUSED=0: round 1's three-kernel tiled sweep (two reads of the signal per sweep)
    if (fused_mode < 0) fused_mode = exp_env_int("HS_IIR_FUSED", 1) ? 1 : 0;
    // look-back records of the single-pass sweep: behind the power tables
    const long long tiles_max = (Lmax + kTile - 1) / kTile;
    double* lb_base = ppow_dev + (size_t)kMaxTiledFilters * kPpowEntries * 4;
    IirLookback LB;
    LB.rec = reinterpret_cast<double2*>(lb_base);
    LB.sweep = 0;
    const size_t lb_set = (size_t)n_sig * tiles_max;          // double2 per sweep
    bool lb_cleared = false;
    std::vector<double> ppow_host;

    if (remove_dc) {
        double* mean_part = mean + (((size_t)n_sig + 31) / 32 * 32);
        mean_partial_kernel<<<dim3(kMeanParts, n_sig), 256, 0, st>>>(d_x, n, sig_stride, t_stride, mean_part);
        int rc = check_launch("mean_partial_kernel");
        if (rc) return rc;
        mean_finish_kernel<<<(n_sig + 127) / 128, 128, 0, st>>>(mean_part, n, n_sig, mean);
        rc = check_launch("mean_finish_kernel");
        if (rc) return rc;
        if (n_filt == 0) {
            const long long tot = (long long)n * n_sig;
            sub_mean_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(d_x, n, sig_stride, t_stride, mean, n_sig);
            return check_launch("sub_mean_kernel");
        }
    }
    // host-side preparation of every filter first (coefficients, zi, power tables), ONE upload, then only launches:
    // the device never waits for the host between sweeps
    std::vector<IirCoef> coefs(n_filt), coefs_tile(n_filt);
    std::vector<char> use_tiled(n_filt, 0), fused_ok(n_filt, 0);
    std::vector<double> all_tables;
    for (int k = 0; k < n_filt; ++k) {
        // trailing zero taps do not change the filter but would change padlen: the caller passes ntaps = max(len(a), len(b))
        const double* bk = h_b + (size_t)k * ntaps;
        const double* ak = h_a + (size_t)k * ntaps;
        // the tables depend on the coefficients only: keep the last few designs (a loader applies the same three filters
        // to every file)
        const PreparedFilter* hit = nullptr;
        {
            std::lock_guard<std::mutex> lock(g_prep_mutex);
            for (const PreparedFilter& e : g_prep_cache)
                if (e.ntaps == ntaps && memcmp(e.b.data(), bk, ntaps * sizeof(double)) == 0 && memcmp(e.a.data(), ak, ntaps * sizeof(double)) == 0) {
                    hit = &e;
                    coefs[k] = e.c;
                    coefs_tile[k] = e.ct;
                    ppow_host = e.ppow;
                    break;
                }
        }
        if (!hit) {
            int rc = prepare_filter(bk, ak, ntaps, &coefs[k]);
            if (rc) return rc;
            if (coefs[k].d <= 2) prepare_tile_tables(coefs[k], ppow_host, &coefs_tile[k]);
            else ppow_host.clear();
            PreparedFilter e;
            e.ntaps = ntaps;
            e.b.assign(bk, bk + ntaps);
            e.a.assign(ak, ak + ntaps);
            e.c = coefs[k];
            e.ct = coefs_tile[k];
            e.ppow = ppow_host;
            std::lock_guard<std::mutex> lock(g_prep_mutex);
            if (g_prep_cache.size() >= 32) g_prep_cache.erase(g_prep_cache.begin());
            g_prep_cache.push_back(e);
        }
        if (causal) {       // lfilter: no extension, zero initial state
            coefs[k].e = coefs_tile[k].e = 0;
            for (int q = 0; q < kMaxOrder; ++q) coefs[k].zi[q] = coefs_tile[k].zi[q] = 0.0;
        }
        const long long L = n + 2LL * coefs[k].e;
        use_tiled[k] = tiled_mode && coefs[k].d <= 2 && t_stride == 1 && k < kMaxTiledFilters && L >= kTile;
        {       // the look-back sums Q^k e(t-1-k) until |Q^k| < 1e-30: worth it only when that is a handful of records
            double q8 = 0.0;
            if (coefs[k].d <= 2)
                for (int i = 0; i < coefs[k].d * coefs[k].d; ++i) q8 = std::max(q8, std::fabs(coefs_tile[k].pw[3][i]));
            fused_ok[k] = q8 < 1e-3;
        }
        if (use_tiled[k]) {
            all_tables.resize((size_t)(k + 1) * kPpowEntries * 4, 0.0);
            memcpy(all_tables.data() + (size_t)k * kPpowEntries * 4, ppow_host.data(), ppow_host.size() * sizeof(double));
        }
    }
    if (!all_tables.empty() &&
        cudaMemcpyAsync(ppow_dev, all_tables.data(), all_tables.size() * sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess)
        return set_error(HS_ERR_CUDA, "filtfilt: table upload failed");
    for (int k = 0; k < n_filt; ++k) {
        const IirCoef& c = coefs[k];
        int rc = 0;
        const long long L = n + 2LL * c.e;
        IirPass P;
        P.n_sig = n_sig;
        P.L = L;
        P.e = c.e;
        P.n_chunks = (L + kChunk - 1) / kChunk;
        P.states = states;
        // forward: x (virtual odd extension, DC removed on the first filter) -> f
        P.in = d_x;
        P.in_sig_stride = sig_stride;
        P.in_t_stride = t_stride;
        P.n_in = n;
        P.out = causal ? causal_out : f;
        P.out_sig_stride = causal ? causal_stride : L;
        P.out_t_stride = 1;
        P.forward = 1;
        P.mean = (remove_dc && k == 0) ? mean : nullptr;
        const bool tiled = use_tiled[k];
        const IirCoef& ct = coefs_tile[k];
        double* d_ppow = ppow_dev + (size_t)k * kPpowEntries * 4;
        auto sweep = [&](const IirPass& Pp) -> int {
            if (!tiled) return run_sweep_d(c.d, Pp, c, st);
            if (!fused_mode || !fused_ok[k]) return c.d == 1 ? run_sweep_tiled<1>(Pp, c, ct, d_ppow, st) : run_sweep_tiled<2>(Pp, c, ct, d_ppow, st);
            if (!lb_cleared) {      // every sweep of the call has its own records: all of them preset to "empty" (all-ones) once
                int n_tiled = 0;
                for (int q = 0; q < n_filt; ++q) n_tiled += use_tiled[q] ? 1 : 0;
                const size_t bytes = (size_t)(causal ? 1 : 2) * n_tiled * lb_set * sizeof(double2);
                if (cudaMemsetAsync(LB.rec, 0xff, bytes, st) != cudaSuccess) return set_error(HS_ERR_CUDA, "filtfilt: memset failed");
                lb_cleared = true;
            }
            IirLookback S = LB;
            S.rec = LB.rec + (size_t)LB.sweep * lb_set;
            const int r = c.d == 1 ? run_sweep_fused<1>(Pp, c, ct, d_ppow, S, st) : run_sweep_fused<2>(Pp, c, ct, d_ppow, S, st);
            ++LB.sweep;
            return r;
        };
        rc = sweep(P);
        if (rc) return rc;
        if (causal) continue;
        // backward: f reversed -> x (middle n samples), in place
        P.in = f;
        P.in_sig_stride = L;
        P.in_t_stride = 1;
        P.n_in = L;
        P.out = d_x;
        P.out_sig_stride = sig_stride;
        P.out_t_stride = t_stride;
        P.forward = 0;
        P.mean = nullptr;
        rc = sweep(P);
        if (rc) return rc;
    }
    return HS_OK;
}

int hs_iir_filtfilt_f64(double* d_x, int n_sig, int64_t n, int64_t sig_stride, int64_t t_stride, const double* h_b,
                        const double* h_a, int n_filt, int ntaps, int remove_dc, void* d_ws, void* stream) {
    return iir_run(d_x, n_sig, n, sig_stride, t_stride, h_b, h_a, n_filt, ntaps, remove_dc, d_ws, stream, nullptr, 0);
}

int hs_iir_lfilter_f64(const double* d_x, int n_sig, int64_t n, int64_t sig_stride, const double* h_b, const double* h_a, int ntaps,
                       int remove_dc, double* d_y, int64_t y_stride, void* d_ws, void* stream) {
    if (!d_y || d_y == d_x) return set_error(HS_ERR_INVALID, "hs_iir_lfilter_f64: the output must be a separate buffer");
    return iir_run(const_cast<double*>(d_x), n_sig, n, sig_stride, 1, h_b, h_a, 1, ntaps, remove_dc, d_ws, stream, d_y, y_stride);
}

int hs_fir_filter_f64(const double* d_x, int n_sig, int64_t n, int64_t sig_stride, int q, int64_t off, const double* d_b, int ntaps,
                      double* d_y, int64_t n_out, int64_t y_stride, void* stream) {
    if (!d_x || !d_b || !d_y) return set_error(HS_ERR_INVALID, "hs_fir_filter_f64: null pointer");
    if (q < 1 || ntaps < 1) return set_error(HS_ERR_INVALID, "hs_fir_filter_f64: q >= 1 and ntaps >= 1 are required");
    if (n_sig <= 0 || n <= 0 || n_out <= 0) return HS_OK;
    const size_t smem = ((size_t)q * dec_phase_stride(q, ntaps) + ntaps) * sizeof(double);
    if (smem > 200 * 1024) return set_error(HS_ERR_UNSUPPORTED, "hs_fir_filter_f64: q=%d, ntaps=%d need %zu B shared memory", q, ntaps, smem);
    if (n_sig >= 5 && mma_dec_ksteps(q, ntaps) <= 56) {      // Toeplitz products on the FP64 tensor pipe
        const size_t smem_m = (size_t)2 * 8 * mma_dec_sig_stride(q, ntaps) * sizeof(double) + (size_t)4 * 2 * 32 * sizeof(double2);
        dim3 grid_m((unsigned)((n_out + (long long)kMmaOut * kMmaTilesPerCta - 1) / ((long long)kMmaOut * kMmaTilesPerCta)), (unsigned)((n_sig + 7) / 8));
        auto launch_m = [&](auto kern) -> int {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_m);
            if (e != cudaSuccess) return set_error(HS_ERR_CUDA, "hs_fir_filter_f64: %s", cudaGetErrorString(e));
            kern<<<grid_m, 256, smem_m, (cudaStream_t)stream>>>(d_x, n, sig_stride, n_sig, q, off, d_b, ntaps, d_y, n_out, y_stride);
            return check_launch("fir_mma_kernel");
        };
        const int kh = mma_dec_kh(q, ntaps);
        if (kh == 7) return launch_m(fir_mma_kernel<7>);
        if (kh == 14) return launch_m(fir_mma_kernel<14>);
        return launch_m(fir_mma_kernel<28>);
    }
    dim3 grid((unsigned)((n_out + kDecOut - 1) / kDecOut), n_sig);
    auto launch = [&](auto kern) -> int {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(HS_ERR_CUDA, "hs_fir_filter_f64: %s", cudaGetErrorString(e));
        kern<<<grid, kDecThreads, smem, (cudaStream_t)stream>>>(d_x, n, sig_stride, q, off, d_b, ntaps, d_y, n_out, y_stride);
        return check_launch("fir_decimate_kernel");
    };
    switch (q) {
        case 1: return launch(fir_decimate_kernel<1>);
        case 2: return launch(fir_decimate_kernel<2>);
        case 4: return launch(fir_decimate_kernel<4>);
        case 8: return launch(fir_decimate_kernel<8>);
        default: return launch(fir_decimate_kernel<0>);
    }
}

int hs_fir_decimate_f64(const double* d_x, int n_sig, int64_t n, int64_t sig_stride, int q, const double* d_b, int ntaps,
                        double* d_y, int64_t y_stride, void* stream) {
    if (ntaps < 1 || (ntaps & 1) == 0) return set_error(HS_ERR_INVALID, "hs_fir_decimate_f64: q >= 1 and an odd tap count are required");
    if (q < 1) return set_error(HS_ERR_INVALID, "hs_fir_decimate_f64: q >= 1 and an odd tap count are required");
    return hs_fir_filter_f64(d_x, n_sig, n, sig_stride, q, (ntaps - 1) / 2, d_b, ntaps, d_y, (n + q - 1) / q, y_stride, stream);
}

}
