// C-ABI entry points (include/hs_b200.h) for the MVAR/DTF path + error plumbing.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "hs_internal.h"
#include "mvar_launch.h"

namespace hs {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(HS_ERR_CUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
    return HS_OK;
}

#ifdef HS_EXPERIMENT
int exp_env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return (e && *e) ? atoi(e) : dflt;
}
#endif

int device_sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Optional CUDA-event bracket around the dominant kernel (the optimistic K5 pass), for bench.py's roofline:
// events are recorded on the stream the kernel is launched on, only while hs_timing_enable(1) is in effect.
// The hook is a bench facility: one timed stream at a time (the mutex only keeps concurrent callers from corrupting it).
static std::atomic<int> g_timing{0};
static std::mutex g_timing_mu;
static cudaEvent_t g_ev[2] = {nullptr, nullptr};
static bool g_ev_valid = false;

static std::atomic<int> g_k5_kernel{0};      // hs_transfer_set_kernel: 0 automatic, 1 transfer_mma_kernel (6 groups), 2 transfer_ws_kernel

static int k5_groups() {
    static int ng = 0;
    if (ng == 0) {
        ng = exp_env_int("HS_K5_GROUPS", 6);
        if (ng < 4 || ng > 8) ng = 6;
    }
    return ng;
}

// doubles reserved for the per-piece row sums: the segment layout of the DFMA kernels (n_win * n_seg * m) or the balanced
// partition of the tensor-pipe kernel (n_win * slots <= CTAs + 2 * n_win, CTAs <= 1024), whichever is larger
static size_t k5_rowpart_bytes(int n_win, int n_seg, int m) {
    size_t a = (size_t)n_win * n_seg, b = (size_t)1024 + 2 * (size_t)n_win;
    return ((a > b ? a : b) * m * sizeof(double) + 255) / 256 * 256;
}

static void k5_segments(int F, int ng, int* n_seg, int* seg_len) {
    int target = exp_env_int("HS_K5_SEG", 8 * ng);
    if (target < 1) target = 8 * ng;
    int ns = (F + target / 2) / target;
    if (ns < 1) ns = 1;
    int sl = (F + ns - 1) / ns;
    sl = (sl + ng - 1) / ng * ng;
    ns = (F + sl - 1) / sl;
    *n_seg = ns;
    *seg_len = sl;
}

// DFMA throughput probe: 16 independent FMA chains per thread, 8 warps x 4 CTAs per SM.
__global__ void dfma_probe_kernel(double* out, double a, double b, int iters) {
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = a + i + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fma(x[i], b, a);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace hs

using namespace hs;

extern "C" {

const char* hs_last_error(void) { return g_err; }
int hs_version(void) { return 100; }
long long hs_launch_count(void) { return g_launches.load(); }

int hs_transfer_set_kernel(int which) {
    if (which < 0 || which > 4)
        return set_error(HS_ERR_INVALID, "hs_transfer_set_kernel: 0 (automatic), 1 (transfer_mma_kernel), 2 / 3 (transfer_ws_kernel with 4 / 6 groups), 4 (pipe-turn lock)");
#ifndef HS_EXPERIMENT
    if (which >= 2) return set_error(HS_ERR_UNSUPPORTED, "hs_transfer_set_kernel: the warp-specialised kernel and the pipe-turn lock are compiled into HS_EXPERIMENT builds only");
#endif
    g_k5_kernel.store(which);
    return HS_OK;
}

void hs_timing_enable(int on) { g_timing.store(on ? 1 : 0); }

int hs_timing_last_k5_ms(double* ms) {
    if (!ms) return set_error(HS_ERR_INVALID, "hs_timing_last_k5_ms: null pointer");
    std::lock_guard<std::mutex> lock(g_timing_mu);
    if (!g_ev_valid) return set_error(HS_ERR_INVALID, "hs_timing_last_k5_ms: no timed launch yet (call hs_timing_enable(1) first)");
    if (cudaEventSynchronize(g_ev[1]) != cudaSuccess) return set_error(HS_ERR_CUDA, "hs_timing_last_k5_ms: event synchronize failed");
    float f = 0.f;
    if (cudaEventElapsedTime(&f, g_ev[0], g_ev[1]) != cudaSuccess) return set_error(HS_ERR_CUDA, "hs_timing_last_k5_ms: elapsed time failed");
    *ms = f;
    return HS_OK;
}

int hs_measure_dfma_tflops(double* tflops, double* d_scratch, int reps) {
    if (!tflops || !d_scratch) return set_error(HS_ERR_INVALID, "hs_measure_dfma_tflops: null pointer");
    const int sms = device_sm_count();
    const int grid = sms * 4, block = 256, iters = 4096;
    cudaEvent_t e0, e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return set_error(HS_ERR_CUDA, "event create");
    dfma_probe_kernel<<<grid, block>>>(d_scratch, 1.0000001, 0.9999999, iters);
    int rc = check_launch("dfma_probe_kernel");
    if (rc) return rc;
    cudaDeviceSynchronize();
    float best = 1e30f;
    if (reps < 1) reps = 1;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        dfma_probe_kernel<<<grid, block>>>(d_scratch, 1.0000001, 0.9999999, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = (double)grid * block * 16.0 * iters * 2.0 / (best * 1e-3) * 1e-12;
    return HS_OK;
}

int hs_lagcov_f64(const double* d_x, const int64_t* d_offsets, int64_t ch_stride, int n_win, int trials, int m, int n,
                  int p, double* d_R, void* stream) {
    if (!d_x || !d_offsets || !d_R) return set_error(HS_ERR_INVALID, "hs_lagcov_f64: null pointer");
    if (n_win < 0 || trials < 1 || m < 1 || n < 1 || p < 0) return set_error(HS_ERR_INVALID, "hs_lagcov_f64: bad sizes");
    if (p >= n) return set_error(HS_ERR_INVALID, "hs_lagcov_f64: model order %d must be < window length %d", p, n);
    if (n_win == 0) return HS_OK;
    K3Params P;
    P.x = d_x;
    P.offsets = reinterpret_cast<const long long*>(d_offsets);
    P.ch_stride = ch_stride;
    P.R = d_R;
    P.n_win = n_win;
    P.trials = trials;
    P.m = m;
    P.n = n;
    P.p = p;
    return launch_lagcov(P, (cudaStream_t)stream);
}

int hs_yw_assemble_f64(const double* d_R, int n_win, int m, int p, double* d_G, double* d_rhs, void* stream) {
    if (!d_R || !d_G || !d_rhs) return set_error(HS_ERR_INVALID, "hs_yw_assemble_f64: null pointer");
    if (n_win <= 0) return HS_OK;
    return launch_toeplitz(d_R, n_win, m, p, d_G, d_rhs, (cudaStream_t)stream);
}

static bool yw_use_batched(int m) {
    static int mode = -1;
    if (mode < 0) mode = exp_env_int("HS_YW_BATCHED", 0);
    return m > kPadMaxHost || mode == 1;
}

size_t hs_yw_ws_bytes(int n_win, int m, int p) {
    if (yw_use_batched(m)) return align_up(lwr_generic_ws_doubles(n_win, m, p) * sizeof(double));
    return align_up(lwr_ws_doubles(lwr_grid(n_win), m, p) * sizeof(double));
}

int hs_yw_solve_f64(const double* d_R, int n_win, int m, int p, double* d_A, double* d_V, double* d_Vall,
                    int32_t* d_status, void* d_ws, void* stream) {
    if (!d_R || !d_A || !d_V || !d_status || !d_ws) return set_error(HS_ERR_INVALID, "hs_yw_solve_f64: null pointer");
    if (m < 1 || p < 1) return set_error(HS_ERR_INVALID, "hs_yw_solve_f64: bad sizes m=%d p=%d", m, p);
    if (n_win <= 0) return HS_OK;
    K4Params P;
    P.R = d_R;
    P.A = d_A;
    P.V = d_V;
    P.Vall = d_Vall;
    P.status = d_status;
    P.ws = reinterpret_cast<double*>(d_ws);
    P.n_win = n_win;
    P.m = m;
    P.p = p;
    if (yw_use_batched(m)) return launch_lwr_generic(P, (cudaStream_t)stream);
    return launch_lwr(P, lwr_grid(n_win), (cudaStream_t)stream);
}

int hs_ztable_f64(const double* d_freqs, int F, int p, double fs, void* d_z, void* stream) {
    if (!d_freqs || !d_z) return set_error(HS_ERR_INVALID, "hs_ztable_f64: null pointer");
    if (F <= 0 || p <= 0) return HS_OK;
    return launch_ztable(d_freqs, F, p, fs, d_z, (cudaStream_t)stream);
}

size_t hs_transfer_ws_bytes(int n_win, int m, int p, int F) {
    int ns, sl;
    k5_segments(F, k5_groups(), &ns, &sl);
    if (m > kPadMaxHost)      // generic path: one partial row sum per bin + per-CTA scratch matrices
        return align_up((size_t)p * F * 16) + align_up((size_t)n_win * F * m * sizeof(double)) + align_up(transfer_generic_scratch_bytes(m)) + 256;
    // z table, row sums of the optimistic pass, per-matrix flags, list of flagged matrices, counter, |H|^2 staging (n_win, m, F, m)
    return align_up((size_t)p * F * 16) + k5_rowpart_bytes(n_win, ns, m) + 2 * align_up((size_t)n_win * F * sizeof(int)) + 512 +
           align_up((size_t)n_win * F * m * m * sizeof(double));
}

size_t hs_transfer_ws_flag_offset(int n_win, int m, int p, int F) {
    if (m > kPadMaxHost) return (size_t)-1;
    int ns, sl;
    k5_segments(F, k5_groups(), &ns, &sl);
    return align_up((size_t)p * F * 16) + k5_rowpart_bytes(n_win, ns, m) + 2 * align_up((size_t)n_win * F * sizeof(int));
}

int hs_transfer_dtf_f64(const double* d_A, const double* d_freqs, int F, double fs, int n_win, int m, int p, void* d_H,
                        void* d_Af, double* d_dtf, double* d_ffdtf, int32_t* d_status, void* d_ws, void* stream) {
    if (!d_A || !d_freqs || !d_status || !d_ws) return set_error(HS_ERR_INVALID, "hs_transfer_dtf_f64: null pointer");
    if (m < 1 || p < 1 || F < 1) return set_error(HS_ERR_INVALID, "hs_transfer_dtf_f64: bad sizes");
    if (n_win <= 0) return HS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int ng = k5_groups();
    int ns, sl;
    k5_segments(F, ng, &ns, &sl);
    unsigned char* ws = reinterpret_cast<unsigned char*>(d_ws);
    double2* z = reinterpret_cast<double2*>(ws);
    double* rowpart = reinterpret_cast<double*>(ws + align_up((size_t)p * F * 16));
    double* rowpart2 = nullptr;
    int* bad = nullptr;
    int* bad_count = nullptr;
    double* stage = nullptr;
    static int k5_mode = -1;       // 1 (default): optimistic elimination + check + pivoted redo; 0: pivoted only
    if (k5_mode < 0) k5_mode = exp_env_int("HS_K5_PIVOT_ONLY", 0) ? 0 : 1;
    int* bad_list = nullptr;
    if (m <= kPadMaxHost) {
        const size_t rp = k5_rowpart_bytes(n_win, ns, m);
        const size_t fl = align_up((size_t)n_win * F * sizeof(int));
        unsigned char* q = ws + align_up((size_t)p * F * 16) + rp;
        bad = reinterpret_cast<int*>(q);
        bad_list = reinterpret_cast<int*>(q + fl);
        bad_count = reinterpret_cast<int*>(q + 2 * fl);
        stage = reinterpret_cast<double*>(q + 2 * fl + 512);
    }
    int rc = launch_ztable(d_freqs, F, p, fs, z, st);
    if (rc) return rc;
    K5Params P;
    P.A = d_A;
    P.z = z;
    const bool want_dtf = d_dtf || d_ffdtf;
    const bool staged = want_dtf && stage;       // m <= 40: |H|^2 goes through the (w, i, f, j) staging buffer
    P.dtf = staged ? stage : (d_dtf ? d_dtf : d_ffdtf);
    P.dtf_fij = staged ? 1 : 0;
    P.rowpart = d_ffdtf ? rowpart : nullptr;
    P.H = reinterpret_cast<double2*>(d_H);
    P.Af = reinterpret_cast<double2*>(d_Af);
    P.status = d_status;
    P.n_win = n_win;
    P.m = m;
    P.p = p;
    P.F = F;
    P.n_seg = ns;
    P.seg_len = sl;
    P.per_cta = 0;
    P.flip = exp_env_int("HS_K5_FLIP", 0);
    P.pipe_turns = (g_k5_kernel.load() == 4) ? 1 : 0;
    P.rowpart2 = nullptr;
    P.bad = bad;
    P.bad_count = bad_count;
    P.bad_list = bad_list;
    P.verify_tol2 = 1e-18;       // relative residual 1e-9 on the probe vector
    bool redo = false;
    if (m > kPadMaxHost) {
        P.n_seg = ns = F;
        P.seg_len = 1;
        void* scratch = ws + align_up((size_t)p * F * 16) + align_up((size_t)n_win * F * m * sizeof(double));
        rc = launch_transfer_generic(P, scratch, st);
    } else if (k5_mode == 0) {
        rc = launch_transfer_dtf(P, ng, 0, st);
    } else {
        redo = true;
        // flags and the counter are cleared; the list needs no clearing (only entries below the counter are read)
        if (cudaMemsetAsync(bad, 0, (size_t)n_win * F * sizeof(int), st) != cudaSuccess || cudaMemsetAsync(bad_count, 0, sizeof(int), st) != cudaSuccess)
            return set_error(HS_ERR_CUDA, "hs_transfer_dtf_f64: memset failed");
        static int use_mma = -1;   // default: blocked elimination on the FP64 tensor pipe; HS_K5_MMA=0 -> register-tile DFMA kernel
        if (use_mma < 0) use_mma = exp_env_int("HS_K5_MMA", 1) ? 1 : 0;
        const bool timed = g_timing.load() != 0;
        if (timed) {
            std::lock_guard<std::mutex> lock(g_timing_mu);
            if (!g_ev[0] && (cudaEventCreate(&g_ev[0]) != cudaSuccess || cudaEventCreate(&g_ev[1]) != cudaSuccess))
                return set_error(HS_ERR_CUDA, "hs_transfer_dtf_f64: cannot create timing events");
            cudaEventRecord(g_ev[0], st);
        }
        const int which = g_k5_kernel.load();          // 2: warp-specialised, 4 groups; 3: warp-specialised, 6 groups
        const int ws_ng = which == 3 ? 6 : 4;
        const bool ws = (which == 2 || which == 3) && transfer_mma_fits(p, ws_ng, sl);
        if (use_mma && (ws || transfer_mma_fits(p, ng, sl))) {
            // balanced partition: a window may be covered by fewer CTAs than it has row-sum slots -> the slots start at zero
            transfer_mma_partition(n_win, F, &P.per_cta, &P.n_seg, ws ? ws_ng : ng);
            ns = P.n_seg;
            if (P.rowpart && cudaMemsetAsync(P.rowpart, 0, (size_t)n_win * ns * m * sizeof(double), st) != cudaSuccess)
                return set_error(HS_ERR_CUDA, "hs_transfer_dtf_f64: memset failed");
            rc = ws ? launch_transfer_ws(P, ws_ng, st) : launch_transfer_mma(P, ng, st);
        } else {
            rc = launch_transfer_dtf(P, ng, 1, st);
        }
        if (timed) {
            std::lock_guard<std::mutex> lock(g_timing_mu);
            cudaEventRecord(g_ev[1], st);
            g_ev_valid = true;
        }
        if (rc) return rc;
        rc = launch_transfer_dtf(P, ng, 2, st);      // returns immediately on the device when nothing was flagged
    }
    if (rc) return rc;
    if (staged) return launch_dtf_finalize(stage, d_ffdtf ? rowpart : nullptr, redo ? bad : nullptr, n_win, m, F, ns, d_dtf, d_ffdtf, st);
    if (d_ffdtf) rc = launch_ffdtf_normalize(P.dtf, rowpart, redo ? rowpart2 : nullptr, n_win, m, F, ns, d_ffdtf, st);
    return rc;
}

size_t hs_mvar_ffdtf_ws_bytes(int n_win, int m, int p, int F) {
    size_t b = 0;
    b += align_up((size_t)n_win * (p + 1) * m * m * sizeof(double));   // R
    b += align_up((size_t)n_win * m * m * p * sizeof(double));         // A
    b += align_up((size_t)n_win * m * m * sizeof(double));             // V
    b += hs_yw_ws_bytes(n_win, m, p);
    b += hs_transfer_ws_bytes(n_win, m, p, F);
    return b + 256;
}

int hs_mvar_ffdtf_f64(const double* d_x, const int64_t* d_offsets, int64_t ch_stride, int n_win, int m, int n, int p,
                      const double* d_freqs, int F, double fs, double* d_ffdtf, double* d_A, double* d_V,
                      int32_t* d_status, void* d_ws, void* stream) {
    if (!d_x || !d_offsets || !d_freqs || !d_ffdtf || !d_status || !d_ws)
        return set_error(HS_ERR_INVALID, "hs_mvar_ffdtf_f64: null pointer");
    if (n_win <= 0) return HS_OK;
    unsigned char* ws = reinterpret_cast<unsigned char*>(d_ws);
    double* R = reinterpret_cast<double*>(ws);
    ws += align_up((size_t)n_win * (p + 1) * m * m * sizeof(double));
    double* A = reinterpret_cast<double*>(ws);
    ws += align_up((size_t)n_win * m * m * p * sizeof(double));
    double* V = reinterpret_cast<double*>(ws);
    ws += align_up((size_t)n_win * m * m * sizeof(double));
    void* yw_ws = ws;
    ws += hs_yw_ws_bytes(n_win, m, p);
    void* tr_ws = ws;
    if (d_A) A = d_A;
    if (d_V) V = d_V;
    int rc = hs_lagcov_f64(d_x, d_offsets, ch_stride, n_win, 1, m, n, p, R, stream);
    if (rc) return rc;
    rc = hs_yw_solve_f64(R, n_win, m, p, A, V, nullptr, d_status, yw_ws, stream);
    if (rc) return rc;
    return hs_transfer_dtf_f64(A, d_freqs, F, fs, n_win, m, p, nullptr, nullptr, nullptr, d_ffdtf, d_status, tr_ws, stream);
}

int hs_spectra_f64(const void* d_H, const double* d_V, int n_win, int m, int F, void* d_S, void* stream) {
    if (!d_H || !d_V || !d_S) return set_error(HS_ERR_INVALID, "hs_spectra_f64: null pointer");
    if (n_win <= 0) return HS_OK;
    return launch_spectra(d_H, d_V, n_win, m, F, d_S, (cudaStream_t)stream);
}

int hs_partial_coherence_f64(const void* d_S, int n_win, int m, int F, void* d_kappa, const double* d_ffdtf, double* d_ddtf,
                             int32_t* d_status, void* stream) {
    if (n_win <= 0) return HS_OK;               // empty batch: nothing to do (and no buffers to point at)
    if (!d_S || !d_status || (!d_kappa && !d_ddtf)) return set_error(HS_ERR_INVALID, "hs_partial_coherence_f64: null pointer");
    if (d_ddtf && !d_ffdtf) return set_error(HS_ERR_INVALID, "hs_partial_coherence_f64: dDTF needs the ffDTF input");
    if (m < 1 || F < 1) return set_error(HS_ERR_INVALID, "hs_partial_coherence_f64: bad shape m=%d F=%d", m, F);
    return launch_pcoh(d_S, n_win, m, F, d_kappa, d_ffdtf, d_ddtf, d_status, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------
// Host-buffer plan: chunked H2D -> K3/K4/K5 -> D2H pipeline on three streams.
// ------------------------------------------------------------------------------------------
struct hs_plan {
    int max_windows, m, n, p, F;
    int64_t max_samples;
    int chunk;                 // windows per full chunk
    int ramp;                  // 1: short first chunks (16, 32, 64 windows) so that the device-to-host stream starts early
    double* d_x = nullptr;
    int64_t* d_offsets = nullptr;
    double* d_freqs = nullptr;
    double* d_out[2] = {nullptr, nullptr};
    void* d_ws[2] = {nullptr, nullptr};
    int32_t* d_status = nullptr;
    int64_t* h_offsets = nullptr;     // pinned
    int32_t* h_status = nullptr;      // pinned
    double* h_result = nullptr;       // pinned (max_windows, m, m, F), allocated on first use: the destination of every device-to-host
                                      // copy unless the caller's own buffer is page-locked
    cudaStream_t s_compute[2] = {nullptr, nullptr};
    cudaStream_t s_copy = nullptr;
    cudaEvent_t ev_in = nullptr, ev_in0 = nullptr, ev_done[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> ev_chunk;     // one per chunk: its device-to-host copy has finished
    size_t ws_bytes = 0;
    std::mutex mu;                    // a plan runs one call at a time; different plans are independent
};

#define PLAN_CUDA(call)                                                                          \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            rc = set_error(e__ == cudaErrorMemoryAllocation ? HS_ERR_NOMEM : HS_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
            goto fail;                                                                           \
        }                                                                                        \
    } while (0)

static int plan_chunk_at(const hs_plan* pl, int idx) {
    // 16, 32, 64, then full chunks: chunk k's compute (~0.5 ms + 15 us/window) fits inside chunk k-1's copy (52 us/window)
    const int c = pl->ramp ? (16 << (idx < 3 ? idx : 3)) : pl->chunk;
    return c < pl->chunk ? c : pl->chunk;
}

static bool host_pointer_is_pinned(const void* ptr) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

int hs_plan_create(hs_plan** plan, int max_windows, int m, int n, int p, int F, int64_t max_samples) {
    if (!plan || max_windows < 1 || m < 1 || n < 1 || p < 1 || F < 1 || max_samples < n)
        return set_error(HS_ERR_INVALID, "hs_plan_create: bad arguments");
    int rc = HS_OK;
    hs_plan* pl = new hs_plan();
    pl->max_windows = max_windows;
    pl->m = m;
    pl->n = n;
    pl->p = p;
    pl->F = F;
    pl->max_samples = max_samples;
    int chunk = exp_env_int("HS_PLAN_CHUNK", 96);
    if (chunk < 1) chunk = 96;
    if (chunk > max_windows) chunk = max_windows;
    pl->chunk = chunk;
    pl->ramp = exp_env_int("HS_PLAN_RAMP", 1) ? 1 : 0;
    *plan = pl;
    PLAN_CUDA(cudaMalloc(&pl->d_x, (size_t)m * max_samples * sizeof(double)));
    PLAN_CUDA(cudaMalloc(&pl->d_offsets, (size_t)max_windows * sizeof(int64_t)));
    PLAN_CUDA(cudaMalloc(&pl->d_freqs, (size_t)F * sizeof(double)));
    PLAN_CUDA(cudaMalloc(&pl->d_status, (size_t)max_windows * sizeof(int32_t)));
    PLAN_CUDA(cudaMallocHost(&pl->h_offsets, (size_t)max_windows * sizeof(int64_t)));
    PLAN_CUDA(cudaMallocHost(&pl->h_status, (size_t)max_windows * sizeof(int32_t)));
    pl->ws_bytes = hs_mvar_ffdtf_ws_bytes(chunk, m, p, F);
    for (int i = 0; i < 2; ++i) {
        PLAN_CUDA(cudaMalloc(&pl->d_out[i], (size_t)chunk * m * m * F * sizeof(double)));
        PLAN_CUDA(cudaMalloc(&pl->d_ws[i], pl->ws_bytes));
        PLAN_CUDA(cudaStreamCreateWithFlags(&pl->s_compute[i], cudaStreamNonBlocking));
        PLAN_CUDA(cudaEventCreateWithFlags(&pl->ev_done[i], cudaEventDisableTiming));
        PLAN_CUDA(cudaEventCreateWithFlags(&pl->ev_free[i], cudaEventDisableTiming));
    }
    PLAN_CUDA(cudaStreamCreateWithFlags(&pl->s_copy, cudaStreamNonBlocking));
    PLAN_CUDA(cudaEventCreateWithFlags(&pl->ev_in, cudaEventDisableTiming));
    PLAN_CUDA(cudaEventCreateWithFlags(&pl->ev_in0, cudaEventDisableTiming));
    {
        int n_chunks = 0;
        for (int w0 = 0, idx = 0; w0 < max_windows; ++idx, ++n_chunks) w0 += plan_chunk_at(pl, idx);
        pl->ev_chunk.assign(n_chunks, nullptr);
        for (int i = 0; i < n_chunks; ++i) PLAN_CUDA(cudaEventCreateWithFlags(&pl->ev_chunk[i], cudaEventDisableTiming));
    }
    return HS_OK;
fail:
    return rc;       // the caller destroys the partially built plan (hs_plan_destroy accepts it)
}

void hs_plan_destroy(hs_plan* pl) {
    if (!pl) return;
    cudaDeviceSynchronize();
    cudaFree(pl->d_x);
    cudaFree(pl->d_offsets);
    cudaFree(pl->d_freqs);
    cudaFree(pl->d_status);
    cudaFreeHost(pl->h_offsets);
    cudaFreeHost(pl->h_status);
    if (pl->h_result) cudaFreeHost(pl->h_result);
    for (int i = 0; i < 2; ++i) {
        cudaFree(pl->d_out[i]);
        cudaFree(pl->d_ws[i]);
        if (pl->s_compute[i]) cudaStreamDestroy(pl->s_compute[i]);
        if (pl->ev_done[i]) cudaEventDestroy(pl->ev_done[i]);
        if (pl->ev_free[i]) cudaEventDestroy(pl->ev_free[i]);
    }
    for (cudaEvent_t e : pl->ev_chunk)
        if (e) cudaEventDestroy(e);
    if (pl->s_copy) cudaStreamDestroy(pl->s_copy);
    if (pl->ev_in) cudaEventDestroy(pl->ev_in);
    if (pl->ev_in0) cudaEventDestroy(pl->ev_in0);
    delete pl;
}

int hs_plan_host_result(hs_plan* pl, double** h_result, size_t* bytes) {
    if (!pl || !h_result) return set_error(HS_ERR_INVALID, "hs_plan_host_result: null pointer");
    std::lock_guard<std::mutex> lock(pl->mu);
    const size_t nbytes = (size_t)pl->max_windows * pl->m * pl->m * pl->F * sizeof(double);
    if (!pl->h_result) {
        cudaError_t e = cudaMallocHost(&pl->h_result, nbytes);
        if (e != cudaSuccess) {
            pl->h_result = nullptr;
            return set_error(HS_ERR_NOMEM, "hs_plan_host_result: cannot page-lock %zu bytes: %s", nbytes, cudaGetErrorString(e));
        }
    }
    *h_result = pl->h_result;
    if (bytes) *bytes = nbytes;
    return HS_OK;
}

int hs_plan_mvar_ffdtf_host(hs_plan* pl, const double* h_x, int64_t t_total, const int64_t* h_starts, int n_win,
                            const double* h_freqs, double fs, double* h_ffdtf, int32_t* h_status) {
    if (!pl || !h_x || !h_starts || !h_freqs || !h_status)
        return set_error(HS_ERR_INVALID, "hs_plan_mvar_ffdtf_host: null pointer");
    if (n_win > pl->max_windows || t_total > pl->max_samples || n_win < 0)
        return set_error(HS_ERR_INVALID, "hs_plan_mvar_ffdtf_host: plan too small (n_win=%d, T=%lld)", n_win, (long long)t_total);
    // where the device-to-host copies land: the caller's buffer when it is page-locked (cudaMemcpyAsync into pageable memory
    // is staged by the driver and serialises with the kernels), else the plan's own pinned buffer, from which finished
    // chunks are copied to the caller's buffer by the host while later chunks are still in flight
    double* h_pinned = h_ffdtf;
    if (!h_ffdtf || !host_pointer_is_pinned(h_ffdtf)) {
        int rc0 = hs_plan_host_result(pl, &h_pinned, nullptr);
        if (rc0) return rc0;
    }
    std::lock_guard<std::mutex> lock(pl->mu);
    for (int w = 0; w < n_win; ++w) {
        if (h_starts[w] < 0 || h_starts[w] + pl->n > t_total)
            return set_error(HS_ERR_INVALID, "hs_plan_mvar_ffdtf_host: window %d [%lld, +%d) outside the signal", w, (long long)h_starts[w], pl->n);
        pl->h_offsets[w] = h_starts[w];
    }
    if (n_win == 0) return HS_OK;
    int rc = HS_OK;
    const int m = pl->m, F = pl->F;
    const int c0 = plan_chunk_at(pl, 0);
    // inputs: the samples the first chunk needs go first (columns [0, t_split) of every channel row), the rest follows
    int64_t t_split = 0;
    for (int w = 0; w < n_win && w < c0; ++w)
        if (h_starts[w] + pl->n > t_split) t_split = h_starts[w] + pl->n;
    if (!pl->ramp || t_split > t_total / 2) t_split = t_total;
    const size_t pitch = (size_t)t_total * sizeof(double);
    const size_t per_win = (size_t)m * m * F;
    int slot = 0, idx = 0, n_chunks = 0;
    bool used[2] = {false, false};
#ifdef HS_EXPERIMENT
    const int dbg = exp_env_int("HS_DEBUG", 0);
    cudaEvent_t tev[64];
    int ntev = 0;
#endif
    PLAN_CUDA(cudaMemcpyAsync(pl->d_offsets, pl->h_offsets, (size_t)n_win * sizeof(int64_t), cudaMemcpyHostToDevice, pl->s_copy));
    PLAN_CUDA(cudaMemcpyAsync(pl->d_freqs, h_freqs, (size_t)F * sizeof(double), cudaMemcpyHostToDevice, pl->s_copy));
    PLAN_CUDA(cudaMemsetAsync(pl->d_status, 0, (size_t)n_win * sizeof(int32_t), pl->s_copy));
    PLAN_CUDA(cudaMemcpy2DAsync(pl->d_x, pitch, h_x, pitch, (size_t)t_split * sizeof(double), m, cudaMemcpyHostToDevice, pl->s_copy));
    PLAN_CUDA(cudaEventRecord(pl->ev_in0, pl->s_copy));
    if (t_split < t_total)
        PLAN_CUDA(cudaMemcpy2DAsync(pl->d_x + t_split, pitch, h_x + t_split, pitch, (size_t)(t_total - t_split) * sizeof(double), m,
                                    cudaMemcpyHostToDevice, pl->s_copy));
    PLAN_CUDA(cudaEventRecord(pl->ev_in, pl->s_copy));
#ifdef HS_EXPERIMENT
    if (dbg) { cudaEventCreate(&tev[0]); cudaEventRecord(tev[0], pl->s_copy); ntev = 1; }
#endif
    for (int w0 = 0; w0 < n_win; slot ^= 1, ++idx) {
        const int want = plan_chunk_at(pl, idx);
        const int nw = (n_win - w0 < want) ? (n_win - w0) : want;
        cudaStream_t sc = pl->s_compute[slot];
        PLAN_CUDA(cudaStreamWaitEvent(sc, idx == 0 ? pl->ev_in0 : pl->ev_in, 0));
        if (used[slot]) PLAN_CUDA(cudaStreamWaitEvent(sc, pl->ev_free[slot], 0));    // previous D2H of this slot finished
        rc = hs_mvar_ffdtf_f64(pl->d_x, pl->d_offsets + w0, t_total, nw, m, pl->n, pl->p, pl->d_freqs, F, fs, pl->d_out[slot],
                               nullptr, nullptr, pl->d_status + w0, pl->d_ws[slot], sc);
        if (rc) goto fail;
        PLAN_CUDA(cudaEventRecord(pl->ev_done[slot], sc));
        PLAN_CUDA(cudaStreamWaitEvent(pl->s_copy, pl->ev_done[slot], 0));
#ifdef HS_EXPERIMENT
        if (dbg && ntev < 62) { cudaEventCreate(&tev[ntev]); cudaEventRecord(tev[ntev++], pl->s_copy); }
#endif
        PLAN_CUDA(cudaMemcpyAsync(h_pinned + (size_t)w0 * per_win, pl->d_out[slot], (size_t)nw * per_win * sizeof(double),
                                  cudaMemcpyDeviceToHost, pl->s_copy));
        PLAN_CUDA(cudaEventRecord(pl->ev_free[slot], pl->s_copy));
        PLAN_CUDA(cudaEventRecord(pl->ev_chunk[idx], pl->s_copy));
#ifdef HS_EXPERIMENT
        if (dbg && ntev < 62) { cudaEventCreate(&tev[ntev]); cudaEventRecord(tev[ntev++], pl->s_copy); }
#endif
        used[slot] = true;
        w0 += nw;
        n_chunks = idx + 1;
    }
    PLAN_CUDA(cudaMemcpyAsync(pl->h_status, pl->d_status, (size_t)n_win * sizeof(int32_t), cudaMemcpyDeviceToHost, pl->s_copy));
    if (h_ffdtf && h_pinned != h_ffdtf) {
        // pageable destination: hand every chunk over as soon as its copy has landed in the pinned buffer
        for (int i = 0, w0 = 0; i < n_chunks; ++i) {
            const int want = plan_chunk_at(pl, i);
            const int nw = (n_win - w0 < want) ? (n_win - w0) : want;
            PLAN_CUDA(cudaEventSynchronize(pl->ev_chunk[i]));
            memcpy(h_ffdtf + (size_t)w0 * per_win, h_pinned + (size_t)w0 * per_win, (size_t)nw * per_win * sizeof(double));
            w0 += nw;
        }
    }
    PLAN_CUDA(cudaStreamSynchronize(pl->s_copy));
    PLAN_CUDA(cudaStreamSynchronize(pl->s_compute[0]));
    PLAN_CUDA(cudaStreamSynchronize(pl->s_compute[1]));
    memcpy(h_status, pl->h_status, (size_t)n_win * sizeof(int32_t));
#ifdef HS_EXPERIMENT
    if (dbg) {
        fprintf(stderr, "[hs] plan timeline (ms since first copy): ");
        for (int i = 1; i < ntev; ++i) { float ms = 0; cudaEventElapsedTime(&ms, tev[0], tev[i]); fprintf(stderr, "%s%.2f", (i & 1) ? " [" : "-", ms); if (!(i & 1)) fprintf(stderr, "]"); }
        fprintf(stderr, "\n");
        for (int i = 0; i < ntev; ++i) cudaEventDestroy(tev[i]);
    }
#endif
    return HS_OK;
fail:
    // nothing of this call may still be in flight when the caller frees (or reuses) its buffers
    cudaStreamSynchronize(pl->s_copy);
    cudaStreamSynchronize(pl->s_compute[0]);
    cudaStreamSynchronize(pl->s_compute[1]);
    return rc;
}

}  // extern "C"
