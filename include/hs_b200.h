/* hs_b200.h -- C ABI of the B200-native connectivity hot path.
 *
 * The reference (SYNCC-IN/hyperscanning-signal-analysis) is pure Python and has no
 * FFI; its "operator interface" for this path is a set of module-level Python
 * functions.  Each entry point below names the reference function (file:line under
 * /root/reference) whose arithmetic it replaces; the Python modules in
 * hyperscanning_signal_analysis_b200/ keep the reference signatures and call these
 * through ctypes (see INTEGRATION.md for the binding a maintainer would add).
 *
 * Conventions
 *   - plain pointers and sizes only; every array is float64 / complex128 / int32 /
 *     int64, C-contiguous unless a stride argument says otherwise;
 *   - "d_" pointers are DEVICE pointers, "h_" pointers are HOST pointers;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - device entry points never allocate and never synchronise: scratch is
 *     caller-provided (sizes from the *_ws_bytes queries);
 *   - return 0 on success, a negative HS_ERR_* code otherwise; hs_last_error()
 *     returns a thread-local message.  Numerically singular windows do not fail the
 *     batch: they set bits in the per-window int32 `status` array
 *     (1 = A(f) singular at some bin, 2 = singular residual covariance in the
 *     Yule-Walker recursion), which the Python layer turns into
 *     numpy.linalg.LinAlgError like np.linalg.solve / inv would raise.
 */
#ifndef HS_B200_H
#define HS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HS_OK 0
#define HS_ERR_INVALID (-1)
#define HS_ERR_UNSUPPORTED (-2)
#define HS_ERR_CUDA (-3)
#define HS_ERR_NOMEM (-4)

#define HS_STATUS_SINGULAR_TRANSFER 1
#define HS_STATUS_SINGULAR_YW 2

const char* hs_last_error(void);
int hs_version(void);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
long long hs_launch_count(void);

/* Measurement hook for bench.py's roofline: while enabled, hs_transfer_dtf_f64 brackets its dominant kernel (the
 * optimistic A(f)^-1 pass, transfer_mma_kernel / transfer_dtf_kernel) with CUDA events on the caller's stream;
 * hs_timing_last_k5_ms synchronises on the closing event and returns that launch's duration.               */
void hs_timing_enable(int on);
int hs_timing_last_k5_ms(double* ms);

/* Roofline probe: achieved FP64 FMA throughput (TFLOP/s, best of `reps` runs of a register-only
 * DFMA loop on every SM) -- the denominator bench.py reports K3-K5 against, because the
 * driver-written MEASURED_PEAKS.json holds no FP64 figure.  d_scratch: >= 8 MiB device memory.
 * Synchronises; not part of the data path.                                                  */
int hs_measure_dfma_tflops(double* tflops, double* d_scratch, int reps);

/* ---------------------------------------------------------------- MVAR / DTF (K3-K5) */

/* Lag covariances R(0..p), biased 1/n, no mean removal, averaged over trials.
 * Replaces count_corr, src/mtmvar.py:35-87 (lags :54-59, lag 0 :72-73, trial mean :78-85).
 *   d_x          base pointer of the signals
 *   d_offsets    (n_win*trials) int64 element offsets of (channel 0, sample 0) of each window/trial
 *   ch_stride    elements between consecutive channels (time stride is 1)
 *   d_R          out (n_win, p+1, m, m)                                                     */
int hs_lagcov_f64(const double* d_x, const int64_t* d_offsets, int64_t ch_stride, int n_win, int trials, int m,
                  int n, int p, double* d_R, void* stream);

/* Block-Toeplitz system as count_corr returns it (src/mtmvar.py:65-76):
 * d_G (n_win, m*p, m*p), d_rhs (n_win, m*p, m); R(0) is d_R[:, 0].                          */
int hs_yw_assemble_f64(const double* d_R, int n_win, int m, int p, double* d_G, double* d_rhs, void* stream);

/* Yule-Walker solve by the Levinson-Wiggins-Robinson recursion.
 * Replaces np.linalg.solve + residual + reshape of ar_coeff, src/mtmvar.py:116-122.
 *   d_A (n_win, m, m, p)   d_V (n_win, m, m)   d_Vall optional (n_win, p, m, m): residual
 *   covariance after every order (what mvar_criterion, mtmvar.py:577-590, refits p times for)
 *   d_ws scratch of hs_yw_ws_bytes(n_win, m, p) bytes.                                       */
size_t hs_yw_ws_bytes(int n_win, int m, int p);
int hs_yw_solve_f64(const double* d_R, int n_win, int m, int p, double* d_A, double* d_V, double* d_Vall,
                    int32_t* d_status, void* d_ws, void* stream);

/* Model-order criteria for every window at once.  Replaces the refit loop of mvar_criterion, src/mtmvar.py:577-590:
 * crit[w][k] = ln det V_{k+1}(w) + penalty * (k+1), penalty = 2 m^2/n (crit_type 0, 'AIC'), 2 ln ln n m^2/n (1, 'HQ'),
 * ln n m^2/n (2, 'SC'); popt[w] = 1 + first index of the minimum (np.argmin).
 *   d_Vall (n_win, P, m, m): residual covariance of every order, as hs_yw_solve_f64 writes it for model order P
 *   d_crit (n_win, P) out;  d_logdet (n_win, P) out or NULL;  d_popt (n_win) int32 out;  n_samples = window length      */
int hs_mvar_criterion_f64(const double* d_Vall, int n_win, int P, int m, int n_samples, int crit_type, double* d_crit,
                          double* d_logdet, int32_t* d_popt, void* stream);

/* z[k][f] = exp(-(k+1) 2 pi i f / fs), src/mtmvar.py:151-153.  d_z (p, F) complex128.        */
int hs_ztable_f64(const double* d_freqs, int F, int p, double fs, void* d_z, void* stream);

/* A(f) = I - sum_k A_k z_k(f), H(f) = A(f)^-1, dtf = |H|^2 and ffDTF.
 * Replaces mvar_transfer_function (src/mtmvar.py:126-162), dtf_multivariate (:232) and the
 * normalisation loop of full_freq_dtf (:278-284).  All outputs optional (NULL to skip):
 *   d_H, d_Af   (n_win, m, m, F) complex128       d_dtf, d_ffdtf (n_win, m, m, F) float64
 * d_ffdtf may alias d_dtf.  d_ws: hs_transfer_ws_bytes(n_win, m, p, F) bytes.                */
size_t hs_transfer_ws_bytes(int n_win, int m, int p, int F);
/* Diagnostics: byte offset inside d_ws of the int32 counter of matrices whose optimistic (unpivoted)
 * elimination failed the a-posteriori check and were redone with pivoting by the last
 * hs_transfer_dtf_f64 call on that workspace; (size_t)-1 when the shape takes the generic path.  */
size_t hs_transfer_ws_flag_offset(int n_win, int m, int p, int F);
/* Test / measurement hook: which optimistic A(f)^-1 kernel the following calls use: 0 automatic, 1 transfer_mma_kernel
 * (6 groups of Re / Im warps); 2 / 3 transfer_ws_kernel (4 / 6 groups with helper warps inverting the pivot blocks one step
 * ahead) and 4 (transfer_mma_kernel with a per-sub-partition tensor-pipe turn lock) exist in `make HS_EXPERIMENT=1` builds only
 * (both measured slower) and return HS_ERR_UNSUPPORTED otherwise.
 * Process-wide.                                                                                                        */
int hs_transfer_set_kernel(int which);
int hs_transfer_dtf_f64(const double* d_A, const double* d_freqs, int F, double fs, int n_win, int m, int p,
                        void* d_H, void* d_Af, double* d_dtf, double* d_ffdtf, int32_t* d_status, void* d_ws,
                        void* stream);

/* S(f) = H(f) V H(f)^T with a plain transpose, src/mtmvar.py:197-199.  d_S (n_win, m, m, F) complex128. */
int hs_spectra_f64(const void* d_H, const double* d_V, int n_win, int m, int F, void* d_S, void* stream);

/* Generalised partial directed coherence from A(f) and diag(V).  Replaces the double loop of gen_partial_directed_coherence,
 * src/mtmvar.py:449-468.  d_Af (n_win, m, m, F) complex128 (hs_transfer_dtf_f64's d_Af), d_V (n_win, m, m), d_gpdc (n_win, m, m, F).  */
int hs_gpdc_f64(const void* d_Af, const double* d_V, int n_win, int m, int F, double* d_gpdc, void* stream);

/* Partial coherence of a spectral matrix and the direct DTF.
 * Replaces partial_coherence (src/mtmvar.py:287-338: determinant of every minor of S(f)) by one pivoted complex
 * inverse per bin (minor_ij = (-1)^(i+j) det S (S^-1)_ji), and the product of direct_dtf (:379-383).
 *   d_S (n_win, m, m, F) complex128 in;  d_kappa (n_win, m, m, F) complex128 out or NULL;
 *   d_ffdtf in / d_ddtf out (n_win, m, m, F) float64, both or neither:  ddtf = ffdtf * |kappa|.
 * d_status[w] |= 4 when S(f) of window w is singular for some bin.  m > 40: in place in d_kappa, which must then be given.  */
int hs_partial_coherence_f64(const void* d_S, int n_win, int m, int F, void* d_kappa, const double* d_ffdtf,
                             double* d_ddtf, int32_t* d_status, void* stream);

/* Fused windows -> ffDTF (the metric path): K3 -> K4 -> K5 on one stream.
 * Equivalent to calling full_freq_dtf(window, freqs, fs, optimal_model_order=p)
 * (src/mtmvar.py:237) for every window of EEG_IBI_FFDTF_Pipeline.run_pipeline's loop
 * (src/eeg_alpha_ibi_ffdtf.py:741-755) and stacking as at :651.
 *   d_A / d_V optional outputs (may be NULL: taken from the workspace).                     */
size_t hs_mvar_ffdtf_ws_bytes(int n_win, int m, int p, int F);
int hs_mvar_ffdtf_f64(const double* d_x, const int64_t* d_offsets, int64_t ch_stride, int n_win, int m, int n, int p,
                      const double* d_freqs, int F, double fs, double* d_ffdtf, double* d_A, double* d_V,
                      int32_t* d_status, void* d_ws, void* stream);

/* Host-buffer variant (what a Python caller holding NumPy arrays uses; e2e in bench.py):
 * h_x (m, T_total) float64 host, h_starts (n_win) int64 window start samples,
 * h_ffdtf (n_win, m, m, F) host output, h_status (n_win).  Copies, computes and copies back
 * in window chunks on two streams so PCIe transfers overlap the kernels; synchronises before
 * returning.  The plan owns its device buffers and pinned staging.                           */
typedef struct hs_plan hs_plan;
int hs_plan_create(hs_plan** plan, int max_windows, int m, int n, int p, int F, int64_t max_samples);
void hs_plan_destroy(hs_plan* plan);
/* h_ffdtf page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory): results are copied straight into it.
 * h_ffdtf pageable: the copies land in the plan's own pinned buffer and every finished chunk is handed to h_ffdtf by the
 * host while later chunks are in flight.  h_ffdtf NULL: results stay in the plan's pinned buffer (hs_plan_host_result),
 * valid until the next call on this plan.  A plan runs one call at a time (internal lock); plans are independent.       */
int hs_plan_mvar_ffdtf_host(hs_plan* plan, const double* h_x, int64_t t_total, const int64_t* h_starts, int n_win,
                            const double* h_freqs, double fs, double* h_ffdtf, int32_t* h_status);
/* the plan's pinned result buffer (max_windows, m, m, F) float64, allocated on first use */
int hs_plan_host_result(hs_plan* plan, double** h_result, size_t* bytes);

/* ---------------------------------------------------------------- multi-GPU result exchange */

/* The reference keeps one result array per (dyad, film) (src/eeg_alpha_ibi_ffdtf.py:647-656); north_star shards the units
 * by dyad over the GPUs of a box and all-gathers the result over NVLink.  Every rank owns a slot of a gather buffer that
 * exists on every GPU; hs_mvar_ffdtf_f64 writes a chunk of windows into the local slot (d_ffdtf points into it) and
 * hs_gather_push_f64 streams the chunk to the same offset on the peers, on another stream, while the next chunk computes.
 *   d_src              local chunk, `count` float64 (even, 16-byte aligned)
 *   d_multicast_dst    same offset inside an NVSwitch multicast mapping of the buffer (one multimem.st per 16 bytes,
 *                      replicated by the switch to every GPU of the group), or NULL
 *   h_peer_dst         HOST array of n_peers device pointers: the same offset inside each peer's mapping of its buffer
 *                      (plain stores over NVLink); used when d_multicast_dst is NULL
 *   n_ctas             CTAs of the push kernel (<= 0: 8).  The caller reserves that many SMs for it with
 *                      hs_set_compute_sm_limit(sm_count - n_ctas): a K5 CTA owns a whole SM's register file.
 * Cross-rank completion (all peers' pushes have landed) is the caller's barrier after the last push.                   */
int hs_set_compute_sm_limit(int n_sms);      /* 0 = all SMs (default); process-wide */
int hs_gather_push_f64(const double* d_src, int64_t count, void* d_multicast_dst, const void* const* h_peer_dst, int n_peers,
                       int n_ctas, void* stream);
/* same exchange on the copy engines (one cudaMemcpyAsync per peer): the measured alternative */
int hs_gather_push_ce(const double* d_src, int64_t count, const void* const* h_peer_dst, int n_peers, void* stream);
/* cudaMalloc + CUDA IPC export / import of a gather buffer, for setups without torch symmetric memory; handle = 64 bytes */
int hs_ipc_alloc(void** d_ptr, size_t bytes, unsigned char* handle64);
int hs_ipc_open(const unsigned char* handle64, void** d_ptr);
int hs_ipc_close(void* d_ptr);
int hs_ipc_free(void* d_ptr);

/* ---------------------------------------------------------------- front end (K1, K2, K6) */

/* Zero-phase IIR: scipy.signal.filtfilt(b, a, x) with SciPy defaults (odd extension,
 * padlen = 3*ntaps, lfilter_zi initial state), applied for each of n_filt filters in series,
 * after optional DC removal.  Replaces the channel loop of _apply_filters,
 * src/dataloader.py:786-803 (IIR branch :789-792) and the filter block of
 * mne_bridge.load_eeg_signals, src/mne_bridge.py:161-184.
 *   d_x (n_sig rows): element (s, t) at d_x[s*sig_stride + t*t_stride]; filtered in place
 *   h_b, h_a  (n_filt, ntaps) HOST coefficient arrays, zero padded to ntaps, a[0] == 1
 *   d_ws: hs_filtfilt_ws_bytes(n_sig, n) bytes.                                              */
size_t hs_filtfilt_ws_bytes(int n_sig, int64_t n);
int hs_iir_filtfilt_f64(double* d_x, int n_sig, int64_t n, int64_t sig_stride, int64_t t_stride, const double* h_b,
                        const double* h_a, int n_filt, int ntaps, int remove_dc, void* d_ws, void* stream);

/* Causal IIR: scipy.signal.lfilter(b, a, x) with zero initial state, one forward sweep of the same scan kernels.
 * Replaces the notch of the FIR branch of _apply_filters, src/dataloader.py:794 (DC removal :788 fused via remove_dc).
 * d_x (n_sig rows, unit time stride) is NOT modified; d_y (n_sig, n) must be a different buffer.
 * d_ws: hs_filtfilt_ws_bytes(n_sig, n) bytes.                                                 */
int hs_iir_lfilter_f64(const double* d_x, int n_sig, int64_t n, int64_t sig_stride, const double* h_b, const double* h_a,
                       int ntaps, int remove_dc, double* d_y, int64_t y_stride, void* d_ws, void* stream);

/* General FIR with decimation and offset:  y[k] = sum_j b[j] x[q k + off - j],  k < n_out,  x = 0 outside [0, n).
 * off = 0, q = 1 is scipy.signal.lfilter(b, 1, x) (src/dataloader.py:795-796, the 201-tap low-pass and 3049-tap
 * high-pass of the loader's default FIR branch); off = delay, n_out = n - delay is lfilter followed by
 * np.roll(y, -delay) (:798-799); off = (ntaps-1)/2 with q > 1 is hs_fir_decimate_f64.  d_b: DEVICE taps.        */
int hs_fir_filter_f64(const double* d_x, int n_sig, int64_t n, int64_t sig_stride, int q, int64_t off, const double* d_b,
                      int ntaps, double* d_y, int64_t n_out, int64_t y_stride, void* stream);

/* Hilbert envelope: np.abs(scipy.signal.hilbert(x, N=N)[:n]) of EEG_IBI_FFDTF_Pipeline._compute_asymmetry,
 * src/eeg_alpha_ibi_ffdtf.py:352-356 (N = scipy.fft.next_fast_len(n)).  Hand-written mixed-radix Stockham FFT
 * (radices 2..31), forward transform, one-sided mask, inverse transform.
 *   d_x (n_sig rows, row stride sig_stride, n samples each; zero padded / truncated to N like fft(x, N))
 *   d_env (n_sig, min(n, N)) float64 with row stride env_stride, or NULL;  d_analytic (n_sig, min(n, N)) complex128
 *   contiguous, or NULL;  d_ws: hs_hilbert_ws_bytes(n_sig, N) bytes.                                                 */
size_t hs_hilbert_ws_bytes(int n_sig, int64_t N);
int hs_hilbert_f64(const double* d_x, int n_sig, int64_t n, int64_t sig_stride, int64_t N, double* d_env, int64_t env_stride,
                   void* d_analytic, void* d_ws, void* stream);

/* scipy.signal.decimate(x, q, ftype='fir', zero_phase=True), src/data_structures.py:792:
 * y[k] = sum_j b[j] x[q k + half - j], half = (ntaps-1)/2, zero outside.  d_y (n_sig, ceil(n/q)). */
int hs_fir_decimate_f64(const double* d_x, int n_sig, int64_t n, int64_t sig_stride, int q, const double* d_b,
                        int ntaps, double* d_y, int64_t y_stride, void* stream);

/* Multitaper PSD (mne.time_frequency.psd_array_multitaper defaults reached through
 * compute_psd_multitaper, src/psd.py:30-32): remove mean, taper, FFT, eigenvalue-weighted power.
 *   d_x (n_sig, n)   d_tapers (K, n)   d_weights (K) = sqrt(eigvals)
 *   bins k_lo..k_hi-1 of the rfft grid are returned in d_psd (n_sig, k_hi-k_lo).
 * Hand-written FFTs (no cuFFT), ANY n < 2^24: n = 4096 / 8192 in registers (radix-16 stages), n = n1 * 2^a with a small odd
 * n1 in shared memory, every other length through the batched global-memory mixed-radix transform -- directly when all
 * prime factors are <= 31, otherwise as Bluestein's chirp-z over a power-of-two length >= 2n - 1.                     */
size_t hs_mt_psd_ws_bytes(int n_sig, int64_t n, int K);
/* Test / measurement hook: force one of the paths for the following calls (0 automatic, 1 register radix-16,
 * 2 shared memory, 3 general); call hs_mt_psd_ws_bytes again after changing it.  Process-wide.                    */
int hs_mt_psd_set_path(int path);
int hs_mt_psd_f64(const double* d_x, int n_sig, int64_t n, const double* d_tapers, const double* d_weights, int K,
                  int k_lo, int k_hi, double* d_psd, void* d_ws, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HS_B200_H */
