// Parameter blocks and host-side launchers of the MVAR kernels (internal).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace hs {

constexpr int kPadMaxHost = 40;

struct K3Params {
    const double* x;             // base pointer
    const long long* offsets;    // (n_win * trials) element offset of (channel 0, sample 0) of each unit
    long long ch_stride;         // elements between channels
    double* R;                   // (n_win, p+1, m, m)
    int n_win, trials, m, n, p;
};

struct K4Params {
    const double* R;     // (n_win, p+1, m, m)
    double* A;           // (n_win, m, m, p)
    double* V;           // (n_win, m, m)
    double* Vall;        // (n_win, p, m, m) residual covariance after each order, or null
    int* status;         // (n_win)
    double* ws;          // grid * 4 * p * m * m doubles
    int n_win, m, p;
};

struct K5Params {
    const double* A;        // (n_win, m, m, p)
    const double2* z;       // (p, F)
    double* dtf;            // (n_win, m, m, F) or null
    double* rowpart;        // (n_win, n_seg, m) partial row sums, or null
    double2* H;             // (n_win, m, m, F) complex or null
    double2* Af;            // (n_win, m, m, F) complex or null
    int* status;            // (n_win) sticky singular flags
    int n_win, m, p, F, n_seg, seg_len;
    int flip;               // 1: alternate the column-owner parity between groups sharing an SMSP pair
    double* rowpart2;       // row sums of the matrices redone with pivoting (mode 2)
    int* bad;               // (n_win * F) flags set by the optimistic pass
    int* bad_count;         // number of flagged matrices
    int* bad_list;          // (n_win * F) compact list of flagged matrices (w * F + f), filled by the optimistic pass
    double verify_tol2;     // squared relative tolerance of the a-posteriori check
    int per_cta;            // tensor-pipe kernel: matrices (w * F + f, contiguous range) per CTA; rowpart is then (n_win, n_seg, m)
                            //    with slot = CTA index - first CTA of the window, n_seg = ceil(F / per_cta) + 1
    int pipe_turns;         // tensor-pipe kernel: 1 = one warp at a time per SM sub-partition streams its block step's DMMAs
    int dtf_fij;            // 1: P.dtf is a staging buffer laid out (n_win, m, F, m) -- row i of every bin's matrix contiguous --
                            //    that launch_dtf_finalize transposes to the reference's (n_win, m, m, F)
};

int launch_lagcov(const K3Params& P, cudaStream_t stream);
int launch_toeplitz(const double* R, int n_win, int m, int p, double* G, double* rhs, cudaStream_t stream);
size_t lwr_ws_doubles(int grid, int m, int p);
int lwr_grid(int n_win);
int launch_lwr(const K4Params& P, int grid, cudaStream_t stream);
int launch_ztable(const double* freqs, int F, int p, double fs, void* z, cudaStream_t stream);
int launch_transfer_dtf(const K5Params& P, int ng, int mode, cudaStream_t stream);
bool transfer_mma_fits(int p, int ng, int seg_len);
void transfer_mma_partition(int n_win, int F, int* per_cta, int* slots, int ng);      // balanced split of the n_win * F matrices over the SMs       // shared memory for the coefficient planes of order p fits next to ng groups
int launch_transfer_mma(const K5Params& P, int ng, cudaStream_t stream);
int launch_transfer_ws(const K5Params& P, int ng, cudaStream_t stream);               // warp-specialised variant: 4 groups x (Re, Im, helper) warps      // optimistic pass on the FP64 tensor pipe (transfer_mma.cu)
int launch_dtf_finalize(const double* stage, const double* rowpart, const int* bad, int n_win, int m, int F, int n_seg,
                        double* dtf_out, double* ffdtf_out, cudaStream_t stream);
int launch_ffdtf_normalize(double* dtf, const double* rowpart, const double* rowpart2, int n_win, int m, int F, int n_seg, double* out,
                           cudaStream_t stream);
size_t lwr_generic_ws_doubles(int n_win, int m, int p);
int launch_lwr_generic(const K4Params& P, cudaStream_t stream);
size_t transfer_generic_scratch_bytes(int m);
int launch_transfer_generic(const K5Params& P, void* scratch, cudaStream_t stream);
int launch_pcoh(const void* S, int n_win, int m, int F, void* kappa, const double* ffdtf, double* ddtf, int* status, cudaStream_t stream);
int launch_spectra(const void* H, const double* V, int n_win, int m, int F, void* S, cudaStream_t stream);

}  // namespace hs
