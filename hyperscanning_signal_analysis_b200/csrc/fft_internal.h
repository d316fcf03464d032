// Batched global-memory FFT shared by the Hilbert envelope (K7) and the general multitaper path (K6): the mixed-radix
// Stockham passes of hilbert_kernels.cu (internal, not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

namespace hs {

// true when every prime factor of N is <= 31 (one Stockham pass per factor)
bool fft_smooth(long long N);
// forward DFT of `batch` contiguous rows of length N (N smooth): ping-pong between a and b, result left in `a`
int fft_forward_batched(double2*& a, double2*& b, int batch, long long N, cudaStream_t st);

}  // namespace hs
