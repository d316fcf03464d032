// Result exchange between the GPUs of one box (north_star: "NCCL over NVLink is used only for the final result
// all-gather"; the reference keeps one result array per (dyad, film): src/eeg_alpha_ibi_ffdtf.py:647-656).
//
// Every rank owns one slot of a gather buffer that exists, at the same size, on every GPU (symmetric allocation).  K5's
// finalize kernel writes a chunk of windows straight into the local slot; gather_push_kernel then streams that chunk to the
// same offset of every peer's buffer while the next chunk is being computed on another stream:
//   * with an NVSwitch multicast mapping of the buffer: ONE multimem.st per 16 bytes, replicated by the switch to all
//     GPUs of the group (egress = the rank's own data, once);
//   * otherwise plain 16-byte stores to each peer's mapped pointer (P2P over NVLink, egress = (N-1) copies).
// Few CTAs drive it (the link, not the SMs, is the limit); the MVAR kernels are told to leave those SMs alone
// (hs_set_compute_sm_limit), because a K5 CTA takes a whole SM's register file and could not share one with a pusher.
#include <cstdint>
#include <cstring>
#include <mutex>
#include <vector>

#include "hs_internal.h"

namespace hs {

constexpr int kPushThreads = 512;
constexpr int kMaxPeers = 16;

struct PushPeers {
    double2* dst[kMaxPeers];
};

__device__ __forceinline__ void multimem_st16(double2* mc, const double2 v) {
    const unsigned a = (unsigned)__double2loint(v.x), b = (unsigned)__double2hiint(v.x);
    const unsigned c = (unsigned)__double2loint(v.y), d = (unsigned)__double2hiint(v.y);
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// src, dst: 16-byte aligned; n16 = number of 16-byte units.  Grid-stride, 4 loads in flight per thread.
template <bool MULTICAST>
__global__ void __launch_bounds__(kPushThreads) gather_push_kernel(const double2* __restrict__ src, double2* __restrict__ mc, const PushPeers peers,
                                                                   const int n_peers, const long long n16) {
    const long long stride = (long long)gridDim.x * kPushThreads;
    long long i = (long long)blockIdx.x * kPushThreads + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        double2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldcs(src + i + u * stride);
        if (MULTICAST) {
#pragma unroll
            for (int u = 0; u < 4; ++u) multimem_st16(mc + i + u * stride, v[u]);
        } else {
            for (int q = 0; q < n_peers; ++q) {
                double2* d = peers.dst[q];
#pragma unroll
                for (int u = 0; u < 4; ++u) d[i + u * stride] = v[u];
            }
        }
    }
    for (; i < n16; i += stride) {
        const double2 v = __ldcs(src + i);
        if (MULTICAST) multimem_st16(mc + i, v);
        else
            for (int q = 0; q < n_peers; ++q) peers.dst[q][i] = v;
    }
}

static int g_sm_limit = 0;       // 0: all SMs

int compute_sm_count() {
    const int sm = device_sm_count();
    return (g_sm_limit > 0 && g_sm_limit < sm) ? g_sm_limit : sm;
}

}  // namespace hs

using namespace hs;

extern "C" {

int hs_set_compute_sm_limit(int n_sms) {
    if (n_sms < 0) return set_error(HS_ERR_INVALID, "hs_set_compute_sm_limit: negative SM count");
    g_sm_limit = n_sms;
    return HS_OK;
}

int hs_gather_push_f64(const double* d_src, int64_t count, void* d_multicast_dst, const void* const* h_peer_dst, int n_peers, int n_ctas,
                       void* stream) {
    if (count == 0) return HS_OK;
    if (!d_src || count < 0) return set_error(HS_ERR_INVALID, "hs_gather_push_f64: bad source");
    if (!d_multicast_dst && (n_peers < 1 || !h_peer_dst)) return set_error(HS_ERR_INVALID, "hs_gather_push_f64: no destination");
    if (n_peers > kMaxPeers) return set_error(HS_ERR_UNSUPPORTED, "hs_gather_push_f64: at most %d peers", kMaxPeers);
    if ((count & 1) || (reinterpret_cast<uintptr_t>(d_src) & 15) || (reinterpret_cast<uintptr_t>(d_multicast_dst) & 15))
        return set_error(HS_ERR_INVALID, "hs_gather_push_f64: buffers must be 16-byte aligned and hold an even number of doubles");
    if (n_ctas < 1) n_ctas = 8;
    if (n_ctas > 148) n_ctas = 148;
    PushPeers pp;
    memset(&pp, 0, sizeof(pp));
    if (!d_multicast_dst) {
        for (int q = 0; q < n_peers; ++q) {
            if (!h_peer_dst[q] || (reinterpret_cast<uintptr_t>(h_peer_dst[q]) & 15))
                return set_error(HS_ERR_INVALID, "hs_gather_push_f64: peer pointer %d is null or unaligned", q);
            pp.dst[q] = reinterpret_cast<double2*>(const_cast<void*>(h_peer_dst[q]));
        }
    }
    const long long n16 = count / 2;
    const double2* src = reinterpret_cast<const double2*>(d_src);
    if (d_multicast_dst)
        gather_push_kernel<true><<<n_ctas, kPushThreads, 0, (cudaStream_t)stream>>>(src, reinterpret_cast<double2*>(d_multicast_dst), pp, 0, n16);
    else
        gather_push_kernel<false><<<n_ctas, kPushThreads, 0, (cudaStream_t)stream>>>(src, nullptr, pp, n_peers, n16);
    return check_launch("gather_push_kernel");
}

// Copy-engine variant of the same exchange (no SM work at all): one peer copy per destination.  Kept as the measured
// alternative to the store kernel (tools/gather_probe.py); the caller chooses.
int hs_gather_push_ce(const double* d_src, int64_t count, const void* const* h_peer_dst, int n_peers, void* stream) {
    if (count == 0) return HS_OK;
    if (!d_src || count < 0 || n_peers < 1 || !h_peer_dst) return set_error(HS_ERR_INVALID, "hs_gather_push_ce: bad arguments");
    for (int q = 0; q < n_peers; ++q) {
        cudaError_t e = cudaMemcpyAsync(const_cast<void*>(h_peer_dst[q]), d_src, (size_t)count * sizeof(double), cudaMemcpyDefault, (cudaStream_t)stream);
        if (e != cudaSuccess) return set_error(HS_ERR_CUDA, "hs_gather_push_ce: copy to peer %d: %s", q, cudaGetErrorString(e));
    }
    return HS_OK;
}

// ---- plain cudaMalloc buffers shared between the ranks of one box through CUDA IPC (used when torch's symmetric
//      memory / the multicast mapping is not available).  The handle is 64 opaque bytes the ranks exchange themselves.
int hs_ipc_alloc(void** d_ptr, size_t bytes, unsigned char* handle64) {
    if (!d_ptr || !handle64 || bytes == 0) return set_error(HS_ERR_INVALID, "hs_ipc_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaError_t e = cudaMalloc(d_ptr, bytes);
    if (e != cudaSuccess) return set_error(e == cudaErrorMemoryAllocation ? HS_ERR_NOMEM : HS_ERR_CUDA, "hs_ipc_alloc: %s", cudaGetErrorString(e));
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, *d_ptr);
    if (e != cudaSuccess) {
        cudaFree(*d_ptr);
        *d_ptr = nullptr;
        return set_error(HS_ERR_CUDA, "hs_ipc_alloc: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    }
    memcpy(handle64, &h, 64);
    return HS_OK;
}

int hs_ipc_open(const unsigned char* handle64, void** d_ptr) {
    if (!handle64 || !d_ptr) return set_error(HS_ERR_INVALID, "hs_ipc_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return set_error(HS_ERR_CUDA, "hs_ipc_open: %s", cudaGetErrorString(e));
    return HS_OK;
}

int hs_ipc_close(void* d_ptr) {
    if (!d_ptr) return HS_OK;
    cudaError_t e = cudaIpcCloseMemHandle(d_ptr);
    if (e != cudaSuccess) return set_error(HS_ERR_CUDA, "hs_ipc_close: %s", cudaGetErrorString(e));
    return HS_OK;
}

int hs_ipc_free(void* d_ptr) {
    if (!d_ptr) return HS_OK;
    cudaError_t e = cudaFree(d_ptr);
    if (e != cudaSuccess) return set_error(HS_ERR_CUDA, "hs_ipc_free: %s", cudaGetErrorString(e));
    return HS_OK;
}

}  // extern "C"
