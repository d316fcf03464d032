"""Drop-in for the reference's ``src/psd.py``: ``compute_psd_multitaper`` (psd.py:7-33) and
``average_psd_across_conditions`` (psd.py:36-52).

The reference delegates to ``mne.time_frequency.psd_array_multitaper`` (mne==1.11.0, not vendored, not
installed here); this module follows that function's defaults as restated in SURVEY.md A.6
(adaptive=False, low_bias=True, normalization='length').  The DPSS tapers are a data-independent design
step and come from ``scipy.signal.windows.dpss`` on the host (cached per length); mean removal, tapering,
the FFTs (hand-written, no cuFFT) and the eigenvalue-weighted power sum run on the GPU.
"""
from __future__ import annotations

import functools

import numpy as np

from . import _lib


@functools.lru_cache(maxsize=16)
def _tapers(n_times, half_nbw):
    from scipy.signal.windows import dpss
    k_max = int(2 * half_nbw)
    tapers, eig = dpss(n_times, half_nbw, k_max, sym=False, norm=2, return_ratios=True)
    keep = eig > 0.9                       # low_bias=True
    if not keep.any():
        keep = np.zeros_like(keep)
        keep[np.argmax(eig)] = True
    return np.ascontiguousarray(tapers[keep]), np.ascontiguousarray(np.sqrt(eig[keep]))


def psd_multitaper_dev(x_dev, sfreq, fmin, fmax, bandwidth):
    """(n_sig, n_times) CUDA float64 tensor -> (freqs ndarray, psd CUDA tensor (n_sig, n_freqs))."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    n_sig, n = x_dev.shape
    half_nbw = float(bandwidth) * n / (2.0 * sfreq)
    tapers, weights = _tapers(int(n), half_nbw)
    freqs = np.fft.rfftfreq(n, 1.0 / sfreq)
    mask = (freqs >= fmin) & (freqs <= fmax)
    idx = np.nonzero(mask)[0]
    if idx.size == 0:
        return freqs[mask], torch.empty((n_sig, 0), dtype=torch.float64, device="cuda")
    k_lo, k_hi = int(idx[0]), int(idx[-1]) + 1
    t_dev = torch.from_numpy(tapers).cuda()
    w_dev = torch.from_numpy(weights).cuda()
    psd = torch.empty((n_sig, k_hi - k_lo), dtype=torch.float64, device="cuda")
    K = tapers.shape[0]
    ws = torch.empty(max(int(lib.hs_mt_psd_ws_bytes(n_sig, n, K)), 16), dtype=torch.uint8, device="cuda")
    if n_sig:
        _lib.check(lib.hs_mt_psd_f64(x_dev.data_ptr(), n_sig, n, t_dev.data_ptr(), w_dev.data_ptr(), K, k_lo, k_hi,
                                     psd.data_ptr(), ws.data_ptr(), torch.cuda.current_stream().cuda_stream), "hs_mt_psd_f64")
    return freqs[mask], psd


def compute_psd_multitaper(data, sfreq, fmin, fmax, bandwidth):
    """Same signature and return order ``(freqs, psd)`` as the reference (psd.py:7-33)."""
    torch = _lib.require_cuda()
    x = torch.from_numpy(np.ascontiguousarray(np.atleast_2d(np.asarray(data, dtype=np.float64)))).cuda()
    freqs, psd = psd_multitaper_dev(x, sfreq, fmin, fmax, bandwidth)
    return freqs, psd.cpu().numpy()


def average_psd_across_conditions(psd_dict):
    """Reference ``average_psd_across_conditions`` (psd.py:36-52); O(n_ch * n_f) host mean."""
    if not psd_dict:
        raise ValueError('psd_dict is empty; no conditions to average PSD over.')
    return np.mean(np.stack(list(psd_dict.values()), axis=0), axis=0)
