"""The windowing contract and the per-window MVAR loop of the reference's
``EEG_IBI_FFDTF_Pipeline`` (src/eeg_alpha_ibi_ffdtf.py), batched.

* ``create_windows`` / ``window_starts``: ``_create_windows`` (eeg_alpha_ibi_ffdtf.py:451-518) -- same start
  positions (``np.linspace(0, T - W, n_windows, dtype=int)`` at :513) and the same ValueErrors (:478-508).
* ``compute_ffdtf_windows``: the serial window loop of ``run_pipeline`` (:741-755) calling ``_compute_ffDTF``
  (:521-634: ``freqs = np.arange(fmin, fmax + step, step)`` :584, ``full_freq_dtf`` :592,
  ``multivariate_spectra`` :599), as ONE batched GPU call; results stacked ``(n_win, m, m, F)`` as at :651-652.
The file discovery, NetCDF loading, Hilbert/FAA, resampling and plotting around it stay in the reference.
"""
from __future__ import annotations

import numpy as np

from . import _lib, mtmvar


def window_starts(T, n_windows=3, window_size=None):
    if window_size is None:
        if T % n_windows != 0:
            raise ValueError(
                f"Cannot evenly divide signal of length {T} into {n_windows} "
                f"non-overlapping windows. Provide a specific window_size.")
        window_size = T // n_windows
    else:
        min_required_size = (T + n_windows - 1) // n_windows
        if window_size < min_required_size:
            raise ValueError(
                f"window_size={window_size} is too short. To cover {T} samples "
                f"with {n_windows} windows without leaving gaps, the minimum "
                f"window_size is {min_required_size}.")
        if window_size > T:
            raise ValueError(f"window_size ({window_size}) cannot exceed signal length ({T}).")
    max_start = T - window_size
    if max_start < n_windows - 1 and n_windows > 1:
        raise ValueError(
            f"window_size={window_size} is too large to generate {n_windows} "
            f"distinct windows. Decrease window_size or n_windows.")
    if n_windows == 1:
        positions = np.zeros(1, dtype=np.int64)
    else:
        positions = np.linspace(0, max_start, n_windows, dtype=int).astype(np.int64)
    return positions, int(window_size)


def create_windows(signals, n_windows=3, window_size=None):
    """List of ``n_windows`` views ``signals[:, s:s+W]`` exactly as the reference returns them."""
    positions, window_size = window_starts(signals.shape[1], n_windows, window_size)
    return [signals[:, s:s + window_size] for s in positions]


def compute_ffdtf_windows(signals, fs, n_windows, window_size=None, ar_p=5, freq_min=0.0, freq_max=None, freq_step=0.125,
                          with_spectra=False, max_model_order=20, crit_type="AIC"):
    """ffDTF (and optionally S = H V H^T) of every window in one batched call.

    Returns dict(ff_dtf_windowed (n_win, m, m, F) float64, spectra_windowed (n_win, m, m, F) complex128 | None,
    freqs, p_opt_w (list), starts).  ``ar_p=None`` selects the order per window with ``mvar_criterion`` like
    the reference (:586-587)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    signals = np.asarray(signals, dtype=np.float64)
    if freq_max is None:
        freq_max = fs / 2.0
    freqs = np.arange(freq_min, freq_max + freq_step, freq_step)
    starts, W = window_starts(signals.shape[1], n_windows, window_size)
    x = torch.from_numpy(np.ascontiguousarray(signals)).cuda()
    m = signals.shape[0]
    if ar_p is None:
        orders = [int(mtmvar.mvar_criterion(signals[:, s:s + W], max_model_order, crit_type, False)[2]) for s in starts]
    else:
        orders = [int(ar_p)] * len(starts)
    ff = torch.empty((len(starts), m, m, len(freqs)), dtype=torch.float64, device="cuda")
    S = torch.empty((len(starts), m, m, len(freqs)), dtype=torch.complex128, device="cuda") if with_spectra else None
    for p in sorted(set(orders)):
        sel = np.array([i for i, o in enumerate(orders) if o == p])
        out, A, V = mtmvar.windowed_ffdtf(x, starts[sel], W, freqs, fs, p, return_model=True)
        ff[torch.from_numpy(sel).cuda()] = out
        if with_spectra:
            res = mtmvar.batched_transfer(A, freqs, fs, want=("H",))
            Ssel = torch.empty_like(res["H"])
            _lib.check(lib.hs_spectra_f64(res["H"].data_ptr(), V.data_ptr(), len(sel), m, len(freqs), Ssel.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream), "hs_spectra_f64")
            S[torch.from_numpy(sel).cuda()] = Ssel
    return {"ff_dtf_windowed": ff.cpu().numpy(), "spectra_windowed": S.cpu().numpy() if with_spectra else None,
            "freqs": freqs, "p_opt_w": orders, "starts": starts}
