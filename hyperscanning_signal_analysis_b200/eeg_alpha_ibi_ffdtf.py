"""The windowing contract and the per-window MVAR loop of the reference's
``EEG_IBI_FFDTF_Pipeline`` (src/eeg_alpha_ibi_ffdtf.py), batched.

* ``create_windows`` / ``window_starts``: ``_create_windows`` (eeg_alpha_ibi_ffdtf.py:451-518) -- same start
  positions (``np.linspace(0, T - W, n_windows, dtype=int)`` at :513) and the same ValueErrors (:478-508).
* ``compute_ffdtf_windows``: the serial window loop of ``run_pipeline`` (:741-755) calling ``_compute_ffDTF``
  (:521-634: ``freqs = np.arange(fmin, fmax + step, step)`` :584, ``full_freq_dtf`` :592,
  ``multivariate_spectra`` :599), as ONE batched GPU call; results stacked ``(n_win, m, m, F)`` as at :651-652.
* ``alpha_bandpass_filter`` / ``compute_asymmetry`` / ``downsample_signal`` / ``crop_signal`` / ``zscore_rows`` /
  ``preprocess_dyad``: the pre-window stage of ``run_pipeline`` (:693-719) -- SOS zero-phase alpha band-pass (:271-311),
  Hilbert-envelope frontal alpha asymmetry (:314-365), anti-aliased integer down-sampling (:368-406), crop (:409-448),
  channel-wise z-score (:719) -- same signatures minus ``self``, same ValueErrors, the arithmetic on the GPU.
The file discovery, NetCDF loading and plotting around it stay in the reference.
"""
from __future__ import annotations

import numpy as np

from . import _lib, mtmvar, frontend


# --------------------------------------------------------------------------- pre-window stage (src/eeg_alpha_ibi_ffdtf.py:271-448, :693-719)
def alpha_bandpass_filter(data, fs, lowcut=8, highcut=12, order=4, axis=-1):
    """``_alpha_bandpass_filter`` (:271-311): butter(order, [low, high], 'band', output='sos') + ``sosfiltfilt``."""
    from scipy.signal import butter
    nyq = 0.5 * fs
    sos = butter(order, [lowcut / nyq, highcut / nyq], btype='band', output='sos')      # design stays host-side SciPy
    return frontend.sosfiltfilt(sos, data, axis=axis)


def compute_asymmetry(filtered_eeg, channel_names, left_chan="F3", right_chan="F4", metric='amp'):
    """``_compute_asymmetry`` (:314-365): FAA = log(env_right + 1e-12) - log(env_left + 1e-12), env = |hilbert| with
    ``N = next_fast_len(n)``.  ``left_chan`` / ``right_chan`` are the pipeline's ``self.left_chan`` / ``self.right_chan``."""
    torch = _lib.require_cuda()
    try:
        left_idx = channel_names.index(left_chan)
        right_idx = channel_names.index(right_chan)
    except ValueError as e:
        raise ValueError(f"Channels {left_chan} or {right_chan} not found: {e}")
    filtered_eeg = np.asarray(filtered_eeg, dtype=np.float64)
    pair = np.ascontiguousarray(np.stack([filtered_eeg[left_idx, :], filtered_eeg[right_idx, :]]))
    orig_len = pair.shape[1]
    env = frontend.hilbert_envelope_dev(torch.from_numpy(pair).cuda(), N=frontend.next_fast_len(orig_len))
    if metric == 'power':
        env = env ** 2
    elif metric != 'amp':
        raise ValueError("metric must be 'power' or 'amp'")
    faa = torch.log(env[1] + 1e-12) - torch.log(env[0] + 1e-12)
    return faa.cpu().numpy()


def downsample_signal(signal, fs=128, fs_new=8):
    """``_downsample_signal`` (:368-406): ``resample_poly(signal, up=1, down=fs/fs_new)`` for a 1-D signal."""
    torch = _lib.require_cuda()
    if fs_new >= fs:
        raise ValueError("fs_new must be lower than fs")
    ratio = fs / fs_new
    if not np.isclose(ratio, round(ratio)):
        raise ValueError("fs must be divisible by fs_new")
    down = int(round(ratio))
    x = np.ascontiguousarray(np.asarray(signal, dtype=np.float64))
    y = frontend.downsample_dev(torch.from_numpy(np.atleast_2d(x)).cuda(), down).cpu().numpy()
    return y[0] if x.ndim == 1 else y


def crop_signal(signal, fs, drop_front_sec=10, keep_duration_sec=60):
    """``_crop_signal`` (:409-448): drop the first seconds, keep the next ``keep_duration_sec`` (a view, like the reference)."""
    required_sec = drop_front_sec + keep_duration_sec
    total_samples = signal.shape[-1]
    total_sec = total_samples / fs
    if total_sec >= required_sec:
        return signal[..., int(drop_front_sec * fs):int(required_sec * fs)]
    raise ValueError(
        f"Cropping failed: The signal is too short. "
        f"It needs to be at least {required_sec} seconds long, "
        f"but the provided signal is only {total_sec:.2f} seconds.")


def zscore_rows(signals):
    """Channel-wise z-score of run_pipeline (:719): (x - mean) / std along axis 1 (population std, ddof = 0)."""
    signals = np.asarray(signals, dtype=np.float64)
    return (signals - np.mean(signals, axis=1, keepdims=True)) / np.std(signals, axis=1, keepdims=True)


def preprocess_dyad(eeg_ch, eeg_cg, channel_names, ibi_ch, ibi_cg, fs_eeg, fs_ibi, fs_ds=8.0, left_chan="F3", right_chan="F4",
                    drop_front_sec=10, keep_duration_sec=60):
    """run_pipeline's pre-window stage (:693-719) for one dyad/film: alpha band-pass of both EEG arrays (n_ch, n), FAA of each
    participant, down-sampling of FAA and IBI to ``fs_ds``, crop, stack [faa_ch, ibi_ch, faa_cg, ibi_cg], z-score.
    Returns the (4, keep_duration_sec * fs_ds) array that ``_create_windows`` / ``compute_ffdtf_windows`` take."""
    rows = []
    for eeg, ibi in ((eeg_ch, ibi_ch), (eeg_cg, ibi_cg)):
        filt = alpha_bandpass_filter(eeg, fs_eeg)
        faa = compute_asymmetry(filt, list(channel_names), left_chan, right_chan, metric="amp")
        faa_ds = downsample_signal(faa, fs_eeg, fs_ds)
        ibi_ds = downsample_signal(np.asarray(ibi).squeeze(), fs_ibi, fs_ds)
        rows.append(crop_signal(faa_ds, fs_ds, drop_front_sec, keep_duration_sec))
        rows.append(crop_signal(ibi_ds, fs_ds, drop_front_sec, keep_duration_sec))
    return zscore_rows(np.vstack(rows))


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
