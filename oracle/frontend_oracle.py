"""CPU oracle for the front end (filter / decimate / multitaper PSD)
--  TEST INFRASTRUCTURE ONLY (see oracle/mvar_oracle.py header for the rule).

* ``design_eeg_filters`` / ``apply_filters_iir``: the SciPy calls of the
  reference's ``dataloader._design_eeg_filters`` (dataloader.py:687,704-708)
  and the IIR branch of ``_apply_filters`` (dataloader.py:788-792).
* ``filtfilt_df2t``: explicit restatement of what ``scipy.signal.filtfilt``
  does with its defaults -- pinned against SciPy
  itself in tests (agrees to ~1e-15 relative; SciPy's C loop contracts FMAs).
* ``decimate_fir``: closed form of ``scipy.signal.decimate(x, q, ftype='fir',
  zero_phase=True)`` (data_structures.py:792; Appendix A.2), pinned against SciPy.
* ``psd_multitaper``: restatement of MNE 1.11 ``psd_array_multitaper`` defaults
  (psd.py:30-32; Appendix A.6).  **parity unpinned**: mne is a third-party
  dependency (requirements.txt:11, mne==1.11.0) that is neither vendored in the
  reference nor installed here; the restatement follows MNE's published
  algorithm and is cross-checked against ``mne`` whenever it is importable.
"""
from __future__ import annotations

import numpy as np
from scipy import signal


# ------------------------------------------------------------------ design
def design_eeg_filters(fs, lowcut, highcut, notch_freq=50, notch_q=30, filter_type="iir"):
    b_n, a_n = signal.iirnotch(notch_freq, notch_q, fs=fs)
    if filter_type == "fir":
        b_l = signal.firwin(201, highcut, fs=fs, pass_zero="lowpass")
        b_h = signal.firwin(3049, lowcut, fs=fs, pass_zero="highpass")
        a_l = a_h = 1.0
    else:
        b_l, a_l = signal.butter(N=2, Wn=highcut, btype="low", fs=fs)
        b_h, a_h = signal.butter(N=2, Wn=lowcut, btype="high", fs=fs)
    return (b_n, a_n), (b_l, a_l), (b_h, a_h), filter_type


def apply_filters_iir(x, filters):
    """Rows of ``x`` (n_ch, N): DC removal then notch -> low -> high filtfilt."""
    (b_n, a_n), (b_l, a_l), (b_h, a_h), _ = filters
    out = np.empty_like(x, dtype=np.float64)
    for r in range(x.shape[0]):
        s = x[r] - np.mean(x[r])
        s = signal.filtfilt(b_n, a_n, s, axis=0)
        s = signal.filtfilt(b_l, a_l, s, axis=0)
        s = signal.filtfilt(b_h, a_h, s, axis=0)
        out[r] = s
    return out


def apply_filters_fir(x, filters):
    """FIR branch of ``_apply_filters`` (/root/reference src/dataloader.py:788, 793-801): DC removal, CAUSAL lfilter
    notch -> 201-tap low-pass -> 3049-tap high-pass, then the two FIR group delays are rolled out and the wrapped tail
    is zeroed (the notch's delay is not compensated: SURVEY.md Appendix B.8)."""
    (b_n, a_n), (b_l, a_l), (b_h, a_h), _ = filters
    out = np.empty_like(x, dtype=np.float64)
    for r in range(x.shape[0]):
        s = x[r] - np.mean(x[r])
        s = signal.lfilter(b_n, a_n, s, axis=0)
        s = signal.lfilter(b_l, a_l, s, axis=0)
        s = signal.lfilter(b_h, a_h, s, axis=0)
        delay = (len(b_l) - 1) // 2 + (len(b_h) - 1) // 2
        s = np.roll(s, -delay)
        s[-delay:] = 0.0
        out[r] = s
    return out


# ------------------------------------------------------------- filtfilt A.1
def _lfilter_zi(b, a):
    """Steady-state DF2T state for a unit step (scipy.signal.lfilter_zi)."""
    b = np.atleast_1d(b).astype(float)
    a = np.atleast_1d(a).astype(float)
    if a[0] != 1.0:
        b = b / a[0]
        a = a / a[0]
    n = max(len(a), len(b))
    a = np.r_[a, np.zeros(n - len(a))]
    b = np.r_[b, np.zeros(n - len(b))]
    comp = np.zeros((n - 1, n - 1))
    comp[0, :] = -a[1:]
    comp[1:, :-1] = np.eye(n - 2)
    IminusA = np.eye(n - 1) - comp.T
    B = b[1:] - a[1:] * b[0]
    return np.linalg.solve(IminusA, B)


def _df2t(b, a, u, z):
    """Direct-form-II-transposed recursion, sequential (order len(a)-1)."""
    n = len(a)
    z = z.copy()
    y = np.empty_like(u)
    for t in range(len(u)):
        ut = u[t]
        yt = b[0] * ut + z[0]
        for k in range(n - 2):
            z[k] = b[k + 1] * ut + z[k + 1] - a[k + 1] * yt
        z[n - 2] = b[n - 1] * ut - a[n - 1] * yt
        y[t] = yt
    return y


def filtfilt_df2t(b, a, x):
    """filtfilt with SciPy defaults: odd extension, padlen 3*ntaps, zi init."""
    b = np.atleast_1d(b).astype(float)
    a = np.atleast_1d(a).astype(float)
    nt = max(len(a), len(b))
    b = np.r_[b, np.zeros(nt - len(b))] / a[0]
    a = np.r_[a, np.zeros(nt - len(a))] / a[0]
    e = 3 * nt
    x = np.asarray(x, dtype=np.float64)
    if x.shape[0] <= e:
        raise ValueError("The length of the input vector x must be greater than padlen, which is %d." % e)
    ext = np.concatenate((2 * x[0] - x[e:0:-1], x, 2 * x[-1] - x[-2:-e - 2:-1]))
    zi = _lfilter_zi(b, a)
    f = _df2t(b, a, ext, zi * ext[0])
    r = _df2t(b, a, f[::-1], zi * f[-1])
    return r[::-1][e:-e]


# ------------------------------------------------------------ decimate A.2
def decimate_taps(q):
    return signal.firwin(20 * q + 1, 1.0 / q, window="hamming")


def decimate_fir(x, q):
    """y[k] = sum_j b[j] x[q k + 10 q - j], zero outside [0, N)."""
    x = np.asarray(x, dtype=np.float64)
    b = decimate_taps(q)
    half = 10 * q
    n = x.shape[-1]
    n_out = -(-n // q)
    xp = np.concatenate((np.zeros(half), x, np.zeros(half + q)))
    out = np.empty(n_out)
    br = b[::-1]
    for k in range(n_out):
        # x[qk+half-j], j=0..2*half  ->  padded index qk+2*half-j
        out[k] = np.dot(br, xp[q * k: q * k + 2 * half + 1])
    return out


def decimate_column(values, q):
    """NaN handling of ``MultimodalData._decimate_signals`` (data_structures.py:779-797)."""
    v = np.asarray(values, dtype=float)
    nan = np.isnan(v)
    if nan.any():
        idx = np.where(~nan, np.arange(len(v)), -1)
        np.maximum.accumulate(idx, out=idx)          # ffill
        filled = np.where(idx >= 0, v[np.maximum(idx, 0)], np.nan)
        if np.isnan(filled).any():                   # leading NaNs -> bfill
            good = np.where(~np.isnan(filled))[0]
            filled = np.where(np.isnan(filled), filled[good[0]] if len(good) else 0.0, filled)
        v = filled
    y = signal.decimate(v, q, ftype="fir", zero_phase=True)
    if nan.any():
        y[nan[::q]] = np.nan
    return y


# ------------------------------------------------------ mne_bridge filters
def bridge_filters(data_tc, fs, low_cutoff_hz=None, high_cutoff_hz=None):
    """Filter block of ``mne_bridge.load_eeg_signals`` (mne_bridge.py:158-184) on (time, channel)."""
    nyq = fs / 2.0
    d = np.asarray(data_tc, dtype=np.float64)
    if low_cutoff_hz is not None:
        wn = float(low_cutoff_hz) / nyq
        if not 0.0 < wn < 1.0:
            raise ValueError(f"Invalid low_cutoff_hz={low_cutoff_hz}. Must satisfy 0 < cutoff < {nyq:.3f} Hz.")
        b, a = signal.butter(4, wn, btype="highpass")
        d = signal.filtfilt(b, a, d, axis=0)
    if high_cutoff_hz is not None:
        wn = float(high_cutoff_hz) / nyq
        if not 0.0 < wn < 1.0:
            raise ValueError(f"Invalid high_cutoff_hz={high_cutoff_hz}. Must satisfy 0 < cutoff < {nyq:.3f} Hz.")
        b, a = signal.butter(4, wn, btype="lowpass")
        d = signal.filtfilt(b, a, d, axis=0)
    if 50.0 < nyq:
        b, a = signal.iirnotch(50.0, Q=15, fs=fs)
        d = signal.filtfilt(b, a, d, axis=0)
    return d


# ------------------------------------------------------- multitaper A.6
def multitaper_params(n_times, sfreq, bandwidth):
    half_nbw = float(bandwidth) * n_times / (2.0 * sfreq)
    k_max = int(2 * half_nbw)
    tapers, eig = signal.windows.dpss(n_times, half_nbw, k_max, sym=False, norm=2, return_ratios=True)
    keep = eig > 0.9
    if not keep.any():
        keep = np.zeros_like(keep)
        keep[np.argmax(eig)] = True
    return np.ascontiguousarray(tapers[keep]), np.ascontiguousarray(eig[keep])


def psd_multitaper(data, sfreq, fmin, fmax, bandwidth):
    """(freqs, psd) as ``psd.compute_psd_multitaper`` returns them (psd.py:30-33)."""
    x = np.asarray(data, dtype=np.float64)
    n = x.shape[-1]
    tapers, eig = multitaper_params(n, sfreq, bandwidth)
    x0 = x - x.mean(axis=-1, keepdims=True)
    xk = np.fft.rfft(x0[:, None, :] * tapers[None, :, :], n=n)
    xk[..., 0] /= np.sqrt(2.0)
    if n % 2 == 0:
        xk[..., -1] /= np.sqrt(2.0)
    w = np.sqrt(eig)
    psd = (np.abs(w[None, :, None] * xk) ** 2).sum(axis=1) * 2.0 / (w ** 2).sum()
    freqs = np.fft.rfftfreq(n, 1.0 / sfreq)
    mask = (freqs >= fmin) & (freqs <= fmax)
    return freqs[mask], psd[:, mask]


def average_psd_across_conditions(psd_dict):
    if not psd_dict:
        raise ValueError("psd_dict is empty; no conditions to average PSD over.")
    return np.mean(np.stack(list(psd_dict.values()), axis=0), axis=0)


# ------------------------------------------------------------------ pre-window stage of EEG_IBI_FFDTF_Pipeline
def alpha_bandpass(data, fs, lowcut=8, highcut=12, order=4, axis=-1):
    """``_alpha_bandpass_filter`` (/root/reference src/eeg_alpha_ibi_ffdtf.py:303-309): SciPy's own calls."""
    nyq = 0.5 * fs
    sos = signal.butter(order, [lowcut / nyq, highcut / nyq], btype="band", output="sos")
    return signal.sosfiltfilt(sos, data, axis=axis)


def sosfiltfilt_explicit(sos, x):
    """Restatement of ``scipy.signal.sosfiltfilt`` defaults along the last axis (odd extension by
    3 * (2 n_sections + 1 - min(#b2 == 0, #a2 == 0)), ``sosfilt_zi`` scaled by the first sample of each sweep)."""
    sos = np.atleast_2d(sos)
    ns = sos.shape[0]
    ntaps = 2 * ns + 1 - min(int((sos[:, 2] == 0).sum()), int((sos[:, 5] == 0).sum()))
    edge = 3 * ntaps
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    ext = np.concatenate([2 * x[:, :1] - x[:, edge:0:-1], x, 2 * x[:, -1:] - x[:, -2:-edge - 2:-1]], axis=1)
    zi = signal.sosfilt_zi(sos)[:, None, :]
    f, _ = signal.sosfilt(sos, ext, axis=1, zi=zi * ext[:, :1][None])
    r = f[:, ::-1]
    b, _ = signal.sosfilt(sos, r, axis=1, zi=zi * r[:, :1][None])
    return b[:, ::-1][:, edge:-edge]


def hilbert_envelope(x, N=None):
    """|scipy.signal.hilbert(x, N)[..., :n]| written out (eeg_alpha_ibi_ffdtf.py:352-356): one-sided spectrum doubling."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[-1]
    N = n if N is None else int(N)
    X = np.fft.fft(x, N, axis=-1)
    h = np.zeros(N)
    if N % 2 == 0:
        h[0] = h[N // 2] = 1
        h[1:N // 2] = 2
    else:
        h[0] = 1
        h[1:(N + 1) // 2] = 2
    return np.abs(np.fft.ifft(X * h, axis=-1)[..., :n])


def frontal_alpha_asymmetry(filtered_eeg, left_idx, right_idx, metric="amp"):
    """``_compute_asymmetry`` (eeg_alpha_ibi_ffdtf.py:343-365)."""
    from scipy.fft import next_fast_len
    n = filtered_eeg.shape[1]
    N = next_fast_len(n)
    left = hilbert_envelope(filtered_eeg[left_idx], N)
    right = hilbert_envelope(filtered_eeg[right_idx], N)
    if metric == "power":
        left, right = left ** 2, right ** 2
    return np.log(right + 1e-12) - np.log(left + 1e-12)


def downsample(x, down):
    """``resample_poly(x, 1, down)`` in closed form (eeg_alpha_ibi_ffdtf.py:403): Kaiser-5 firwin(20 down + 1, 1/down),
    y[k] = sum_j h[j] x[down k + 10 down - j], zero outside."""
    x = np.asarray(x, dtype=np.float64)
    h = signal.firwin(20 * down + 1, 1.0 / down, window=("kaiser", 5.0))
    full = np.convolve(x, h)
    n_out = -(-len(x) // down)
    idx = down * np.arange(n_out) + 10 * down
    return full[idx]
