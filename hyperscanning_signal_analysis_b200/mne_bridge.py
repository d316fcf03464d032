"""Filter block of the reference's ``mne_bridge.load_eeg_signals`` (src/mne_bridge.py:158-184) on the GPU.
The NetCDF/xarray I/O around it is out of scope; ``filter_time_channel`` takes the
``(time, channel)`` float64 array the reference builds at :158 and returns what it holds after :184."""
from __future__ import annotations

import numpy as np
from scipy.signal import butter, iirnotch

from . import frontend


def filter_time_channel(data_tc, fs, low_cutoff_hz=None, high_cutoff_hz=None):
    nyquist = fs / 2.0
    filters = []
    if low_cutoff_hz is not None:
        wn_hp = float(low_cutoff_hz) / nyquist
        if not 0.0 < wn_hp < 1.0:
            raise ValueError(f"Invalid low_cutoff_hz={low_cutoff_hz}. Must satisfy 0 < cutoff < {nyquist:.3f} Hz.")
        filters.append(butter(4, wn_hp, btype="highpass"))
    if high_cutoff_hz is not None:
        wn_lp = float(high_cutoff_hz) / nyquist
        if not 0.0 < wn_lp < 1.0:
            raise ValueError(f"Invalid high_cutoff_hz={high_cutoff_hz}. Must satisfy 0 < cutoff < {nyquist:.3f} Hz.")
        filters.append(butter(4, wn_lp, btype="lowpass"))
    if 50.0 < nyquist:
        filters.append(iirnotch(50.0, Q=15, fs=fs))
    data_tc = np.asarray(data_tc, dtype=np.float64)
    if not filters:
        return data_tc.copy()
    return frontend.filtfilt_cascade(data_tc, filters, remove_dc=False, axis=0)


def zscore_channels(signals):
    """z-score of mne_bridge.py:215-218 (host side; O(n), after trimming)."""
    stds = np.std(signals, axis=1, keepdims=True)
    stds[stds == 0] = 1.0
    return (signals - np.mean(signals, axis=1, keepdims=True)) / stds
