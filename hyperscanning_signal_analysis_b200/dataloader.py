"""Drop-in for the filter functions of the reference's ``src/dataloader.py``:
``_design_eeg_filters`` (dataloader.py:662-764) and ``_apply_filters`` (dataloader.py:767-814)."""
from __future__ import annotations

import numpy as np
from scipy.signal import butter, firwin, iirnotch

from . import frontend


def _design_eeg_filters(multimodal_data, lowcut, highcut, notch_freq=50, notch_q=30, filter_type="fir", plot_flag=False):
    """Same design calls, same ``eeg_filtration`` bookkeeping and same return tuple as the reference
    (dataloader.py:687-764).  ``plot_flag`` is accepted and ignored (no matplotlib on this path)."""
    b_notch, a_notch = iirnotch(notch_freq, notch_q, fs=multimodal_data.fs)
    if filter_type == "fir":
        numtaps_low = 201
        b_low = firwin(numtaps_low, highcut, fs=multimodal_data.fs, pass_zero="lowpass")
        numtaps_high = 3049
        b_high = firwin(numtaps_high, lowcut, fs=multimodal_data.fs, pass_zero="highpass")
        a_low = a_high = 1.0
        low_order = numtaps_low - 1
        high_order = numtaps_high - 1
        f_type = "firwin"
    else:
        butter_order = 2
        b_low, a_low = butter(N=butter_order, Wn=highcut, btype="low", fs=multimodal_data.fs)
        b_high, a_high = butter(N=butter_order, Wn=lowcut, btype="high", fs=multimodal_data.fs)
        low_order = butter_order
        high_order = butter_order
        f_type = "butter"
    filt = multimodal_data.eeg_filtration
    filt.low_pass.update({"type": filter_type, "cut_f": highcut, "order": low_order, "f_type": f_type, "a": a_low, "b": b_low})
    filt.high_pass.update({"type": filter_type, "cut_f": lowcut, "order": high_order, "f_type": f_type, "a": a_high, "b": b_high})
    filt.notch.update({"Q": notch_q, "freq": notch_freq, "a": a_notch, "b": b_notch})
    return (b_notch, a_notch), (b_low, a_low), (b_high, a_high), filter_type


def _apply_filters(multimodal_data, filters, raw_eeg_data, plot_flag=False):
    """Reference ``_apply_filters`` (dataloader.py:767-814): every mapped channel row of ``raw_eeg_data`` gets DC removal
    and either filtfilt notch -> low -> high (``'iir'``, :789-792) or the causal lfilter chain with the FIR delays rolled
    out (any other ``filter_type``, the reference's default ``'fir'``, :793-801), written back IN PLACE (cast to the
    array's dtype like the reference's assignment at :803), ``applied`` flags set (:812-814).
    All channels go through batched GPU calls instead of the per-channel Python loop (:786)."""
    (b_notch, a_notch), (b_low, a_low), (b_high, a_high), filter_type = filters
    print(f"Applying {filter_type} filters to EEG data.")
    rows = [multimodal_data.eeg_channel_mapping[ch] for ch in multimodal_data.eeg_channel_names_all()]
    if rows:
        block = np.ascontiguousarray(raw_eeg_data[rows, :], dtype=np.float64)
        if filter_type == "iir":
            out = frontend.filtfilt_cascade(block, [(b_notch, a_notch), (b_low, a_low), (b_high, a_high)], remove_dc=True)
        elif np.size(a_low) == 1 and np.size(a_high) == 1:
            bl = np.atleast_1d(np.asarray(b_low, dtype=np.float64)) / float(np.ravel(a_low)[0])
            bh = np.atleast_1d(np.asarray(b_high, dtype=np.float64)) / float(np.ravel(a_high)[0])
            import torch
            out = frontend.lfilter_fir_chain_dev(torch.from_numpy(block).cuda(), (b_notch, a_notch), bl, bh, remove_dc=True).cpu().numpy()
        else:
            # any other filter_type with recursive low-/high-pass coefficients (e.g. the Butterworth pair _design_eeg_filters
            # returns for filter_type='butter'): the same causal chain, every stage a zero-state IIR sweep (:794-801)
            import torch
            out = frontend.lfilter_iir_chain_dev(torch.from_numpy(block).cuda(), [(b_notch, a_notch), (b_low, a_low), (b_high, a_high)],
                                                 delay=(np.size(b_low) - 1) // 2 + (np.size(b_high) - 1) // 2, remove_dc=True).cpu().numpy()
        raw_eeg_data[rows, :] = out
    multimodal_data.eeg_filtration.notch["applied"] = True
    multimodal_data.eeg_filtration.low_pass["applied"] = True
    multimodal_data.eeg_filtration.high_pass["applied"] = True
