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
    x = np.asarray(signal, dtype=np.float64)
    if x.ndim == 1:
        return frontend.downsample_dev(torch.from_numpy(np.ascontiguousarray(x[None, :])).cuda(), down).cpu().numpy()[0]
    # resample_poly's default axis is 0 (the reference passes no axis): every trailing index is its own signal along axis 0
    moved = np.ascontiguousarray(np.moveaxis(x, 0, -1).reshape(-1, x.shape[0]))
    y = frontend.downsample_dev(torch.from_numpy(moved).cuda(), down).cpu().numpy()
    return np.moveaxis(y.reshape(x.shape[1:] + (y.shape[-1],)), -1, 0)


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
        # per-window order search (:586-587) for ALL windows in one batched call: one LWR recursion per window to
        # max_model_order gives every order's residual covariance; ln det + penalty + argmin on the device
        T = signals.shape[1]
        st_d = torch.from_numpy(np.asarray(starts, dtype=np.int64)).cuda()
        _, popt, status = mtmvar.batched_mvar_criterion(x, st_d, T, len(starts), m, W, max_model_order, crit_type)
        mtmvar._raise_if_singular(status, "mvar_criterion")
        orders = [int(v) for v in popt.cpu().numpy()]
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


# --------------------------------------------------------------------------- the class callers import
class EEG_IBI_FFDTF_Pipeline:
    """Drop-in for the reference class (src/eeg_alpha_ibi_ffdtf.py:29): same constructor keywords and attributes
    (:96-125), same method names and signatures for the numeric stages -- ``_alpha_bandpass_filter`` (:271),
    ``_compute_asymmetry`` (:314), ``_downsample_signal`` (:368), ``_crop_signal`` (:409), ``_create_windows`` (:451),
    ``_compute_ffDTF`` (:521) -- so call sites written against the reference need only the import changed.

    What differs, on purpose: the window loop of ``run_pipeline`` (:741-755) is available as ONE batched GPU call
    (``compute_windows``); plotting arguments are accepted and ignored (no matplotlib on this path); file discovery and
    NetCDF loading (:128-268) are thin host code that needs ``xarray`` and is only touched by ``run_pipeline``.
    ``cleaned_signals_folder=None`` skips the directory scan, for callers that bring their own arrays
    (``process_dyad``)."""

    def __init__(self, cleaned_signals_folder=None, output_ffDTF_folder=None, target_events=(), smoke_test=False, smoke_dyads_n=1,
                 left_frontal_eeg_channel="F3", right_frontal_eeg_channel="F4", fs_downsampled=8.0, n_windows=3, window_size=None,
                 ar_p=5, plot_global_enabled=True, save_global_enabled=True, plot_windowed_enabled=True, save_windowed_enabled=True):
        from pathlib import Path
        self.cleaned_signals_folder = Path(cleaned_signals_folder) if cleaned_signals_folder is not None else None
        self.output_ffDTF_folder = Path(output_ffDTF_folder) if output_ffDTF_folder is not None else None
        self.target_events = list(target_events)
        self.smoke_test = smoke_test
        self.smoke_dyads_n = smoke_dyads_n
        self.left_chan = left_frontal_eeg_channel
        self.right_chan = right_frontal_eeg_channel
        self.fs_ds = float(fs_downsampled)
        self.freq_min = 1.0
        self.freq_step = 0.1
        self.freq_max = self.fs_ds / 2.0 - self.freq_step
        self.n_windows = n_windows
        self.window_size = window_size
        self.ar_p = ar_p
        self.plot_global_enabled = plot_global_enabled
        self.save_global_enabled = save_global_enabled
        self.plot_windowed_enabled = plot_windowed_enabled
        self.save_windowed_enabled = save_windowed_enabled
        self.eeg_files = []
        self.ibi_files = []
        self.dyads_to_process = []
        if self.cleaned_signals_folder is not None:
            self._prepare_file_lists()

    # ---- file handling (host only; mirrors :128-268 in behaviour, not on the GPU path)
    @staticmethod
    def _dyad_of(path):
        parts = path.stem.split("_")
        return "_".join(parts[:2]) if len(parts) >= 2 else path.stem

    def _prepare_file_lists(self):
        found = {}
        dyads = set()
        for kind in ("EEG", "IBI"):
            folder = self.cleaned_signals_folder / kind
            files = sorted(q for q in folder.rglob("*.nc") if kind in q.name and any(ev in q.name for ev in self.target_events))
            if not files:
                raise FileNotFoundError(f"No {kind} files found for events {self.target_events} under: {folder}")
            found[kind] = files
            dyads.update(self._dyad_of(q) for q in files)
        every = sorted(dyads)
        self.dyads_to_process = every[: self.smoke_dyads_n] if self.smoke_test else every
        self.eeg_files = [q for q in found["EEG"] if self._dyad_of(q) in self.dyads_to_process]
        self.ibi_files = [q for q in found["IBI"] if self._dyad_of(q) in self.dyads_to_process]

    def _find_file(self, file_list, dyad, film, role):
        hits = [f for f in file_list if dyad in f.name and f"_{film}" in f.name and f"_{role}_" in f.name]
        if not hits:
            return None, False
        if len(hits) > 1:
            raise ValueError(f"Found multiple files for dyad: {dyad}, film: {film}, role: {role} -> {hits}")
        return hits[0], True

    def _load_eeg_and_ibi(self, eeg_file, ibi_file, role):
        from .export import open_dataarray      # xarray where it exists, the classic-NetCDF reader of export.py otherwise
        with open_dataarray(eeg_file) as da:
            eeg = da.values.T.copy()
            time_s = da.coords["time"].values.copy()
            names = [str(c) for c in da.coords["channel"].values.tolist()]
            duration = float(da.attrs["event_duration_s"])
            raw_fs = da.attrs.get("sampling_freq") or da.attrs.get("sfreq")
            fs_eeg = float(raw_fs) if raw_fs is not None else 128.0
        with open_dataarray(ibi_file) as da:
            ibi = da.values.T.copy()
            raw_fs = da.attrs.get("sampling_freq") or da.attrs.get("sfreq")
            fs_ibi = float(raw_fs) if raw_fs is not None else fs_eeg
        return time_s, eeg, fs_eeg, names, ibi, fs_ibi, duration

    # ---- numeric stages: the reference's method names, delegating to the module functions above
    def _alpha_bandpass_filter(self, data, fs, lowcut=8, highcut=12, order=4, axis=-1):
        return alpha_bandpass_filter(data, fs, lowcut, highcut, order, axis)

    def _compute_asymmetry(self, filtered_eeg, channel_names, metric='amp'):
        return compute_asymmetry(filtered_eeg, channel_names, self.left_chan, self.right_chan, metric)

    def _downsample_signal(self, signal, fs=128, fs_new=8):
        return downsample_signal(signal, fs, fs_new)

    def _crop_signal(self, signal, fs, drop_front_sec=10, keep_duration_sec=60):
        return crop_signal(signal, fs, drop_front_sec, keep_duration_sec)

    def _create_windows(self, signals, n_windows=3, window_size=None):
        return create_windows(signals, n_windows, window_size)

    def _freqs(self):
        return np.arange(self.freq_min, self.freq_max + self.freq_step, self.freq_step)      # :584

    def _compute_ffDTF(self, dyad, signals, chan_names, fs, max_model_order=20, crit_type="AIC", plot=True, save_plot=False,
                       save_path=None, fig_name=None):
        """(ff_dtf, spectra, p_opt) of ONE segment, as the reference returns them (:586-634); figures are not drawn."""
        signals = np.asarray(signals, dtype=np.float64)
        res = compute_ffdtf_windows(signals, fs, 1, signals.shape[1], ar_p=self.ar_p, freq_min=self.freq_min, freq_max=self.freq_max,
                                    freq_step=self.freq_step, with_spectra=True, max_model_order=max_model_order, crit_type=crit_type)
        return res["ff_dtf_windowed"][0], res["spectra_windowed"][0], res["p_opt_w"][0]

    def compute_windows(self, signals, fs=None, max_model_order=20, crit_type="AIC"):
        """The window loop of ``run_pipeline`` (:741-755) as one batched call: lists (ff_dtf, spectra, p_opt) per window."""
        res = compute_ffdtf_windows(np.asarray(signals, dtype=np.float64), self.fs_ds if fs is None else fs, self.n_windows, self.window_size,
                                    ar_p=self.ar_p, freq_min=self.freq_min, freq_max=self.freq_max, freq_step=self.freq_step,
                                    with_spectra=True, max_model_order=max_model_order, crit_type=crit_type)
        return list(res["ff_dtf_windowed"]), list(res["spectra_windowed"]), list(res["p_opt_w"])

    def process_dyad(self, dyad, film, eeg_ch, ibi_ch, eeg_cg, ibi_cg, fs_eeg, fs_ibi, channel_names):
        """Everything ``run_pipeline`` does for one (dyad, film) after loading (:693-790): returns the ``result`` dict it builds."""
        from datetime import datetime
        sig = preprocess_dyad(eeg_ch, eeg_cg, channel_names, ibi_ch, ibi_cg, fs_eeg, fs_ibi, self.fs_ds, self.left_chan, self.right_chan)
        names = ["faa_ch", "ibi_ch", "faa_cg", "ibi_cg"]
        if self.ar_p is not None:
            _, _, suggested = mtmvar.mvar_criterion(sig, 20, "AIC", plot=False)
            print(f" [INFO] AIC suggested p={suggested} for global signal. Forcing fixed p={self.ar_p}.")
        ff_w, sp_w, p_w = self.compute_windows(sig)
        ff_g, sp_g, p_g = self._compute_ffDTF(dyad, sig, names, self.fs_ds, plot=False)
        return {"mvar": {"ff_dtf_global": ff_g, "spectra_global": sp_g, "ff_dtf_windowed": ff_w, "spectra_windowed": sp_w,
                         "p_opt_g": p_g, "p_opt_w": p_w},
                "meta": {"dyad": dyad, "film": film, "fs": self.fs_ds, "fs_original": fs_eeg, "chan_names": names,
                         "faa_chan_names": (self.left_chan, self.right_chan),
                         "windowing": {"n_windows": self.n_windows, "window_size": self.window_size},
                         "computed_at": datetime.now().isoformat()}}

    def _save_single_result(self, dyad, film, result):
        """Same ``.npz`` keys as the reference writes (:637-658)."""
        import json
        from pathlib import Path
        dyad_dir = Path(self.output_ffDTF_folder) / dyad
        dyad_dir.mkdir(parents=True, exist_ok=True)
        file_path = dyad_dir / f"{dyad}_{film}_ffDTF.npz"
        np.savez_compressed(file_path, ff_dtf_global=result["mvar"]["ff_dtf_global"], spectra_global=result["mvar"]["spectra_global"],
                            ff_dtf_windowed=np.array(result["mvar"]["ff_dtf_windowed"]), spectra_windowed=np.array(result["mvar"]["spectra_windowed"]),
                            p_opt_g=result["mvar"]["p_opt_g"], p_opt_w=result["mvar"]["p_opt_w"], meta=json.dumps(result["meta"]))
        print(f"[SAVED] {dyad} | {film} --> {file_path}\n")
        return file_path

    def run_pipeline(self):
        if not self.dyads_to_process:
            raise RuntimeError("No loaded dyads. Check the files.")
        for dyad in self.dyads_to_process:
            for film in self.target_events:
                print(f"--- Processing dyad: {dyad} | Film: {film} ---")
                picks = {(kind, role): self._find_file(files, dyad, film, role)
                         for kind, files in (("EEG", self.eeg_files), ("IBI", self.ibi_files)) for role in ("ch", "cg")}
                missing = [f"{kind} ({role})" for (kind, role), (_, ok) in picks.items() if not ok]
                if missing:
                    print(f" [SKIP] Missing files: {', '.join(missing)} -> Skipping {film}")
                    continue
                _, eeg_ch, fs_eeg, names, ibi_ch, fs_ibi, _ = self._load_eeg_and_ibi(picks[("EEG", "ch")][0], picks[("IBI", "ch")][0], role="Child")
                _, eeg_cg, fs_eeg, names, ibi_cg, fs_ibi, _ = self._load_eeg_and_ibi(picks[("EEG", "cg")][0], picks[("IBI", "cg")][0], role="Care Giver")
                result = self.process_dyad(dyad, film, eeg_ch, ibi_ch, eeg_cg, ibi_cg, fs_eeg, fs_ibi, names)
                self._save_single_result(dyad, film, result)
