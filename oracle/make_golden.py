"""Generate tests/golden/*.npz by running the REFERENCE ITSELF on seeded inputs.

Run only in the build container (needs /root/reference; the GPU box has no
reference tree):  ``python oracle/make_golden.py``

What is imported from the reference, unmodified:
  * ``src/mtmvar.py``            (matplotlib stubbed -- numerics never touch it)
  * ``src/dataloader.py``        ``_design_eeg_filters`` / ``_apply_filters``
  * ``src/data_structures.py``   ``MultimodalData._decimate_signals``
  * ``src/eeg_alpha_ibi_ffdtf.py`` ``EEG_IBI_FFDTF_Pipeline._create_windows``
Third-party modules that are absent here (mne, xarray, neurokit2, xmltodict,
matplotlib, ...) are replaced by empty stubs: none of them is touched by the
functions called below.  ``mne_bridge.load_eeg_signals`` needs a NetCDF file
and xarray, so its filter block (mne_bridge.py:158-184) is reproduced with the
same SciPy calls.  ``psd.compute_psd_multitaper`` is a one-line call into mne
(absent) -> no golden vector from the reference; see frontend_oracle header.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Stub(self.__name__ + "." + name)
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):
        return None


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "mne", "mne.time_frequency", "xmltodict", "neurokit2",
                 "xarray", "joblib", "plotly", "plotly.graph_objects", "plotly.subplots", "netCDF4",
                 "autoreject", "mne_icalabel", "specparam", "seaborn"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _Stub(name)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from src import mtmvar, dataloader, data_structures
    try:
        from src import eeg_alpha_ibi_ffdtf
    except Exception as exc:          # pragma: no cover
        print("eeg_alpha_ibi_ffdtf not importable:", exc)
        eeg_alpha_ibi_ffdtf = None
    return mtmvar, dataloader, data_structures, eeg_alpha_ibi_ffdtf


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def gen_pcoh():
    """Partial coherence / dDTF from the reference (mtmvar.py:287-385): the m=4 AR(1) example and two cfg-2 windows
    (m=38, production band) on a 12-bin grid -- the reference takes 38*38*12 determinants of 37x37 minors per window."""
    sys.path.insert(0, ROOT)
    import scipy
    mtmvar = import_reference()[0]
    versions = f"numpy {np.__version__} scipy {scipy.__version__}"
    g4 = np.load(os.path.join(OUT, "mvar_m4.npz"))
    x4, p4, fs4, fr4 = g4["x"], int(g4["p"]), float(g4["fs"]), g4["freqs"]
    S4 = quiet(mtmvar.multivariate_spectra, x4, fr4, fs4, optimal_model_order=p4)
    rec = dict(versions=versions, m4_S=S4, m4_kappa=mtmvar.partial_coherence(S4),
               m4_ddtf=quiet(mtmvar.direct_dtf, x4, fr4, fs4, optimal_model_order=p4))
    g = np.load(os.path.join(OUT, "mvar_cfg2_windows.npz"))
    f12 = np.linspace(0, 128, 12, endpoint=False)
    w0 = g["windows"][0]
    S = quiet(mtmvar.multivariate_spectra, w0, f12, 256.0, optimal_model_order=8)
    rec.update(freqs=f12, w0_S=S, w0_kappa=mtmvar.partial_coherence(S),
               w0_ddtf=quiet(mtmvar.direct_dtf, w0, f12, 256.0, optimal_model_order=8))
    np.savez_compressed(os.path.join(OUT, "mvar_pcoh.npz"), **rec)
    print("mvar_pcoh.npz", os.path.getsize(os.path.join(OUT, "mvar_pcoh.npz")))


def gen_prewindow():
    """Pre-window stage of EEG_IBI_FFDTF_Pipeline, run through the REFERENCE's own methods (eeg_alpha_ibi_ffdtf.py:271-448,
    :693-719) on a synthetic dyad: 19 ch x 75 s @ 128 Hz per participant + IBI series at 128 Hz."""
    sys.path.insert(0, ROOT)
    import scipy
    from hyperscanning_signal_analysis_b200 import synth
    ffd = import_reference()[3]
    versions = f"numpy {np.__version__} scipy {scipy.__version__}"
    pipe = ffd.EEG_IBI_FFDTF_Pipeline.__new__(ffd.EEG_IBI_FFDTF_Pipeline)
    pipe.left_chan, pipe.right_chan, pipe.fs_ds = "F3", "F4", 8.0
    names = ["Fp1", "Fp2", "F7", "F3", "Fz", "F4", "F8", "T3", "C3", "Cz", "C4", "T4", "T5", "P3", "Pz", "P4", "T6", "O1", "O2"]
    fs = 128.0
    n = int(75 * fs) + 37                     # 9637 samples: next_fast_len -> 9680 = 2^4 5 11^2 (radix-11 passes)
    x = synth.dyad_eeg(seed=77, m=38, fs=fs, n_samples=n, line_amp=0.0)
    rng = np.random.default_rng(78)
    t = np.arange(n) / fs
    ibi = np.stack([0.8 + 0.05 * np.sin(2 * np.pi * 0.1 * t + ph) + 0.01 * rng.standard_normal(n) for ph in (0.0, 1.0)])
    rec = dict(versions=versions, fs=fs, names=np.array(names), n=n, eeg_sum=float(np.sum(x)), ibi=ibi)      # eeg: synth.dyad_eeg(seed=77, ...)
    rows = []
    for who, sl in (("ch", slice(0, 19)), ("cg", slice(19, 38))):
        eeg = x[sl]
        filt = pipe._alpha_bandpass_filter(eeg, fs)
        faa = pipe._compute_asymmetry(filt, names, metric="amp")
        faa_p = pipe._compute_asymmetry(filt, names, metric="power")
        faa_ds = pipe._downsample_signal(faa, fs, 8.0)
        ibi_ds = pipe._downsample_signal(ibi[0 if who == "ch" else 1], fs, 8.0)
        rows += [pipe._crop_signal(faa_ds, 8.0, 10, 60), pipe._crop_signal(ibi_ds, 8.0, 10, 60)]
        rec.update({f"filt_{who}": filt[[3, 5, 18]], f"faa_{who}": faa, f"faa_power_{who}": faa_p, f"faa_ds_{who}": faa_ds,
                    f"ibi_ds_{who}": ibi_ds})
    sig = np.vstack(rows)
    sig = (sig - np.mean(sig, axis=1, keepdims=True)) / np.std(sig, axis=1, keepdims=True)      # eeg_alpha_ibi_ffdtf.py:719
    rec["signals_to_ffDTF"] = sig
    # short odd-length case: N = next_fast_len(1001) = 1008 = 2^4 3^2 7; hilbert with default N (prime factor 13: 1001 = 7 11 13)
    from scipy.signal import hilbert
    xs = x[:3, :1001]
    rec["short_env_fast"] = np.abs(hilbert(xs, N=1008, axis=-1)[:, :1001])
    rec["short_env_n"] = np.abs(hilbert(xs, axis=-1))
    np.savez_compressed(os.path.join(OUT, "prewindow.npz"), **rec)
    print("prewindow.npz", os.path.getsize(os.path.join(OUT, "prewindow.npz")))


def gen_round2():
    """Round-2 additions, all from the REFERENCE's own code: count_corr(iwhat=2) (mtmvar.py:60-63), the causal branch of
    _apply_filters with recursive low/high-pass coefficients (dataloader.py:793-801, filter_type='butter'), per-window order
    search (mvar_criterion, mtmvar.py:551-601) and EEG_IBI_FFDTF_Pipeline._compute_ffDTF (eeg_alpha_ibi_ffdtf.py:521-634)
    called on an instance, with a fixed order and with ar_p=None."""
    sys.path.insert(0, ROOT)
    import scipy
    from hyperscanning_signal_analysis_b200 import synth
    mtmvar, dataloader, ds, ffd = import_reference()
    versions = f"numpy {np.__version__} scipy {scipy.__version__}"
    rec = dict(versions=versions)
    g4 = np.load(os.path.join(OUT, "mvar_m4.npz"))
    rl, rr, r0 = mtmvar.count_corr(g4["x"][:, :, None], int(g4["p"]), 2)
    rec.update(m4_iwhat2_left=rl, m4_iwhat2_right=rr, m4_iwhat2_zero=r0)
    gt = np.load(os.path.join(OUT, "mvar_trials.npz"))
    rl, rr, r0 = mtmvar.count_corr(gt["x"], int(gt["p"]), 2)
    rec.update(tr_iwhat2_left=rl, tr_iwhat2_right=rr, tr_iwhat2_zero=r0)
    # causal chain with Butterworth low/high-pass: any filter_type other than 'iir' takes the lfilter branch
    raw = synth.dyad_eeg(seed=synth.BASE_SEED + 21, m=38, fs=256.0, n_samples=5000)
    md = ds.MultimodalData()
    md.fs = 256.0
    names_ch = [f"c{i}" for i in range(19)]
    names_cg = [f"c{i}_cg" for i in range(19)]
    md.eeg_channel_names_ch = names_ch
    md.eeg_channel_names_cg = names_cg
    md.eeg_channel_mapping = {nm: i for i, nm in enumerate(names_ch + names_cg)}
    fb = dataloader._design_eeg_filters(md, lowcut=1.0, highcut=40.0, filter_type="butter")
    assert fb[3] == "butter" and np.size(fb[1][1]) == 3
    out = raw.copy()
    quiet(dataloader._apply_filters, md, fb, out)
    sel = [0, 19, 37]
    rec.update(butter_raw=raw[sel], butter_out=out[sel])
    # the pipeline's real shape: 4 signals x 480 samples @ 8 Hz (prewindow.npz), 3 windows of 160, and 5 windows of 200
    sig = np.load(os.path.join(OUT, "prewindow.npz"))["signals_to_ffDTF"]
    pipe = ffd.EEG_IBI_FFDTF_Pipeline.__new__(ffd.EEG_IBI_FFDTF_Pipeline)
    pipe.fs_ds = 8.0
    pipe.freq_min, pipe.freq_step = 1.0, 0.1
    pipe.freq_max = pipe.fs_ds / 2.0 - pipe.freq_step
    names = ["faa_ch", "ibi_ch", "faa_cg", "ibi_cg"]
    for tag, nw, ws in (("w3", 3, None), ("w5", 5, 200)):
        wins = pipe._create_windows(sig, nw, ws)
        for ar_p, ptag in ((5, "p5"), (None, "auto")):
            pipe.ar_p = ar_p
            ff, sp, po = [], [], []
            for w in wins:
                a, b, c = pipe._compute_ffDTF("D", w, names, 8.0, plot=False, save_plot=False)
                ff.append(a); sp.append(b); po.append(c)
            rec[f"{tag}_{ptag}_ffdtf"] = np.stack(ff)
            rec[f"{tag}_{ptag}_spectra"] = np.stack(sp)
            rec[f"{tag}_{ptag}_popt"] = np.array(po)
        for crit in ("AIC", "HQ", "SC"):
            rec[f"{tag}_crit_{crit}"] = np.stack([mtmvar.mvar_criterion(w, 20, crit, False)[0] for w in wins])
            rec[f"{tag}_popt_{crit}"] = np.array([mtmvar.mvar_criterion(w, 20, crit, False)[2] for w in wins])
    pipe.ar_p = 5
    a, b, c = pipe._compute_ffDTF("D", sig, names, 8.0, plot=False, save_plot=False)
    rec.update(global_p5_ffdtf=a, global_p5_spectra=b, freqs=np.arange(pipe.freq_min, pipe.freq_max + pipe.freq_step, pipe.freq_step))
    # m = 38 windows (cfg2 fixture): criterion up to order 12
    w38 = np.load(os.path.join(OUT, "mvar_cfg2_windows.npz"))["windows"]
    rec["w38_crit_AIC"] = np.stack([mtmvar.mvar_criterion(w, 12, "AIC", False)[0] for w in w38])
    np.savez_compressed(os.path.join(OUT, "round2.npz"), **rec)
    print("round2.npz", os.path.getsize(os.path.join(OUT, "round2.npz")))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "round2":
        return gen_round2()
    if len(sys.argv) > 1 and sys.argv[1] == "pcoh":
        return gen_pcoh()
    if len(sys.argv) > 1 and sys.argv[1] == "prewindow":
        return gen_prewindow()
    sys.path.insert(0, ROOT)
    from hyperscanning_signal_analysis_b200 import synth
    from scipy import signal
    import scipy

    mtmvar, dataloader, ds, ffd = import_reference()
    os.makedirs(OUT, exist_ok=True)
    versions = f"numpy {np.__version__} scipy {scipy.__version__}"

    # ---------------------------------------------------------------- MVAR, small (the reference's real use: m=4, fs=8, p=5)
    rng = np.random.default_rng(7)
    m, n, p, fs = 4, 480, 5, 8.0
    x = np.zeros((m, n))
    e = rng.standard_normal((m, n))
    a1 = np.array([[0.5, 0.2, 0, 0], [0, 0.4, 0.3, 0], [0, 0, -0.3, 0.2], [0.25, 0, 0, 0.35]])
    for t in range(1, n):
        x[:, t] = a1 @ x[:, t - 1] + e[:, t]
    freqs = np.arange(0.1, 4.0 + 0.13, 0.13)
    A, V = mtmvar.ar_coeff(x, p)
    H, Af = mtmvar.mvar_transfer_function(A, freqs, fs)
    rl, rr, r0 = mtmvar.count_corr(x[:, :, None], p, 1)
    np.savez_compressed(
        os.path.join(OUT, "mvar_m4.npz"), versions=versions, x=x, p=p, fs=fs, freqs=freqs,
        r_left=rl, r_right=rr, r_zero=r0, A=A, V=V, H=H, Af=Af,
        dtf=quiet(mtmvar.dtf_multivariate, x, freqs, fs, optimal_model_order=p),
        ffdtf=quiet(mtmvar.full_freq_dtf, x, freqs, fs, optimal_model_order=p),
        S=quiet(mtmvar.multivariate_spectra, x, freqs, fs, optimal_model_order=p),
        gpdc=quiet(mtmvar.gen_partial_directed_coherence, x, freqs, fs, optimal_model_order=p),
        crit_aic=mtmvar.mvar_criterion(x, 8, "AIC", False)[0],
        crit_hq=mtmvar.mvar_criterion(x, 8, "HQ", False)[0],
        crit_sc=mtmvar.mvar_criterion(x, 8, "SC", False)[0],
        popt=np.array([mtmvar.mvar_criterion(x, 8, c, False)[2] for c in ("AIC", "HQ", "SC")]),
    )

    # ---------------------------------------------------------------- MVAR, multi-trial (3-D ar_coeff path)
    rng = np.random.default_rng(11)
    xt = rng.standard_normal((7, 96, 5))
    xt[:, 1:, :] += 0.6 * xt[:, :-1, :]
    A, V = mtmvar.ar_coeff(xt, 3)
    rl, rr, r0 = mtmvar.count_corr(xt, 3, 1)
    np.savez_compressed(os.path.join(OUT, "mvar_trials.npz"), versions=versions, x=xt, p=3, A=A, V=V,
                        r_left=rl, r_right=rr, r_zero=r0)

    # ---------------------------------------------------------------- filters (cfg1 generator, 4 channels x 3000 samples)
    raw = synth.dyad_eeg(seed=synth.BASE_SEED, m=38, fs=256.0, n_samples=3000)
    md = ds.MultimodalData()
    md.fs = 256.0
    names_ch = [f"c{i}" for i in range(19)]
    names_cg = [f"c{i}_cg" for i in range(19)]
    md.eeg_channel_names_ch = names_ch
    md.eeg_channel_names_cg = names_cg
    md.eeg_channel_mapping = {nm: i for i, nm in enumerate(names_ch + names_cg)}
    filters = dataloader._design_eeg_filters(md, lowcut=1.0, highcut=40.0, filter_type="iir")
    filt = raw.copy()
    quiet(dataloader._apply_filters, md, filters, filt)
    sel = [0, 7, 19, 37]
    (bn, an), (bl, al), (bh, ah), _ = filters
    np.savez_compressed(os.path.join(OUT, "filters_iir.npz"), versions=versions, fs=256.0, raw=raw[sel], out=filt[sel],
                        b_notch=bn, a_notch=an, b_low=bl, a_low=al, b_high=bh, a_high=ah,
                        applied=np.array([md.eeg_filtration.notch["applied"], md.eeg_filtration.low_pass["applied"],
                                          md.eeg_filtration.high_pass["applied"]]))
    # default branch of the loader: filter_type='fir' (causal lfilter chain + delay roll), 3 channels x 9000 samples
    rawf = synth.dyad_eeg(seed=synth.BASE_SEED + 9, m=38, fs=256.0, n_samples=9000)
    ffir = dataloader._design_eeg_filters(md, lowcut=1.0, highcut=40.0)
    assert ffir[3] == "fir"
    filt_f = rawf.copy()
    quiet(dataloader._apply_filters, md, ffir, filt_f)
    self_ = [0, 19, 37]
    np.savez_compressed(os.path.join(OUT, "filters_fir.npz"), versions=versions, fs=256.0, raw=rawf[self_], out=filt_f[self_],
                        b_low=ffir[1][0], b_high=ffir[2][0])
    # the reference's own unit-test input (tests/test_dataloader.py:181-197): 10 Hz + 60 Hz, fs 256, n 1000
    tt = np.arange(1000) / 256.0
    sig = np.sin(2 * np.pi * 10 * tt) + 0.5 * np.sin(2 * np.pi * 60 * tt) + 3.0
    md2 = ds.MultimodalData()
    md2.fs = 256.0
    md2.eeg_channel_names_ch = ["Fz"]
    md2.eeg_channel_names_cg = []
    md2.eeg_channel_mapping = {"Fz": 0}
    f2 = dataloader._design_eeg_filters(md2, lowcut=1.0, highcut=40.0, filter_type="iir")
    buf = sig[None, :].copy()
    quiet(dataloader._apply_filters, md2, f2, buf)
    np.savez_compressed(os.path.join(OUT, "filters_unit.npz"), versions=versions, fs=256.0, raw=sig[None, :], out=buf)

    # mne_bridge filter block (order-4 Butterworth + notch Q=15), (time, channel) layout, same SciPy calls as mne_bridge.py:158-184
    tc = raw[[1, 20, 30]].T.copy()
    d = tc.copy()
    b, a = signal.butter(4, 1.0 / 128.0, btype="highpass"); d = signal.filtfilt(b, a, d, axis=0)
    b, a = signal.butter(4, 45.0 / 128.0, btype="lowpass"); d = signal.filtfilt(b, a, d, axis=0)
    b, a = signal.iirnotch(50.0, Q=15, fs=256.0); d = signal.filtfilt(b, a, d, axis=0)
    np.savez_compressed(os.path.join(OUT, "filters_bridge.npz"), versions=versions, fs=256.0, low=1.0, high=45.0, raw=tc, out=d)

    # ---------------------------------------------------------------- decimation through the real MultimodalData._decimate_signals
    import pandas as pd
    raw4 = synth.dyad_eeg(seed=synth.BASE_SEED + 4, m=4, fs=1024.0, n_samples=4099, drift=True)
    md3 = ds.MultimodalData()
    md3.fs = 1024.0
    md3.id = "W_000"
    md3.eeg_channel_names_ch = ["Fz", "Cz"]
    md3.eeg_channel_names_cg = ["Fz_cg"]
    md3.eeg_channel_mapping = {"Fz": 0, "Cz": 1, "Fz_cg": 2}
    withnan = raw4[3].copy()
    withnan[:5] = np.nan
    withnan[1000:1040] = np.nan
    md3.data = pd.DataFrame({
        "time": np.arange(4099) / 1024.0, "time_idx": np.arange(4099),
        "EEG_ch_Fz": raw4[0], "EEG_ch_Cz": raw4[1], "EEG_cg_Fz": raw4[2], "IBI_ch": withnan,
        "events": np.array([None] * 4099, dtype=object), "diode": (np.arange(4099) % 7).astype(float),
    })
    dec = quiet(md3._decimate_signals, q=8)
    np.savez_compressed(os.path.join(OUT, "decimate_q8.npz"), versions=versions, q=8, fs_in=1024.0, fs_out=dec.fs,
                        raw=np.stack([raw4[0], raw4[1], raw4[2], withnan]),
                        out=np.stack([dec.data[c].values for c in ("EEG_ch_Fz", "EEG_ch_Cz", "EEG_cg_Fz", "IBI_ch")]),
                        time=dec.data["time"].values, diode=dec.data["diode"].values)
    # the reference unit-test shape (tests/test_data_structures.py:305-339): n=1000, q=4
    md4 = ds.MultimodalData()
    md4.fs = 256.0
    xs = np.sin(2 * np.pi * 5 * np.arange(1000) / 256.0)
    md4.data = pd.DataFrame({"time": np.arange(1000) / 256.0, "time_idx": np.arange(1000), "EEG_ch_Fz": xs})
    dec4 = quiet(md4._decimate_signals, q=4)
    np.savez_compressed(os.path.join(OUT, "decimate_q4.npz"), versions=versions, q=4, raw=xs[None], out=dec4.data["EEG_ch_Fz"].values[None],
                        fs_out=dec4.fs)

    # ---------------------------------------------------------------- MVAR at the BASELINE shape: cfg1 pipeline, two 2-s windows, F=32 grid
    # Band 1-64 Hz = the reference's production values (scripts/export_dyade_to_ncdf_by_task_batch.py:27-31): cond(G) ~ 3e7.
    # A second fixture with the 40 Hz low-pass is kept as an ill-conditioned stress case (cond ~ 1.5e9, where two runs
    # of the SAME LU solver on inputs differing by 1 ulp already differ by ~1e-8).
    x1 = synth.cfg1_raw()
    fgrid = np.linspace(0, 128, 32, endpoint=False)
    f16 = np.linspace(0, 128, 16, endpoint=False)
    starts = np.array([1024, 9000])
    for tag, highcut in (("", 64.0), ("_lp40", 40.0)):
        md5 = ds.MultimodalData()
        md5.fs = 256.0
        md5.eeg_channel_names_ch = names_ch
        md5.eeg_channel_names_cg = names_cg
        md5.eeg_channel_mapping = {nm: i for i, nm in enumerate(names_ch + names_cg)}
        f5 = dataloader._design_eeg_filters(md5, lowcut=1.0, highcut=highcut, filter_type="iir")
        y1 = x1.copy()
        quiet(dataloader._apply_filters, md5, f5, y1)
        wins = np.stack([y1[:, s:s + 512] for s in starts])
        As, Vs, ffs, conds = [], [], [], []
        for w in wins:
            A, V = mtmvar.ar_coeff(w, 8)
            As.append(A); Vs.append(V)
            ffs.append(quiet(mtmvar.full_freq_dtf, w, fgrid, 256.0, optimal_model_order=8))
            conds.append(np.linalg.cond(mtmvar.count_corr(w[:, :, None], 8, 1)[0]))
        np.savez_compressed(os.path.join(OUT, f"mvar_cfg2_windows{tag}.npz"), versions=versions, windows=wins, starts=starts, p=8,
                            fs=256.0, highcut=highcut, freqs=fgrid, A=np.stack(As), V=np.stack(Vs), ffdtf=np.stack(ffs),
                            cond=np.array(conds))
        if tag == "":
            # whole 60-s segment (cfg1), F=16 grid -> small file; the input is regenerated from synth + filters in the test
            A, V = mtmvar.ar_coeff(y1, 8)
            np.savez_compressed(os.path.join(OUT, "mvar_cfg1.npz"), versions=versions, p=8, fs=256.0, highcut=highcut, freqs=f16, A=A, V=V,
                                filtered_rows=y1[[0, 18, 19, 37]], filtered_sum=float(np.sum(y1)), raw_sum=float(np.sum(x1)),
                                ffdtf=quiet(mtmvar.full_freq_dtf, y1, f16, 256.0, optimal_model_order=8),
                                S=quiet(mtmvar.multivariate_spectra, y1, f16, 256.0, optimal_model_order=8),
                                cond=np.linalg.cond(mtmvar.count_corr(y1[:, :, None], 8, 1)[0]))

    # ---------------------------------------------------------------- window starts from the real _create_windows
    if ffd is not None:
        pipe = ffd.EEG_IBI_FFDTF_Pipeline.__new__(ffd.EEG_IBI_FFDTF_Pipeline)
        cases = [(153600, 599, 512), (15360, 59, 512), (999, 3, None), (999, 3, 400), (777, 10, 100), (512, 1, 512)]
        rec = {}
        for (T, nw, ws) in cases:
            sig = np.arange(T, dtype=float)[None, :]
            wl = pipe._create_windows(sig, n_windows=nw, window_size=ws)
            rec[f"T{T}_n{nw}_w{ws}"] = np.array([int(w[0, 0]) for w in wl] + [wl[0].shape[1]])
        np.savez_compressed(os.path.join(OUT, "window_starts.npz"), versions=versions, **rec)

    gen_pcoh()
    gen_prewindow()
    gen_round2()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
