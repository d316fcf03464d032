"""Parity of the GPU front end (zero-phase IIR cascade, FIR decimation) against the reference's golden
vectors and SciPy.  Tolerance: 1e-9 relative, norm-wise (north_star: filtered signals)."""
import contextlib
import io

import numpy as np
import pytest
from scipy import signal

from conftest import golden, relerr, TOL_SIGNAL
from oracle import frontend_oracle as fo

pytestmark = pytest.mark.gpu


def _md(fs, n_ch):
    from hyperscanning_signal_analysis_b200.data_structures import MultimodalData
    md = MultimodalData()
    md.fs = fs
    md.eeg_channel_names_ch = [f"c{i}" for i in range(n_ch)]
    md.eeg_channel_names_cg = []
    md.eeg_channel_mapping = {f"c{i}": i for i in range(n_ch)}
    return md


def test_apply_filters_against_reference_golden(capsys):
    from hyperscanning_signal_analysis_b200 import dataloader
    g = golden("filters_iir.npz")
    md = _md(float(g["fs"]), 4)
    filters = dataloader._design_eeg_filters(md, lowcut=1.0, highcut=40.0, filter_type="iir")
    for got, key in zip((filters[0][0], filters[0][1], filters[1][0], filters[1][1], filters[2][0], filters[2][1]),
                        ("b_notch", "a_notch", "b_low", "a_low", "b_high", "a_high")):
        np.testing.assert_array_equal(got, g[key])
    assert filters[3] == "iir" and md.eeg_filtration.low_pass["cut_f"] == 40.0 and md.eeg_filtration.notch["Q"] == 30
    assert md.eeg_filtration.notch["applied"] is False
    buf = g["raw"].copy()
    ret = dataloader._apply_filters(md, filters, buf)
    assert ret is None and "Applying iir filters to EEG data." in capsys.readouterr().out
    assert relerr(buf, g["out"]) < TOL_SIGNAL
    assert md.eeg_filtration.notch["applied"] and md.eeg_filtration.low_pass["applied"] and md.eeg_filtration.high_pass["applied"]
    # the reference's own unit-test input (tests/test_dataloader.py:181-197)
    u = golden("filters_unit.npz")
    md1 = _md(256.0, 1)
    f1 = dataloader._design_eeg_filters(md1, 1.0, 40.0, filter_type="iir")
    b1 = u["raw"].copy()
    with contextlib.redirect_stdout(io.StringIO()):
        dataloader._apply_filters(md1, f1, b1)
    assert not np.array_equal(b1, u["raw"])            # what the reference test asserts
    assert relerr(b1, u["out"]) < TOL_SIGNAL           # what it should have asserted
    # default design branch stays 'fir' like the reference (tests/test_dataloader.py:129-173)
    ffir = dataloader._design_eeg_filters(md1, 1.0, 40.0)
    assert ffir[3] == "fir" and len(ffir[1][0]) == 201 and len(ffir[2][0]) == 3049


def test_fir_branch_against_reference_golden():
    """filter_type='fir' (the reference's default): causal lfilter notch + 201-tap low-pass + 3049-tap high-pass, the FIR
    delays rolled out, tail zeroed (dataloader.py:793-801)."""
    from hyperscanning_signal_analysis_b200 import dataloader
    g = golden("filters_fir.npz")
    md = _md(float(g["fs"]), 3)
    filters = dataloader._design_eeg_filters(md, lowcut=1.0, highcut=40.0)
    np.testing.assert_array_equal(filters[1][0], g["b_low"])
    np.testing.assert_array_equal(filters[2][0], g["b_high"])
    buf = g["raw"].copy()
    with contextlib.redirect_stdout(io.StringIO()) as out:
        dataloader._apply_filters(md, filters, buf)
    assert "Applying fir filters to EEG data." in out.getvalue()
    assert relerr(buf, g["out"]) < TOL_SIGNAL
    delay = 100 + 1524
    assert np.all(buf[:, -delay:] == 0.0)
    assert md.eeg_filtration.high_pass["applied"] and md.eeg_filtration.high_pass["f_type"] == "firwin"
    # a signal shorter than the combined delay comes back all zero, like np.roll + s[-delay:] = 0 does
    short = g["raw"][:, :1000].copy()
    with contextlib.redirect_stdout(io.StringIO()):
        dataloader._apply_filters(md, filters, short)
    assert np.all(short == 0.0)
    # long signal: the notch takes the tiled scan path; compare with the oracle
    rng = np.random.default_rng(5)
    x = 20.0 * rng.standard_normal((3, 30000)) + 40.0
    ref = fo.apply_filters_fir(x, fo.design_eeg_filters(256.0, 1.0, 40.0, filter_type="fir"))
    with contextlib.redirect_stdout(io.StringIO()):
        dataloader._apply_filters(md, filters, x)
    assert relerr(x, ref) < TOL_SIGNAL


def test_float32_storage_quirk():
    """Raw SVAROG data is float32 and the filtered row is written back into it (dataloader.py:620-630, 803)."""
    from hyperscanning_signal_analysis_b200 import dataloader
    g = golden("filters_iir.npz")
    md = _md(256.0, 4)
    filters = dataloader._design_eeg_filters(md, 1.0, 40.0, filter_type="iir")
    buf32 = g["raw"].astype(np.float32)
    with contextlib.redirect_stdout(io.StringIO()):
        dataloader._apply_filters(md, filters, buf32)
    ref = fo.apply_filters_iir(g["raw"].astype(np.float32).astype(np.float64), filters).astype(np.float32)
    assert buf32.dtype == np.float32 and relerr(buf32, ref) < 1e-6


def test_bridge_filters_against_golden():
    from hyperscanning_signal_analysis_b200 import mne_bridge
    g = golden("filters_bridge.npz")
    out = mne_bridge.filter_time_channel(g["raw"], float(g["fs"]), float(g["low"]), float(g["high"]))
    assert out.shape == g["raw"].shape and relerr(out, g["out"]) < TOL_SIGNAL
    with pytest.raises(ValueError):
        mne_bridge.filter_time_channel(g["raw"], 256.0, low_cutoff_hz=500.0)
    # no notch below 100 Hz sampling (mne_bridge.py:181-182)
    low = mne_bridge.filter_time_channel(g["raw"], 64.0, 1.0, None)
    b, a = signal.butter(4, 1.0 / 32.0, btype="highpass")
    assert relerr(low, signal.filtfilt(b, a, g["raw"], axis=0)) < TOL_SIGNAL


@pytest.mark.parametrize("n", [10, 28, 513, 1024, 5000, 40000])
def test_filtfilt_lengths_and_edges(n):
    from hyperscanning_signal_analysis_b200 import frontend
    rng = np.random.default_rng(n)
    x = rng.standard_normal((3, n)).cumsum(axis=1) + 5.0
    filt = fo.design_eeg_filters(256.0, 1.0, 64.0)[:3]
    if n <= 9:
        with pytest.raises(ValueError):
            frontend.filtfilt_cascade(x, filt)
        return
    got = frontend.filtfilt_cascade(x, filt, remove_dc=True)
    ref = fo.apply_filters_iir(x, fo.design_eeg_filters(256.0, 1.0, 64.0))
    assert relerr(got, ref) < TOL_SIGNAL


def test_filtfilt_padlen_error_and_constant():
    from hyperscanning_signal_analysis_b200 import frontend
    b, a = signal.butter(2, 0.2)
    with pytest.raises(ValueError, match="padlen"):
        frontend.filtfilt_cascade(np.ones((1, 9)), [(b, a)])
    c = np.full((2, 300), 3.25)
    np.testing.assert_allclose(frontend.filtfilt_cascade(c, [(b, a)]), c, rtol=1e-12)   # odd extension + zi keep constants
    # slow poles (|p| = 0.9957 at 1024 Hz) over a long record: the chunk carries matter here
    x = np.random.default_rng(1).standard_normal((2, 200000)) + 100.0
    hb, ha = signal.butter(2, 1.0, "high", fs=1024.0)
    assert relerr(frontend.filtfilt_cascade(x, [(hb, ha)]), signal.filtfilt(hb, ha, x, axis=1)) < TOL_SIGNAL


def test_decimate_against_reference_golden():
    from hyperscanning_signal_analysis_b200 import frontend
    g = golden("decimate_q8.npz")
    got = frontend.decimate(g["raw"][:3], 8)
    assert got.shape == g["out"][:3].shape and relerr(got, g["out"][:3]) < TOL_SIGNAL
    g4 = golden("decimate_q4.npz")
    got4 = frontend.decimate(g4["raw"][0], 4)
    assert got4.shape == (250,) and relerr(got4, g4["out"][0]) < TOL_SIGNAL
    for n, q in ((1, 2), (7, 8), (161, 8), (1000, 3), (4097, 5)):
        x = np.random.default_rng(n).standard_normal(n)
        assert relerr(frontend.decimate(x, q), signal.decimate(x, q, ftype="fir", zero_phase=True)) < TOL_SIGNAL


def test_decimate_signals_object_semantics():
    import pandas as pd
    from hyperscanning_signal_analysis_b200.data_structures import MultimodalData
    g = golden("decimate_q8.npz")
    md = MultimodalData()
    md.fs = float(g["fs_in"])
    md.id = "W_000"
    md.eeg_channel_names_ch = ["Fz", "Cz"]
    md.eeg_channel_names_cg = ["Fz_cg"]
    n = g["raw"].shape[1]
    md.data = pd.DataFrame({"time": np.arange(n) / 1024.0, "time_idx": np.arange(n), "EEG_ch_Fz": g["raw"][0], "EEG_ch_Cz": g["raw"][1],
                            "EEG_cg_Fz": g["raw"][2], "IBI_ch": g["raw"][3], "diode": (np.arange(n) % 7).astype(float)})
    with contextlib.redirect_stdout(io.StringIO()):
        dec = md._decimate_signals(q=8)
    assert dec is not md and md.fs == 1024.0 and dec.fs == float(g["fs_out"]) and dec.id == "W_000"
    assert len(md.data) == n                                              # original untouched
    assert list(dec.data.columns)[:2] == ["time", "time_idx"]
    np.testing.assert_array_equal(dec.data["time"].values, g["time"])
    np.testing.assert_array_equal(dec.data["diode"].values, g["diode"])
    for r, col in enumerate(("EEG_ch_Fz", "EEG_ch_Cz", "EEG_cg_Fz", "IBI_ch")):
        y, ref = dec.data[col].values, g["out"][r]
        assert np.array_equal(np.isnan(y), np.isnan(ref))
        ok = ~np.isnan(ref)
        assert relerr(y[ok], ref[ok]) < TOL_SIGNAL


@pytest.mark.parametrize("n,fs", [(8192 * 3 + 517, 1024.0), (8192, 256.0), (50000, 256.0)])
def test_long_signals_take_the_tiled_path(n, fs):
    """Signals longer than one tile run through the coalesced tiled kernels (biquads, unit time stride);
    compare with SciPy's filtfilt cascade as dataloader.py:788-792 applies it, and with the thread-per-chunk kernels."""
    import os
    import torch
    from hyperscanning_signal_analysis_b200 import frontend
    rng = np.random.default_rng(int(n))
    t = np.arange(n) / fs
    x = 20.0 * rng.standard_normal((3, n)) + 5.0 * np.sin(2 * np.pi * 50.0 * t) + 30.0 + 3.0 * t[None, :] / t[-1]
    filt = fo.design_eeg_filters(fs, 1.0, min(64.0, 0.4 * fs))
    ref = fo.apply_filters_iir(x, filt)
    got = frontend.filtfilt_cascade(x, list(filt[:3]), remove_dc=True)
    assert relerr(got, ref) < TOL_SIGNAL
    # decimation of the long result (interior and edge spans of the polyphase kernel)
    for q in (2, 8, 5):
        d_ref = signal.decimate(ref, q, ftype="fir", zero_phase=True, axis=1)
        assert relerr(frontend.decimate(got, q), d_ref) < TOL_SIGNAL


@pytest.mark.parametrize("fc", [0.3, 0.02])
def test_slowly_decaying_filter_over_many_tiles(fc):
    """A high-pass far below the loader's 1 Hz: the state forgets slowly, so the tile start states reach back over many tiles
    (fc = 0.3 Hz: a look-back of ~20 aggregates in rounds of 8; fc = 0.02 Hz: |Q^8| >= 1e-3, the three-kernel path with its carry
    kernel).  Against scipy.signal.filtfilt, and bit-reproducible."""
    import torch
    from hyperscanning_signal_analysis_b200 import frontend
    fs, n = 1024.0, 4096 * 40 + 333
    rng = np.random.default_rng(5)
    x = rng.standard_normal((4, n)).cumsum(axis=1) * 0.05 + 10.0 * rng.standard_normal((4, n))
    b, a = signal.butter(2, fc, "high", fs=fs)
    ref = signal.filtfilt(b, a, x, axis=1)
    xd = torch.from_numpy(x).cuda()
    y1, y2 = xd.clone(), xd.clone()
    frontend.filtfilt_cascade_(y1, [(b, a)], remove_dc=False)
    frontend.filtfilt_cascade_(y2, [(b, a)], remove_dc=False)
    assert torch.equal(y1, y2)
    # filtfilt through a pole at 1 - 1e-4 amplifies rounding differently in any two implementations: SciPy's own result moves by
    # ~1e-8 relative when the same recurrence is evaluated in a different order
    assert relerr(y1.cpu().numpy(), ref) < (TOL_SIGNAL if fc > 0.1 else 1e-6)


def test_cfg4_full_size_properties():
    """BASELINE cfg4 at full size (2 x 19 ch x 1 h at 1024 Hz = 1.12 GB of float64): filtfilt cascade + decimate by 8 on the
    GPU; two channels against SciPy over the full hour, the rest through size-independent properties."""
    import torch
    from hyperscanning_signal_analysis_b200 import frontend, synth
    fs, n_ch, n = 1024.0, 38, 3600 * 1024
    x = synth.dyad_eeg(seed=4, m=n_ch, fs=fs, n_samples=n, drift=True)
    filt = fo.design_eeg_filters(fs, 1.0, 64.0)
    f3 = list(filt[:3])
    xd = torch.from_numpy(x).cuda()
    yd = xd.clone()
    frontend.filtfilt_cascade_(yd, f3, remove_dc=True)
    assert bool(torch.isfinite(yd).all())
    chk = [0, 37]
    ref = fo.apply_filters_iir(x[chk], filt)
    assert relerr(yd[chk].cpu().numpy(), ref) < TOL_SIGNAL
    dd = frontend.decimate_dev(yd, 8)
    assert dd.shape == (n_ch, n // 8)
    assert relerr(dd[chk].cpu().numpy(), signal.decimate(ref, 8, ftype="fir", zero_phase=True, axis=1)) < TOL_SIGNAL
    # run-to-run determinism
    y2 = xd.clone()
    frontend.filtfilt_cascade_(y2, f3, remove_dc=True)
    assert torch.equal(yd, y2)
    # linearity of the whole cascade: F(a x_i + b x_j) = a F(x_i) + b F(x_j)   (DC removal is linear too)
    mix = (0.75 * xd[3] - 1.5 * xd[21])[None].contiguous()
    frontend.filtfilt_cascade_(mix, f3, remove_dc=True)
    lin = 0.75 * yd[3] - 1.5 * yd[21]
    assert float((mix[0] - lin).abs().max() / lin.abs().max()) < TOL_SIGNAL
    # zero phase: away from the two ends (where forward-then-backward and backward-then-forward start-up transients differ; the
    # slowest pole, |z| = 0.9957, has decayed below 1e-9 after ~5000 samples) filtering the time-reversed signal gives the
    # time-reversed result
    rev = xd[5:6].flip(1).contiguous()
    frontend.filtfilt_cascade_(rev, f3, remove_dc=True)
    mid = slice(50000, n - 50000)
    assert float((rev.flip(1)[0, mid] - yd[5, mid]).abs().max() / yd[5].abs().max()) < TOL_SIGNAL
    # a constant is removed by the DC step and stays zero; a pure 50 Hz line is notched away in the interior
    c = torch.full((1, 100000), 7.25, dtype=torch.float64, device="cuda")
    frontend.filtfilt_cascade_(c, f3, remove_dc=True)
    assert float(c.abs().max()) < 1e-9
    t = torch.arange(200000, dtype=torch.float64, device="cuda") / fs
    line = (5.0 * torch.sin(2 * np.pi * 50.0 * t))[None].contiguous()
    frontend.filtfilt_cascade_(line, f3, remove_dc=True)
    assert float(line[0, 50000:150000].abs().max()) < 1e-6
