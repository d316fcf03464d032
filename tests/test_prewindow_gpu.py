"""Pre-window stage of EEG_IBI_FFDTF_Pipeline on the GPU (alpha SOS band-pass, Hilbert envelope / FAA, resample_poly,
crop, z-score) against vectors produced by the reference's own methods (tests/golden/prewindow.npz) and the oracle."""
import numpy as np
import pytest
from scipy import signal

from conftest import golden, relerr, TOL_SIGNAL
from oracle import frontend_oracle as fo
from test_oracle_cpu import _prewindow_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pl():
    from hyperscanning_signal_analysis_b200 import eeg_alpha_ibi_ffdtf
    return eeg_alpha_ibi_ffdtf


@pytest.fixture(scope="module")
def fe():
    from hyperscanning_signal_analysis_b200 import frontend
    return frontend


def test_alpha_bandpass_against_reference(pl):
    g, x = _prewindow_inputs()
    fs = float(g["fs"])
    for who, sl in (("ch", slice(0, 19)), ("cg", slice(19, 38))):
        filt = pl.alpha_bandpass_filter(x[sl], fs)
        assert filt.shape == (19, x.shape[1]) and filt.dtype == np.float64
        assert relerr(filt[[3, 5, 18]], g[f"filt_{who}"]) < TOL_SIGNAL
        assert relerr(filt, fo.alpha_bandpass(x[sl], fs)) < TOL_SIGNAL
    # 1-D input and axis=0 like scipy.signal.sosfiltfilt
    one = pl.alpha_bandpass_filter(x[3], fs)
    assert one.shape == (x.shape[1],) and relerr(one, g["filt_ch"][0]) < TOL_SIGNAL
    col = pl.alpha_bandpass_filter(np.ascontiguousarray(x[:4].T), fs, axis=0)
    assert relerr(col.T, fo.alpha_bandpass(x[:4], fs)) < TOL_SIGNAL
    with pytest.raises(ValueError):
        pl.alpha_bandpass_filter(x[:2, :27], fs)          # padlen = 27 for four sections


def test_sosfiltfilt_other_designs(fe):
    rng = np.random.default_rng(5)
    x = rng.standard_normal((5, 4000)) * 30 + 7
    for sos in (signal.butter(2, 0.2, output="sos"), signal.butter(6, [0.05, 0.4], btype="band", output="sos"),
                signal.cheby1(3, 1, 0.3, btype="high", output="sos"), signal.ellip(4, 0.5, 40, 0.25, output="sos")):
        assert relerr(fe.sosfiltfilt(sos, x), signal.sosfiltfilt(sos, x, axis=-1)) < TOL_SIGNAL


def test_hilbert_envelope_and_faa(pl, fe):
    import torch
    g, x = _prewindow_inputs()
    fs = float(g["fs"])
    names = [str(s) for s in g["names"]]
    xs = np.ascontiguousarray(x[:3, :1001])
    env = fe.hilbert_envelope_dev(torch.from_numpy(xs).cuda(), N=1008).cpu().numpy()       # 2^4 3^2 7
    assert relerr(env, g["short_env_fast"]) < TOL_SIGNAL
    env = fe.hilbert_envelope_dev(torch.from_numpy(xs).cuda()).cpu().numpy()               # N = 1001 = 7 11 13 (generic radix)
    assert relerr(env, g["short_env_n"]) < TOL_SIGNAL
    for N in (1000, 1024, 1215, 2187, 900):                                                # 5-smooth, 2^k, 3^5 5, 3^7, N < n (truncation)
        ref = np.abs(signal.hilbert(xs, N=N, axis=-1)[:, :1001])
        got = fe.hilbert_envelope_dev(torch.from_numpy(xs).cuda(), N=N).cpu().numpy()
        assert got.shape == ref.shape and relerr(got, ref) < TOL_SIGNAL
    for who, sl in (("ch", slice(0, 19)), ("cg", slice(19, 38))):
        filt = fo.alpha_bandpass(x[sl], fs)
        faa = pl.compute_asymmetry(filt, names, metric="amp")                              # N = next_fast_len(9637) = 9680 = 2^4 5 11^2
        assert np.max(np.abs(faa - g[f"faa_{who}"])) < 1e-9 * max(1.0, np.max(np.abs(g[f"faa_{who}"])))
        faa_p = pl.compute_asymmetry(filt, names, metric="power")
        assert np.max(np.abs(faa_p - g[f"faa_power_{who}"])) < 1e-9 * max(1.0, np.max(np.abs(g[f"faa_power_{who}"])))
    with pytest.raises(ValueError):
        pl.compute_asymmetry(x[:19], names, left_chan="XX")
    with pytest.raises(ValueError):
        pl.compute_asymmetry(x[:19], names, metric="phase")


def test_downsample_crop_zscore_and_whole_stage(pl):
    g, x = _prewindow_inputs()
    fs = float(g["fs"])
    names = [str(s) for s in g["names"]]
    for who, k in (("ch", 0), ("cg", 1)):
        assert relerr(pl.downsample_signal(g[f"faa_{who}"], fs, 8.0), g[f"faa_ds_{who}"]) < TOL_SIGNAL
        assert relerr(pl.downsample_signal(g["ibi"][k], fs, 8.0), g[f"ibi_ds_{who}"]) < TOL_SIGNAL
    rng = np.random.default_rng(2)
    y = rng.standard_normal(1000)
    assert relerr(pl.downsample_signal(y, 12, 4), signal.resample_poly(y, 1, 3)) < TOL_SIGNAL
    with pytest.raises(ValueError):
        pl.downsample_signal(y, 8, 8)
    with pytest.raises(ValueError):
        pl.downsample_signal(y, 128, 50)
    c = pl.crop_signal(np.arange(700.0), 8.0, 10, 60)
    assert c[0] == 80 and c.shape == (480,)
    with pytest.raises(ValueError):
        pl.crop_signal(np.arange(500.0), 8.0, 10, 60)
    sig = pl.preprocess_dyad(x[:19], x[19:], names, g["ibi"][0], g["ibi"][1], fs, fs, fs_ds=8.0)
    assert sig.shape == (4, 480)
    assert relerr(sig, g["signals_to_ffDTF"]) < 1e-7          # z-scored log-ratios of envelopes: condition of log near 0 envelope
    np.testing.assert_allclose(sig.mean(axis=1), 0.0, atol=1e-12)
    np.testing.assert_allclose(sig.std(axis=1), 1.0, rtol=1e-12)
    # ... and on into the windows: the stage's output feeds compute_ffdtf_windows like run_pipeline (:729-755)
    res = pl.compute_ffdtf_windows(sig, 8.0, 3, None, ar_p=5, freq_min=1.0, freq_max=3.9, freq_step=0.1)
    from oracle import mvar_oracle as mo
    ref = mo.full_freq_dtf(g["signals_to_ffDTF"][:, :160], res["freqs"], 8.0, optimal_model_order=5)
    assert relerr(res["ff_dtf_windowed"][0], ref) < 1e-6


def test_prewindow_edge_cases(pl, fe):
    import torch
    rng = np.random.default_rng(11)
    fs = 128.0
    # shortest signal sosfiltfilt accepts: padlen + 1 = 28 samples (four sections)
    x = rng.standard_normal((2, 28))
    assert relerr(pl.alpha_bandpass_filter(x, fs), fo.alpha_bandpass(x, fs)) < TOL_SIGNAL
    # a single section and a section with a zero a2 / b2 (padlen shrinks by one tap: scipy's ntaps correction)
    sos1 = signal.butter(1, 0.3, output="sos")
    y = rng.standard_normal((3, 50))
    assert relerr(fe.sosfiltfilt(sos1, y), signal.sosfiltfilt(sos1, y, axis=-1)) < TOL_SIGNAL
    # Hilbert: n = 1, n = 2, odd n, a prime length (generic radix) and an all-zero signal
    for n in (1, 2, 3, 31, 97 * 2):
        xs = rng.standard_normal((2, n))
        if n == 97 * 2:
            with pytest.raises(Exception):          # 97 > 31: no radix for it, the library says so instead of guessing
                fe.hilbert_envelope_dev(torch.from_numpy(xs).cuda())
            continue
        got = fe.hilbert_envelope_dev(torch.from_numpy(xs).cuda()).cpu().numpy()
        assert relerr(got, np.abs(signal.hilbert(xs, axis=-1))) < TOL_SIGNAL
    z = fe.hilbert_envelope_dev(torch.zeros((1, 64), dtype=torch.float64, device="cuda"))
    assert float(z.abs().max()) == 0.0
    assert fe.hilbert_envelope_dev(torch.zeros((0, 64), dtype=torch.float64, device="cuda")).shape == (0, 64)
    # resample_poly: length not a multiple of the factor, factor larger than the signal
    for n, down in ((1001, 16), (10, 16), (17, 2)):
        v = rng.standard_normal(n)
        assert relerr(pl.downsample_signal(v, float(down), 1.0), signal.resample_poly(v, 1, down)) < TOL_SIGNAL
    # z-score of a constant row is NaN, like NumPy's 0 / 0 in the reference (:719)
    with np.errstate(invalid="ignore", divide="ignore"):
        zc = pl.zscore_rows(np.vstack([np.ones(8), np.arange(8.0)]))
    assert np.isnan(zc[0]).all() and abs(zc[1].std() - 1.0) < 1e-12
