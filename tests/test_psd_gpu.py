"""Multitaper PSD (hand-written FFT) against the oracle restatement of MNE's psd_array_multitaper
(parity with MNE itself is unpinned: mne is not installable here; cross-checked when importable)."""
import numpy as np
import pytest

from conftest import relerr
from oracle import frontend_oracle as fo

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,sfreq,bw,fmin,fmax", [
    (8192, 128.0, 2.0, 1.0, 30.0),      # cfg4 segment: 64 s @128 Hz -> 126 tapers
    (7680, 128.0, 2.0, 1.0, 30.0),      # non power of two: 2^9 * 15
    (1024, 128.0, 4.0, 0.0, 64.0),      # full range incl. DC and Nyquist
    (1000, 250.0, 4.0, 0.0, 125.0),     # 2^3 * 125, even n
    (999, 100.0, 3.0, 2.0, 40.0),       # odd n: direct stage only
    (4096, 256.0, 1.0, 5.0, 20.0),
])
def test_psd_against_oracle(n, sfreq, bw, fmin, fmax):
    from hyperscanning_signal_analysis_b200 import psd
    rng = np.random.default_rng(n)
    x = rng.standard_normal((5, n)).cumsum(axis=1) * 0.1 + rng.standard_normal((5, n)) + 3.0
    freqs, p = psd.compute_psd_multitaper(x, sfreq, fmin, fmax, bw)
    fr, pr = fo.psd_multitaper(x, sfreq, fmin, fmax, bw)
    assert np.array_equal(freqs, fr) and p.shape == pr.shape == (5, len(fr))
    assert relerr(p, pr) < 1e-9
    try:
        from mne.time_frequency import psd_array_multitaper
    except Exception:
        return
    pm, fm = psd_array_multitaper(x, sfreq=sfreq, fmin=fmin, fmax=fmax, bandwidth=bw, verbose=False)
    assert relerr(p, pm) < 1e-9


def test_psd_white_noise_level_and_average():
    from hyperscanning_signal_analysis_b200 import psd
    x = np.random.default_rng(0).standard_normal((19, 2048))
    freqs, p = psd.compute_psd_multitaper(x, 128.0, 1.0, 30.0, 2.0)
    assert abs(p.mean() - 2.0) < 0.1                  # unit-variance white noise, 'length' normalisation
    avg = psd.average_psd_across_conditions({"a": p, "b": 3 * p})
    np.testing.assert_allclose(avg, 2 * p)
    with pytest.raises(ValueError):
        psd.average_psd_across_conditions({})
    # many signals, one group per signal
    big = np.random.default_rng(1).standard_normal((400, 512))
    f2, p2 = psd.compute_psd_multitaper(big, 64.0, 0.0, 32.0, 2.0)
    assert relerr(p2, fo.psd_multitaper(big, 64.0, 0.0, 32.0, 2.0)[1]) < 1e-9


@pytest.mark.parametrize("n,sfreq,bw,fmin,fmax", [
    (8191, 128.0, 1.0, 1.0, 30.0),      # prime > 4096: Bluestein over a 16384-point transform
    (23041, 128.0, 0.5, 1.0, 30.0),     # a 3-minute movie segment plus one odd sample (src/io_utils.py:131: arbitrary lengths)
    (23040, 128.0, 0.5, 1.0, 30.0),     # 3 minutes: 2^9 * 45, smooth, too long for shared memory
    (4099, 100.0, 2.0, 0.0, 50.0),      # prime, full range incl. DC
    (62, 31.0, 4.0, 0.0, 15.5),         # 2 * 31: smooth, tiny, Nyquist bin
    (74, 37.0, 4.0, 0.0, 18.5),         # 2 * 37: Bluestein, tiny, Nyquist bin
])
def test_psd_any_length(n, sfreq, bw, fmin, fmax):
    from hyperscanning_signal_analysis_b200 import psd
    rng = np.random.default_rng(n)
    x = rng.standard_normal((3, n)).cumsum(axis=1) * 0.1 + rng.standard_normal((3, n)) - 2.0
    freqs, p = psd.compute_psd_multitaper(x, sfreq, fmin, fmax, bw)
    fr, pr = fo.psd_multitaper(x, sfreq, fmin, fmax, bw)
    assert np.array_equal(freqs, fr) and p.shape == pr.shape
    assert relerr(p, pr) < 1e-9


@pytest.mark.parametrize("n", [4096, 8192])
def test_psd_paths_agree(n):
    """The register radix-16 kernel, the shared-memory kernel and the general (global-memory FFT) path on the same input."""
    import torch
    from hyperscanning_signal_analysis_b200 import _lib, psd
    lib = _lib.load()
    rng = np.random.default_rng(7 + n)
    x = rng.standard_normal((7, n)).cumsum(axis=1) * 0.05 + rng.standard_normal((7, n)) + 1.5
    ref = fo.psd_multitaper(x, 128.0, 0.0, 64.0, 2.0)[1]           # every bin incl. DC and Nyquist
    got = {}
    try:
        for path in (1, 2, 3):
            _lib.check(lib.hs_mt_psd_set_path(path), "set_path")
            got[path] = psd.compute_psd_multitaper(x, 128.0, 0.0, 64.0, 2.0)[1]
    finally:
        lib.hs_mt_psd_set_path(0)
    for path, p in got.items():
        assert relerr(p, ref) < 1e-9, path
    # a narrow band (only some rows of the spectrum are exchanged) and an odd number of tapers (last pair has one member)
    for fmin, fmax, bw in ((1.0, 30.0, 2.0), (10.0, 12.0, 0.75), (50.0, 64.0, 1.0)):
        f, p = psd.compute_psd_multitaper(x, 128.0, fmin, fmax, bw)
        assert relerr(p, fo.psd_multitaper(x, 128.0, fmin, fmax, bw)[1]) < 1e-9, (fmin, fmax, bw)
    assert lib.hs_mt_psd_set_path(9) != 0
