"""Parity of the CUDA MVAR/DTF path (through the C ABI) against the reference's golden
vectors and the oracle.  Tolerances are north_star's: 1e-9 covariances, 1e-7 AR / DTF,
norm-wise (max|a-b| / max|b|)."""
import contextlib
import io

import numpy as np
import pytest

from conftest import golden, relerr, TOL_MODEL, TOL_SIGNAL
from oracle import mvar_oracle as mo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mv():
    from hyperscanning_signal_analysis_b200 import mtmvar
    return mtmvar


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def test_m4_against_reference_golden(mv, capsys):
    g = golden("mvar_m4.npz")
    x, p, fs, freqs = g["x"], int(g["p"]), float(g["fs"]), g["freqs"]
    rl, rr, r0 = mv.count_corr(x[:, :, None], p, 1)
    assert relerr(rl, g["r_left"]) < TOL_SIGNAL and relerr(rr, g["r_right"]) < TOL_SIGNAL and relerr(r0, g["r_zero"]) < TOL_SIGNAL
    A, V = mv.ar_coeff(x, p)
    assert A.shape == (4, 4, p) and V.shape == (4, 4) and A.dtype == np.float64
    assert relerr(A, g["A"]) < TOL_MODEL and relerr(V, g["V"]) < TOL_MODEL
    H, Af = mv.mvar_transfer_function(g["A"], freqs, fs)
    assert H.dtype == np.complex128 and H.shape == (4, 4, len(freqs))
    assert relerr(H, g["H"]) < TOL_MODEL and relerr(Af, g["Af"]) < 1e-12
    dtf = mv.dtf_multivariate(x, freqs, fs, optimal_model_order=p)
    assert "Using provided model order: p = 5" in capsys.readouterr().out       # reference print kept (mtmvar.py:229)
    assert relerr(dtf, g["dtf"]) < TOL_MODEL
    ff = quiet(mv.full_freq_dtf, x, freqs, fs, optimal_model_order=p)
    assert relerr(ff, g["ffdtf"]) < TOL_MODEL
    np.testing.assert_allclose(ff.sum(axis=(1, 2)), 1.0, rtol=1e-12)
    S = quiet(mv.multivariate_spectra, x, freqs, fs, optimal_model_order=p)
    assert relerr(S, g["S"]) < TOL_MODEL
    assert relerr(S, S.transpose(1, 0, 2)) < 1e-10                               # H V H^T quirk
    gp = quiet(mv.gen_partial_directed_coherence, x, freqs, fs, optimal_model_order=p)
    assert relerr(gp, g["gpdc"]) < TOL_MODEL


def test_criterion_against_reference_golden(mv):
    g = golden("mvar_m4.npz")
    for c, key in (("AIC", "crit_aic"), ("HQ", "crit_hq"), ("SC", "crit_sc")):
        crit, rng, popt = mv.mvar_criterion(g["x"], 8, c)
        assert relerr(crit, g[key]) < TOL_MODEL
        assert rng.tolist() == list(range(1, 9))
    assert [int(mv.mvar_criterion(g["x"], 8, c)[2]) for c in ("AIC", "HQ", "SC")] == g["popt"].tolist()
    with pytest.raises(ValueError):
        mv.mvar_criterion(g["x"], 3, "BIC")
    # order selection path of full_freq_dtf (optimal_model_order=None, mtmvar.py:224-227)
    ff = quiet(mv.full_freq_dtf, g["x"], g["freqs"], float(g["fs"]), max_model_order=8)
    ref = mo.full_freq_dtf(g["x"], g["freqs"], float(g["fs"]), max_model_order=8)
    assert relerr(ff, ref) < TOL_MODEL


def test_multi_trial_against_reference_golden(mv):
    g = golden("mvar_trials.npz")
    A, V = mv.ar_coeff(g["x"], int(g["p"]))
    assert relerr(A, g["A"]) < TOL_MODEL and relerr(V, g["V"]) < TOL_MODEL
    rl, rr, r0 = mv.count_corr(g["x"], int(g["p"]), 1)
    assert relerr(rl, g["r_left"]) < TOL_SIGNAL and relerr(rr, g["r_right"]) < TOL_SIGNAL


def test_cfg2_windows_against_reference_golden(mv):
    g = golden("mvar_cfg2_windows.npz")
    wins = g["windows"]
    for w in range(wins.shape[0]):
        A, V = mv.ar_coeff(wins[w], 8)
        assert relerr(A, g["A"][w]) < TOL_MODEL, g["cond"][w]
        assert relerr(V, g["V"][w]) < TOL_MODEL
        ff = quiet(mv.full_freq_dtf, wins[w], g["freqs"], 256.0, optimal_model_order=8)
        assert relerr(ff, g["ffdtf"][w]) < TOL_MODEL
    # batched call on the two windows laid out back to back
    sig = np.concatenate([wins[0], wins[1]], axis=1)
    out = mv.windowed_ffdtf(sig, [0, 512], 512, g["freqs"], 256.0, 8).cpu().numpy()
    assert relerr(out, g["ffdtf"]) < TOL_MODEL


def test_cfg2_stress_ill_conditioned(mv):
    """cond(G) ~ 1.5e9: the reference's own LU result moves by ~1e-8 under 1-ulp input changes, so the gate
    is cond-aware here (10 * cond * eps ~ 3e-6); ffDTF is still required within 1e-6."""
    g = golden("mvar_cfg2_windows_lp40.npz")
    for w in range(g["windows"].shape[0]):
        A, V = mv.ar_coeff(g["windows"][w], 8)
        assert relerr(A, g["A"][w]) < 10 * g["cond"][w] * 2.2e-16
        ff = quiet(mv.full_freq_dtf, g["windows"][w], g["freqs"], 256.0, optimal_model_order=8)
        assert relerr(ff, g["ffdtf"][w]) < 1e-6


def test_cfg1_full_segment(mv):
    from hyperscanning_signal_analysis_b200 import synth
    from oracle import frontend_oracle as fo
    g = golden("mvar_cfg1.npz")
    x = synth.cfg1_raw()
    assert abs(float(np.sum(x)) - float(g["raw_sum"])) < 1e-6 * abs(float(g["raw_sum"])) + 1e-6
    y = fo.apply_filters_iir(x, fo.design_eeg_filters(256.0, 1.0, float(g["highcut"])))
    assert relerr(y[[0, 18, 19, 37]], g["filtered_rows"]) < 1e-12
    A, V = mv.ar_coeff(y, 8)
    assert relerr(A, g["A"]) < TOL_MODEL and relerr(V, g["V"]) < TOL_MODEL
    ff = quiet(mv.full_freq_dtf, y, g["freqs"], 256.0, optimal_model_order=8)
    assert relerr(ff, g["ffdtf"]) < TOL_MODEL
    S = quiet(mv.multivariate_spectra, y, g["freqs"], 256.0, optimal_model_order=8)
    assert relerr(S, g["S"]) < TOL_MODEL


@pytest.mark.parametrize("m,n,p,F", [(1, 64, 2, 5), (2, 100, 1, 7), (5, 200, 3, 33), (8, 256, 4, 16), (9, 300, 5, 50),
                                      (16, 400, 6, 31), (17, 400, 2, 64), (24, 512, 8, 40), (33, 600, 4, 13), (38, 512, 8, 256),
                                      (40, 700, 10, 20), (19, 512, 20, 12)])
def test_shapes_against_oracle(mv, m, n, p, F):
    rng = np.random.default_rng(100 + m + p)
    x = rng.standard_normal((m, n))
    x[:, 1:] += 0.5 * x[:, :-1]
    x += 0.2 * rng.standard_normal((m, m)) @ x
    freqs = np.sort(rng.uniform(0, 64, F))          # arbitrary, non-uniform grid
    A, V = mv.ar_coeff(x, p)
    Ar, Vr = mo.ar_coeff(x, p)
    assert relerr(A, Ar) < TOL_MODEL and relerr(V, Vr) < TOL_MODEL
    H, Af = mv.mvar_transfer_function(Ar, freqs, 128.0)
    Hr, Afr = mo.mvar_transfer_function(Ar, freqs, 128.0)
    assert relerr(H, Hr) < TOL_MODEL and relerr(Af, Afr) < 1e-12
    # H(f) A(f) = I
    for fi in (0, F - 1):
        assert np.max(np.abs(H[:, :, fi] @ Af[:, :, fi] - np.eye(m))) < 1e-9
    ff = quiet(mv.full_freq_dtf, x, freqs, 128.0, optimal_model_order=p)
    assert relerr(ff, mo.full_freq_dtf(x, freqs, 128.0, optimal_model_order=p)) < TOL_MODEL
    np.testing.assert_allclose(ff.sum(axis=(1, 2)), 1.0, rtol=1e-12)


def test_windowed_batch_against_oracle(mv):
    from hyperscanning_signal_analysis_b200 import synth
    from oracle import frontend_oracle as fo
    x = synth.dyad_eeg(seed=synth.BASE_SEED + 9, n_samples=256 * 20)
    y = fo.apply_filters_iir(x, fo.design_eeg_filters(256.0, 1.0, 64.0))
    starts, wsz = mo.window_starts(y.shape[1], 19, 512)
    freqs = np.linspace(0, 128, 24, endpoint=False)
    out, A, V = mv.windowed_ffdtf(y, starts, wsz, freqs, 256.0, 8, return_model=True)
    out = out.cpu().numpy()
    assert out.shape == (19, 38, 38, 24)
    for w in (0, 7, 18):
        seg = y[:, starts[w]:starts[w] + wsz]
        Ar, Vr = mo.ar_coeff(seg, 8)
        assert relerr(A[w].cpu().numpy(), Ar) < TOL_MODEL
        assert relerr(V[w].cpu().numpy(), Vr) < TOL_MODEL
        assert relerr(out[w], mo.full_freq_dtf(seg, freqs, 256.0, optimal_model_order=8)) < TOL_MODEL
    np.testing.assert_allclose(out.sum(axis=(2, 3)), 1.0, rtol=1e-12)
    # run-to-run determinism (fixed summation order everywhere)
    out2 = mv.windowed_ffdtf(y, starts, wsz, freqs, 256.0, 8).cpu().numpy()
    assert np.array_equal(out, out2)


def test_host_plan_matches_device_path(mv):
    from hyperscanning_signal_analysis_b200 import synth
    x = synth.dyad_eeg(seed=5, n_samples=4096, line_amp=0.0)
    starts = np.arange(0, 4096 - 512 + 1, 256)
    freqs = np.linspace(0, 128, 32, endpoint=False)
    plan = mv.FfdtfPlan(len(starts), 38, 512, 8, 32, 4096)
    got = plan.run(x, starts, freqs, 256.0)
    ref = mv.windowed_ffdtf(x, starts, 512, freqs, 256.0, 8).cpu().numpy()
    assert np.array_equal(got, ref)
    got2 = plan.run(x, starts[:3], freqs, 256.0)
    assert np.array_equal(got2, ref[:3])
    plan.close()


def test_singular_and_bad_arguments(mv):
    x = np.zeros((4, 100))
    with pytest.raises(np.linalg.LinAlgError):
        mv.ar_coeff(x, 2)
    with pytest.raises(ValueError):
        mv.windowed_ffdtf(np.zeros((4, 100)), [90], 20, [1.0], 8.0, 2)
    with pytest.raises(ValueError):
        mv.windowed_ffdtf(np.zeros((4, 100)), [0], 20, [1.0], 8.0, 20)
    # empty batch
    out = mv.windowed_ffdtf(np.ones((4, 100)), [], 20, [1.0, 2.0], 8.0, 2)
    assert tuple(out.shape) == (0, 4, 4, 2)


def test_window_driver_batched_matches_reference_loop():
    """compute_ffdtf_windows == the reference's serial loop (eeg_alpha_ibi_ffdtf.py:741-755) at its real-use shape
    (m = 4, fs = 8 Hz, p = 5, freqs = arange(fmin, fmax + step, step))."""
    from hyperscanning_signal_analysis_b200.eeg_alpha_ibi_ffdtf import compute_ffdtf_windows
    rng = np.random.default_rng(42)
    m, T, fs = 4, 2400, 8.0
    x = rng.standard_normal((m, T))
    x[:, 1:] += 0.7 * x[:, :-1]
    x[1, 2:] += 0.4 * x[0, :-2]
    res = compute_ffdtf_windows(x, fs, n_windows=7, window_size=600, ar_p=5, freq_min=0.1, freq_max=4.0, freq_step=0.1, with_spectra=True)
    freqs = np.arange(0.1, 4.0 + 0.1, 0.1)
    assert np.array_equal(res["freqs"], freqs) and res["ff_dtf_windowed"].shape == (7, 4, 4, len(freqs))
    starts, W = mo.window_starts(T, 7, 600)
    assert res["starts"].tolist() == starts.tolist()
    for w, s in enumerate(starts):
        seg = x[:, s:s + W]
        assert relerr(res["ff_dtf_windowed"][w], mo.full_freq_dtf(seg, freqs, fs, optimal_model_order=5)) < TOL_MODEL
        assert relerr(res["spectra_windowed"][w], mo.multivariate_spectra(seg, freqs, fs, optimal_model_order=5)) < TOL_MODEL
    # per-window order selection (ar_p=None -> mvar_criterion, eeg_alpha_ibi_ffdtf.py:586-587)
    res2 = compute_ffdtf_windows(x, fs, n_windows=3, window_size=None, ar_p=None, freq_min=0.1, freq_max=4.0, freq_step=0.5, max_model_order=6)
    st2, W2 = mo.window_starts(T, 3, None)
    for w, s in enumerate(st2):
        popt = int(mo.mvar_criterion(x[:, s:s + W2], 6, "AIC")[2])
        assert res2["p_opt_w"][w] == popt
        assert relerr(res2["ff_dtf_windowed"][w], mo.full_freq_dtf(x[:, s:s + W2], res2["freqs"], fs, optimal_model_order=popt)) < TOL_MODEL


@pytest.mark.parametrize("m,n,p,F,trials", [(41, 300, 2, 9, 1), (48, 400, 3, 17, 1), (64, 512, 4, 12, 3), (90, 700, 2, 5, 1)])
def test_generic_path_large_m(mv, m, n, p, F, trials):
    """m > 40 goes through the block-GEMM LWR and the scratch-matrix Gauss-Jordan (generic_kernels.cu)."""
    rng = np.random.default_rng(m * 7 + p)
    shape = (m, n) if trials == 1 else (m, n, trials)
    x = rng.standard_normal(shape)
    x[:, 1:] += 0.5 * x[:, :-1]
    freqs = np.sort(rng.uniform(0, 64, F))
    A, V = mv.ar_coeff(x, p)
    Ar, Vr = mo.ar_coeff(x, p)
    assert relerr(A, Ar) < TOL_MODEL and relerr(V, Vr) < TOL_MODEL
    H, Af = mv.mvar_transfer_function(Ar, freqs, 128.0)
    Hr, Afr = mo.mvar_transfer_function(Ar, freqs, 128.0)
    assert relerr(H, Hr) < TOL_MODEL and relerr(Af, Afr) < 1e-12
    if trials == 1:
        ff = quiet(mv.full_freq_dtf, x, freqs, 128.0, optimal_model_order=p)
        assert relerr(ff, mo.full_freq_dtf(x, freqs, 128.0, optimal_model_order=p)) < TOL_MODEL
        np.testing.assert_allclose(ff.sum(axis=(1, 2)), 1.0, rtol=1e-12)
        crit, _, popt = mv.mvar_criterion(x, p, "SC")
        cr, _, pr = mo.mvar_criterion(x, p, "SC")
        assert relerr(crit, cr) < TOL_MODEL and int(popt) == int(pr)


def test_cfg5_shape_high_channel_multi_trial(mv):
    """BASELINE configs[4]: 2 x 64 channels, p = 15, 512 bins, covariances averaged over 100 epochs per window."""
    from hyperscanning_signal_analysis_b200 import synth
    ep = synth.cfg5_epochs(n_windows=1)[0]                       # (128, 512, 100)
    A, V = mv.ar_coeff(ep, 15)
    Ar, Vr = mo.ar_coeff(ep, 15)
    assert A.shape == (128, 128, 15) and relerr(A, Ar) < TOL_MODEL and relerr(V, Vr) < TOL_MODEL
    freqs = np.linspace(0, 128, 512, endpoint=False)
    H, _ = mv.mvar_transfer_function(Ar, freqs, 256.0)
    sel = [0, 100, 511]
    Hr, _ = mo.mvar_transfer_function(Ar, freqs[sel], 256.0)
    assert relerr(H[:, :, sel], Hr) < TOL_MODEL


def test_optimistic_elimination_falls_back_to_pivoting(mv):
    """A(f) with a zero diagonal at f = 0 (A_1 = I - P, P a cyclic shift): unpivoted elimination breaks down, the
    a-posteriori check must flag those bins and the pivoted redo must deliver np.linalg.inv's answer."""
    m, p = 38, 2
    rng = np.random.default_rng(5)
    P = np.roll(np.eye(m), 1, axis=1)
    A = np.zeros((m, m, p))
    A[:, :, 0] = np.eye(m) - P
    A[:, :, 1] = 0.01 * rng.standard_normal((m, m))
    freqs = np.array([0.0, 3.0, 17.0, 40.0, 64.0, 90.0, 127.5])
    H, Af = mv.mvar_transfer_function(A, freqs, 256.0)
    Hr, Afr = mo.mvar_transfer_function(A, freqs, 256.0)
    assert abs(np.diagonal(Afr[:, :, 0])).max() < 0.05            # (nearly) zero diagonal at f = 0
    assert relerr(Af, Afr) < 1e-12
    for fi in range(len(freqs)):
        assert relerr(H[:, :, fi], Hr[:, :, fi]) < TOL_MODEL, fi
    # ffDTF over a grid that mixes flagged and unflagged bins (row sums come from both passes)
    from hyperscanning_signal_analysis_b200 import mtmvar
    import torch
    res = mtmvar.batched_transfer(torch.from_numpy(A[None]).cuda(), freqs, 256.0, want=("ffdtf", "dtf"))
    dtf_r = np.abs(Hr) ** 2
    ff_r = dtf_r / dtf_r.sum(axis=(1, 2), keepdims=True)
    assert relerr(res["dtf"][0].cpu().numpy(), dtf_r) < TOL_MODEL
    assert relerr(res["ffdtf"][0].cpu().numpy(), ff_r) < TOL_MODEL


def test_partial_coherence_and_ddtf_against_reference_golden(mv, capsys):
    """partial_coherence / direct_dtf (mtmvar.py:287-385): one pivoted inverse per bin instead of m^2 minor determinants."""
    g = golden("mvar_pcoh.npz")
    k4 = mv.partial_coherence(g["m4_S"])
    assert k4.dtype == np.complex128 and k4.shape == g["m4_kappa"].shape
    assert relerr(k4, g["m4_kappa"]) < TOL_MODEL
    assert np.all(k4[np.arange(4), np.arange(4), :] == 1.0)
    k38 = mv.partial_coherence(g["w0_S"])
    assert relerr(k38, g["w0_kappa"]) < TOL_MODEL
    assert np.max(np.abs(k38 - g["w0_kappa"])) < TOL_MODEL          # |kappa| <= ~1: absolute = element-wise here
    m4 = golden("mvar_m4.npz")
    capsys.readouterr()
    dd = mv.direct_dtf(m4["x"], m4["freqs"], float(m4["fs"]), optimal_model_order=int(m4["p"]))
    out = capsys.readouterr().out
    assert out.count("Using provided model order") == 2            # both nested reference calls print (mtmvar.py:191, :266)
    assert relerr(dd, g["m4_ddtf"]) < TOL_MODEL
    w = golden("mvar_cfg2_windows.npz")["windows"][0]
    dd38 = quiet(mv.direct_dtf, w, g["freqs"], 256.0, optimal_model_order=8)
    assert dd38.shape == (38, 38, 12) and dd38.dtype == np.float64
    assert relerr(dd38, g["w0_ddtf"]) < TOL_MODEL
    # 1 x 1 input: the minor is defined as 1 (mtmvar.py:320-321)
    assert np.all(mv.partial_coherence(np.full((1, 1, 3), 2.0 + 1j)) == 1.0)
    # singular spectrum -> LinAlgError, like np.linalg.det/inv based code would misbehave; the kernel flags it
    with pytest.raises(np.linalg.LinAlgError):
        mv.partial_coherence(np.zeros((3, 3, 2), dtype=complex))


def test_cfg2_full_size_properties(mv):
    """BASELINE cfg2 at full size (599 windows x 38 ch x 512 samples, p = 8, 256 bins: 153 344 complex inversions) through
    size-independent properties, plus the oracle on a handful of windows."""
    import torch
    import bench
    y, starts, freqs = bench.make_task(20260101)
    assert len(starts) == 599 and freqs.size == 256 and y.shape == (38, 153600)
    ff, A, V = mv.windowed_ffdtf(y, starts, 512, freqs, 256.0, 8, return_model=True)
    assert ff.shape == (599, 38, 38, 256)
    # (ii) every row of every window sums to 1 over (j, f)  (mtmvar.py:281-283)
    rows = ff.sum(dim=(2, 3))
    assert float((rows - 1.0).abs().max()) < 1e-12
    assert bool(torch.isfinite(ff).all()) and float(ff.min()) >= 0.0
    # run-to-run determinism at full size (fixed summation order, no atomics on the data path)
    ff2 = mv.windowed_ffdtf(y, starts, 512, freqs, 256.0, 8)
    assert torch.equal(ff, ff2)
    del ff2
    # scale invariance: x * 2^k scales every covariance by exactly 4^k, so A, H and the ffDTF are unchanged to the last bit
    ff3, A3, V3 = mv.windowed_ffdtf(y * 8.0, starts, 512, freqs, 256.0, 8, return_model=True)
    assert torch.equal(A, A3) and torch.equal(V * 64.0, V3) and torch.equal(ff, ff3)
    del ff3
    # (iii) H(f) A(f) = I on a slice of windows (complex outputs of the same kernels)
    sel = torch.tensor([0, 123, 300, 598], device="cuda")
    res = mv.batched_transfer(A[sel].contiguous(), freqs, 256.0, want=("H", "Af", "dtf"))
    H, Af = res["H"].permute(0, 3, 1, 2), res["Af"].permute(0, 3, 1, 2)         # (w, f, m, m)
    eye = torch.eye(38, dtype=torch.complex128, device="cuda")
    assert float((H @ Af - eye).abs().max()) < 1e-9
    assert float((res["dtf"] - res["H"].abs() ** 2).abs().max() / res["dtf"].max()) < 1e-12
    # overlapping windows share samples: window w+1 starts 256 samples after window w (hop = W/2)
    assert np.all(np.diff(starts) == 256)
    # ... and the oracle on a few windows
    ffh = ff[sel].cpu().numpy()
    for k, w in enumerate(sel.tolist()):
        seg = y[:, starts[w]:starts[w] + 512]
        assert relerr(ffh[k], mo.full_freq_dtf(seg, freqs, 256.0, optimal_model_order=8)) < TOL_MODEL


def test_partial_coherence_batch_and_edges(mv):
    import torch
    rng = np.random.default_rng(3)
    # batched call == per-matrix calls; F = 1; m = 2 .. 9 (every tile count below the dyad's)
    for m in (2, 3, 8, 9):
        H = rng.standard_normal((m, m, 5)) + 1j * rng.standard_normal((m, m, 5))
        S = np.einsum("ikf,jkf->ijf", H, H) + np.eye(m)[:, :, None] * 0.1          # H H^T: complex symmetric like the reference's spectra
        ref = mo.partial_coherence(S)
        assert relerr(mv.partial_coherence(S), ref) < TOL_MODEL
        assert relerr(mv.partial_coherence(S[:, :, :1]), ref[:, :, :1]) < TOL_MODEL
    S2 = torch.from_numpy(np.stack([S, 2.0 * S])).cuda()
    k, dd, st = mv.batched_partial_coherence(S2, ffdtf=torch.ones((2, 9, 9, 5), dtype=torch.float64, device="cuda"))
    assert int(st.max()) == 0
    assert relerr(k[0].cpu().numpy(), ref) < TOL_MODEL and relerr(k[1].cpu().numpy(), ref) < TOL_MODEL      # scale invariant
    assert relerr(dd[0].cpu().numpy(), np.abs(ref)) < TOL_MODEL                                              # ffDTF = 1 -> dDTF = |kappa|
    # empty batch
    k0, _, st0 = mv.batched_partial_coherence(torch.empty((0, 4, 4, 3), dtype=torch.complex128, device="cuda"))
    assert k0.shape == (0, 4, 4, 3) and st0.numel() == 0
