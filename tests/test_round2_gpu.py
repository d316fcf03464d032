"""Round-2 coverage on the GPU, against vectors written by the reference's own code (tests/golden/round2.npz):
count_corr(iwhat=2), the causal lfilter branch with recursive coefficients, batched model-order criteria, and the
EEG_IBI_FFDTF_Pipeline class used through an INSTANCE (methods, not free functions); plus element-wise ffDTF parity."""
import contextlib
import io

import numpy as np
import pytest
from scipy import signal

from conftest import TOL_MODEL, TOL_SIGNAL, golden, relerr
from test_oracle_cpu import _prewindow_inputs, elementwise_relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mv():
    from hyperscanning_signal_analysis_b200 import mtmvar
    return mtmvar


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def test_count_corr_unbiased_variant(mv):
    g = golden("round2.npz")
    m4 = golden("mvar_m4.npz")
    rl, rr, r0 = mv.count_corr(m4["x"][:, :, None], int(m4["p"]), 2)
    assert relerr(rl, g["m4_iwhat2_left"]) < TOL_SIGNAL and relerr(rr, g["m4_iwhat2_right"]) < TOL_SIGNAL
    assert relerr(r0, g["m4_iwhat2_zero"]) < TOL_SIGNAL
    tr = golden("mvar_trials.npz")
    rl, rr, r0 = mv.count_corr(tr["x"], int(tr["p"]), 2)
    assert relerr(rl, g["tr_iwhat2_left"]) < TOL_SIGNAL and relerr(rr, g["tr_iwhat2_right"]) < TOL_SIGNAL
    with pytest.raises(ValueError):
        mv.count_corr(m4["x"], 2, 3)


def test_causal_branch_with_recursive_lowpass_highpass():
    """filter_type='butter': not 'iir', so the reference runs the causal lfilter chain + delay roll (dataloader.py:793-801)."""
    from hyperscanning_signal_analysis_b200 import dataloader
    from test_frontend_gpu import _md
    g = golden("round2.npz")
    md = _md(256.0, 3)
    fb = dataloader._design_eeg_filters(md, lowcut=1.0, highcut=40.0, filter_type="butter")
    assert fb[3] == "butter" and np.size(fb[1][1]) == 3
    buf = g["butter_raw"].copy()
    quiet(dataloader._apply_filters, md, fb, buf)
    assert relerr(buf, g["butter_out"]) < TOL_SIGNAL
    assert np.all(buf[:, -2:] == 0.0)              # delay = 1 + 1 samples zeroed at the tail


def test_batched_order_criteria(mv):
    import torch
    g = golden("round2.npz")
    sig = golden("prewindow.npz")["signals_to_ffDTF"]
    from hyperscanning_signal_analysis_b200.eeg_alpha_ibi_ffdtf import window_starts
    x = torch.from_numpy(np.ascontiguousarray(sig)).cuda()
    for tag, nw, ws in (("w3", 3, None), ("w5", 5, 200)):
        starts, W = window_starts(sig.shape[1], nw, ws)
        st = torch.from_numpy(starts).cuda()
        for c in ("AIC", "HQ", "SC"):
            crit, popt, status = mv.batched_mvar_criterion(x, st, sig.shape[1], nw, 4, W, 20, c)
            assert int(status.max()) == 0
            assert relerr(crit.cpu().numpy(), g[f"{tag}_crit_{c}"]) < TOL_MODEL
            assert popt.cpu().numpy().tolist() == g[f"{tag}_popt_{c}"].tolist()
            # the per-call reference signature agrees with the batch
            c0, rng, p0 = mv.mvar_criterion(sig[:, starts[0]:starts[0] + W], 20, c)
            assert np.array_equal(c0, crit[0].cpu().numpy()) and int(p0) == int(popt[0]) and rng.tolist() == list(range(1, 21))
    # m = 38: orders 1..8 are well posed (tight), the nearly singular high orders to what conditioning allows
    w38 = golden("mvar_cfg2_windows.npz")["windows"]
    for k in range(w38.shape[0]):
        crit = mv.mvar_criterion(w38[k], 12, "AIC")[0]
        assert relerr(crit[:8], g["w38_crit_AIC"][k][:8]) < TOL_MODEL
        assert relerr(crit, g["w38_crit_AIC"][k]) < 1e-4
    with pytest.raises(ValueError):
        mv.mvar_criterion(sig, 3, "BIC")


def test_pipeline_class_methods_on_an_instance():
    """Call sites written against the reference use the methods of an EEG_IBI_FFDTF_Pipeline instance."""
    from hyperscanning_signal_analysis_b200.eeg_alpha_ibi_ffdtf import EEG_IBI_FFDTF_Pipeline
    g = golden("round2.npz")
    pw, x = _prewindow_inputs()
    fs = float(pw["fs"])
    names = [str(s) for s in pw["names"]]
    pipe = EEG_IBI_FFDTF_Pipeline(None, None, ["M1"], fs_downsampled=8.0, n_windows=3, window_size=None, ar_p=5)
    assert (pipe.left_chan, pipe.right_chan, pipe.fs_ds, pipe.freq_min, pipe.freq_step) == ("F3", "F4", 8.0, 1.0, 0.1)
    assert abs(pipe.freq_max - 3.9) < 1e-12
    rows = []
    for who, sl, k in (("ch", slice(0, 19), 0), ("cg", slice(19, 38), 1)):
        filt = pipe._alpha_bandpass_filter(x[sl], fs)
        assert relerr(filt[[3, 5, 18]], pw[f"filt_{who}"]) < TOL_SIGNAL
        faa = pipe._compute_asymmetry(filt, names, metric="amp")
        assert relerr(faa, pw[f"faa_{who}"]) < 1e-7
        faa_ds = pipe._downsample_signal(faa, fs, 8.0)
        ibi_ds = pipe._downsample_signal(pw["ibi"][k], fs, 8.0)
        assert relerr(faa_ds, pw[f"faa_ds_{who}"]) < 1e-7 and relerr(ibi_ds, pw[f"ibi_ds_{who}"]) < TOL_SIGNAL
        rows += [pipe._crop_signal(faa_ds, 8.0, 10, 60), pipe._crop_signal(ibi_ds, 8.0, 10, 60)]
    sig = np.vstack(rows)
    sig = (sig - np.mean(sig, axis=1, keepdims=True)) / np.std(sig, axis=1, keepdims=True)
    assert relerr(sig, pw["signals_to_ffDTF"]) < 1e-6
    sig = pw["signals_to_ffDTF"]                   # continue from the reference's own array: isolates the MVAR stage
    chan = ["faa_ch", "ibi_ch", "faa_cg", "ibi_cg"]
    for tag, nw, ws in (("w3", 3, None), ("w5", 5, 200)):
        pipe.n_windows, pipe.window_size = nw, ws
        wins = pipe._create_windows(sig, nw, ws)
        assert len(wins) == nw and all(w.base is not None for w in wins)          # views, like the reference
        for ar_p, ptag in ((5, "p5"), (None, "auto")):
            pipe.ar_p = ar_p
            for k, w in enumerate(wins):
                ff, sp, po = pipe._compute_ffDTF("D", w, chan, 8.0, plot=False, save_plot=False)
                assert ff.shape == (4, 4, 30) and sp.dtype == np.complex128
                assert int(po) == int(g[f"{tag}_{ptag}_popt"][k])
                assert relerr(ff, g[f"{tag}_{ptag}_ffdtf"][k]) < TOL_MODEL
                assert elementwise_relerr(ff, g[f"{tag}_{ptag}_ffdtf"][k]) < 1e-6
                assert relerr(sp, g[f"{tag}_{ptag}_spectra"][k]) < TOL_MODEL
            # the same windows as ONE batched call
            ffw, spw, pw_ = pipe.compute_windows(sig)
            assert [int(v) for v in pw_] == g[f"{tag}_{ptag}_popt"].tolist()
            assert relerr(np.stack(ffw), g[f"{tag}_{ptag}_ffdtf"]) < TOL_MODEL
            assert relerr(np.stack(spw), g[f"{tag}_{ptag}_spectra"]) < TOL_MODEL
    pipe.ar_p = 5
    ffg, spg, pg = pipe._compute_ffDTF("D", sig, chan, 8.0, plot=False)
    assert pg == 5 and relerr(ffg, g["global_p5_ffdtf"]) < TOL_MODEL and relerr(spg, g["global_p5_spectra"]) < TOL_MODEL
    with pytest.raises(ValueError):
        pipe._create_windows(sig, 3, 100)
    with pytest.raises(ValueError):
        pipe._crop_signal(sig, 8.0, 10, 600)
    with pytest.raises(ValueError):
        pipe._compute_asymmetry(sig, ["a", "b", "c", "d"])
    with pytest.raises(RuntimeError):
        pipe.run_pipeline()                        # no files were scanned: same error as the reference (:662-663)


def test_process_dyad_builds_the_reference_result(tmp_path):
    from hyperscanning_signal_analysis_b200.eeg_alpha_ibi_ffdtf import EEG_IBI_FFDTF_Pipeline
    g = golden("round2.npz")
    pw, x = _prewindow_inputs()
    names = [str(s) for s in pw["names"]]
    pipe = EEG_IBI_FFDTF_Pipeline(None, tmp_path, ["M1"], n_windows=3, ar_p=5)
    res = quiet(pipe.process_dyad, "W_001", "M1", x[:19], pw["ibi"][0][:, None], x[19:], pw["ibi"][1][:, None], float(pw["fs"]), float(pw["fs"]), names)
    assert sorted(res["mvar"]) == ["ff_dtf_global", "ff_dtf_windowed", "p_opt_g", "p_opt_w", "spectra_global", "spectra_windowed"]
    assert relerr(np.stack(res["mvar"]["ff_dtf_windowed"]), g["w3_p5_ffdtf"]) < 1e-5      # whole chain from raw EEG (FAA through a log)
    path = quiet(pipe._save_single_result, "W_001", "M1", res)
    z = np.load(path)
    assert sorted(z.files) == ["ff_dtf_global", "ff_dtf_windowed", "meta", "p_opt_g", "p_opt_w", "spectra_global", "spectra_windowed"]
    assert z["ff_dtf_windowed"].shape == (3, 4, 4, 30)


def test_run_pipeline_from_netcdf_files(tmp_path):
    """run_pipeline (src/eeg_alpha_ibi_ffdtf.py:661-790) end to end: files in the on-disk container (export.write_netcdf3) ->
    discovery -> _load_eeg_and_ibi -> pre-window stage -> windows -> ffDTF -> the reference's .npz; same numbers as process_dyad."""
    from hyperscanning_signal_analysis_b200 import export
    from hyperscanning_signal_analysis_b200.eeg_alpha_ibi_ffdtf import EEG_IBI_FFDTF_Pipeline
    g = golden("round2.npz")
    pw, x = _prewindow_inputs()
    names = [str(s) for s in pw["names"]]
    fs = float(pw["fs"])
    t = np.arange(x.shape[1]) / fs
    for role, eeg, ibi in (("ch", x[:19], pw["ibi"][0]), ("cg", x[19:], pw["ibi"][1])):
        for kind, arr, chans in (("EEG", eeg.T, names), ("IBI", ibi[:, None], ["IBI"])):
            folder = tmp_path / "in" / kind / "W_001"
            folder.mkdir(parents=True, exist_ok=True)
            export.write_netcdf3(folder / f"W_001_{kind}_{role}_M1.nc", arr, t, chans, {"sampling_freq": fs, "event_duration_s": float(t[-1]), "who": role})
    pipe = EEG_IBI_FFDTF_Pipeline(tmp_path / "in", tmp_path / "out", ["M1"], n_windows=3, ar_p=5)
    quiet(pipe.run_pipeline)
    z = np.load(tmp_path / "out" / "W_001" / "W_001_M1_ffDTF.npz")
    assert z["ff_dtf_windowed"].shape == (3, 4, 4, 30)
    assert relerr(z["ff_dtf_windowed"], g["w3_p5_ffdtf"]) < 1e-5


def _transfer_with_flag_count(A, freqs, fs):
    """hs_transfer_dtf_f64 through the C ABI on one window: (|H|^2 (m, m, F), number of matrices the optimistic pass flagged)."""
    import torch
    from hyperscanning_signal_analysis_b200 import _lib
    lib = _lib.load()
    m, _, p = A.shape
    F = len(freqs)
    Ad = torch.from_numpy(np.ascontiguousarray(A[None])).cuda()
    fr = torch.from_numpy(np.ascontiguousarray(freqs, dtype=np.float64)).cuda()
    out = torch.empty((1, m, m, F), dtype=torch.float64, device="cuda")
    st = torch.zeros(1, dtype=torch.int32, device="cuda")
    ws = torch.zeros(lib.hs_transfer_ws_bytes(1, m, p, F), dtype=torch.uint8, device="cuda")
    sp = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.hs_transfer_dtf_f64(Ad.data_ptr(), fr.data_ptr(), F, float(fs), 1, m, p, None, None, out.data_ptr(), None, st.data_ptr(),
                                       ws.data_ptr(), sp), "transfer")
    torch.cuda.synchronize()
    off = lib.hs_transfer_ws_flag_offset(1, m, p, F)
    return out[0].cpu().numpy(), int(ws[off:off + 4].view(torch.int32).item())


@pytest.mark.parametrize("m", [7, 8, 28, 29, 31, 32, 33, 35, 36, 37, 39, 40])
def test_a_posteriori_check_at_every_padding_geometry(m):
    """The optimistic elimination verifies every matrix through the padding column when there is one (m < 8 T: the column carries
    A(f) u; it lies inside the last eliminated block for 8 T - 4 < m and outside it otherwise) and through an explicit H v product
    when m = 8 T.  Benign matrices must not be flagged, a matrix that needs pivoting must be -- and be repaired -- in each case."""
    from oracle import mvar_oracle as mo
    rng = np.random.default_rng(400 + m)
    p, fs = 3, 256.0
    freqs = np.array([0.0, 5.0, 31.0, 64.0, 100.0, 127.0])
    A = 0.1 * rng.standard_normal((m, m, p)) / np.sqrt(m)                       # A(f) close to I: no pivoting needed
    Hr, _ = mo.mvar_transfer_function(A, freqs, fs)
    dtf, flagged = _transfer_with_flag_count(A, freqs, fs)
    assert flagged == 0
    assert relerr(dtf, np.abs(Hr) ** 2) < TOL_MODEL
    B = np.zeros((m, m, p))                          # A(0) = a cyclic shift exactly: the first pivot block is singular, the unpivoted pass must fail
    B[:, :, 0] = np.eye(m) - np.roll(np.eye(m), 1, axis=1)
    Hb, _ = mo.mvar_transfer_function(B, freqs, fs)
    dtf_b, flagged_b = _transfer_with_flag_count(B, freqs, fs)
    assert 1 <= flagged_b <= len(freqs)
    assert relerr(dtf_b, np.abs(Hb) ** 2) < TOL_MODEL


@pytest.mark.parametrize("m,p", [(6, 2), (38, 8)])
def test_more_windows_than_resident_slots(mv, m, p):
    """More windows than the Yule-Walker kernel has resident CTAs (5 per SM: 740 on a B200): CTAs then walk several windows through
    the same scratch and panels.  Sampled windows against the oracle, all of them through the row sums."""
    from oracle import mvar_oracle as mo
    rng = np.random.default_rng(900 + m)
    n, hop, n_win = 64 if m == 6 else 512, 16 if m == 6 else 64, 1700 if m == 6 else 1000
    T = n + hop * (n_win - 1)
    x = rng.standard_normal((m, T))
    x[:, 1:] += 0.6 * x[:, :-1]
    x += 0.1 * rng.standard_normal((m, m)) @ x
    starts = np.arange(n_win) * hop
    freqs = np.linspace(0.0, 64.0, 12, endpoint=False)
    ff, A, V = mv.windowed_ffdtf(x, starts, n, freqs, 128.0, p, return_model=True)
    ff = ff.cpu().numpy()
    np.testing.assert_allclose(ff.sum(axis=(2, 3)), 1.0, rtol=1e-11)
    for w in (0, 739, 740, 741, n_win - 1):
        Ar, Vr = mo.ar_coeff(x[:, starts[w]:starts[w] + n], p)
        assert relerr(A[w].cpu().numpy(), Ar) < TOL_MODEL and relerr(V[w].cpu().numpy(), Vr) < TOL_MODEL, w
        ref = mo.full_freq_dtf(x[:, starts[w]:starts[w] + n], freqs, 128.0, optimal_model_order=p)
        assert relerr(ff[w], ref) < TOL_MODEL, w


@pytest.mark.parametrize("n_win", [1, 7, 75, 149, 160, 222, 296, 303])
def test_lagcov_last_wave_split_is_bit_identical(mv, n_win):
    """K3 splits the windows of the last, partial wave of its grid over several CTAs (1 ... 15 parts, depending on how many windows
    are left over after the full waves of one CTA per SM).  Every part computes the same tiles with the same instruction sequence
    as a whole-window CTA: R of a window must not depend on how many other windows the launch holds -- bit for bit -- and must be
    the reference's biased lag covariance (mtmvar.py:57-59, 72-73)."""
    import torch
    m, n, p, hop = 38, 512, 8, 64
    rng = np.random.default_rng(123)
    x = rng.standard_normal((m, hop * 303 + n))
    xd = torch.from_numpy(x).cuda()
    offs = torch.arange(n_win, dtype=torch.int64, device="cuda") * hop
    R = mv.batched_lagcov(xd, offs, x.shape[1], n_win, 1, m, n, p)
    for w in sorted({0, n_win // 2, max(n_win - 8, 0), n_win - 1}):
        alone = mv.batched_lagcov(xd, offs[w:w + 1].clone(), x.shape[1], 1, 1, m, n, p)
        assert torch.equal(R[w], alone[0]), w
    w = n_win - 1
    xw = x[:, w * hop:w * hop + n]
    ref = np.stack([xw[:, :n - L] @ xw[:, L:].T / n for L in range(p + 1)])
    assert relerr(R[w].cpu().numpy(), ref) < 1e-12


@pytest.mark.parametrize("m,F", [(41, 3), (48, 4), (64, 2)])
def test_partial_coherence_more_than_40_channels(mv, m, F):
    """partial_coherence / dDTF beyond the register-tile limit (pcoh_generic_kernel: pivoted Gauss-Jordan in place in the output, the
    phase of the determinant instead of its value) against the reference's minor determinants (oracle, mtmvar.py:300-321)."""
    import torch
    from oracle import mvar_oracle as mo
    rng = np.random.default_rng(70 + m)
    H = rng.standard_normal((m, m, F)) + 1j * rng.standard_normal((m, m, F))
    V = rng.standard_normal((m, m))
    V = V @ V.T + m * np.eye(m)
    S = np.stack([H[:, :, f] @ (V @ H[:, :, f].T) for f in range(F)], axis=2) * 1e-9       # the reference's S = H V H^T, in "volts"
    # kappa is invariant under S -> c S.  The reference's minors are ~1e-225 at this scale and the PRODUCT of two underflows to 0 (it then
    # returns kappa = 0 everywhere off the diagonal: INTEGRATION.md, known differences); the oracle is therefore run on a copy of S scaled
    # to unit geometric-mean singular value per bin, the GPU path on the data as they are.
    gm = np.array([np.exp(np.mean(np.log(np.linalg.svd(S[:, :, f], compute_uv=False)))) for f in range(F)])
    ref = mo.partial_coherence(S / gm)
    got = mv.partial_coherence(S)
    assert got.shape == ref.shape and np.all(got[np.arange(m), np.arange(m), :] == 1.0)
    scale = max(1.0, np.max(np.abs(ref)))                       # S = H V H^T (plain transpose) is not Hermitian: |kappa| may exceed 1
    assert np.max(np.abs(got - ref)) < 1e-8 * scale              # det of (m-1) x (m-1) minors on the CPU side
    ff = rng.uniform(0.0, 1.0, (1, m, m, F))
    kap, dd, status = mv.batched_partial_coherence(torch.from_numpy(S[None]).cuda(), torch.from_numpy(ff).cuda())
    assert int(status.max()) == 0
    assert np.max(np.abs(dd[0].cpu().numpy() - ff[0] * np.abs(ref))) < 1e-8 * scale
    Ssing = S.copy()
    Ssing[:, 5, :] = 0.0                                          # a zero column stays exactly zero through the elimination: zero pivot
    with pytest.raises(np.linalg.LinAlgError):
        mv.partial_coherence(Ssing)


def test_ffdtf_elementwise_above_floor(mv):
    """Norm-wise 1e-7 leaves small entries unchecked (ffDTF spans many decades): element-wise check above 1e-6 max."""
    g = golden("mvar_cfg2_windows.npz")
    worst = 0.0
    for w in range(g["windows"].shape[0]):
        ff = quiet(mv.full_freq_dtf, g["windows"][w], g["freqs"], 256.0, optimal_model_order=8)
        worst = max(worst, elementwise_relerr(ff, g["ffdtf"][w], 1e-6))
    print("max element-wise relative error of ffDTF above 1e-6*max:", worst)
    assert worst < 1e-5        # cond(G) ~ 3e7: an entry 1e-6 of the maximum keeps ~5 digits of the 1e-11 norm-wise agreement
    m4 = golden("mvar_m4.npz")
    ff4 = quiet(mv.full_freq_dtf, m4["x"], m4["freqs"], float(m4["fs"]), optimal_model_order=int(m4["p"]))
    assert elementwise_relerr(ff4, m4["ffdtf"], 1e-9) < 1e-9


def test_downsample_two_dimensional_is_axis_zero():
    """resample_poly's default axis is 0 (the reference passes none, eeg_alpha_ibi_ffdtf.py:403)."""
    from hyperscanning_signal_analysis_b200.eeg_alpha_ibi_ffdtf import downsample_signal
    rng = np.random.default_rng(3)
    x = rng.standard_normal((640, 3))
    got = downsample_signal(x, 128, 8)
    ref = signal.resample_poly(x, up=1, down=16)
    assert got.shape == ref.shape == (40, 3) and relerr(got, ref) < TOL_SIGNAL


@pytest.mark.parametrize("which", [1, 2, 3, 4])
def test_both_transfer_kernels_against_reference(mv, which):
    """transfer_mma_kernel (6 groups of Re/Im warps) and transfer_ws_kernel (4 groups of Re/Im/helper warps) deliver the same
    H / ffDTF: reference goldens at m = 4 and m = 38, odd shapes against the oracle, and the forced-pivot fallback."""
    import torch
    from hyperscanning_signal_analysis_b200 import _lib
    from oracle import mvar_oracle as mo
    lib = _lib.load()
    if lib.hs_transfer_set_kernel(which) != 0:
        assert which >= 2 and "HS_EXPERIMENT" in _lib.last_error()
        pytest.skip("warp-specialised kernel: HS_EXPERIMENT builds only")
    try:
        g = golden("mvar_cfg2_windows.npz")
        for w in range(g["windows"].shape[0]):
            ff = quiet(mv.full_freq_dtf, g["windows"][w], g["freqs"], 256.0, optimal_model_order=8)
            assert relerr(ff, g["ffdtf"][w]) < TOL_MODEL
        m4 = golden("mvar_m4.npz")
        H, Af = mv.mvar_transfer_function(m4["A"], m4["freqs"], float(m4["fs"]))
        assert relerr(H, m4["H"]) < TOL_MODEL and relerr(Af, m4["Af"]) < 1e-12
        rng = np.random.default_rng(17)
        for m, p, F in ((9, 3, 11), (17, 5, 40), (33, 2, 7), (38, 8, 130), (40, 4, 64)):
            A = rng.standard_normal((m, m, p)) * (0.4 / np.sqrt(m))
            freqs = np.linspace(0.0, 100.0, F)
            H, _ = mv.mvar_transfer_function(A, freqs, 256.0)
            Hr, _ = mo.mvar_transfer_function(A, freqs, 256.0)
            assert relerr(H, Hr) < TOL_MODEL, (m, p, F)
        # zero diagonal at f = 0: the optimistic pass must flag, the pivoted redo must deliver
        m, p = 38, 2
        Pm = np.roll(np.eye(m), 1, axis=1)
        A = np.zeros((m, m, p))
        A[:, :, 0] = np.eye(m) - Pm
        A[:, :, 1] = 0.01 * rng.standard_normal((m, m))
        freqs = np.array([0.0, 3.0, 17.0, 40.0, 64.0, 90.0, 127.5])
        H, _ = mv.mvar_transfer_function(A, freqs, 256.0)
        Hr, _ = mo.mvar_transfer_function(A, freqs, 256.0)
        for fi in range(len(freqs)):
            assert relerr(H[:, :, fi], Hr[:, :, fi]) < TOL_MODEL, fi
        # many windows: row sums and determinism
        from hyperscanning_signal_analysis_b200 import mtmvar, synth
        x = synth.dyad_eeg(seed=5, n_samples=8192, line_amp=0.0)
        starts = np.arange(0, 8192 - 512 + 1, 128)
        freqs = np.linspace(0, 128, 256, endpoint=False)
        a = mtmvar.windowed_ffdtf(x, starts, 512, freqs, 256.0, 8)
        b = mtmvar.windowed_ffdtf(x, starts, 512, freqs, 256.0, 8)
        assert torch.equal(a, b)
        assert float((a.sum(dim=(2, 3)) - 1).abs().max()) < 1e-9
    finally:
        lib.hs_transfer_set_kernel(0)
    assert lib.hs_transfer_set_kernel(7) != 0 and lib.hs_transfer_set_kernel(-1) != 0


def test_partial_coherence_is_scale_invariant(mv):
    """Unscaled data (volts): det S underflows (|S_ii| ~ 1e-12, m = 38) unless S is pre-scaled by its diagonal; kappa must not change."""
    g = golden("mvar_pcoh.npz")
    S = g["w0_S"]
    k1 = mv.partial_coherence(S)
    assert relerr(k1, g["w0_kappa"]) < TOL_MODEL
    for scale in (1e-12, 1e+9):
        ks = mv.partial_coherence(S * scale)
        assert np.isfinite(ks).all()
        assert relerr(ks, k1) < 1e-9, scale          # S * scale rounds every entry: a relative 1e-16 perturbation times cond(S)
    # channel-wise scaling (different units per channel) leaves kappa unchanged as well
    d = np.logspace(-8, 6, S.shape[0])
    kd = mv.partial_coherence(S * d[:, None, None] * d[None, :, None])
    assert relerr(kd, k1) < 1e-9
