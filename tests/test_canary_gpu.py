"""Out-of-bounds guard for the main entry points (compute-sanitizer is closed on this GPU pool, profiles/README.md):
every output and workspace is allocated inside a larger buffer whose margins hold a sentinel; a kernel that writes one
element past its array, or uses more workspace than its *_ws_bytes query promises, trips the check."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GUARD = 4096          # bytes on each side
SENT = 0x5A


class Guarded:
    """`nbytes` usable bytes with sentinel margins, 256-byte aligned like a caller's own allocation."""

    def __init__(self, torch, nbytes):
        self.torch = torch
        self.n = int(nbytes)
        self.raw = torch.full((self.n + 2 * GUARD,), SENT, dtype=torch.uint8, device="cuda")
        self.ptr = self.raw.data_ptr() + GUARD

    def view(self, dtype, shape):
        t = self.raw[GUARD:GUARD + self.n].view(dtype)
        return t.view(shape)

    def intact(self):
        a = self.raw[:GUARD]
        b = self.raw[GUARD + self.n:]
        return bool((a == SENT).all().item()) and bool((b == SENT).all().item())


@pytest.fixture(scope="module")
def env():
    import torch
    from hyperscanning_signal_analysis_b200 import _lib
    return torch, _lib, _lib.load()


@pytest.mark.parametrize("m,n,p,F,n_win", [(38, 512, 8, 64, 7), (4, 160, 5, 30, 3), (19, 300, 3, 17, 5), (48, 256, 2, 9, 2)])
def test_mvar_path_stays_inside_its_buffers(env, m, n, p, F, n_win):
    torch, _lib, lib = env
    rng = np.random.default_rng(m * 1000 + n)
    T = n + 64 * (n_win - 1)
    x = rng.standard_normal((m, T))
    x[:, 1:] += 0.5 * x[:, :-1]
    xd = torch.from_numpy(x).cuda()
    starts = torch.arange(n_win, dtype=torch.int64, device="cuda") * 64
    freqs = torch.linspace(0.0, 100.0, F, dtype=torch.float64, device="cuda")
    out = Guarded(torch, n_win * m * m * F * 8)
    A = Guarded(torch, n_win * m * m * p * 8)
    V = Guarded(torch, n_win * m * m * 8)
    status = Guarded(torch, n_win * 4)
    ws = Guarded(torch, lib.hs_mvar_ffdtf_ws_bytes(n_win, m, p, F))
    status.view(torch.int32, (n_win,)).zero_()
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.hs_mvar_ffdtf_f64(xd.data_ptr(), starts.data_ptr(), T, n_win, m, n, p, freqs.data_ptr(), F, 256.0, out.ptr, A.ptr, V.ptr,
                                     status.ptr, ws.ptr, st), "hs_mvar_ffdtf_f64")
    torch.cuda.synchronize()
    for g in (out, A, V, status, ws):
        assert g.intact()
    ff = out.view(torch.float64, (n_win, m, m, F))
    assert float((ff.sum(dim=(2, 3)) - 1).abs().max()) < 1e-9
    # stage-wise entry points with H / A(f) / dtf outputs and the criterion kernel
    R = Guarded(torch, n_win * (p + 1) * m * m * 8)
    _lib.check(lib.hs_lagcov_f64(xd.data_ptr(), starts.data_ptr(), T, n_win, 1, m, n, p, R.ptr, st), "lagcov")
    Vall = Guarded(torch, n_win * p * m * m * 8)
    yws = Guarded(torch, lib.hs_yw_ws_bytes(n_win, m, p))
    _lib.check(lib.hs_yw_solve_f64(R.ptr, n_win, m, p, A.ptr, V.ptr, Vall.ptr, status.ptr, yws.ptr, st), "yw")
    crit = Guarded(torch, n_win * p * 8)
    popt = Guarded(torch, n_win * 4)
    _lib.check(lib.hs_mvar_criterion_f64(Vall.ptr, n_win, p, m, n, 0, crit.ptr, None, popt.ptr, st), "criterion")
    H = Guarded(torch, n_win * m * m * F * 16)
    Af = Guarded(torch, n_win * m * m * F * 16)
    dtf = Guarded(torch, n_win * m * m * F * 8)
    tws = Guarded(torch, lib.hs_transfer_ws_bytes(n_win, m, p, F))
    _lib.check(lib.hs_transfer_dtf_f64(A.ptr, freqs.data_ptr(), F, 256.0, n_win, m, p, H.ptr, Af.ptr, dtf.ptr, out.ptr, status.ptr, tws.ptr, st), "transfer")
    S = Guarded(torch, n_win * m * m * F * 16)
    _lib.check(lib.hs_spectra_f64(H.ptr, V.ptr, n_win, m, F, S.ptr, st), "spectra")
    torch.cuda.synchronize()
    for g in (R, Vall, yws, crit, popt, H, Af, dtf, tws, S, out, A, V, status):
        assert g.intact()


@pytest.mark.parametrize("n_sig,n", [(3, 20000), (5, 8192 + 18), (2, 9000), (4, 700)])
def test_front_end_stays_inside_its_buffers(env, n_sig, n):
    torch, _lib, lib = env
    from scipy import signal
    rng = np.random.default_rng(n)
    x = Guarded(torch, n_sig * n * 8)
    xv = x.view(torch.float64, (n_sig, n))
    xv.copy_(torch.from_numpy(rng.standard_normal((n_sig, n)) * 20 + 3))
    ref_in = xv.cpu().numpy().copy()
    ws = Guarded(torch, lib.hs_filtfilt_ws_bytes(n_sig, n))
    filt = [signal.iirnotch(50.0, 30.0, fs=1024.0), signal.butter(2, 64.0, "low", fs=1024.0), signal.butter(2, 1.0, "high", fs=1024.0)]
    b = np.ascontiguousarray(np.stack([f[0] for f in filt]))
    a = np.ascontiguousarray(np.stack([f[1] for f in filt]))
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.hs_iir_filtfilt_f64(x.ptr, n_sig, n, n, 1, b.ctypes.data, a.ctypes.data, 3, 3, 1, ws.ptr, st), "filtfilt")
    torch.cuda.synchronize()
    assert x.intact() and ws.intact()
    y = ref_in - ref_in.mean(axis=1, keepdims=True)
    for bb, aa in filt:
        y = signal.filtfilt(bb, aa, y, axis=1)
    assert float(np.max(np.abs(xv.cpu().numpy() - y)) / np.max(np.abs(y))) < 1e-9
    q = 8
    taps = torch.from_numpy(signal.firwin(20 * q + 1, 1.0 / q, window="hamming")).cuda()
    n_out = (n + q - 1) // q
    dec = Guarded(torch, n_sig * n_out * 8)
    _lib.check(lib.hs_fir_decimate_f64(x.ptr, n_sig, n, n, q, taps.data_ptr(), taps.numel(), dec.ptr, n_out, st), "decimate")
    torch.cuda.synchronize()
    assert dec.intact() and x.intact()


@pytest.mark.parametrize("n,K", [(8192, 7), (4096, 4), (1000, 5), (4099, 3), (23040, 3)])
def test_psd_stays_inside_its_buffers(env, n, K):
    torch, _lib, lib = env
    from scipy.signal.windows import dpss
    n_sig = 3
    rng = np.random.default_rng(n + K)
    x = torch.from_numpy(rng.standard_normal((n_sig, n))).cuda()
    tapers, eig = dpss(n, 4.0, K, sym=False, norm=2, return_ratios=True)
    t_dev = torch.from_numpy(np.ascontiguousarray(tapers)).cuda()
    w_dev = torch.from_numpy(np.sqrt(eig)).cuda()
    k_lo, k_hi = 1, n // 2 + 1
    psd = Guarded(torch, n_sig * (k_hi - k_lo) * 8)
    ws = Guarded(torch, lib.hs_mt_psd_ws_bytes(n_sig, n, K))
    _lib.check(lib.hs_mt_psd_f64(x.data_ptr(), n_sig, n, t_dev.data_ptr(), w_dev.data_ptr(), K, k_lo, k_hi, psd.ptr, ws.ptr,
                                 torch.cuda.current_stream().cuda_stream), "psd")
    torch.cuda.synchronize()
    assert psd.intact() and ws.intact()
    got = psd.view(torch.float64, (n_sig, k_hi - k_lo)).cpu().numpy()
    X = np.fft.rfft((x.cpu().numpy() - x.cpu().numpy().mean(axis=1, keepdims=True))[:, None, :] * tapers[None], axis=-1)
    w2 = eig[None, :, None]
    ref = (np.abs(X) ** 2 * w2).sum(axis=1) * 2.0 / eig.sum()
    if n % 2 == 0:
        ref[:, -1] *= 0.5
    assert float(np.max(np.abs(got - ref[:, k_lo:k_hi])) / np.max(np.abs(ref))) < 1e-9


@pytest.mark.parametrize("n_sig,n,q,ntaps,off,shift", [
    (5, 5000, 8, 161, 80, 0), (13, 4099, 8, 161, 80, 0), (38, 3000, 4, 81, 40, 0), (8, 2048, 2, 41, 20, 0),
    (6, 1500, 1, 201, 0, 0), (7, 1500, 1, 201, 100, 0), (9, 3001, 8, 160, 79, 0), (5, 700, 8, 161, 80, 0),
    (16, 100, 8, 161, 80, 0), (5, 20000, 8, 161, 80, 1), (11, 70000, 8, 161, 80, 0), (8, 9000, 1, 217, 3, 0)])
def test_fir_on_the_tensor_pipe_matches_direct_evaluation(env, n_sig, n, q, ntaps, off, shift):
    """hs_fir_filter_f64 with >= 5 signals and 7 q + ntaps <= 224 runs fir_mma_kernel (Toeplitz products, cp.async staging): every
    staging path (16-byte interior copies, 8-byte copies for odd strides / unaligned bases / even tap counts, zero-filled edges,
    missing signals of the last group of 8, signals shorter than the filter) against y[k] = sum_j b[j] x[q k + off - j], inside
    sentinel-guarded buffers."""
    torch, _lib, lib = env
    rng = np.random.default_rng(n_sig * 1000 + n + ntaps)
    xbuf = Guarded(torch, (n_sig * n + shift) * 8)
    xall = xbuf.view(torch.float64, (n_sig * n + shift,))
    xall.copy_(torch.from_numpy(rng.standard_normal(n_sig * n + shift) * 10 + 1))
    xh = xall[shift:].cpu().numpy().reshape(n_sig, n)
    b = rng.standard_normal(ntaps)
    b_d = torch.from_numpy(b).cuda()
    n_out = (n + q - 1) // q if off < n else 1
    n_out = max(1, min(n_out, (n - 1 + ntaps - 1 - off) // q + 1)) if off <= n - 1 + ntaps - 1 else 1
    ybuf = Guarded(torch, n_sig * n_out * 8)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.hs_fir_filter_f64(xbuf.ptr + 8 * shift, n_sig, n, n, q, off, b_d.data_ptr(), ntaps, ybuf.ptr, n_out, n_out, st), "fir")
    torch.cuda.synchronize()
    assert ybuf.intact() and xbuf.intact()
    got = ybuf.view(torch.float64, (n_sig, n_out)).cpu().numpy()
    ref = np.zeros((n_sig, n_out))
    for s in range(n_sig):
        full = np.convolve(xh[s], b)                     # full[i] = sum_j b[j] x[i - j]
        idx = q * np.arange(n_out) + off
        ok = idx < full.size
        ref[s, ok] = full[idx[ok]]
    assert float(np.max(np.abs(got - ref)) / np.max(np.abs(ref))) < 1e-12
