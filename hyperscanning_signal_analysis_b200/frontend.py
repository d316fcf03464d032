"""Array-level front end: zero-phase IIR filtering (K1) and FIR decimation (K2) on the GPU.
NumPy (or CUDA tensors) in, same kind out.  Filter *design* (iirnotch / butter / firwin) stays host-side
SciPy: it is O(taps), data independent, and the coefficients are stored verbatim in ``Filtration``."""
from __future__ import annotations

import numpy as np

from . import _lib


def _torch():
    return _lib.require_cuda()


def _pack_filters(filters):
    """[(b, a), ...] -> (n_filt, ntaps) float64 host arrays, zero padded; ntaps = max(len(a), len(b)) per SciPy."""
    bs, as_ = [], []
    for b, a in filters:
        b = np.atleast_1d(np.asarray(b, dtype=np.float64))
        a = np.atleast_1d(np.asarray(a, dtype=np.float64))
        bs.append(b)
        as_.append(a)
    groups = []       # consecutive filters with the same ntaps share one C call
    for b, a in zip(bs, as_):
        nt = max(len(a), len(b))
        bb = np.zeros(nt)
        aa = np.zeros(nt)
        bb[:len(b)] = b
        aa[:len(a)] = a
        if groups and groups[-1][0] == nt:
            groups[-1][1].append(bb)
            groups[-1][2].append(aa)
        else:
            groups.append((nt, [bb], [aa]))
    return [(nt, np.ascontiguousarray(np.stack(b)), np.ascontiguousarray(np.stack(a))) for nt, b, a in groups]


def filtfilt_cascade_(x_dev, filters, remove_dc=False, axis=-1):
    """In-place ``scipy.signal.filtfilt`` cascade on a 2-D CUDA float64 tensor.

    ``axis=-1``: rows are signals (dataloader layout, dataloader.py:787); ``axis=0``: columns are signals
    ((time, channel) layout of mne_bridge.py:161-184)."""
    torch = _torch()
    lib = _lib.load()
    assert x_dev.is_cuda and x_dev.dtype == torch.float64 and x_dev.dim() == 2 and x_dev.is_contiguous()
    if axis in (-1, 1):
        n_sig, n = x_dev.shape
        sig_stride, t_stride = n, 1
    else:
        n, n_sig = x_dev.shape
        sig_stride, t_stride = 1, n_sig
    ws = torch.empty(max(int(lib.hs_filtfilt_ws_bytes(n_sig, n)), 16), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    groups = _pack_filters(filters)
    if not groups:
        groups = [(2, np.zeros((0, 2)), np.zeros((0, 2)))]
    first = True
    for nt, b, a in groups:
        rc = lib.hs_iir_filtfilt_f64(x_dev.data_ptr(), n_sig, n, sig_stride, t_stride, b.ctypes.data, a.ctypes.data,
                                     b.shape[0], nt, int(remove_dc and first), ws.data_ptr(), stream)
        if rc != 0:
            msg = _lib.last_error()
            if "padlen" in msg:
                raise ValueError(msg)
            raise _lib.HsError(f"hs_iir_filtfilt_f64 failed ({rc}): {msg}")
        first = False
    return x_dev


def filtfilt_cascade(x, filters, remove_dc=False, axis=-1):
    """NumPy wrapper: returns a new float64 array."""
    torch = _torch()
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).cuda()
    filtfilt_cascade_(t, filters, remove_dc=remove_dc, axis=axis)
    return t.cpu().numpy()


def lfilter_fir_chain_dev(x_dev, notch, b_low, b_high, remove_dc=True):
    """FIR branch of the loader (dataloader.py:788, 793-801) on a (n_sig, n) CUDA float64 tensor; returns a new tensor:
    causal ``lfilter`` notch (IIR) -> ``lfilter`` low-pass FIR -> ``lfilter`` high-pass FIR, ``np.roll`` by the two FIR
    group delays and a zeroed tail.  The roll is folded into the last FIR pass (offset = delay)."""
    torch = _torch()
    lib = _lib.load()
    assert x_dev.is_cuda and x_dev.dtype == torch.float64 and x_dev.dim() == 2 and x_dev.is_contiguous()
    n_sig, n = x_dev.shape
    b_n = np.ascontiguousarray(np.atleast_1d(notch[0]), dtype=np.float64)
    a_n = np.ascontiguousarray(np.atleast_1d(notch[1]), dtype=np.float64)
    nt = max(len(a_n), len(b_n))
    bb = np.zeros(nt)
    aa = np.zeros(nt)
    bb[:len(b_n)] = b_n
    aa[:len(a_n)] = a_n
    stream = torch.cuda.current_stream().cuda_stream
    out = torch.zeros_like(x_dev)
    if n_sig == 0 or n == 0:
        return out
    ws = torch.empty(max(int(lib.hs_filtfilt_ws_bytes(n_sig, n)), 16), dtype=torch.uint8, device="cuda")
    y1 = torch.empty_like(x_dev)
    _lib.check(lib.hs_iir_lfilter_f64(x_dev.data_ptr(), n_sig, n, n, bb.ctypes.data, aa.ctypes.data, nt, int(remove_dc),
                                      y1.data_ptr(), n, ws.data_ptr(), stream), "hs_iir_lfilter_f64")
    bl = torch.from_numpy(np.ascontiguousarray(b_low, dtype=np.float64)).cuda()
    bh = torch.from_numpy(np.ascontiguousarray(b_high, dtype=np.float64)).cuda()
    y2 = torch.empty_like(x_dev)
    _lib.check(lib.hs_fir_filter_f64(y1.data_ptr(), n_sig, n, n, 1, 0, bl.data_ptr(), bl.numel(), y2.data_ptr(), n, n, stream),
               "hs_fir_filter_f64")
    delay = (bl.numel() - 1) // 2 + (bh.numel() - 1) // 2
    # np.roll(s, -delay); s[-delay:] = 0  ->  out[t] = s[t + delay] for t < n - delay, 0 after.  delay == 0 zeroes the whole
    # signal in the reference (s[-0:] is s[0:]), and so does delay >= n.
    if 0 < delay < n:
        _lib.check(lib.hs_fir_filter_f64(y2.data_ptr(), n_sig, n, n, 1, delay, bh.data_ptr(), bh.numel(), out.data_ptr(),
                                         n - delay, n, stream), "hs_fir_filter_f64")
    return out


def lfilter_iir_chain_dev(x_dev, filters, delay, remove_dc=True):
    """Causal branch of the loader for recursive filters (dataloader.py:794-801): ``lfilter`` with every (b, a) in turn (zero
    initial state), then ``np.roll(s, -delay)`` and ``s[-delay:] = 0``.  (n_sig, n) CUDA float64 in, new tensor out."""
    torch = _torch()
    lib = _lib.load()
    assert x_dev.is_cuda and x_dev.dtype == torch.float64 and x_dev.dim() == 2 and x_dev.is_contiguous()
    n_sig, n = x_dev.shape
    out = torch.zeros_like(x_dev)
    if n_sig == 0 or n == 0:
        return out
    stream = torch.cuda.current_stream().cuda_stream
    ws = torch.empty(max(int(lib.hs_filtfilt_ws_bytes(n_sig, n)), 16), dtype=torch.uint8, device="cuda")
    cur, nxt = x_dev, torch.empty_like(x_dev)
    for k, (b, a) in enumerate(filters):
        b = np.atleast_1d(np.asarray(b, dtype=np.float64))
        a = np.atleast_1d(np.asarray(a, dtype=np.float64))
        nt = max(len(a), len(b), 2)
        bb, aa = np.zeros(nt), np.zeros(nt)
        bb[:len(b)] = b
        aa[:len(a)] = a
        _lib.check(lib.hs_iir_lfilter_f64(cur.data_ptr(), n_sig, n, n, bb.ctypes.data, aa.ctypes.data, nt, int(remove_dc and k == 0),
                                          nxt.data_ptr(), n, ws.data_ptr(), stream), "hs_iir_lfilter_f64")
        cur, nxt = nxt, (torch.empty_like(x_dev) if cur is x_dev else cur)
    # delay == 0 zeroes the whole signal in the reference (s[-0:] is s[0:]), and so does delay >= n
    if 0 < delay < n:
        out[:, : n - delay] = cur[:, delay:]
    return out


def decimate_taps(q):
    """Anti-alias FIR of ``scipy.signal.decimate(..., ftype='fir')``: firwin(20 q + 1, 1/q, window='hamming')."""
    from scipy.signal import firwin
    return firwin(20 * q + 1, 1.0 / q, window="hamming")


_TAPS_DEV = {}      # (q, device) -> CUDA tensor of the default anti-alias taps (designed once per factor)


def decimate_dev(x_dev, q, taps=None):
    """(n_sig, n) CUDA float64 -> (n_sig, ceil(n/q)) CUDA float64."""
    torch = _torch()
    lib = _lib.load()
    assert x_dev.is_cuda and x_dev.dtype == torch.float64 and x_dev.dim() == 2 and x_dev.is_contiguous()
    n_sig, n = x_dev.shape
    if taps is None:
        key = (int(q), x_dev.device.index)
        b = _TAPS_DEV.get(key)
        if b is None:
            b = torch.from_numpy(np.ascontiguousarray(decimate_taps(q), dtype=np.float64)).to(x_dev.device)
            _TAPS_DEV[key] = b
    elif isinstance(taps, torch.Tensor):
        b = taps.to(x_dev.device, torch.float64).contiguous()
    else:
        b = torch.from_numpy(np.ascontiguousarray(taps, dtype=np.float64)).to(x_dev.device)
    n_out = -(-n // q)
    y = torch.empty((n_sig, n_out), dtype=torch.float64, device="cuda")
    if n_sig and n:
        _lib.check(lib.hs_fir_decimate_f64(x_dev.data_ptr(), n_sig, n, n, int(q), b.data_ptr(), b.numel(), y.data_ptr(), n_out,
                                           torch.cuda.current_stream().cuda_stream), "hs_fir_decimate_f64")
    return y


def decimate(x, q):
    """``scipy.signal.decimate(x, q, ftype='fir', zero_phase=True)`` along the last axis (data_structures.py:792)."""
    torch = _torch()
    x = np.asarray(x, dtype=np.float64)
    one_d = x.ndim == 1
    t = torch.from_numpy(np.ascontiguousarray(np.atleast_2d(x))).cuda()
    y = decimate_dev(t, q).cpu().numpy()
    return y[0] if one_d else y


# ------------------------------------------------------------------------------------------------------------------
# Pre-window stage of EEG_IBI_FFDTF_Pipeline (src/eeg_alpha_ibi_ffdtf.py:271-448): SOS zero-phase band-pass,
# anti-aliased integer down-sampling, Hilbert envelope.
# ------------------------------------------------------------------------------------------------------------------
def sosfiltfilt_dev(x_dev, sos):
    """``scipy.signal.sosfiltfilt(sos, x, axis=-1)`` (defaults: odd extension, padlen 3 * effective taps, ``sosfilt_zi``)
    on a (n_sig, n) CUDA float64 tensor; returns a new tensor.  Replaces eeg_alpha_ibi_ffdtf.py:307-309.

    Each sweep runs the sections in cascade with the scan kernels of K1 (``hs_iir_lfilter_f64``, zero state); the initial
    state ``zi * x0`` SciPy starts every section from enters by superposition: its zero-input response ``x0 * g_s[t]`` is
    signal independent up to the scalar ``x0``, so ``g_s`` is tabulated once per section (host, O(n), SciPy's own
    recursion) and added on the device."""
    from scipy import signal as _sig
    torch = _torch()
    lib = _lib.load()
    assert x_dev.is_cuda and x_dev.dtype == torch.float64 and x_dev.dim() == 2 and x_dev.is_contiguous()
    sos = np.atleast_2d(np.asarray(sos, dtype=np.float64))
    if sos.ndim != 2 or sos.shape[1] != 6:
        raise ValueError("sos array must be shape (n_sections, 6)")
    n_sections = sos.shape[0]
    ntaps = 2 * n_sections + 1
    ntaps -= min(int((sos[:, 2] == 0).sum()), int((sos[:, 5] == 0).sum()))
    edge = ntaps * 3
    n_sig, n = x_dev.shape
    if n <= edge:       # scipy.signal._validate_pad
        raise ValueError("The length of the input vector x must be greater than padlen, which is %d." % edge)
    left = 2.0 * x_dev[:, :1] - x_dev[:, 1:edge + 1].flip(1)
    right = 2.0 * x_dev[:, -1:] - x_dev[:, -edge - 1:-1].flip(1)
    ext = torch.cat([left, x_dev, right], dim=1).contiguous()
    n_ext = ext.shape[1]
    zi = _sig.sosfilt_zi(sos)
    g = np.stack([_sig.lfilter(sos[s, :3], sos[s, 3:], np.zeros(n_ext), zi=zi[s])[0] for s in range(n_sections)])
    g_dev = torch.from_numpy(g).cuda()
    ws = torch.empty(max(int(lib.hs_filtfilt_ws_bytes(n_sig, n_ext)), 16), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    coef = [(np.ascontiguousarray(sos[s, :3] / sos[s, 3]), np.ascontiguousarray(sos[s, 3:] / sos[s, 3])) for s in range(n_sections)]

    def sweep(u):
        x0 = u[:, :1].clone()
        for s in range(n_sections):
            y = torch.empty_like(u)
            _lib.check(lib.hs_iir_lfilter_f64(u.data_ptr(), n_sig, n_ext, n_ext, coef[s][0].ctypes.data, coef[s][1].ctypes.data, 3, 0,
                                              y.data_ptr(), n_ext, ws.data_ptr(), stream), "hs_iir_lfilter_f64")
            y.addcmul_(x0, g_dev[s][None, :])
            u = y
        return u

    fwd = sweep(ext)
    bwd = sweep(fwd.flip(1).contiguous())
    return bwd.flip(1)[:, edge:n_ext - edge].contiguous()


def sosfiltfilt(sos, x, axis=-1):
    """NumPy wrapper of :func:`sosfiltfilt_dev` (1-D or 2-D input, filtering along ``axis``)."""
    torch = _torch()
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        return sosfiltfilt_dev(torch.from_numpy(np.ascontiguousarray(x[None])).cuda(), sos)[0].cpu().numpy()
    if x.ndim != 2:
        raise ValueError("sosfiltfilt: 1-D or 2-D input")
    if axis in (-1, 1):
        return sosfiltfilt_dev(torch.from_numpy(np.ascontiguousarray(x)).cuda(), sos).cpu().numpy()
    return np.ascontiguousarray(sosfiltfilt_dev(torch.from_numpy(np.ascontiguousarray(x.T)).cuda(), sos).cpu().numpy().T)


def resample_poly_taps(down):
    """Anti-alias FIR of ``scipy.signal.resample_poly(x, 1, down)``: firwin(20 down + 1, 1/down, window=('kaiser', 5.0))."""
    from scipy.signal import firwin
    return firwin(2 * 10 * down + 1, 1.0 / down, window=("kaiser", 5.0))


def downsample_dev(x_dev, down):
    """``scipy.signal.resample_poly(x, up=1, down=down)`` (eeg_alpha_ibi_ffdtf.py:403): same polyphase kernel as
    ``decimate`` (K2), Kaiser-5 taps instead of Hamming; y[k] = sum_j h[j] x[down k + 10 down - j], zero outside."""
    return decimate_dev(x_dev, int(down), taps=resample_poly_taps(int(down)))


def next_fast_len(n):
    """``scipy.fft.next_fast_len(n)`` (complex transforms: 2-3-5-7-11 smooth), as eeg_alpha_ibi_ffdtf.py:350 calls it."""
    from scipy.fft import next_fast_len as _nfl
    return int(_nfl(int(n)))


def hilbert_envelope_dev(x_dev, N=None):
    """``np.abs(scipy.signal.hilbert(x, N=N)[..., :n])`` along the last axis of a (n_sig, n) CUDA float64 tensor
    (eeg_alpha_ibi_ffdtf.py:352-356); ``N`` defaults to ``n`` like SciPy.  Hand-written mixed-radix FFT (K7)."""
    torch = _torch()
    lib = _lib.load()
    assert x_dev.is_cuda and x_dev.dtype == torch.float64 and x_dev.dim() == 2 and x_dev.is_contiguous()
    n_sig, n = x_dev.shape
    N = n if N is None else int(N)
    if N <= 0:
        raise ValueError("N must be positive.")
    n_out = min(n, N)
    env = torch.empty((n_sig, n_out), dtype=torch.float64, device="cuda")
    if n_sig == 0:
        return env
    ws = torch.empty(int(lib.hs_hilbert_ws_bytes(n_sig, N)), dtype=torch.uint8, device="cuda")
    _lib.check(lib.hs_hilbert_f64(x_dev.data_ptr(), n_sig, n, n, N, env.data_ptr(), n_out, None, ws.data_ptr(),
                                  torch.cuda.current_stream().cuda_stream), "hs_hilbert_f64")
    return env
