"""Drop-in for the numeric functions of the reference's ``src/mtmvar.py``.

Same names, arguments, return shapes/dtypes, prints and exceptions as
``/root/reference/src/mtmvar.py`` (file:line cited per function); the arithmetic
runs in the sm_100a kernels behind the C ABI (``include/hs_b200.h``).  NumPy in,
NumPy out.  The ``batched_*`` / ``windowed_*`` functions at the bottom are the
same computations for many windows at once -- what a B200 is for -- and accept /
return CUDA ``torch.Tensor`` objects to keep data resident.

Quirks reproduced on purpose: ``dtf_multivariate``
returns un-normalised |H|^2; ``full_freq_dtf`` normalises rows over (j, f);
``multivariate_spectra`` uses H V H^T (plain transpose); lag covariances are
biased and not mean-removed.  The Yule-Walker system is solved by the LWR block
recursion instead of the dense LU the reference uses (mtmvar.py:116) -- same
solution to rounding (tests pin 1e-7 norm-wise).
"""
from __future__ import annotations

import numpy as np

from . import _lib

__all__ = [
    "count_corr", "ar_coeff", "mvar_transfer_function", "multivariate_spectra", "dtf_multivariate",
    "full_freq_dtf", "mvar_criterion", "gen_partial_directed_coherence", "partial_coherence", "direct_dtf",
    "batched_partial_coherence", "batched_gpdc", "batched_mvar_criterion", "batched_lagcov", "batched_ar_coeff", "batched_transfer", "windowed_ffdtf", "FfdtfPlan",
]


# ----------------------------------------------------------------------------- helpers
def _torch():
    return _lib.require_cuda()


def _dev(arr, dtype=None):
    """Host array -> contiguous CUDA float64 tensor (copy)."""
    torch = _torch()
    if isinstance(arr, torch.Tensor):
        t = arr
        if t.device.type != "cuda":
            t = t.cuda()
        return t.to(dtype or torch.float64).contiguous()
    a = np.ascontiguousarray(np.asarray(arr), dtype=np.float64 if dtype is None else None)
    t = torch.from_numpy(a).cuda()
    if dtype is not None:
        t = t.to(dtype)
    return t


def _stream():
    torch = _torch()
    return torch.cuda.current_stream().cuda_stream


def _raise_if_singular(status, what):
    st = status.cpu().numpy()
    if st.any():
        bad = np.nonzero(st)[0]
        raise np.linalg.LinAlgError(f"Singular matrix ({what}; window(s) {bad[:8].tolist()})")


def _ws(nbytes):
    torch = _torch()
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device="cuda")


# ----------------------------------------------------------------------------- batched (device) API
def batched_lagcov(x, offsets, ch_stride, n_win, trials, m, n, p):
    """R (n_win, p+1, m, m) on device.  ``x``: CUDA f64 tensor, ``offsets``: CUDA int64 (n_win*trials)."""
    torch = _torch()
    lib = _lib.load()
    if p >= n:
        raise ValueError(f"model order {p} must be smaller than the window length {n}")
    R = torch.empty((n_win, p + 1, m, m), dtype=torch.float64, device="cuda")
    _lib.check(lib.hs_lagcov_f64(x.data_ptr(), offsets.data_ptr(), int(ch_stride), n_win, trials, m, n, p,
                                 R.data_ptr(), _stream()), "hs_lagcov_f64")
    return R


def batched_yw_solve(R, want_all_orders=False):
    """LWR solve of every window's Yule-Walker system.  Returns (A (w,m,m,p), V (w,m,m), Vall|None, status)."""
    torch = _torch()
    lib = _lib.load()
    n_win, p1, m, _ = R.shape
    p = p1 - 1
    A = torch.empty((n_win, m, m, p), dtype=torch.float64, device="cuda")
    V = torch.empty((n_win, m, m), dtype=torch.float64, device="cuda")
    Vall = torch.empty((n_win, p, m, m), dtype=torch.float64, device="cuda") if want_all_orders else None
    status = torch.zeros((n_win,), dtype=torch.int32, device="cuda")
    ws = _ws(lib.hs_yw_ws_bytes(n_win, m, p))
    _lib.check(lib.hs_yw_solve_f64(R.data_ptr(), n_win, m, p, A.data_ptr(), V.data_ptr(),
                                   Vall.data_ptr() if Vall is not None else None, status.data_ptr(), ws.data_ptr(),
                                   _stream()), "hs_yw_solve_f64")
    return A, V, Vall, status


def _window_tensor(data):
    """(m, n) or (m, n, trials) host/device array -> device tensor (trials, m, n) + offsets."""
    torch = _torch()
    t = _dev(data)
    if t.ndim == 2:
        t = t.unsqueeze(0)
    elif t.ndim == 3:
        t = t.permute(2, 0, 1).contiguous()
    else:
        raise ValueError("data must be (channels, samples) or (channels, samples, trials)")
    trials, m, n = t.shape
    offsets = torch.arange(trials, dtype=torch.int64, device="cuda") * (m * n)
    return t, offsets, trials, m, n


def batched_ar_coeff(x, offsets, ch_stride, n_win, m, n, p, trials=1, want_all_orders=False):
    R = batched_lagcov(x, offsets, ch_stride, n_win, trials, m, n, p)
    A, V, Vall, status = batched_yw_solve(R, want_all_orders)
    return A, V, Vall, status, R


def batched_transfer(A, freqs, fs, want=("H",)):
    """A (n_win, m, m, p) device -> dict with any of 'H', 'Af' (complex128), 'dtf', 'ffdtf' (float64) + 'status'."""
    torch = _torch()
    lib = _lib.load()
    n_win, m, _, p = A.shape
    fr = _dev(np.asarray(freqs, dtype=np.float64).ravel())
    F = fr.numel()
    out = {}
    ptr = {"H": None, "Af": None, "dtf": None, "ffdtf": None}
    for k in want:
        if k in ("H", "Af"):
            out[k] = torch.empty((n_win, m, m, F), dtype=torch.complex128, device="cuda")
        else:
            out[k] = torch.empty((n_win, m, m, F), dtype=torch.float64, device="cuda")
        ptr[k] = out[k].data_ptr()
    status = torch.zeros((n_win,), dtype=torch.int32, device="cuda")
    ws = _ws(lib.hs_transfer_ws_bytes(n_win, m, p, F))
    _lib.check(lib.hs_transfer_dtf_f64(A.data_ptr(), fr.data_ptr(), F, float(fs), n_win, m, p, ptr["H"], ptr["Af"],
                                       ptr["dtf"], ptr["ffdtf"], status.data_ptr(), ws.data_ptr(), _stream()),
               "hs_transfer_dtf_f64")
    out["status"] = status
    return out


def windowed_ffdtf(signals, starts, window_size, freqs, fs, p, return_model=False):
    """ffDTF of every window ``signals[:, s:s+window_size]``: the loop of
    ``EEG_IBI_FFDTF_Pipeline.run_pipeline`` (eeg_alpha_ibi_ffdtf.py:741-755) as ONE batched call.

    ``signals``: (m, T) NumPy array or CUDA tensor; ``starts``: window start samples.
    Returns a CUDA tensor (n_win, m, m, F) (and A, V if ``return_model``); raises LinAlgError on singular windows.
    """
    torch = _torch()
    lib = _lib.load()
    x = _dev(signals)
    m, T = x.shape
    st = torch.as_tensor(np.asarray(starts, dtype=np.int64)).cuda() if not isinstance(starts, torch.Tensor) else starts.to("cuda", torch.int64)
    n_win = int(st.numel())
    if n_win and (int(st.min()) < 0 or int(st.max()) + window_size > T):
        raise ValueError("window outside the signal")
    if p >= window_size:
        raise ValueError(f"model order {p} must be smaller than the window length {window_size}")
    fr = _dev(np.asarray(freqs, dtype=np.float64).ravel())
    F = fr.numel()
    out = torch.empty((n_win, m, m, F), dtype=torch.float64, device="cuda")
    A = torch.empty((n_win, m, m, p), dtype=torch.float64, device="cuda") if return_model else None
    V = torch.empty((n_win, m, m), dtype=torch.float64, device="cuda") if return_model else None
    if n_win == 0:
        return (out, A, V) if return_model else out
    status = torch.zeros((n_win,), dtype=torch.int32, device="cuda")
    ws = _ws(lib.hs_mvar_ffdtf_ws_bytes(n_win, m, p, F))
    _lib.check(lib.hs_mvar_ffdtf_f64(x.data_ptr(), st.data_ptr(), T, n_win, m, window_size, p, fr.data_ptr(), F, float(fs),
                                     out.data_ptr(), A.data_ptr() if A is not None else None,
                                     V.data_ptr() if V is not None else None, status.data_ptr(), ws.data_ptr(), _stream()),
               "hs_mvar_ffdtf_f64")
    _raise_if_singular(status, "windowed_ffdtf")
    if return_model:
        return out, A, V
    return out


class FfdtfPlan:
    """Host-buffer pipeline (chunked copies overlapped with compute): NumPy (m, T) + window starts -> NumPy
    (n_win, m, m, F).  Wraps hs_plan_*.

    ``run(..., out=None)`` returns a VIEW of the plan's own page-locked result buffer (zero copy; valid until the next
    ``run`` / ``close`` -- ``.copy()`` it to keep it).  ``run(..., out=array)`` fills the caller's array: directly when it is
    page-locked (e.g. ``torch.empty(...).pin_memory().numpy()``), otherwise through the plan's pinned buffer with the host
    copy of finished chunks overlapped with the transfers still in flight.  One call at a time per plan (internal lock)."""

    def __init__(self, max_windows, m, window_size, p, n_freqs, max_samples):
        import ctypes as C
        _torch()
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.shape = (max_windows, m, window_size, p, n_freqs, max_samples)
        rc = self._lib.hs_plan_create(C.byref(self._h), max_windows, m, window_size, p, n_freqs, int(max_samples))
        if rc != 0:
            msg = _lib.last_error()
            self.close()
            raise _lib.HsError(f"hs_plan_create failed ({rc}): {msg}")

    def _own_result(self, n_win):
        import ctypes as C
        max_windows, m, _, _, F, _ = self.shape
        ptr, nbytes = C.c_void_p(), C.c_size_t()
        _lib.check(self._lib.hs_plan_host_result(self._h, C.byref(ptr), C.byref(nbytes)), "hs_plan_host_result")
        buf = (C.c_double * (max_windows * m * m * F)).from_address(ptr.value)
        buf._plan = self                       # the view keeps the plan (and with it the pinned buffer) alive
        return np.frombuffer(buf, dtype=np.float64).reshape(max_windows, m, m, F)[:n_win]

    def run(self, signals, starts, freqs, fs, out=None):
        max_windows, m, n, p, F, max_samples = self.shape
        if self._h is None:
            raise _lib.HsError("FfdtfPlan is closed")
        x = np.ascontiguousarray(signals, dtype=np.float64)
        st = np.ascontiguousarray(starts, dtype=np.int64).ravel()
        fr = np.ascontiguousarray(freqs, dtype=np.float64).ravel()
        if x.ndim != 2 or x.shape[0] != m or fr.size != F:
            raise ValueError("signals / freqs do not match the plan")
        n_win = st.size
        if n_win > max_windows or x.shape[1] > max_samples:
            raise ValueError("more windows / samples than the plan was created for")
        if out is None:
            res, out_ptr = self._own_result(n_win), None
        else:
            if not isinstance(out, np.ndarray) or out.dtype != np.float64 or not out.flags["C_CONTIGUOUS"] or not out.flags["WRITEABLE"] \
                    or out.shape != (n_win, m, m, F):
                raise ValueError(f"out must be a writable C-contiguous float64 ndarray of shape {(n_win, m, m, F)}")
            res, out_ptr = out, out.ctypes.data
        status = np.zeros(max(n_win, 1), dtype=np.int32)
        rc = self._lib.hs_plan_mvar_ffdtf_host(self._h, x.ctypes.data, x.shape[1], st.ctypes.data, n_win, fr.ctypes.data,
                                               float(fs), out_ptr, status.ctypes.data)
        _lib.check(rc, "hs_plan_mvar_ffdtf_host")
        if status[:n_win].any():
            raise np.linalg.LinAlgError(f"Singular matrix (window(s) {np.nonzero(status[:n_win])[0][:8].tolist()})")
        return res

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.hs_plan_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ----------------------------------------------------------------------------- reference-signature functions
def count_corr(x, ip, iwhat):
    """Reference ``count_corr`` (mtmvar.py:35-87): (r_left, r_right, r) for (m, n, trials) data."""
    if iwhat not in (1, 2):
        raise ValueError("iwhat must be 1 (biased, 1/n) or 2 (1/(n-k)); the reference leaves the lag blocks undefined otherwise")
    torch = _torch()
    lib = _lib.load()
    t, offsets, trials, m, n = _window_tensor(x)
    R = batched_lagcov(t, offsets, n, 1, trials, m, n, ip)
    if iwhat == 2:
        # mtmvar.py:60-63: lag L = k + 1 is scaled by 1 / (n - k) instead of 1 / n; R(0) keeps 1 / n (:72-73)
        scale = torch.tensor([1.0] + [n / (n - k) for k in range(ip)], dtype=torch.float64, device="cuda")
        R = R * scale[None, :, None, None]
    mp = m * ip
    G = torch.empty((1, mp, mp), dtype=torch.float64, device="cuda")
    rhs = torch.empty((1, mp, m), dtype=torch.float64, device="cuda")
    _lib.check(lib.hs_yw_assemble_f64(R.data_ptr(), 1, m, ip, G.data_ptr(), rhs.data_ptr(), _stream()), "hs_yw_assemble_f64")
    return G[0].cpu().numpy(), rhs[0].cpu().numpy(), R[0, 0].cpu().numpy()


def _fit(data, model_order, want_all_orders=False):
    t, offsets, trials, m, n = _window_tensor(data)
    A, V, Vall, status, _ = batched_ar_coeff(t, offsets, n, 1, m, n, int(model_order), trials, want_all_orders)
    _raise_if_singular(status, "ar_coeff")
    return A, V, Vall


def ar_coeff(data, model_order=5):
    """Reference ``ar_coeff`` (mtmvar.py:90-123): (ar_coeffs (m, m, p), variance (m, m))."""
    A, V, _ = _fit(data, model_order)
    return A[0].cpu().numpy(), V[0].cpu().numpy()


def mvar_transfer_function(ar_coeffs, freqs, fs):
    """Reference ``mvar_transfer_function`` (mtmvar.py:126-162): (H, A(f)), both (m, m, F) complex128."""
    A = _dev(ar_coeffs).unsqueeze(0)
    res = batched_transfer(A, freqs, fs, want=("H", "Af"))
    _raise_if_singular(res["status"], "mvar_transfer_function")
    return res["H"][0].cpu().numpy(), res["Af"][0].cpu().numpy()


_CRIT_CODES = {'AIC': 0, 'HQ': 1, 'SC': 2}


def batched_mvar_criterion(x, offsets, ch_stride, n_win, m, n, max_model_order, crit_type='AIC', trials=1):
    """Order criteria of ``n_win`` windows in one pass: lag covariances to ``max_model_order`` (K3), ONE LWR recursion per
    window that yields the residual covariance of every order (K4), ``ln det`` + penalty + argmin in ``criterion_kernel``.
    Returns CUDA tensors (crit (n_win, P) float64, popt (n_win) int32, status (n_win) int32)."""
    torch = _torch()
    lib = _lib.load()
    if crit_type not in _CRIT_CODES:
        raise ValueError("Invalid criterion type. Choose from 'AIC', 'HQ', 'SC'.")
    P = int(max_model_order)
    _, _, Vall, status, _ = batched_ar_coeff(x, offsets, ch_stride, n_win, m, n, P, trials, want_all_orders=True)
    crit = torch.empty((n_win, P), dtype=torch.float64, device="cuda")
    popt = torch.empty((n_win,), dtype=torch.int32, device="cuda")
    _lib.check(lib.hs_mvar_criterion_f64(Vall.data_ptr(), n_win, P, m, n, _CRIT_CODES[crit_type], crit.data_ptr(), None,
                                         popt.data_ptr(), _stream()), "hs_mvar_criterion_f64")
    return crit, popt, status


def mvar_criterion(data, max_model_order, crit_type='AIC', plot=False):
    """Reference ``mvar_criterion`` (mtmvar.py:551-601).  One LWR recursion to ``max_model_order`` yields the
    residual covariance of every lower order, so no refits.  ``plot`` is accepted and ignored (no matplotlib)."""
    if crit_type not in _CRIT_CODES:
        raise ValueError("Invalid criterion type. Choose from 'AIC', 'HQ', 'SC'.")
    data = np.asarray(data)
    n_channels, n_samples = data.shape
    model_order_range = np.arange(1, max_model_order + 1, dtype=int)
    t, offsets, trials, m, n = _window_tensor(data)
    crit, popt, status = batched_mvar_criterion(t, offsets, n, 1, m, n, max_model_order, crit_type, trials)
    _raise_if_singular(status, "mvar_criterion")
    crit = crit[0].cpu().numpy()
    optimal_model_range = model_order_range[int(popt[0].item()) - 1]
    return crit, model_order_range, optimal_model_range


def multivariate_spectra(signals, freqs, fs, max_model_order=20, optimal_model_order=None, crit_type='AIC'):
    """Reference ``multivariate_spectra`` (mtmvar.py:165-201): S(f) = H V H^T (plain transpose), (m, m, F) complex128."""
    torch = _torch()
    lib = _lib.load()
    if optimal_model_order is None:
        _, _, optimal_model_order = mvar_criterion(signals, max_model_order, crit_type, True)
        print('Optimal model order for all channels: p = ', str(optimal_model_order))
    else:
        print('Using provided model order: p = ', str(optimal_model_order))
    A, V, _ = _fit(signals, optimal_model_order)
    res = batched_transfer(A, freqs, fs, want=("H",))
    _raise_if_singular(res["status"], "multivariate_spectra")
    H = res["H"]
    S = torch.empty_like(H)
    _, m, _, F = H.shape
    _lib.check(lib.hs_spectra_f64(H.data_ptr(), V.data_ptr(), 1, m, F, S.data_ptr(), _stream()), "hs_spectra_f64")
    return S[0].cpu().numpy()


def dtf_multivariate(signals, freqs, fs, max_model_order=20, optimal_model_order=None, crit_type='AIC', comment=None):
    """Reference ``dtf_multivariate`` (mtmvar.py:204-234): |H|^2, NOT normalised (:232)."""
    if optimal_model_order is None:
        _, _, optimal_model_order = mvar_criterion(signals, max_model_order, crit_type, False)
        comment_str = '' if comment is None else comment + ' '
        print(f'Optimal model order for all {comment_str}channels: p = {optimal_model_order}')
    else:
        print(f'Using provided model order: p = {optimal_model_order}')
    A, _, _ = _fit(signals, optimal_model_order)
    res = batched_transfer(A, freqs, fs, want=("dtf",))
    _raise_if_singular(res["status"], "dtf_multivariate")
    return res["dtf"][0].cpu().numpy()


def full_freq_dtf(signals, freqs, fs, max_model_order=20, optimal_model_order=None, crit_type='AIC'):
    """Reference ``full_freq_dtf`` (mtmvar.py:237-284): rows normalised over (j, f)."""
    if optimal_model_order is None:
        _, _, optimal_model_order = mvar_criterion(signals, max_model_order, crit_type, False)
        print(f'Optimal model order for all channels: p = {optimal_model_order}')
    else:
        print(f'Using provided model order: p = {optimal_model_order}')
    A, _, _ = _fit(signals, optimal_model_order)
    res = batched_transfer(A, freqs, fs, want=("ffdtf",))
    _raise_if_singular(res["status"], "full_freq_dtf")
    return res["ffdtf"][0].cpu().numpy()


def gen_partial_directed_coherence(signals, freqs, fs, max_model_order=20, optimal_model_order=None, crit_type='AIC'):
    """Reference GPDC (mtmvar.py:388-468) from A(f) and diag(V); elementwise, done on the device tensors."""
    torch = _torch()
    if optimal_model_order is None:
        _, _, optimal_model_order = mvar_criterion(signals, max_model_order, crit_type, False)
        print('Optimal model order for all channels: p = ', str(optimal_model_order))
    else:
        print('Using provided model order: p = ', str(optimal_model_order))
    A, V, _ = _fit(signals, optimal_model_order)
    res = batched_transfer(A, freqs, fs, want=("Af",))
    return batched_gpdc(res["Af"], V)[0].cpu().numpy()


def batched_gpdc(Af, V):
    """GPDC of every window: Af (n_win, m, m, F) complex128 and V (n_win, m, m) on the device -> (n_win, m, m, F) float64."""
    torch = _torch()
    lib = _lib.load()
    Af = Af.contiguous()
    V = V.contiguous()
    n_win, m, _, F = Af.shape
    out = torch.empty((n_win, m, m, F), dtype=torch.float64, device="cuda")
    _lib.check(lib.hs_gpdc_f64(Af.data_ptr(), V.data_ptr(), n_win, m, F, out.data_ptr(), _stream()), "hs_gpdc_f64")
    return out


def batched_partial_coherence(S, ffdtf=None, want_kappa=True):
    """S (n_win, m, m, F) complex128 on the device -> (kappa or None, ddtf or None, status).

    One pivoted complex inverse per (window, bin) replaces the m^2 minor determinants of the reference
    (mtmvar.py:300-321): minor_ij = (-1)^(i+j) det(S) (S^-1)_ji.  With ``ffdtf`` (n_win, m, m, F) float64 the
    kernel also returns dDTF = ffDTF * |kappa| (mtmvar.py:383)."""
    torch = _torch()
    lib = _lib.load()
    S = S.contiguous()
    n_win, m, _, F = S.shape
    kappa = torch.empty_like(S) if want_kappa else None
    ddtf = None
    if ffdtf is not None:
        ffdtf = ffdtf.contiguous()
        ddtf = torch.empty_like(ffdtf)
    status = torch.zeros((n_win,), dtype=torch.int32, device="cuda")
    _lib.check(lib.hs_partial_coherence_f64(S.data_ptr(), n_win, m, F, kappa.data_ptr() if want_kappa else None,
                                            ffdtf.data_ptr() if ffdtf is not None else None,
                                            ddtf.data_ptr() if ddtf is not None else None, status.data_ptr(), _stream()),
               "hs_partial_coherence_f64")
    return kappa, ddtf, status


def partial_coherence(spectra):
    """Reference ``partial_coherence`` (mtmvar.py:287-338): (m, m, F) complex128 in and out, diagonal = 1."""
    torch = _torch()
    sp = np.ascontiguousarray(np.asarray(spectra, dtype=np.complex128))
    n_chan, _, n_f = sp.shape
    if n_chan == 1:          # mtmvar.py:320-321: the minor of a 1 x 1 matrix is 1; the diagonal is 1
        return np.ones((1, 1, n_f), dtype=np.complex128)
    S = torch.from_numpy(sp).cuda()[None]
    kappa, _, status = batched_partial_coherence(S)
    _raise_if_singular(status, "partial_coherence")
    return kappa[0].cpu().numpy()


def direct_dtf(signals, freqs, fs, max_model_order=20, optimal_model_order=None, crit_type='AIC'):
    """Reference ``direct_dtf`` (mtmvar.py:341-385): dDTF = ffDTF * |partial coherence|.

    The reference fits the model twice (``multivariate_spectra`` and ``full_freq_dtf`` each refit and each print);
    here one fit feeds both, the prints are kept."""
    torch = _torch()
    lib = _lib.load()
    # the two nested calls of the reference each resolve the order and print (mtmvar.py:186-191 and :262-267)
    if optimal_model_order is None:
        _, _, optimal_model_order_ = mvar_criterion(signals, max_model_order, crit_type, True)
        print('Optimal model order for all channels: p = ', str(optimal_model_order_))
        print(f'Optimal model order for all channels: p = {optimal_model_order_}')
    else:
        optimal_model_order_ = optimal_model_order
        print('Using provided model order: p = ', str(optimal_model_order))
        print(f'Using provided model order: p = {optimal_model_order}')
    A, V, _ = _fit(signals, optimal_model_order_)
    res = batched_transfer(A, freqs, fs, want=("H", "ffdtf"))
    _raise_if_singular(res["status"], "direct_dtf")
    H = res["H"]
    S = torch.empty_like(H)
    _, m, _, F = H.shape
    _lib.check(lib.hs_spectra_f64(H.data_ptr(), V.data_ptr(), 1, m, F, S.data_ptr(), _stream()), "hs_spectra_f64")
    _, ddtf, status = batched_partial_coherence(S, res["ffdtf"], want_kappa=False)
    _raise_if_singular(status, "direct_dtf")
    return ddtf[0].cpu().numpy()
