"""CPU oracle for the MVAR / DTF hot path  --  TEST INFRASTRUCTURE ONLY.

A NumPy restatement of the arithmetic of the reference's ``src/mtmvar.py``
(SURVEY.md Appendix A.3-A.5).  It exists to *check* the CUDA path.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it; the product package never does.

Pinned: every function here is compared against the *imported reference*
(``/root/reference/src/mtmvar.py``) by ``oracle/make_golden.py`` and the
resulting vectors are committed under ``tests/golden/`` (the reference's own
test-suite holds no numbers for this path).

The per-frequency Python loops are kept on purpose: the reference runs exactly
these loops (one ``np.linalg.inv`` per frequency bin), and this module is also
the "port" timed as the CPU baseline.
"""
from __future__ import annotations

import numpy as np


def lag_covariances(x, p):
    """Biased lag covariances R(0..p), no mean removal, averaged over trials.

    R(L)[i, j] = 1/n * sum_{t=0}^{n-1-L} x_i(t) x_j(t+L)

    Follows reference ``count_corr`` (mtmvar.py:54-59 lags 1..p, :72-73 lag 0,
    :78-85 trial mean) for ``iwhat == 1``.  ``x`` is (m, n) or (m, n, trials).
    Returns (p+1, m, m).
    """
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 2:
        x = x[:, :, None]
    m, n, trials = x.shape
    acc = np.zeros((p + 1, m, m))
    for tr in range(trials):
        xt = x[:, :, tr]
        for lag in range(p + 1):
            acc[lag] += (xt[:, : n - lag] @ xt[:, lag:].T) * (1.0 / n)
    if trials > 1:
        acc = acc / trials
    return acc


def yule_walker_system(R):
    """Block-Toeplitz system of reference ``count_corr`` (mtmvar.py:65-76).

    G[a, b] = R(a-b) for a > b, R(b-a)^T for a < b, R(0) on the diagonal;
    rhs = [R(1); ...; R(p)].  Returns (G (mp, mp), rhs (mp, m), R0 (m, m)).
    """
    p = R.shape[0] - 1
    m = R.shape[1]
    G = np.zeros((m * p, m * p))
    rhs = np.zeros((m * p, m))
    for a in range(p):
        rhs[a * m:(a + 1) * m] = R[a + 1]
        for b in range(p):
            if a > b:
                blk = R[a - b]
            elif a < b:
                blk = R[b - a].T
            else:
                blk = R[0]
            G[a * m:(a + 1) * m, b * m:(b + 1) * m] = blk
    return G, rhs, R[0].copy()


def count_corr(x, ip, iwhat=1):
    """Drop-in signature of reference ``count_corr`` (mtmvar.py:35).  ``iwhat == 2`` (mtmvar.py:60-63) scales lag
    L = k + 1 by 1 / (n - k) instead of 1 / n; lag 0 keeps 1 / n (:72-73)."""
    if iwhat not in (1, 2):
        raise ValueError("iwhat must be 1 or 2")
    R = lag_covariances(x, ip)
    if iwhat == 2:
        n = np.shape(x)[1]
        x3 = np.asarray(x, dtype=np.float64)
        if x3.ndim == 2:
            x3 = x3[:, :, None]
        R = R.copy()
        for k in range(ip):
            acc = np.zeros_like(R[0])
            for tr in range(x3.shape[2]):
                xt = x3[:, :, tr]
                acc += (xt[:, : n - k - 1] @ xt[:, k + 1:].T) * (1.0 / (n - k))
            R[k + 1] = acc / x3.shape[2] if x3.shape[2] > 1 else acc
    return yule_walker_system(R)


def ar_coeff(data, model_order=5):
    """Yule-Walker MVAR fit, reference ``ar_coeff`` (mtmvar.py:90-123).

    LU solve of the full system (:116), V = R0 - X rhs (:119),
    A[i, j, k] = X[i, k*m + j] (:122).
    """
    data = np.asarray(data, dtype=np.float64)
    m = data.shape[0]
    G, rhs, r0 = count_corr(data, model_order, 1)
    X = np.linalg.solve(G, rhs).T
    V = r0 - X @ rhs
    A = X.reshape(m, model_order, m).transpose(0, 2, 1)
    return A, V


def mvar_transfer_function(ar_coeffs, freqs, fs):
    """A(f) = I - sum_k A_k exp(-2 pi i k f / fs); H(f) = inv(A(f)).

    Reference ``mvar_transfer_function`` (mtmvar.py:126-162): same phase
    expression (:153) and one LAPACK inverse per bin (:159).
    """
    m, _, p = ar_coeffs.shape
    freqs = np.asarray(freqs, dtype=np.float64)
    F = len(freqs)
    z = np.empty((p, F), dtype=complex)
    for k in range(1, p + 1):
        z[k - 1] = np.exp(-k * 2 * np.pi * 1j * freqs / fs)
    H = np.empty((m, m, F), dtype=complex)
    Af = np.empty((m, m, F), dtype=complex)
    eye = np.eye(m, dtype=complex)
    for fi in range(F):
        a = eye.copy()
        for k in range(p):
            a -= ar_coeffs[:, :, k] * z[k, fi]
        Af[:, :, fi] = a
        H[:, :, fi] = np.linalg.inv(a)
    return H, Af


def mvar_criterion(data, max_model_order, crit_type="AIC"):
    """ln det V_p + penalty, reference ``mvar_criterion`` (mtmvar.py:551-601)."""
    m, n = data.shape[:2]
    orders = np.arange(1, max_model_order + 1, dtype=int)
    crit = np.zeros(max_model_order)
    for p in orders:
        _, V = ar_coeff(data, int(p))
        ld = np.log(np.linalg.det(V))
        if crit_type == "AIC":
            crit[p - 1] = ld + 2 * p * m ** 2 / n
        elif crit_type == "HQ":
            crit[p - 1] = ld + 2 * np.log(np.log(n)) * p * m ** 2 / n
        elif crit_type == "SC":
            crit[p - 1] = ld + np.log(n) * p * m ** 2 / n
        else:
            raise ValueError("Invalid criterion type. Choose from 'AIC', 'HQ', 'SC'.")
    return crit, orders, orders[np.argmin(crit)]


def _resolve_order(signals, max_model_order, optimal_model_order, crit_type):
    if optimal_model_order is None:
        return int(mvar_criterion(signals, max_model_order, crit_type)[2])
    return int(optimal_model_order)


def dtf_multivariate(signals, freqs, fs, max_model_order=20, optimal_model_order=None, crit_type="AIC"):
    """|H(f)|^2, **un-normalised** as the reference returns it (mtmvar.py:232)."""
    p = _resolve_order(signals, max_model_order, optimal_model_order, crit_type)
    A, _ = ar_coeff(signals, p)
    H, _ = mvar_transfer_function(A, freqs, fs)
    return np.abs(H) ** 2


def full_freq_dtf(signals, freqs, fs, max_model_order=20, optimal_model_order=None, crit_type="AIC"):
    """ffDTF[i,j,f] = |H_ij(f)|^2 / sum_{j',f'} |H_ij'(f')|^2 (mtmvar.py:281-283)."""
    dtf = dtf_multivariate(signals, freqs, fs, max_model_order, optimal_model_order, crit_type)
    m = dtf.shape[0]
    out = np.empty_like(dtf)
    for i in range(m):
        denom = np.sum(dtf[i, :, :])
        for j in range(m):
            out[i, j, :] = dtf[i, j, :] / denom
    return out


def multivariate_spectra(signals, freqs, fs, max_model_order=20, optimal_model_order=None, crit_type="AIC"):
    """S(f) = H V H^T with a PLAIN transpose (quirk, mtmvar.py:199)."""
    p = _resolve_order(signals, max_model_order, optimal_model_order, crit_type)
    A, V = ar_coeff(signals, p)
    H, _ = mvar_transfer_function(A, freqs, fs)
    m, _, F = H.shape
    S = np.empty((m, m, F), dtype=complex)
    for fi in range(F):
        h = H[:, :, fi]
        S[:, :, fi] = h @ (V @ h.T)
    return S


def gen_partial_directed_coherence(signals, freqs, fs, max_model_order=20, optimal_model_order=None, crit_type="AIC"):
    """GPDC from A(f) and diag(V), reference mtmvar.py:444-468."""
    p = _resolve_order(signals, max_model_order, optimal_model_order, crit_type)
    A, V = ar_coeff(signals, p)
    _, Af = mvar_transfer_function(A, freqs, fs)
    s2 = np.diag(V)
    absA = np.abs(Af)
    denom = np.sqrt(np.sum(absA ** 2 / s2[:, None, None], axis=0))     # (j, f)
    num = absA / np.sqrt(s2)[:, None, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        g = np.where(denom[None] != 0, num / denom[None], 0.0)
    return g


def partial_coherence(spectra):
    """Partial coherence from the minors of S(f), reference mtmvar.py:287-338 (det of every minor, as the reference)."""
    n_chan, _, n_f = spectra.shape
    minors = np.zeros((n_chan, n_chan, n_f), dtype=np.complex128)
    idx = np.arange(n_chan)
    for i in range(n_chan):
        for j in range(n_chan):
            if n_chan > 1:
                sub = spectra[np.ix_(idx != i, idx != j)]                   # (m-1, m-1, F)
                minors[i, j, :] = np.linalg.det(np.moveaxis(sub, 2, 0))     # mtmvar.py:318
            else:
                minors[i, j, :] = 1.0
    kappa = np.zeros((n_chan, n_chan, n_f), dtype=np.complex128)
    for i in range(n_chan):
        for j in range(n_chan):
            if i != j:
                den = np.sqrt(minors[i, i, :] * minors[j, j, :])
                with np.errstate(divide="ignore", invalid="ignore"):
                    kappa[i, j, :] = np.where(den != 0, minors[i, j, :] / den, 0)
            else:
                kappa[i, j, :] = 1.0
    return kappa


def direct_dtf(signals, freqs, fs, max_model_order=20, optimal_model_order=None, crit_type="AIC"):
    """dDTF = ffDTF * |partial coherence|, reference mtmvar.py:341-385."""
    S = multivariate_spectra(signals, freqs, fs, max_model_order, optimal_model_order, crit_type)
    kappa = partial_coherence(S)
    ff = full_freq_dtf(signals, freqs, fs, max_model_order, optimal_model_order, crit_type)
    return ff * np.abs(kappa)


# ---------------------------------------------------------------- windows
def window_starts(T, n_windows=3, window_size=None):
    """Start samples of ``EEG_IBI_FFDTF_Pipeline._create_windows``
    (eeg_alpha_ibi_ffdtf.py:451-518), including its ValueErrors."""
    if window_size is None:
        if T % n_windows != 0:
            raise ValueError(
                f"Cannot evenly divide signal of length {T} into {n_windows} "
                f"non-overlapping windows. Provide a specific window_size.")
        window_size = T // n_windows
    else:
        need = (T + n_windows - 1) // n_windows
        if window_size < need:
            raise ValueError(
                f"window_size={window_size} is too short. To cover {T} samples "
                f"with {n_windows} windows without leaving gaps, the minimum "
                f"window_size is {need}.")
        if window_size > T:
            raise ValueError(f"window_size ({window_size}) cannot exceed signal length ({T}).")
    last = T - window_size
    if last < n_windows - 1 and n_windows > 1:
        raise ValueError(
            f"window_size={window_size} is too large to generate {n_windows} "
            f"distinct windows. Decrease window_size or n_windows.")
    if n_windows == 1:
        starts = np.zeros(1, dtype=np.int64)
    else:
        starts = np.linspace(0, last, n_windows, dtype=int).astype(np.int64)
    return starts, int(window_size)


def windowed_ffdtf(signals, starts, window_size, freqs, fs, p):
    """Serial window loop of ``run_pipeline`` (eeg_alpha_ibi_ffdtf.py:741-755),
    ffDTF only; stacked (n_win, m, m, F) as at :651."""
    out = [full_freq_dtf(signals[:, s:s + window_size], freqs, fs, optimal_model_order=p) for s in starts]
    return np.array(out)
