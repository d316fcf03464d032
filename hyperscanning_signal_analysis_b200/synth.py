"""Seeded synthetic dyad-EEG workloads.

The reference ships no data and pins no numbers, so every parity test and
bench run uses these generators.  They only *define the workload*; nothing
here is on the measured path.

cfg1  38 ch (0-18 child, 19-37 caregiver) x 60 s @ 256 Hz
cfg2  same generator, one 600 s task -> 599 windows of 512 @ hop 256
cfg3  64 dyads x 3 tasks, seed = 20260101 + 3*dyad + task
cfg4  38 ch x 1 h @ 1024 Hz + slow drift (front end)
cfg5  128 ch epochs (128, 512, 100) per window
"""
from __future__ import annotations

import numpy as np

BASE_SEED = 20260101


def _stable_var(rng, m, order=3, coupling=0.05, rho_max=0.95, fs=256.0):
    """Random stable VAR(order): per-channel resonances + weak cross terms.

    Returns coefficient stack ``(order, m, m)`` with spectral radius of the
    companion matrix <= ``rho_max``.
    """
    half = m // 2
    a = np.zeros((order, m, m))
    for i in range(m):
        f0 = rng.uniform(4.0, 45.0)          # theta .. low gamma
        r = rng.uniform(0.55, 0.80)          # broad resonances keep the spectrum shallow
        a[0, i, i] = 2.0 * r * np.cos(2.0 * np.pi * f0 / fs)
        a[1, i, i] = -r * r
    for k in range(order):
        c = rng.standard_normal((m, m)) * (0.04 / (k + 1))
        np.fill_diagonal(c, 0.0)
        # 5 % cross-brain coupling relative to within-brain coupling
        c[:half, half:] *= coupling
        c[half:, :half] *= coupling
        a[k] += c
    comp = np.zeros((m * order, m * order))
    comp[:m, :] = np.concatenate(list(a), axis=1)
    comp[m:, :-m] = np.eye(m * (order - 1))
    rho = np.max(np.abs(np.linalg.eigvals(comp)))
    if rho > rho_max:
        s = rho_max / rho
        for k in range(order):
            a[k] *= s ** (k + 1)
    return a


def dyad_eeg(seed=BASE_SEED, m=38, fs=256.0, n_samples=15360, line_amp=5.0,
             floor_sigma=0.1, drift=False, sigma_uv=20.0):
    """Raw (unfiltered) synthetic dyad EEG, float64, shape ``(m, n_samples)``."""
    rng = np.random.default_rng(seed)
    order = 3
    a = _stable_var(rng, m, order=order, fs=min(fs, 256.0))
    burn = 512
    n_tot = n_samples + burn
    e = rng.standard_normal((n_tot, m))
    x = np.zeros((n_tot, m))
    acat = np.concatenate([a[k].T for k in range(order)], axis=0)  # (order*m, m)
    hist = np.zeros(order * m)
    for t in range(n_tot):
        xt = e[t] + hist @ acat
        x[t] = xt
        hist[m:] = hist[:-m]
        hist[:m] = xt
    x = x[burn:].T.copy()
    x *= sigma_uv / x.std()
    t = np.arange(n_samples) / fs
    x += line_amp * np.sin(2.0 * np.pi * 50.0 * t)[None, :]
    x += floor_sigma * sigma_uv * rng.standard_normal((m, n_samples))
    if drift:
        ph = rng.uniform(0, 2 * np.pi, size=(m, 1))
        x += 3.0 * sigma_uv * np.sin(2.0 * np.pi * 0.05 * t[None, :] + ph)
        x += rng.uniform(-50, 50, size=(m, 1))
    return np.ascontiguousarray(x)


def cfg1_raw(seed=BASE_SEED):
    return dyad_eeg(seed, m=38, fs=256.0, n_samples=15360)


def cfg2_raw(seed=BASE_SEED, seconds=600):
    return dyad_eeg(seed, m=38, fs=256.0, n_samples=int(seconds * 256))


def cfg3_seed(dyad, task):
    return BASE_SEED + 3 * dyad + task


def cfg4_raw(seed=BASE_SEED, seconds=3600, m=38):
    return dyad_eeg(seed, m=m, fs=1024.0, n_samples=int(seconds * 1024), drift=True)


def cfg5_epochs(seed=BASE_SEED, m=128, n=512, trials=100, n_windows=8):
    """Multi-trial epochs ``(n_windows, m, n, trials)``; white + weak mixing."""
    rng = np.random.default_rng(seed)
    mix = np.eye(m) + 0.05 * rng.standard_normal((m, m))
    out = np.empty((n_windows, m, n, trials))
    b = np.array([1.0, 0.6, 0.2])
    for w in range(n_windows):
        z = rng.standard_normal((m, n + 2, trials))
        s = b[0] * z[:, 2:] + b[1] * z[:, 1:-1] + b[2] * z[:, :-2]
        out[w] = np.einsum("ij,jnt->int", mix, s)
    return out


def default_freqs(n_freqs=256, fs=256.0):
    """``np.linspace(0, fs/2, F, endpoint=False)``."""
    return np.linspace(0.0, fs / 2.0, n_freqs, endpoint=False)
