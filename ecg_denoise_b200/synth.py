"""Synthetic MIT-BIH-style ECG windows (host-side data preparation, numpy).

The reference trains on un-shipped `data/dict_data/*.npy` (local_utils/data_utils.py:88-117), so
benchmarks and tests use a seeded synthetic generator shaped like that data:

  * windows of `leads x length` samples cut from 360 Hz beat trains, the R peak of the middle
    beat at `length/2 +- 8` (the R-wave bias of the model is centred, model/transformer.py:541);
  * per-lead z-normalisation like `np_norm` (local_utils/local_utils.py:261-266);
  * baseline-wander / muscle-artifact / electrode-motion style noise mixed to a target SNR with
    the reference's formula (local_utils/local_utils.py:176-192):
        noise *= sqrt(P_signal / (10**(snr/10) * P_noise)).

Uses numpy's frozen legacy `RandomState` stream so fixtures are reproducible everywhere.
"""
from __future__ import annotations

import numpy as np

FS = 360.0


def _beat(t: np.ndarray, r_time: float, rs: np.random.RandomState, lead_gain: float) -> np.ndarray:
    """P-QRS-T complex as a sum of Gaussians centred relative to the R time (seconds)."""
    amp = lead_gain * (1.0 + 0.1 * rs.standard_normal())
    waves = (  # (offset s, width s, amplitude)
        (-0.20, 0.025, 0.12), (-0.035, 0.010, -0.14), (0.0, 0.011, 1.0),
        (0.035, 0.012, -0.22), (0.25, 0.055, 0.30),
    )
    y = np.zeros_like(t)
    for off, wid, a in waves:
        y += amp * a * np.exp(-0.5 * ((t - r_time - off) / wid) ** 2)
    return y


def clean_windows(n: int, leads: int = 2, length: int = 256, seed: int = 2023) -> np.ndarray:
    """(n, leads, length) float32 clean windows, z-normalised per lead."""
    rs = np.random.RandomState(seed)
    t = np.arange(length) / FS
    out = np.empty((n, leads, length), np.float32)
    for i in range(n):
        r_mid = (length / 2 + rs.randint(-8, 9)) / FS
        rr = rs.uniform(0.6, 1.2)
        r_times = [r_mid + k * rr * (1 + 0.03 * rs.standard_normal()) for k in range(-3, 4)]
        for l in range(leads):
            gain = 1.0 if l == 0 else rs.uniform(0.4, 0.9) * rs.choice([-1.0, 1.0])
            y = np.zeros(length)
            for rt in r_times:
                y += _beat(t, rt, rs, gain)
            y = (y - y.mean()) / (y.std() + 1e-8)
            out[i, l] = y
    return out


def _noise(kind: str, shape, rs: np.random.RandomState) -> np.ndarray:
    n, leads, length = shape
    t = np.arange(length) / FS
    if kind == "bw":       # baseline wander: sum of slow sinusoids
        z = np.zeros(shape)
        for _ in range(3):
            f = rs.uniform(0.05, 0.5, size=(n, leads, 1))
            ph = rs.uniform(0, 2 * np.pi, size=(n, leads, 1))
            z += rs.uniform(0.3, 1.0, size=(n, leads, 1)) * np.sin(2 * np.pi * f * t + ph)
        return z
    if kind == "ma":       # muscle artifact: band-limited Gaussian
        z = rs.standard_normal(shape)
        ker = np.hanning(9)
        ker /= ker.sum()
        return z - np.apply_along_axis(lambda v: np.convolve(v, ker, mode="same"), -1, z) * 0.7
    if kind == "em":       # electrode motion: random-walk steps
        steps = rs.standard_normal(shape) * (rs.uniform(size=shape) < 0.05)
        return np.cumsum(steps, -1) + 0.05 * rs.standard_normal(shape)
    raise ValueError(kind)


def add_noise(clean: np.ndarray, kind: str = "bw", snr_db: float = -4.0, seed: int = 500) -> np.ndarray:
    """noisy = clean + noise scaled to `snr_db` per window-lead (local_utils/local_utils.py:176-192)."""
    rs = np.random.RandomState(seed)
    kinds = ["bw", "ma", "em"] if kind == "emb" else [kind]
    noise = sum(_noise(k, clean.shape, rs) for k in kinds)
    noise = noise - noise.mean(-1, keepdims=True)
    ps = (clean.astype(np.float64) ** 2).mean(-1, keepdims=True)
    pn = (noise ** 2).mean(-1, keepdims=True) + 1e-12
    noise *= np.sqrt(ps / (10 ** (snr_db / 10.0) * pn))
    return (clean + noise).astype(np.float32)


def make_batch(n: int, leads: int = 2, length: int = 256, seed: int = 2023,
               kind: str = "emb", snr_db: float = -4.0):
    """(noisy, clean) float32 arrays of shape (n, leads, length)."""
    clean = clean_windows(n, leads, length, seed)
    return add_noise(clean, kind, snr_db, seed + 1), clean
