"""Seeded synthetic workloads of the shapes BASELINE.json names.

Used by the parity tests, the golden-vector generator and ``bench.py``.  numpy's
``default_rng`` (PCG64) is used everywhere so that the same seed gives the same
signal on every machine.
"""

from __future__ import annotations

import numpy as np


def utterance_lengths(
    n_utts: int, sample_rate: int, hop: int, seed: int, min_s: float = 1.0, max_s: float = 10.0
) -> np.ndarray:
    """Lengths in samples, ``(floor(U(min_s, max_s) * sr) // hop) * hop`` -- i.e. already
    truncated to a multiple of hop like ``Preprocessor.process_audio`` does
    (preprocessor.py:216-218)."""
    rng = np.random.default_rng(seed)
    secs = rng.uniform(min_s, max_s, size=n_utts)
    L = (np.floor(secs * sample_rate).astype(np.int64) // hop) * hop
    return L


def white_noise(n: int, seed: int) -> np.ndarray:
    """U(-0.95, 0.95) float32."""
    rng = np.random.default_rng(seed)
    return rng.uniform(-0.95, 0.95, size=n).astype(np.float32)


def speech_like(n: int, sample_rate: int, seed: int) -> np.ndarray:
    """A harmonic stack (5-20 harmonics of f0 in [80, 300] Hz, 1/k amplitude decay) under a
    slow AM envelope plus -60 dB noise, peak-normalised to 0.95: exercises the dynamic
    range that white noise does not."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / sample_rate
    f0 = rng.uniform(80.0, 300.0)
    n_h = int(rng.integers(5, 21))
    vib = 1.0 + 0.02 * np.sin(2 * np.pi * rng.uniform(3.0, 7.0) * t)
    phase = 2 * np.pi * f0 * np.cumsum(vib) / sample_rate
    x = np.zeros(n, dtype=np.float64)
    for k in range(1, n_h + 1):
        if k * f0 * 1.02 >= sample_rate / 2:
            break
        x += np.sin(k * phase + rng.uniform(0, 2 * np.pi)) / k
    env = 0.55 + 0.45 * np.sin(2 * np.pi * rng.uniform(0.5, 3.0) * t + rng.uniform(0, 2 * np.pi))
    x *= env
    x /= max(np.abs(x).max(), 1e-12)
    x = 0.95 * x + 1e-3 * rng.standard_normal(n)
    return np.clip(x, -1.0, 1.0).astype(np.float32)


def synthetic_durations(n_frames: int, seed: int, frames_per_phone: int = 7) -> np.ndarray:
    """A random composition of ``n_frames`` into ``P = max(1, n_frames // 7)`` int64 parts
    with ~5 % zeros; 10 % of utterances get their sum perturbed by +-{1, 2} to exercise
    the overrun / underrun behaviour of ``average_data_by_durations``
    (preprocessor.py:287-300)."""
    rng = np.random.default_rng(seed)
    P = max(1, n_frames // frames_per_phone)
    cuts = np.sort(rng.integers(0, n_frames + 1, size=P - 1))
    d = np.diff(np.concatenate([[0], cuts, [n_frames]])).astype(np.int64)
    # ~5 % zeros: merge a phone's frames into its right neighbour
    for i in np.nonzero(rng.uniform(size=P) < 0.05)[0]:
        if i + 1 < P:
            d[i + 1] += d[i]
            d[i] = 0
    if rng.uniform() < 0.10:
        delta = int(rng.choice([-2, -1, 1, 2]))
        j = int(np.argmax(d))
        if d[j] + delta > 0:
            d[j] += delta
    return d


def pack_ragged(arrays) -> tuple[np.ndarray, np.ndarray]:
    """Concatenate 1-D arrays; returns ``(packed, offsets[B+1] int64)``."""
    lens = np.array([len(a) for a in arrays], dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    packed = np.concatenate(arrays) if len(arrays) else np.zeros(0, dtype=np.float32)
    return packed, offsets
