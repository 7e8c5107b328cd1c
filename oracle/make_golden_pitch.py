"""Generate ``tests/golden/pitch.npz``: the LIVE reference's ``Preprocessor.extract_pitch``
(everyvoice/preprocessor/preprocessor.py:244-285) run unmodified on seeded pitch tracks.

    python -m oracle.make_golden_pitch          (build container only)

pyworld is not installed here, and DIO / StoneMask are outside this row anyway: a stub ``pyworld`` module hands the
reference the seeded track (``pitch_tracks``) from ``dio`` and passes it through ``stonemask``; everything the
reference does AFTER them -- zeros -> NaN, ``_interpolate`` (np.interp), the pitch-less ValueError branch, the cast
to float32 -- is the reference's own code.  Only its outputs are stored."""

from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"


def pitch_tracks() -> dict:
    """Seeded float64 tracks as pyworld would return them: 0 where unvoiced."""
    rng = np.random.default_rng(4242)
    out = {}

    def track(n, p_unvoiced, run):
        f0 = 80 + 220 * rng.random(n)
        f0 = np.convolve(f0, np.ones(5) / 5, mode="same")
        t = 0
        while t < n:
            L = int(rng.integers(1, run + 1))
            if rng.random() < p_unvoiced:
                f0[t:t + L] = 0.0
            t += L
        return f0.astype(np.float64)

    out["mixed"] = track(431, 0.35, 20)
    out["leading_trailing"] = track(200, 0.3, 12)
    out["leading_trailing"][:17] = 0.0
    out["leading_trailing"][-9:] = 0.0
    out["all_voiced"] = track(64, 0.0, 5)
    out["all_unvoiced"] = np.zeros(75)
    out["single_voiced"] = np.zeros(40)
    out["single_voiced"][23] = 151.25
    out["one_frame_voiced"] = np.array([212.5])
    out["one_frame_unvoiced"] = np.array([0.0])
    out["long_gap"] = track(900, 0.1, 8)
    out["long_gap"][100:700] = 0.0
    return out


def main():
    from oracle import ev_oracle as O
    from oracle.make_golden import _import_reference

    _, Preprocessor, _ = _import_reference()
    store = {}
    for name, f0 in pitch_tracks().items():
        stub = types.ModuleType("pyworld")
        stub.dio = lambda x, fs, frame_period, speed, _f0=f0: (_f0.copy(), np.arange(len(_f0)) * frame_period / 1000)
        stub.stonemask = lambda x, f0_, t, fs: f0_
        sys.modules["pyworld"] = stub
        self = types.SimpleNamespace(
            input_sampling_rate=22050, audio_config=types.SimpleNamespace(fft_hop_size=256),
            _interpolate=Preprocessor._interpolate)
        ref = Preprocessor.extract_pitch(self, torch.zeros(1, 256 * len(f0)))
        ora = O.postprocess_pitch(f0)
        assert ref.dtype == torch.float32 and np.array_equal(ref.numpy(), ora), name
        store[name] = ref.numpy()
        print(f"{name:20s} T={len(f0):4d} voiced={int((f0 != 0).sum()):4d} oracle == reference")
    np.savez_compressed(GOLD / "pitch.npz", **store)


if __name__ == "__main__":
    main()
