"""Generate ``tests/golden/gates.npz``: keep / skip decisions of the LIVE reference's ``Preprocessor.process_audio``
(everyvoice/preprocessor/preprocessor.py:148-186) for utterances placed AT its gates.

    python -m oracle.make_golden_gate          (build container only)

* loudness gate (``< -36`` LKFS or NaN => "audio_empty", :177-186): seeded signals at several sampling rates are scaled
  so that torchaudio's own ``Loudness`` lands at -36 + d for d in +-{1e-4, 1e-3, 1e-2} (and as close to 0 as float32
  allows); stored: the float32 scale factor, torchaudio's loudness of the scaled signal and the reference's decision;
* length gates (``> max_audio_length`` / ``< min_audio_length`` seconds, :163-176): lengths one sample either side of
  0.4 s and 11.0 s.

``process_audio`` runs unmodified with ``load_audio`` handing it the tensor (see make_golden_frontend.py).  The tests
regenerate the inputs from their seeds (``gate_inputs``)."""

from __future__ import annotations

import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"

# name: (sampling rate, seconds, kind, seed)
SIGNALS = {
    "speech_22k": (22050, 1.7, "speech", 4101),
    "white_22k": (22050, 0.9, "white", 4102),
    "speech_44k": (44100, 1.2, "speech", 4103),
    "speech_48k": (48000, 1.1, "speech", 4104),
    "speech_16k": (16000, 2.1, "speech", 4105),
    "speech_11k": (11025, 1.6, "speech", 4106),   # round(0.4 sr) is not four 100 ms steps here
}
DELTAS = [-1e-2, -1e-3, -1e-4, 0.0, 1e-4, 1e-3, 1e-2]
LENGTHS = {22050: [8819, 8820, 242550, 242551], 16000: [6399, 6400, 176000, 176001]}


def gate_inputs(name: str) -> tuple[np.ndarray, int]:
    from everyvoice_b200 import synth

    sr, seconds, kind, seed = SIGNALS[name]
    n = int(round(seconds * sr)) + 11
    x = synth.speech_like(n, sr, seed=seed) if kind == "speech" else synth.white_noise(n, seed=seed)
    return x.astype(np.float32), sr


def length_inputs(sr: int, n: int) -> np.ndarray:
    from everyvoice_b200 import synth

    return (synth.white_noise(n, seed=5000 + n % 97) * np.float32(0.3)).astype(np.float32)


def _reference_decision(Preprocessor, ref_mod, x, sr):
    ref_mod.load_audio = lambda path, _x=x, _sr=sr: (torch.from_numpy(_x.copy())[None], _sr, len(_x) / _sr)
    self = types.SimpleNamespace(
        audio_config=types.SimpleNamespace(max_audio_length=11.0, min_audio_length=0.4),
        counters=types.SimpleNamespace(increment=lambda k, v=1: None), multichannel_files_list=[])
    audio, _ = Preprocessor.process_audio(self, "seeded.wav", resample_rate=sr, hop_size=256)
    return audio is not None


def main():
    import torchaudio

    from oracle import ev_oracle as O
    from oracle.make_golden import _import_reference

    _, Preprocessor, _ = _import_reference()
    import everyvoice.preprocessor.preprocessor as ref_mod

    loud = lambda x, sr: float(torchaudio.functional.loudness(torch.from_numpy(x)[None], sr))  # noqa: E731
    store = {}
    for name in SIGNALS:
        x, sr = gate_inputs(name)
        for d in DELTAS:
            target = -36.0 + d
            g = 1.0
            best = None
            for _ in range(60):  # float32 loudness is only approximately scale-equivariant: home in
                y = (x * np.float32(g)).astype(np.float32)
                lk = loud(y, sr)
                err = lk - target
                if best is None or abs(err) < abs(best[1]):
                    best = (np.float32(g), err, lk)
                if abs(err) < 4e-6:
                    break
                g *= 10.0 ** (-err / 20.0)
            gf, err, lk = best
            y = (x * gf).astype(np.float32)
            keep = _reference_decision(Preprocessor, ref_mod, y, sr)
            assert keep == (not (np.isnan(lk) or lk < -36)), (name, d, lk, keep)
            o_lk = O.loudness(y, sr)
            key = f"{name}/{d:+.0e}"
            store[key + "/scale"] = gf
            store[key + "/loudness"] = np.float32(lk)
            store[key + "/keep"] = np.int32(keep)
            print(f"{key:24s} scale {float(gf):.8f} torchaudio {lk:.6f} (target {target:.4f}, off by {err:+.1e}) "
                  f"oracle {o_lk:.6f} keep {keep}")
    for sr, lens in LENGTHS.items():
        for n in lens:
            x = length_inputs(sr, n)
            keep = _reference_decision(Preprocessor, ref_mod, x, sr)
            store[f"length/{sr}/{n}/keep"] = np.int32(keep)
            print(f"length sr {sr} n {n} ({n / sr:.6f} s): keep {keep}")
    np.savez_compressed(GOLD / "gates.npz", **store)


if __name__ == "__main__":
    main()
