"""Generate ``tests/golden/frontend.npz`` by running the LIVE reference's ``Preprocessor.process_audio``
(everyvoice/preprocessor/preprocessor.py:131-218) from ``/root/reference`` on seeded waveforms.

    python -m oracle.make_golden_frontend          (build container only)

``process_audio`` is executed unmodified; only ``load_audio`` (torchaudio.load needs torchcodec, absent here) is
replaced by a function that hands it the seeded tensor, which is what torchaudio.load would return for a float wav.
``torchaudio.functional.resample`` / ``transforms.Loudness`` are this container's torchaudio (2.11; the reference
pins 2.7.1).  Only the reference's OUTPUTS are stored; inputs are regenerated from their seeds by the tests
(``frontend_inputs``)."""

from __future__ import annotations

import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"

# name: (input rate, resample_rate, hop, seconds, kind, gain)
CASES = {
    "same_rate_speech": (22050, 22050, 256, 1.30, "speech", 0.6),
    "same_rate_white": (22050, None, 256, 0.90, "white", 0.3),
    "down_2to1_speech": (44100, 22050, 256, 1.10, "speech", 0.8),
    "down_48k_speech": (48000, 22050, 256, 0.75, "speech", 0.5),
    "up_16k_white": (16000, 22050, 256, 1.00, "white", 0.2),
    "down_44k_hop512": (48000, 44100, 512, 0.60, "speech", 0.7),
    "quiet_gated": (22050, 22050, 256, 1.00, "white", 0.004),      # about -50 LKFS: skipped as "audio_empty"
    "silence_gated": (22050, 22050, 256, 0.80, "zeros", 0.0),       # loudness NaN/-inf
    "too_short": (22050, 22050, 256, 0.30, "white", 0.5),
    "too_long": (16000, 22050, 256, 11.5, "white", 0.5),
}


def frontend_inputs(name: str) -> tuple[np.ndarray, int]:
    """The seeded input waveform of a golden case (float32 mono) and its sampling rate."""
    from everyvoice_b200 import synth

    sr, _, _, seconds, kind, gain = CASES[name]
    n = int(round(seconds * sr)) + 37  # not a multiple of any hop
    seed = 9000 + sorted(CASES).index(name)
    if kind == "speech":
        x = synth.speech_like(n, sr, seed=seed)
    elif kind == "white":
        x = synth.white_noise(n, seed=seed)
    else:
        x = np.zeros(n, np.float32)
    return (x * np.float32(gain)).astype(np.float32), sr


def main():
    import torchaudio

    from oracle import ev_oracle as O
    from oracle.make_golden import _import_reference

    _, Preprocessor, _ = _import_reference()
    import everyvoice.preprocessor.preprocessor as ref_mod

    store = {}
    for name, (sr, rs, hop, *_rest) in CASES.items():
        x, sr = frontend_inputs(name)
        ref_mod.load_audio = lambda path, _x=x, _sr=sr: (torch.from_numpy(_x.copy())[None], _sr, len(_x) / _sr)
        counters = types.SimpleNamespace(counts={}, increment=lambda k, v=1: None)
        self = types.SimpleNamespace(
            audio_config=types.SimpleNamespace(max_audio_length=11.0, min_audio_length=0.4),
            counters=counters, multichannel_files_list=[])
        audio, out_sr = Preprocessor.process_audio(self, "seeded.wav", resample_rate=rs, hop_size=hop)
        lk = float(torchaudio.functional.loudness(torch.from_numpy(x)[None], sr)) if len(x) >= int(round(0.4 * sr)) else float("nan")
        o_audio, o_sr = O.process_audio_tensor(x, sr, resample_rate=rs, hop_size=hop)
        o_lk = O.loudness(x, sr)
        store[f"{name}/loudness"] = np.float32(lk)
        if audio is None:
            assert o_audio is None, (name, o_sr)
            store[f"{name}/skipped"] = np.int32(1)
            print(f"{name:20s} skipped ({o_sr}); loudness ref {lk:.4f} oracle {o_lk:.4f}")
            continue
        assert o_audio is not None, name
        a = audio.numpy()
        assert a.shape == o_audio.shape and out_sr == o_sr, (name, a.shape, o_audio.shape)
        d = float(np.abs(a - o_audio).max())
        store[f"{name}/skipped"] = np.int32(0)
        store[f"{name}/audio"] = a
        store[f"{name}/sr"] = np.int32(out_sr)
        print(f"{name:20s} L={len(x):6d} -> {len(a):6d} @ {out_sr}  oracle-vs-reference max|d|={d:.3e}  "
              f"loudness ref {lk:.4f} oracle {o_lk:.4f}")
    np.savez_compressed(GOLD / "frontend.npz", **store)


if __name__ == "__main__":
    main()
