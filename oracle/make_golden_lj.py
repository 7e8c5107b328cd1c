"""Generate ``tests/golden/lj_config1.npz``: BASELINE.json configs[0] -- the reference's own bundled LJ-style
dataset (``everyvoice/tests/data/metadata.psv``: LJ050-0269 .. 0273, 32.6 s at 22.05 kHz) through the numeric
flow of ``everyvoice preprocess`` with the default parameters (n_fft 1024, hop 256, 80 mels), run by the LIVE
reference from ``/root/reference``.

    python -m oracle.make_golden_lj          (build container only)

What the CLI does per file (``tests/preprocessed_audio_fixture.py:20-99`` drives exactly this) and which reference
function is executed here, unmodified:

* ``Preprocessor.process_audio`` (preprocessor.py:131-218): gates, peak normalisation, truncation.  Only
  ``load_audio`` is replaced (torchaudio.load needs torchcodec, absent here) by the wav's int16 samples / 32768,
  which is what torchaudio.load returns for a PCM16 file;
* the wav round trip ``save_wav`` -> ``load_audio`` (helpers.py:31-44, preprocessor.py:883-887): PCM16
  quantisation with the oracle's restated rule (``lrintf(x * 32768)``; parity unpinned, see ev_oracle.pcm16);
* ``process_spec``'s core (preprocessor.py:917-928): ``extract_spectral_features(audio, transform)[:, :L // hop]``
  for the default ``mel-librosa`` transform and for ``mel``;
* ``process_energy`` (preprocessor.py:632-651): ``extract_energy`` and ``average_data_by_durations`` with the
  dataset's REAL ``duration.pt`` files (phone level);
* ``compute_stats`` / ``normalize_stats`` (preprocessor.py:378-490): ``Scaler`` over the five energy files.

The five wavs themselves are stored too (int16, LJ Speech is public domain): the GPU box has no /root/reference.
"""

from __future__ import annotations

import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
GOLD = ROOT / "tests" / "golden"
NAMES = ["LJ050-0269", "LJ050-0270", "LJ050-0271", "LJ050-0272", "LJ050-0273"]


def main():
    from oracle import ev_oracle as O
    from oracle.make_golden import _import_reference, read_wav_int16

    heavy, Preprocessor, RefScaler = _import_reference()
    import everyvoice.preprocessor.preprocessor as ref_mod

    sr, n_fft, win, hop, n_mels, f_min, f_max = 22050, 1024, 1024, 256, 80, 0, 8000
    store = {}
    pcm_after = {}
    for name in NAMES:
        wav, wav_sr = read_wav_int16(REF / f"everyvoice/tests/data/lj/wavs/{name}.wav")
        assert wav_sr == sr
        x = (wav.astype(np.float32) / np.float32(32768.0)).astype(np.float32)
        ref_mod.load_audio = lambda path, _x=x: (torch.from_numpy(_x.copy())[None], sr, len(_x) / sr)
        self = types.SimpleNamespace(
            audio_config=types.SimpleNamespace(max_audio_length=11.0, min_audio_length=0.4),
            counters=types.SimpleNamespace(increment=lambda k, v=1: None), multichannel_files_list=[])
        audio, out_sr = Preprocessor.process_audio(self, f"{name}.wav", resample_rate=sr, hop_size=hop)
        assert audio is not None and out_sr == sr
        o_audio, _ = O.process_audio_tensor(x, sr, resample_rate=sr, hop_size=hop)
        assert np.array_equal(audio.numpy(), o_audio), name
        pcm = O.pcm16(audio.numpy())
        store[f"{name}/wav"] = wav
        store[f"{name}/pcm16"] = pcm
        pcm_after[name] = pcm
        print(f"{name}: {len(wav)} samples -> {len(pcm)} kept ({len(pcm) // hop} frames)")

    for st in ("mel-librosa", "mel"):
        tf = heavy.get_spectral_transform(st, n_fft, win, hop, sr, n_mels, f_min, f_max)
        otf = O.get_spectral_transform(st, n_fft, win, hop, sr, n_mels, f_min, f_max)
        scaler = RefScaler()
        phones = {}
        for name in NAMES:
            a = torch.from_numpy(pcm_after[name].astype(np.float32) / np.float32(32768.0))
            T = a.numel() // hop
            spec = Preprocessor.extract_spectral_features(None, a, tf)[:, :T]
            energy = Preprocessor.extract_energy(None, spec)
            dur = torch.load(REF / f"everyvoice/tests/data/lj/preprocessed/duration/{name}--default--default--duration.pt",
                             weights_only=True)
            phone = Preprocessor.average_data_by_durations(None, energy, dur)
            o_spec, o_energy, o_phone = O.features_one(a, otf, hop, dur)
            assert torch.equal(spec, o_spec) and torch.equal(energy, o_energy)
            assert torch.equal(torch.nan_to_num(phone, nan=-7.0), torch.nan_to_num(o_phone, nan=-7.0))
            store[f"{st}/{name}/spec"] = spec.contiguous().numpy()
            store[f"{st}/{name}/energy"] = energy.numpy()
            store[f"{st}/{name}/phone"] = phone.numpy()
            scaler.append(phone)
            phones[name] = phone
            print(f"{st:11s} {name}: T={T} P={len(dur)} sum(d)={int(dur.sum())} nan={int(torch.isnan(phone).sum())}")
        stats = scaler.calculate_stats()
        for k, v in stats.items():
            store[f"{st}/stats/{k}"] = np.float64(v)
        for name in NAMES:
            store[f"{st}/{name}/phone_norm"] = scaler.normalize(phones[name]).numpy()
        print(st, stats)
    np.savez_compressed(GOLD / "lj_config1.npz", **store)
    print("wrote", GOLD / "lj_config1.npz", (GOLD / "lj_config1.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
