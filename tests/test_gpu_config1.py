"""BASELINE.json configs[0] on the GPU: the reference's bundled 5-utterance LJ dataset (32.6 s, 22.05 kHz, default
parameters) through the whole numeric flow of `everyvoice preprocess` -- process_audio -> PCM16 -> process_spec ->
process_energy with the dataset's real duration.pt files -> compute_stats / normalize_stats -- against what the LIVE
reference produced (tests/golden/lj_config1.npz, oracle/make_golden_lj.py)."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

NAMES = ["LJ050-0269", "LJ050-0270", "LJ050-0271", "LJ050-0272", "LJ050-0273"]
FRAMES = [445, 603, 643, 644, 469]


@pytest.mark.parametrize("spec_type", ["mel-librosa", "mel"])
def test_bundled_lj_dataset_end_to_end(cuda_device, golden_dir, spec_type):
    import everyvoice_b200 as ev

    gold = np.load(golden_dir / "lj_config1.npz")
    durs = np.load(golden_dir / "lj_durations.npz")
    pre = ev.Preprocessor(ev.AudioConfig(spec_type=spec_type), device=cuda_device)   # all defaults: 1024 / 256 / 80
    wavs = [torch.from_numpy(gold[f"{n}/wav"]) for n in NAMES]                        # the files' own int16 PCM
    audio = pre.process_audio_batch(wavs, 22050, resample_rate=22050, hop_size=256, out_dtype=torch.int16)
    assert audio.kept == list(range(5)) and audio.sr == 22050
    for b, n in enumerate(NAMES):   # same-rate path: peak division, * 0.95, truncation, PCM16 -- bit exact
        assert np.array_equal(audio.utterance(b).cpu().numpy(), gold[f"{n}/pcm16"]), n
    feats = pre.process_spec_batch(audio.samples, audio.offsets)
    assert np.diff(feats.frame_offsets).tolist() == FRAMES
    phone, p_off = pre.process_energy_batch(feats, [torch.from_numpy(durs[n]) for n in NAMES])
    raw_phone = phone.clone()
    e_scaler, _ = pre.compute_stats(energy=phone, n_energy_files=5)
    stats = pre.normalize_stats(e_scaler, None, distributed=False)["energy"]
    worst = 0.0
    for b, n in enumerate(NAMES):
        spec = feats.utterance(b).cpu()
        ref = torch.from_numpy(gold[f"{spec_type}/{n}/spec"])
        assert tuple(spec.shape) == tuple(ref.shape)
        worst = max(worst, float((spec - ref).abs().max()))
        assert float((feats.utterance_energy(b).cpu() - torch.from_numpy(gold[f"{spec_type}/{n}/energy"])).abs().max()) <= 1e-3
        got = raw_phone[p_off[b]:p_off[b + 1]].cpu().numpy()
        want = gold[f"{spec_type}/{n}/phone"]
        assert np.array_equal(np.isnan(got), np.isnan(want))          # Python slice semantics: same NaN positions
        assert np.allclose(got, want, atol=1e-3, equal_nan=True)
        got_n = phone[p_off[b]:p_off[b + 1]].cpu().numpy()
        assert np.allclose(got_n, gold[f"{spec_type}/{n}/phone_norm"], atol=1e-3, equal_nan=True)
    assert worst <= 1e-3, worst
    assert stats["sample_size"] == 5
    for k in ("min", "max", "mean", "std", "norm_min", "norm_max"):
        assert stats[k] == pytest.approx(float(gold[f"{spec_type}/stats/{k}"]), abs=1e-3), k
