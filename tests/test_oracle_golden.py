"""Pins the CPU oracle (oracle/ev_oracle.py) against the golden vectors that
oracle/make_golden.py produced by running the LIVE reference from /root/reference.

CPU only; runs in the driver's `-m "not gpu"` pass.  If these fail the oracle can no longer
be trusted as the checker for the CUDA path.
"""

import numpy as np
import pytest
import torch

from conftest import CONFIGS, SPEC_TYPES, golden_inputs

from oracle import ev_oracle as O

# torch.stft on CPU (MKL / pocketfft) is deterministic for a given build, but the golden
# vectors may have been produced with another BLAS thread count: allow float32 round-off.
ATOL_SAME_ALGO = 2e-5


def _tf(config, spec_type):
    sr, n_fft, win, hop, n_mels, f_min, f_max = CONFIGS[config]
    return O.get_spectral_transform(spec_type, n_fft, win, hop, sr, n_mels, f_min, f_max), hop


@pytest.mark.parametrize("config", list(CONFIGS))
@pytest.mark.parametrize("spec_type", ["mel", "mel-librosa", "linear"])
def test_oracle_reproduces_reference_log_spec_and_energy(golden_dir, config, spec_type):
    gold = np.load(golden_dir / f"spectral_{config}.npz")
    tf, hop = _tf(config, spec_type)
    n = 0
    for name, x in golden_inputs(config).items():
        if f"{spec_type}/{name}/spec" not in gold:
            continue
        spec, energy, _ = O.features_one(torch.from_numpy(x), tf, hop)
        ref = torch.from_numpy(gold[f"{spec_type}/{name}/spec"])
        assert tuple(spec.shape) == tuple(ref.shape) == (ref.shape[0], len(x) // hop)
        # `linear` has bins at the fp32 FFT noise floor where even the same algorithm on
        # another thread count moves by more than 2e-5 in the log domain: compare those
        # in the linear domain instead
        if spec_type == "linear":
            a, b = torch.exp(spec), torch.exp(ref)
            assert float((a - b).abs().max()) <= 1e-5 * float(b.max())
        else:
            assert float((spec - ref).abs().max()) <= ATOL_SAME_ALGO
        ref_e = torch.from_numpy(gold[f"{spec_type}/{name}/energy"])
        assert float((energy - ref_e).abs().max()) <= (1e-2 if spec_type == "linear" else 1e-4)
        lin_last = tf(torch.from_numpy(x))[:, -1]
        ref_last = torch.from_numpy(gold[f"{spec_type}/{name}/lin_last"])
        assert float((lin_last - ref_last).abs().max()) <= 1e-5 * max(1.0, float(ref_last.abs().max()))
        n += 1
    assert n >= 1 or (config == "Bfull" and spec_type == "linear")


@pytest.mark.parametrize("config", [c for c in CONFIGS if c != "Bfull"])
def test_oracle_reproduces_reference_raw_stft(golden_dir, config):
    gold = np.load(golden_dir / f"spectral_{config}.npz")
    tf, hop = _tf(config, "raw")
    n = 0
    for name, x in golden_inputs(config).items():
        if f"raw/{name}/re" not in gold:
            continue
        out = tf(torch.from_numpy(x))
        ref = torch.complex(torch.from_numpy(gold[f"raw/{name}/re"]), torch.from_numpy(gold[f"raw/{name}/im"]))
        assert out.dtype == torch.complex64 and tuple(out.shape) == tuple(ref.shape)
        n_fft = CONFIGS[config][1]
        assert out.shape[-1] == 1 + (len(x) + 2 * (n_fft // 2) - n_fft) // hop   # == L // hop + 1 for even n_fft
        assert float((out - ref).abs().max()) <= 1e-6 * float(ref.abs().max()) + 1e-6
        n += 1
    assert n >= 1


def test_oracle_average_by_durations_is_bit_exact(golden_dir):
    gold = np.load(golden_dir / "average_by_durations.npz")
    names = sorted({k.split("/")[0] for k in gold.files})
    assert len(names) >= 9
    n_nan = 0
    for name in names:
        out = O.average_data_by_durations(torch.from_numpy(gold[f"{name}/values"]),
                                          torch.from_numpy(gold[f"{name}/durations"]))
        ref = torch.from_numpy(gold[f"{name}/out"])
        assert torch.equal(torch.isnan(out), torch.isnan(ref))
        assert torch.equal(torch.nan_to_num(out, nan=-7.0), torch.nan_to_num(ref, nan=-7.0)), name
        n_nan += int(torch.isnan(ref).sum())
    assert n_nan >= 1  # the overrun case (clipped to empty -> NaN) is in the fixture set


def test_oracle_scaler_matches_reference(golden_dir):
    gold = np.load(golden_dir / "scaler.npz")
    s = O.Scaler()
    chunks = [torch.from_numpy(gold[f"chunk{i}"]) for i in range(7)]
    for c in chunks:
        s.append(c)
    stats = s.calculate_stats()
    assert stats["sample_size"] == int(gold["stats/sample_size"]) == 7
    for k in ("min", "max", "mean", "std", "norm_min", "norm_max"):
        assert stats[k] == pytest.approx(float(gold[f"stats/{k}"]), rel=1e-6), k
    for i, c in enumerate(chunks):
        out, ref = s.normalize(c), torch.from_numpy(gold[f"norm{i}"])
        assert torch.equal(torch.isnan(out), torch.isnan(ref))
        assert float((torch.nan_to_num(out) - torch.nan_to_num(ref)).abs().max()) <= 1e-5


def test_reference_duration_fixtures_have_the_documented_shape(golden_dir):
    """the bundled LJ duration fixtures (int64, 60-88 phones, frames - sum <= 10)."""
    d = np.load(golden_dir / "lj_durations.npz")
    assert sorted(d.files) == [f"LJ050-{n:04d}" for n in range(269, 274)]
    assert int(d["LJ050-0271"].sum()) == 644 and int(d["LJ050-0269"].sum()) == 443
    for k in d.files:
        assert d[k].dtype == np.int64 and 50 <= len(d[k]) <= 100 and d[k].min() >= 0


@pytest.mark.parametrize("config", ["A", "B", "Bfull"])
def test_librosa_mel_restatement_cross_check(config):
    """librosa 0.11.0 is not installed: filters.mel is restated from its published algorithm
    (fp64) and cross-checked against the independent fp32 Slaney/Slaney construction."""
    sr, n_fft, win, hop, n_mels, f_min, f_max = CONFIGS[config]
    w = O.librosa_mel(sr, n_fft, n_mels, f_min, f_max)
    assert w.shape == (n_mels, n_fft // 2 + 1) and w.dtype == np.float32
    fb = O.torchaudio_melscale_fbanks(n_fft // 2 + 1, float(f_min), float(f_max), n_mels, sr, "slaney", "slaney")
    assert float(np.abs(w - fb.numpy().T).max()) <= 2e-7 + 5e-6 * float(w.max())
    # triangular, adjacent, non-negative, every filter non-empty at these resolutions
    assert (w >= 0).all() and ((w > 0).sum(axis=1) >= 1).all()
    assert ((w > 0).sum(axis=0) <= 2).all()


def test_torchaudio_fbank_restatement_against_installed_torchaudio():
    ta = pytest.importorskip("torchaudio")
    for config in ("A", "B", "Bfull"):
        sr, n_fft, win, hop, n_mels, f_min, f_max = CONFIGS[config]
        ref = ta.functional.melscale_fbanks(n_fft // 2 + 1, float(f_min), float(f_max), n_mels, sr,
                                            norm="slaney", mel_scale="htk")
        ours = O.torchaudio_melscale_fbanks(n_fft // 2 + 1, float(f_min), float(f_max), n_mels, sr)
        assert torch.equal(ref, ours)


@pytest.mark.parametrize("spec_type", ["mel", "mel-librosa", "linear"])
def test_fp64_truth_agrees_with_oracle(spec_type):
    """The independent numpy fp64 evaluation and the fp32 restatement describe the same
    function (separates kernel error from reference fp32 noise in the GPU tests)."""
    from everyvoice_b200 import synth

    sr, n_fft, win, hop, n_mels, f_min, f_max = CONFIGS["A"]
    x = synth.white_noise(40 * hop, seed=11)
    tf, _ = _tf("A", spec_type)
    spec, energy, _ = O.features_one(torch.from_numpy(x), tf, hop)
    t_spec, t_energy = O.truth_features(x, spec_type, n_fft, win, hop, sr, n_mels, f_min, f_max)
    assert spec.shape == t_spec.shape
    assert float(np.abs(spec.numpy() - t_spec).max()) <= (5e-3 if spec_type == "linear" else 2e-4)
    assert float(np.abs(energy.numpy() - t_energy).max()) <= 1e-2


def test_process_spec_drops_last_frame_and_requires_reflectable_input():
    tf, hop = _tf("A", "mel")
    x = torch.zeros(10 * hop + 7)
    assert O.process_spec(x, tf, hop).shape == (80, 10)
    assert tf(x).shape == (80, 11)
    with pytest.raises(RuntimeError):  # torch.stft reflect padding needs L > n_fft // 2
        tf(torch.zeros(512))
    assert tf(torch.zeros(513)).shape == (80, 3)
    assert O.get_spectral_transform("istft", 1024, 1024, 256) is None
    assert set(SPEC_TYPES) == set(O.SPEC_TYPES)


LJ_NAMES = ["LJ050-0269", "LJ050-0270", "LJ050-0271", "LJ050-0272", "LJ050-0273"]


@pytest.mark.parametrize("spec_type", ["mel-librosa", "mel"])
def test_oracle_reproduces_config1_bundled_lj_dataset(golden_dir, spec_type):
    """BASELINE configs[0]: the reference's own 5-utterance LJ dataset through process_audio -> PCM16 wav ->
    process_spec -> process_energy (real duration.pt) -> Scaler, generated by the live reference
    (oracle/make_golden_lj.py)."""
    gold = np.load(golden_dir / "lj_config1.npz")
    durs = np.load(golden_dir / "lj_durations.npz")
    tf = O.get_spectral_transform(spec_type, 1024, 1024, 256, 22050, 80, 0, 8000)
    scaler = O.Scaler()
    phones = {}
    for name in LJ_NAMES:
        x = gold[f"{name}/wav"].astype(np.float32) / np.float32(32768.0)
        audio, sr = O.process_audio_tensor(x, 22050, resample_rate=22050, hop_size=256)
        pcm = O.pcm16(audio)
        assert sr == 22050 and np.array_equal(pcm, gold[f"{name}/pcm16"])
        spec, energy, phone = O.features_one(torch.from_numpy(pcm.astype(np.float32) / np.float32(32768.0)), tf, 256,
                                             torch.from_numpy(durs[name]))
        assert float((spec - torch.from_numpy(gold[f"{spec_type}/{name}/spec"])).abs().max()) <= ATOL_SAME_ALGO
        assert float((energy - torch.from_numpy(gold[f"{spec_type}/{name}/energy"])).abs().max()) <= 1e-4
        assert np.allclose(phone.numpy(), gold[f"{spec_type}/{name}/phone"], atol=1e-4, equal_nan=True)
        scaler.append(phone)
        phones[name] = phone
    stats = scaler.calculate_stats()
    assert stats["sample_size"] == 5
    for k in ("min", "max", "mean", "std", "norm_min", "norm_max"):
        assert stats[k] == pytest.approx(float(gold[f"{spec_type}/stats/{k}"]), abs=1e-4)
    for name in LJ_NAMES:
        assert np.allclose(scaler.normalize(phones[name]).numpy(), gold[f"{spec_type}/{name}/phone_norm"], atol=1e-4,
                           equal_nan=True)
