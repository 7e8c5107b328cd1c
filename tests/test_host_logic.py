"""Host-side logic of the product (no GPU): table construction, frame arithmetic, sharding,
statistics finalisation, the reference's operator names -- and that nothing computes on the CPU."""

import ast
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import CONFIGS

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "everyvoice_b200"


# ------------------------------------------------------------------------------------------
# tables the plan uploads == what the reference builds
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("config", list(CONFIGS))
def test_mel_tables_equal_the_oracles(config):
    from everyvoice_b200 import filterbanks
    from oracle import ev_oracle as O

    sr, n_fft, win, hop, n_mels, f_min, f_max = CONFIGS[config]
    fb = filterbanks.melscale_fbanks_htk_slaney(n_fft // 2 + 1, float(f_min), float(f_max), n_mels, sr)
    assert torch.equal(fb, O.torchaudio_melscale_fbanks(n_fft // 2 + 1, float(f_min), float(f_max), n_mels, sr))
    lb = filterbanks.librosa_mel_basis(sr, n_fft, n_mels, f_min, f_max)
    assert np.array_equal(lb.numpy(), O.librosa_mel(sr, n_fft, n_mels, f_min, f_max).T)
    assert lb.shape == fb.shape == (n_fft // 2 + 1, n_mels) and lb.is_contiguous()
    # the compressed two-weights-per-bin table of the kernel relies on this structure
    for m in (fb.numpy(), lb.numpy()):
        nz = m != 0
        assert (nz.sum(axis=1) <= 2).all()
        rows = np.nonzero(nz.sum(axis=1) == 2)[0]
        cols = [np.nonzero(nz[r])[0] for r in rows]
        assert all(c[1] == c[0] + 1 for c in cols)


def test_window_matches_torch_stft_padding():
    from everyvoice_b200 import filterbanks

    assert torch.equal(filterbanks.hann_window_padded(1024, 1024), torch.hann_window(1024))
    w = filterbanks.hann_window_padded(800, 1024)
    assert w.shape == (1024,) and torch.equal(w[112:912], torch.hann_window(800))
    assert float(w[:112].abs().sum()) == 0.0 and float(w[912:].abs().sum()) == 0.0
    with pytest.raises(ValueError):
        filterbanks.hann_window_padded(2048, 1024)


def test_frame_arithmetic_is_the_references():
    """T = L // hop (process_spec, preprocessor.py:921) and T + 1 for the bare transform."""
    import everyvoice_b200 as ev

    tf = ev.get_spectral_transform("mel", 1024, 1024, 256, 22050, 80, 0, 8000)
    for L in (513, 1024, 25600, 25601, 25855, 25856, 220500):
        assert tf.num_frames(L) == L // 256
        assert tf.num_frames(L, keep_last=True) == L // 256 + 1
    assert tf.n_rows == 80 and tf.n_freqs == 513 and not tf.is_complex
    lin = ev.get_spectral_transform("linear", 2048, 2048, 512)
    assert lin.n_rows == 1025 and lin.mel_fb is None
    # odd n_fft: torch.stft pads n_fft // 2 on both sides, so the bare transform has 1 + (L - 1) // hop frames
    from oracle import ev_oracle as O

    for n_fft, hop in ((1001, 250), (15, 4), (401, 160)):
        odd = ev.get_spectral_transform("linear", n_fft, n_fft, hop)
        ref = O.get_spectral_transform("linear", n_fft, n_fft, hop)
        for L in (n_fft // 2 + 1, 4 * hop, 4 * hop + 1, 7 * hop - 1, 1234):
            if L <= n_fft // 2:
                continue
            assert odd.num_frames(L, keep_last=True) == ref(torch.zeros(L)).shape[-1], (n_fft, hop, L)
            assert odd.num_frames(L) == L // hop <= odd.num_frames(L, keep_last=True)
    assert ev.get_spectral_transform("raw", 1024, 1024, 256).is_complex


def test_get_spectral_transform_surface():
    """Same names / return convention as everyvoice/utils/heavy.py:47-119."""
    import everyvoice_b200 as ev

    for st in ("mel", "mel-librosa", "linear", "raw"):
        assert ev.get_spectral_transform(st, 1024, 1024, 256, 22050, 80, 0, 8000) is not None
    assert ev.get_spectral_transform(ev.AudioSpecTypeEnum.mel, 1024, 1024, 256, 22050, 80) is not None
    assert ev.get_spectral_transform("istft", 1024, 1024, 256) is None
    assert ev.get_spectral_transform("bogus", 1024, 1024, 256) is None
    with pytest.raises(ev.ConfigError):
        ev.Preprocessor(ev.AudioConfig(spec_type="bogus"))
    ac = ev.AudioConfig()
    assert (ac.n_fft, ac.fft_window_size, ac.fft_hop_size, ac.n_mels, ac.f_min, ac.f_max, ac.spec_type) == \
        (1024, 1024, 256, 80, 0, 8000, "mel-librosa")  # preprocessing_config.py:38-85 defaults
    # output transform scales n_fft / win / hop by output_sr // input_sr and keeps the INPUT rate for the basis
    pre = ev.Preprocessor(ev.AudioConfig(spec_type="mel", output_sampling_rate=44100))
    o = pre.output_spectral_transform
    assert (o.n_fft, o.win_length, o.hop_length, o.sample_rate) == (2048, 2048, 512, 22050)
    # a config object shaped like the reference's (config.preprocessing.audio) is accepted
    class _P:  # noqa: E306
        audio = ev.AudioConfig(spec_type="linear")
    class _C:  # noqa: E306
        preprocessing = _P()
    assert ev.Preprocessor(_C()).input_spectral_transform.spec_type == "linear"


# ------------------------------------------------------------------------------------------
# no CPU path
# ------------------------------------------------------------------------------------------
def test_product_raises_without_cuda():
    import everyvoice_b200 as ev

    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU-only pass")
    tf = ev.get_spectral_transform("mel", 1024, 1024, 256, 22050, 80, 0, 8000)
    pre = ev.Preprocessor(ev.AudioConfig(spec_type="mel"))
    x = torch.zeros(4096)
    for call in (lambda: tf(x), lambda: pre.extract_spectral_features(x, tf),
                 lambda: pre.extract_energy(torch.zeros(80, 10)),
                 lambda: pre.average_data_by_durations(torch.zeros(10), torch.tensor([5, 5])),
                 lambda: ev.dynamic_range_compression_torch(x),
                 lambda: ev.Scaler().normalize(x) if False else ev.Scaler().partial_stats()):
        with pytest.raises(RuntimeError, match="no CPU"):
            call()


def test_product_never_imports_the_oracle_or_a_fallback():
    banned = re.compile(r"\b(oracle|torchaudio|librosa|triton)\b")
    for py in PKG.glob("*.py"):
        tree = ast.parse(py.read_text())
        for node in ast.walk(tree):
            mods = []
            if isinstance(node, ast.Import):
                mods = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                mods = [node.module or ""]
            for m in mods:
                assert not banned.search(m), f"{py.name} imports {m}"
            # no torch spectral op anywhere in the product (docstrings may cite them)
            if isinstance(node, ast.Attribute) and node.attr in ("stft", "fft", "rfft", "istft", "matmul"):
                raise AssertionError(f"{py.name}:{node.lineno} uses .{node.attr}")


# ------------------------------------------------------------------------------------------
# sharding and statistics finalisation
# ------------------------------------------------------------------------------------------
def test_shard_utterances_is_a_balanced_partition():
    from everyvoice_b200 import synth
    from everyvoice_b200.distributed import shard_utterances

    lens = synth.utterance_lengths(1000, 22050, 256, seed=1234)
    for world in (1, 2, 4, 8):
        shards = shard_utterances(lens, world)
        assert sorted(i for s in shards for i in s) == list(range(1000))
        loads = np.array([lens[s].sum() for s in shards])
        assert loads.max() - loads.min() <= lens.max()
        assert loads.max() / loads.mean() < 1.01
        assert shards == shard_utterances(lens, world)  # deterministic: every rank computes the same split
    assert shard_utterances([], 4) == [[], [], [], []]


def _five(x: np.ndarray):
    v = x[~np.isnan(x)].astype(np.float64)
    return [float(v.size), float(v.sum()), float((v * v).sum()), float(v.min()), float(v.max())]


def test_finalize_stats_equals_reference_scaler():
    from everyvoice_b200.distributed import finalize_stats
    from oracle import ev_oracle as O

    rng = np.random.default_rng(3)
    chunks = [rng.normal(5.0, 2.0, size=n).astype(np.float32) for n in (50, 77, 1, 300)]
    chunks[1][3] = np.nan
    s = O.Scaler()
    for c in chunks:
        s.append(torch.from_numpy(c))
    ref = s.calculate_stats()
    ours = finalize_stats(_five(np.concatenate(chunks)), sample_size=len(chunks))
    assert set(ours) == set(ref)
    assert ours["sample_size"] == ref["sample_size"] == 4
    assert ours["min"] == ref["min"] and ours["max"] == ref["max"]
    for k in ("mean", "std", "norm_min", "norm_max"):
        assert ours[k] == pytest.approx(ref[k], rel=2e-6)
    with pytest.raises(ValueError):
        finalize_stats([0.0, 0.0, 0.0, float("inf"), float("-inf")], 0)


def test_scaler_surface_matches_reference():
    import everyvoice_b200 as ev

    s = ev.Scaler()
    assert len(s) == 0 and s.calculate_stats(distributed=False) is None  # helpers.py:87-88
    s.append(torch.zeros(3))
    assert len(s) == 1 and len(s.data) == 1
    with pytest.raises(ValueError):
        s.data = []  # helpers.py:64-68
    s.clear_data()
    assert len(s) == 0 and s.mean is None
    s.mean, s.std = torch.tensor(2.0), torch.tensor(4.0)
    assert float(s.denormalize(torch.tensor(0.5))) == 4.0


def test_synthetic_workload_shapes():
    """1k utterances of 1-10 s at 22.05 kHz, lengths multiples of hop."""
    from everyvoice_b200 import synth

    lens = synth.utterance_lengths(1000, 22050, 256, seed=1234)
    assert (lens % 256 == 0).all() and lens.min() >= 22050 - 256 and lens.max() <= 220500
    assert 5000 < lens.sum() / 22050 < 6000
    d = synth.synthetic_durations(500, seed=4)
    assert d.dtype == np.int64 and len(d) == 500 // 7 and abs(int(d.sum()) - 500) <= 2 and (d == 0).any()
    x = synth.speech_like(4000, 22050, seed=1)
    assert x.dtype == np.float32 and np.abs(x).max() <= 1.0
    packed, off = synth.pack_ragged([np.ones(3), np.ones(5)])
    assert off.tolist() == [0, 3, 8] and packed.size == 8
