"""The on-disk artefact writer (everyvoice_b200/artefacts.py): file names, directory
layout and formats are the reference's (create_path preprocessor.py:502-508; save_tensor / save_wav
helpers.py:23-44; stats.json fs2/cli/preprocess.py:68-77) and every file loads with
``torch.load(path, weights_only=True)`` like the reference's datasets read it.  CPU only: the writer is host code
that receives batched results."""
import json
import wave

import numpy as np
import torch

from everyvoice_b200.artefacts import ArtefactWriter, create_path, load_ragged, spec_filename
from everyvoice_b200.audio import ProcessedAudio
from everyvoice_b200.heavy import RaggedFeatures

ITEMS = [
    {"basename": "LJ050-0269", "speaker": "default", "language": "eng"},
    {"basename": "LJ050-0270", "speaker": "spk two", "language": "und"},
    {"basename": "utt.with.dots", "speaker": "default", "language": "fra"},
]


def _feats(rng, n_rows=80, complex_=False):
    T = np.array([7, 1, 12], dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(T)]).astype(np.int64)
    width = 2 * n_rows if complex_ else n_rows
    spec = torch.from_numpy(rng.normal(size=(int(off[-1]), width)).astype(np.float32))
    energy = torch.from_numpy(rng.uniform(0, 60, size=int(off[-1])).astype(np.float32))
    return RaggedFeatures(spec, energy, off, n_rows, complex_), T, off


def test_names_and_tensor_files(tmp_path):
    rng = np.random.default_rng(0)
    feats, T, off = _feats(rng)
    with ArtefactWriter(tmp_path, workers=4) as w:
        w.write_specs(ITEMS, feats, 22050, "mel-librosa")
        w.write_energy(ITEMS, feats.energy, off)
        phone = torch.tensor([1.0, float("nan"), 1e-7, 3.5], dtype=torch.float32)
        w.write_pitch(ITEMS, phone, np.array([0, 2, 2, 4]))
    assert w.files_written == 9
    for b, item in enumerate(ITEMS):
        p = tmp_path / "spec" / f"{item['basename']}--{item['speaker']}--{item['language']}--spec-22050-mel-librosa.pt"
        assert p == create_path(tmp_path, item, "spec", spec_filename(22050, "mel-librosa")) and p.exists()
        s = torch.load(p, weights_only=True)
        assert s.dtype == torch.float32 and tuple(s.shape) == (80, int(T[b])) and s.is_contiguous()
        assert torch.equal(s, feats.utterance(b))
        assert s.untyped_storage().nbytes() == 80 * int(T[b]) * 4  # the file holds this utterance only
        e = torch.load(tmp_path / "energy" / f"{item['basename']}--{item['speaker']}--{item['language']}--energy.pt",
                       weights_only=True)
        assert torch.equal(e, feats.utterance_energy(b)) and e.untyped_storage().nbytes() == int(T[b]) * 4
    pitch1 = torch.load(create_path(tmp_path, ITEMS[1], "pitch", "pitch.pt"), weights_only=True)
    assert pitch1.shape == (0,)
    pitch0 = torch.load(create_path(tmp_path, ITEMS[0], "pitch", "pitch.pt"), weights_only=True)
    assert pitch0[0] == 1.0 and torch.isnan(pitch0[1])
    packed, o = load_ragged([create_path(tmp_path, it, "energy", "energy.pt") for it in ITEMS])
    assert torch.equal(packed, feats.energy) and np.array_equal(o, off)


def test_raw_complex_spec_files(tmp_path):
    feats, T, _ = _feats(np.random.default_rng(1), n_rows=513, complex_=True)
    with ArtefactWriter(tmp_path) as w:
        w.write_specs(ITEMS, feats, 44100, "raw")
    s = torch.load(create_path(tmp_path, ITEMS[2], "spec", "spec-44100-raw.pt"), weights_only=True)
    assert s.dtype == torch.complex64 and tuple(s.shape) == (513, int(T[2])) and torch.equal(s, feats.utterance(2))


def test_wav_files_and_skipped_items(tmp_path):
    rng = np.random.default_rng(2)
    pcm = [rng.integers(-30000, 30000, size=n).astype(np.int16) for n in (512, 1024)]
    processed = ProcessedAudio(torch.from_numpy(np.concatenate(pcm)), np.array([0, 512, 1536]), 22050, kept=[0, 2],
                               skipped={1: "audio_empty"})
    with ArtefactWriter(tmp_path) as w:
        w.write_audio(ITEMS, processed)
    assert not create_path(tmp_path, ITEMS[1], "audio", "audio-22050.wav").exists()
    for j, i in enumerate((0, 2)):
        with wave.open(str(create_path(tmp_path, ITEMS[i], "audio", "audio-22050.wav")), "rb") as f:
            assert (f.getnchannels(), f.getsampwidth(), f.getframerate()) == (1, 2, 22050)
            assert np.array_equal(np.frombuffer(f.readframes(f.getnframes()), dtype="<i2"), pcm[j])


def test_stats_json_is_merged(tmp_path):
    w = ArtefactWriter(tmp_path)
    e = {"sample_size": 5, "norm_min": -1.5, "norm_max": 2.0, "min": 1e-7, "max": 60.0, "mean": 30.0, "std": 9.0}
    w.write_stats({"energy": e})
    w.write_stats({"pitch": {**e, "mean": 180.0}})
    w.close()
    got = json.loads((tmp_path / "stats.json").read_text())
    assert got["energy"] == e and got["pitch"]["mean"] == 180.0 and set(got) == {"energy", "pitch"}


def test_worker_errors_surface(tmp_path):
    feats, _, off = _feats(np.random.default_rng(3))
    w = ArtefactWriter(tmp_path)
    w.write_energy(ITEMS, feats.energy, off)
    (tmp_path / "spec").write_text("not a directory")  # makes the spec folder impossible
    import pytest

    with pytest.raises(Exception):
        w.write_specs(ITEMS, feats, 22050, "mel")
        w.close()
