"""Generate ``tests/golden/*.npz`` by running the LIVE reference from ``/root/reference``.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python -m oracle.make_golden

The reference's own functions are executed unmodified:
``everyvoice.utils.heavy.get_spectral_transform`` / ``dynamic_range_compression_torch``,
``Preprocessor.extract_spectral_features`` / ``extract_energy`` /
``average_data_by_durations`` (called unbound: they use no ``self`` state) and
``everyvoice.preprocessor.helpers.Scaler``.  Modules the reference imports for *text*
processing that are not installed here (g2p, ipatok, grapheme, nltk, panphon) are
stubbed; they are never called on this path.  ``librosa`` is not installed either:
``librosa.filters.mel`` is provided by the restatement in ``oracle.ev_oracle`` so that
the rest of the reference's ``mel-librosa`` closure (heavy.py:69-100) still runs as-is.

Inputs are seeded (``everyvoice_b200.synth``) or the committed LJ excerpt; only the
reference's OUTPUTS are stored.
"""

from __future__ import annotations

import sys
import types
import wave
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
GOLD = ROOT / "tests" / "golden"

CONFIGS = {
    # name: (sample_rate, n_fft, win_length, hop, n_mels, f_min, f_max)
    "A": (22050, 1024, 1024, 256, 80, 0, 8000),
    "B": (44100, 2048, 2048, 512, 128, 0, 8000),
    "Bfull": (44100, 2048, 2048, 512, 128, 0, 22050),
    "W": (22050, 1024, 800, 200, 80, 0, 8000),  # win_length < n_fft (window centre-padded)
    # ---- the rest of the config-field domain (n_fft 512 / 256: warp kernel with 2 / 4 packed jobs; others: any-size kernel) ----
    "S512": (16000, 512, 512, 128, 80, 0, 8000),      # 16 kHz corpora
    "W512": (16000, 512, 400, 160, 80, 0, 8000),      # win < n_fft, hop not a power of two
    "R3": (16000, 3072, 3072, 768, 80, 0, 8000),      # "output" transform of a 16 -> 48 kHz vocoder config: 1024 * 3
    "W3072": (16000, 3072, 2400, 600, 80, 0, 8000),
    "H4096": (44100, 4096, 4096, 1024, 128, 0, 8000),
    "W4096": (44100, 4096, 3000, 750, 128, 0, 11025),
    "OddHop": (22050, 2048, 2048, 441, 80, 0, 8000),  # n_fft 2048 with an odd hop (20 ms)
    "ST2": (16000, 2048, 1200, 300, 80, 0, 8000),     # StyleTTS2's spect_params (styletts2/utils.py:12-21)
    "N400": (16000, 400, 400, 160, 80, 0, 8000),      # 25 ms / 10 ms frames: n_fft = 2^4 * 5^2
    "O1001": (22050, 1001, 1001, 250, 40, 0, 8000),   # odd n_fft = 7 * 11 * 13: direct-DFT stages, 1 + (L - 1) // hop frames
    "BigHop": (22050, 1024, 1024, 1024, 80, 0, 8000), # hop = n_fft: the warp kernel's input ring does not fit
    "Gap": (22050, 512, 512, 700, 80, 0, 8000),       # hop > n_fft: samples between frames are skipped
    "S256": (8000, 256, 256, 64, 40, 0, 4000),        # 8 kHz telephone speech: four packed jobs per warp
    "W256": (8000, 256, 200, 80, 40, 0, 3800),        # win < n_fft, hop not a power of two
}
SPEC_TYPES = ("mel", "mel-librosa", "linear", "raw")


def _import_reference():
    class _Any(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return lambda *a, **kw: None

    def stub(name, **attrs):
        m = _Any(name)
        m.__dict__.update(attrs)
        m.__path__ = []
        sys.modules[name] = m
        return m

    stub("g2p", get_arpabet_langs=lambda: ([], {}), make_g2p=lambda *a, **k: None)
    for n in ("g2p.mappings", "g2p.transducer", "ipatok", "grapheme", "nltk", "nltk.tokenize",
              "panphon", "panphon.featuretable"):
        stub(n)
    from oracle.ev_oracle import librosa_mel

    def mel(sr, n_fft, n_mels, fmin, fmax):
        return librosa_mel(sr, n_fft, n_mels, fmin, fmax)

    stub("librosa")
    stub("librosa.filters", mel=mel)
    sys.path.insert(0, str(REF))
    import everyvoice.utils.heavy as heavy
    from everyvoice.preprocessor.helpers import Scaler
    from everyvoice.preprocessor.preprocessor import Preprocessor

    return heavy, Preprocessor, Scaler


def read_wav_int16(path: Path) -> tuple[np.ndarray, int]:
    with wave.open(str(path), "rb") as w:
        assert w.getsampwidth() == 2 and w.getnchannels() == 1
        sr = w.getframerate()
        data = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").copy()
    return data, sr


def inputs_for(config: str) -> dict[str, np.ndarray]:
    """The seeded inputs of the golden cases (regenerated identically by the tests)."""
    from everyvoice_b200 import synth

    sr, n_fft, win, hop, *_ = CONFIGS[config]
    n_white = 37 * hop + 17  # deliberately NOT a multiple of hop
    n_speech = 45 * hop
    out = {
        "white": synth.white_noise(n_white, seed=101),
        "speech": synth.speech_like(n_speech, sr, seed=202),
        "short": synth.white_noise(n_fft // 2 + 1 + hop, seed=303),  # just above the reflect-pad limit
    }
    if config in ("A", "W"):
        lj = np.load(GOLD / "lj_excerpt_int16.npy")
        out["lj"] = (lj.astype(np.float32) / 32768.0).astype(np.float32)
    return out


def _keep(cname: str, st: str, iname: str) -> bool:
    """Keep the fixtures small: the wide (513/1025-bin) outputs are stored for fewer inputs."""
    if st in ("linear", "raw"):
        if cname == "Bfull":  # identical to B (only the mel basis differs)
            return False
        if cname in ("B", "W", "S512", "N400", "O1001", "Gap"):
            return iname in ("speech", "short")
        if cname not in ("A",):  # the wide transforms: a few frames are enough to pin the bins
            return iname == "short"
    elif cname not in ("A", "B", "Bfull", "W"):
        return iname != "white" or cname in ("S512", "O1001")
    return True


def main():
    heavy, Preprocessor, RefScaler = _import_reference()
    from everyvoice_b200 import synth
    from oracle import ev_oracle as O

    GOLD.mkdir(parents=True, exist_ok=True)

    # ---- real-speech excerpt + the reference's bundled duration fixtures ----------------
    wav, sr = read_wav_int16(REF / "everyvoice/tests/data/lj/wavs/LJ050-0269.wav")
    assert sr == 22050
    np.save(GOLD / "lj_excerpt_int16.npy", wav[20000 : 20000 + 110 * 256])
    durs = {}
    for p in sorted((REF / "everyvoice/tests/data/lj/preprocessed/duration").glob("*.pt")):
        durs[p.name.split("--")[0]] = torch.load(p, weights_only=True).numpy().astype(np.int64)
    np.savez(GOLD / "lj_durations.npz", **durs)

    # ---- spectral features ----------------------------------------------------------------
    worst = 0.0
    for cname, (sr, n_fft, win, hop, n_mels, f_min, f_max) in CONFIGS.items():
        ins = inputs_for(cname)
        store = {}
        for st in SPEC_TYPES:
            ref_tf = heavy.get_spectral_transform(st, n_fft, win, hop, sr, n_mels, f_min, f_max)
            ora_tf = O.get_spectral_transform(st, n_fft, win, hop, sr, n_mels, f_min, f_max)
            for iname, x in ins.items():
                if not _keep(cname, st, iname):
                    continue
                xt = torch.from_numpy(x)
                T = len(x) // hop
                if st == "raw":
                    ref = ref_tf(xt)  # complex, all T+1 frames, no log
                    ora = ora_tf(xt)
                    d = float((ref - ora).abs().max())
                    store[f"{st}/{iname}/re"] = ref.real.numpy()
                    store[f"{st}/{iname}/im"] = ref.imag.numpy()
                else:
                    lin = ref_tf(xt)  # linear-domain, T+1 frames
                    spec = Preprocessor.extract_spectral_features(None, xt, ref_tf)[:, :T]
                    energy = Preprocessor.extract_energy(None, spec)
                    o_spec, o_energy, _ = O.features_one(xt, ora_tf, hop)
                    d = max(float((spec - o_spec).abs().max()), float((energy - o_energy).abs().max()))
                    store[f"{st}/{iname}/lin_last"] = lin[:, -1].numpy()  # the frame process_spec drops
                    store[f"{st}/{iname}/spec"] = spec.contiguous().numpy()
                    store[f"{st}/{iname}/energy"] = energy.numpy()
                worst = max(worst, d)
                print(f"{cname:5s} {st:11s} {iname:6s} T={T:4d} oracle-vs-reference max|d|={d:.3e}")
        np.savez_compressed(GOLD / f"spectral_{cname}.npz", **store)

    # ---- phone-level averaging --------------------------------------------------------------
    store = {}
    rng = np.random.default_rng(404)
    cases = {}
    lj = torch.from_numpy(np.load(GOLD / "lj_excerpt_int16.npy").astype(np.float32) / 32768.0)
    tf = heavy.get_spectral_transform("mel", 1024, 1024, 256, 22050, 80, 0, 8000)
    e = Preprocessor.extract_energy(None, Preprocessor.extract_spectral_features(None, lj, tf)[:, :110])
    cases["lj_energy_synthdur"] = (e.numpy(), synth.synthetic_durations(110, seed=1))
    for name, d in durs.items():  # bundled fixtures drive a synthetic value vector of their length
        T = int(d.sum()) - (1 if name.endswith("0271") else -2)  # overrun by 1 / underrun by 2
        cases[f"fixture_{name}"] = (rng.uniform(0, 60, size=T).astype(np.float32), d)
    cases["zeros_and_overrun"] = (
        rng.uniform(80, 300, size=50).astype(np.float32),
        np.array([0, 5, 0, 0, 10, 30, 4, 0, 7, 3, 2], dtype=np.int64),  # sum 61 > 50: clip then NaN
    )
    cases["single_phone"] = (rng.uniform(-1, 1, size=300).astype(np.float32), np.array([300], dtype=np.int64))
    cases["negative_duration"] = (
        rng.uniform(-1, 1, size=40).astype(np.float32),
        np.array([10, -3, 5, 8, 0, 12], dtype=np.int64),
    )
    for name, (vals, d) in cases.items():
        ref = Preprocessor.average_data_by_durations(None, torch.from_numpy(vals), torch.from_numpy(d))
        ora = O.average_data_by_durations(torch.from_numpy(vals), torch.from_numpy(d))
        assert torch.equal(torch.nan_to_num(ref, nan=-7.0), torch.nan_to_num(ora, nan=-7.0)), name
        store[f"{name}/values"] = vals
        store[f"{name}/durations"] = d
        store[f"{name}/out"] = ref.numpy()
        print(f"avg  {name:28s} P={len(d):3d} T={len(vals):4d} nan={int(torch.isnan(ref).sum())}")
    np.savez_compressed(GOLD / "average_by_durations.npz", **store)

    # ---- Scaler -----------------------------------------------------------------------------
    store = {}
    rng = np.random.default_rng(505)
    chunks = [rng.normal(30.0, 9.0, size=int(n)).astype(np.float32) for n in rng.integers(40, 120, size=7)]
    chunks[2][5] = np.nan
    chunks[4][[0, 17]] = np.nan
    ref_s, ora_s = RefScaler(), O.Scaler()
    for c in chunks:
        ref_s.append(torch.from_numpy(c))
        ora_s.append(torch.from_numpy(c))
    rs, os_ = ref_s.calculate_stats(), ora_s.calculate_stats()
    assert rs == os_, (rs, os_)
    for i, c in enumerate(chunks):
        store[f"chunk{i}"] = c
        store[f"norm{i}"] = ref_s.normalize(torch.from_numpy(c)).numpy()
    for k, v in rs.items():
        store[f"stats/{k}"] = np.float64(v)
    np.savez_compressed(GOLD / "scaler.npz", **store)
    print("scaler", rs)
    print(f"worst oracle-vs-reference deviation on spectral cases: {worst:.3e}")


if __name__ == "__main__":
    main()
