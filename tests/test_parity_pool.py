"""The every-utterance checker (tests/parity_pool.py) itself: fed the oracle's own output it must report no failure,
and a perturbed spectrogram / a moved NaN must be caught.  CPU only."""
import numpy as np
import torch

from conftest import CONFIGS


def _case(spec_type):
    from everyvoice_b200 import synth
    from oracle import ev_oracle as O

    sr, n_fft, win, hop, n_mels, f_min, f_max = CONFIGS["A"]
    xs = [synth.speech_like(hop * n + 3, sr, seed=i) for i, n in enumerate((9, 14, 7, 11))]
    durs = [synth.synthetic_durations(len(x) // hop, seed=i) for i, x in enumerate(xs)]
    durs[1] = np.append(durs[1], [4, 3]).astype(np.int64)   # overruns the frames: a NaN after clipping
    tf = O.get_spectral_transform(spec_type, n_fft, win, hop, sr, n_mels, f_min, f_max)
    outs = [O.features_one(torch.from_numpy(x), tf, hop, torch.from_numpy(d)) for x, d in zip(xs, durs)]
    x, off = synth.pack_ragged(xs)
    d, p_off = synth.pack_ragged(durs)
    spec = np.concatenate([o[0].numpy().T for o in outs])
    energy = np.concatenate([o[1].numpy() for o in outs])
    phone = np.concatenate([o[2].numpy() for o in outs])
    f_off = np.concatenate([[0], np.cumsum([len(x) // hop for x in xs])])
    return x, off, spec, energy, f_off, d, p_off, phone


def test_checker_accepts_the_oracle_and_catches_deviations():
    from parity_pool import compare_all

    x, off, spec, energy, f_off, d, p_off, phone = _case("mel")
    assert np.isnan(phone).any()
    ok = compare_all(x, off, spec, energy, f_off, CONFIGS["A"], "mel", d, p_off, phone, workers=2)
    assert ok["failures"] == [] and ok["utterances"] == 4 and ok["frames"] == int(f_off[-1]) and ok["max_spec"] <= 2e-5
    bad = spec.copy()
    bad[int(f_off[2]) + 3, 17] += 2e-3
    r = compare_all(x, off, bad, energy, f_off, CONFIGS["A"], "mel", d, p_off, phone, workers=1)
    assert [b for b, _ in r["failures"]] == [2]
    moved = phone.copy()
    k = int(np.flatnonzero(np.isnan(moved))[0])
    moved[k], moved[k - 1] = moved[k - 1], np.nan
    r = compare_all(x, off, spec, energy, f_off, CONFIGS["A"], "mel", d, p_off, moved, workers=1)
    assert [b for b, _ in r["failures"]] == [1] and "NaN" in r["failures"][0][1]


def test_checker_linear_criterion_and_pcm_input():
    from parity_pool import compare_all

    x, off, spec, energy, f_off, *_ = _case("linear")
    assert compare_all(x, off, spec, energy, f_off, CONFIGS["A"], "linear", workers=1)["failures"] == []
    strong = spec.copy()
    strong[5, int(np.argmax(spec[5]))] += 5e-3          # a strong bin off by more than 1e-3
    assert len(compare_all(x, off, strong, energy, f_off, CONFIGS["A"], "linear", workers=1)["failures"]) == 1
