"""CPU: the oracle's restatement of the audio front-end (oracle/ev_oracle.py: resample, loudness,
process_audio_tensor) against the outputs of the LIVE reference's ``Preprocessor.process_audio`` stored in
tests/golden/frontend.npz by oracle/make_golden_frontend.py."""
import math

import numpy as np
import pytest

from oracle import ev_oracle as O
from oracle.make_golden_frontend import CASES, frontend_inputs

# float32 FIR accumulation order differs between implementations; after x / max|x| * 0.95 of a low-gain input the
# absolute deviation is a few 1e-5 (about one PCM16 step = 3.05e-5)
ATOL_RESAMPLED = 5e-5
ATOL_LOUDNESS = 2e-3  # LKFS


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(golden_dir / "frontend.npz")


@pytest.mark.parametrize("name", sorted(CASES))
def test_process_audio_restatement_matches_live_reference(golden, name):
    sr_in, rs, hop, *_ = CASES[name]
    x, sr = frontend_inputs(name)
    audio, out_sr = O.process_audio_tensor(x, sr, resample_rate=rs, hop_size=hop)
    if int(golden[f"{name}/skipped"]):
        assert audio is None
        return
    ref = golden[f"{name}/audio"]
    assert audio is not None and audio.shape == ref.shape and out_sr == int(golden[f"{name}/sr"])
    assert len(audio) % hop == 0
    if rs is None or rs == sr_in:
        assert np.array_equal(audio, ref)  # division by the peak, * 0.95, truncation: bit-exact
    else:
        assert float(np.abs(audio - ref).max()) <= ATOL_RESAMPLED


@pytest.mark.parametrize("name", sorted(CASES))
def test_loudness_restatement(golden, name):
    x, sr = frontend_inputs(name)
    ref = float(golden[f"{name}/loudness"])
    got = O.loudness(x, sr)
    if math.isnan(ref):
        assert math.isnan(got) or got == -math.inf
    else:
        assert abs(got - ref) <= ATOL_LOUDNESS


def test_resample_length_and_dc_gain():
    for orig, new in ((44100, 22050), (48000, 22050), (16000, 22050), (22050, 44100)):
        for L in (1, 7, 1000, 12345):
            y = O.resample(np.ones(L, np.float32), orig, new)
            assert len(y) == math.ceil(new * L / orig)
        y = O.resample(np.ones(4000, np.float32), orig, new)
        mid = y[len(y) // 4 : 3 * len(y) // 4]
        assert float(np.abs(mid - 1.0).max()) < 2e-2  # windowed-sinc pass band


def test_pcm16_rule():
    x = np.array([0.0, 0.5, -0.5, 0.95, -1.0, 1.0, 1.5 / 32768, 2.5 / 32768, -1.5 / 32768], np.float32)
    assert O.pcm16(x).tolist() == [0, 16384, -16384, 31130, -32768, 32767, 2, 2, -2]


def test_k_weighting_is_bit_exact_with_torchaudio():
    """The oracle's biquads follow torchaudio's float32 arithmetic operation by operation (normalised coefficients,
    conv1d's fma chain, the C++ recursion's two roundings): the K-weighted signal is IDENTICAL at every rate, and the
    product's host-side coefficients (torch ops, everyvoice_b200.audio.k_weighting_coefficients) are the oracle's."""
    torchaudio = pytest.importorskip("torchaudio")
    import torch

    from everyvoice_b200 import synth
    from everyvoice_b200.audio import k_weighting_coefficients

    for sr in (11025, 16000, 22050, 44100, 48000):
        x = (synth.speech_like(int(0.5 * sr), sr, seed=3) * np.float32(0.3)).astype(np.float32)
        xt = torch.from_numpy(x)[None]
        (tb, ta), (hb, ha) = O.k_weighting_coeffs(sr)
        y1 = torchaudio.functional.treble_biquad(xt, sr, 4.0, 1500.0, 1 / math.sqrt(2))
        o1 = O._biquad_f32(x, tb, ta)
        y2 = torchaudio.functional.highpass_biquad(y1, sr, 38.0, 0.5)
        o2 = O._biquad_f32(o1, hb, ha)
        assert np.array_equal(y1[0].numpy(), o1) and np.array_equal(y2[0].numpy(), o2), sr
        f = np.float32
        want = [f(tb[0] / ta[0]), f(tb[1] / ta[0]), f(tb[2] / ta[0]), f(ta[1] / ta[0]), f(ta[2] / ta[0]),
                f(hb[0] / ha[0]), f(hb[1] / ha[0]), f(hb[2] / ha[0]), f(ha[1] / ha[0]), f(ha[2] / ha[0])]
        assert np.array_equal(k_weighting_coefficients(sr), np.asarray(want, dtype=np.float32)), sr


def test_gate_decisions_of_the_live_reference(golden_dir):
    """tests/golden/gates.npz (oracle/make_golden_gate.py): utterances scaled to -36 LKFS +- {1e-4, 1e-3, 1e-2} and
    lengths one sample either side of 0.4 s / 11 s -- the oracle takes the reference's keep / skip decision."""
    from oracle.make_golden_gate import DELTAS, LENGTHS, SIGNALS, gate_inputs, length_inputs

    gold = np.load(golden_dir / "gates.npz")
    for name in SIGNALS:
        x, sr = gate_inputs(name)
        for d in DELTAS:
            key = f"{name}/{d:+.0e}"
            y = (x * gold[key + "/scale"]).astype(np.float32)
            lk = O.loudness(y, sr)
            assert abs(lk - float(gold[key + "/loudness"])) <= 2e-5, key
            if d != 0.0:
                audio, _ = O.process_audio_tensor(y, sr, resample_rate=sr, hop_size=256)
                assert (audio is not None) == bool(gold[key + "/keep"]), key
    for sr, lens in LENGTHS.items():
        for n in lens:
            audio, _ = O.process_audio_tensor(length_inputs(sr, n), sr, resample_rate=sr, hop_size=256)
            assert (audio is not None) == bool(gold[f"length/{sr}/{n}/keep"]), (sr, n)
