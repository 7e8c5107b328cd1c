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
