import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"

# name: (sample_rate, n_fft, win_length, hop, n_mels, f_min, f_max) -- same table as oracle/make_golden.py
CONFIGS = {
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


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 (B200) device and libevfeat.so")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def golden_inputs(config: str) -> dict:
    """The seeded inputs of the golden cases; must match oracle.make_golden.inputs_for."""
    from everyvoice_b200 import synth

    sr, n_fft, win, hop, *_ = CONFIGS[config]
    out = {
        "white": synth.white_noise(37 * hop + 17, seed=101),
        "speech": synth.speech_like(45 * hop, sr, seed=202),
        "short": synth.white_noise(n_fft // 2 + 1 + hop, seed=303),
    }
    if config in ("A", "W"):
        lj = np.load(GOLDEN / "lj_excerpt_int16.npy")
        out["lj"] = (lj.astype(np.float32) / 32768.0).astype(np.float32)
    return out


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from everyvoice_b200 import _lib

    _lib.load()  # raises (does not skip) if the extension is missing on a GPU box
    return torch.device("cuda", 0)
