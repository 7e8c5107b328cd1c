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
    "W": (22050, 1024, 800, 200, 80, 0, 8000),
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
