"""everyvoice_b200 -- B200 (sm_100a) implementation of EveryVoice's preprocessing
feature-extraction hot path behind the reference's own operator names.

    from everyvoice_b200 import get_spectral_transform, Preprocessor, Scaler

All numerics run in ``libevfeat.so`` (hand-written CUDA, C ABI in ``include/evfeat.h``);
importing this package never falls back to a CPU implementation.
"""

from .artefacts import ArtefactWriter, create_path, load_ragged  # noqa: F401
from .audio import AudioFrontEnd, ProcessedAudio, Resampler, loudness_batch  # noqa: F401
from .config import AudioConfig, AudioSpecTypeEnum, ConfigError  # noqa: F401
from .heavy import (  # noqa: F401
    RaggedBatch,
    RaggedFeatures,
    SpectralTransform,
    dynamic_range_compression_torch,
    get_spectral_transform,
)
from .helpers import Scaler  # noqa: F401
from .pipeline import CorpusPipeline  # noqa: F401
from .preprocessor import Preprocessor  # noqa: F401

__all__ = [
    "AudioConfig", "AudioSpecTypeEnum", "ConfigError", "RaggedBatch", "RaggedFeatures",
    "SpectralTransform", "dynamic_range_compression_torch", "get_spectral_transform",
    "Scaler", "Preprocessor", "CorpusPipeline", "AudioFrontEnd", "ProcessedAudio", "Resampler", "loudness_batch",
    "ArtefactWriter", "create_path", "load_ragged",
]
