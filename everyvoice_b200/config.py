"""The parameter contract of the hot path: a plain-dataclass mirror of the fields of
``everyvoice/config/preprocessing_config.py:18-91`` (``AudioSpecTypeEnum``, ``AudioConfig``)
that ``get_spectral_transform`` / ``Preprocessor`` read.  The reference's pydantic models,
YAML loading and path handling are out of scope; an object with these attribute names
(including the reference's own ``AudioConfig``) can be passed wherever this one is."""

from __future__ import annotations

from dataclasses import dataclass
from enum import Enum


class AudioSpecTypeEnum(str, Enum):
    mel = "mel"  # TorchAudio implementation
    mel_librosa = "mel-librosa"  # Librosa implementation
    linear = "linear"  # TorchAudio Linear Spectrogram
    raw = "raw"  # TorchAudio Complex Spectrogram


class ConfigError(Exception):
    """Reference: ``everyvoice/exceptions.py`` -- raised for an unsupported spec_type
    (preprocessor/preprocessor.py:123-129)."""


@dataclass
class AudioConfig:
    min_audio_length: float = 0.4
    max_audio_length: float = 11.0
    max_wav_value: float = 32767.0
    input_sampling_rate: int = 22050
    output_sampling_rate: int = 22050
    alignment_sampling_rate: int = 22050
    target_bit_depth: int = 16
    n_fft: int = 1024
    fft_window_size: int = 1024
    fft_hop_size: int = 256
    f_min: int = 0
    f_max: int = 8000
    n_mels: int = 80
    spec_type: str = AudioSpecTypeEnum.mel_librosa.value
    vocoder_segment_size: int = 8192
