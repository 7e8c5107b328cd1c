"""Host-side construction of the tables a plan uploads: analysis window and mel basis.

These are one-off, a few thousand floats, and define *which* numbers the kernel uses, so
they reproduce the reference's constructions operation by operation:

* ``hann_window_padded``  -- ``torch.hann_window(win_length)`` (periodic) centre-padded to
  ``n_fft`` as ``torch.stft`` does when ``win_length < n_fft``;
* ``melscale_fbanks_htk_slaney`` -- ``torchaudio.functional.melscale_fbanks(n_freqs, f_min,
  f_max, n_mels, sample_rate, norm="slaney", mel_scale="htk")`` as built by
  ``T.MelSpectrogram`` at ``everyvoice/utils/heavy.py:57-68`` (fp32 torch ops, same order);
* ``librosa_mel_basis`` -- ``librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)`` (Slaney
  scale, Slaney norm, float64 math, float32 result) as used at ``heavy.py:84-91``.
"""

from __future__ import annotations

import math

import numpy as np
import torch


def hann_window_padded(win_length: int, n_fft: int) -> torch.Tensor:
    if not 0 < win_length <= n_fft:
        raise ValueError(f"need 0 < win_length <= n_fft, got {win_length}, {n_fft}")
    w = torch.hann_window(win_length, periodic=True, dtype=torch.float32)
    if win_length < n_fft:
        left = (n_fft - win_length) // 2
        w = torch.nn.functional.pad(w, (left, n_fft - win_length - left))
    return w.contiguous()


def _hz_to_mel(freq: float, mel_scale: str) -> float:
    if mel_scale == "htk":
        return 2595.0 * math.log10(1.0 + (freq / 700.0))
    f_sp = 200.0 / 3
    if freq >= 1000.0:
        return 1000.0 / f_sp + math.log(freq / 1000.0) / (math.log(6.4) / 27.0)
    return freq / f_sp


def melscale_fbanks(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int,
                    norm: str | None = None, mel_scale: str = "htk") -> torch.Tensor:
    """``torchaudio.functional.melscale_fbanks`` with the same float32 torch operations in the same order
    (``mel_scale`` "htk" | "slaney", ``norm`` None | "slaney").  Returns ``fb[n_freqs, n_mels]`` float32.
    ``T.MelSpectrogram`` defaults are htk / no norm (StyleTTS2: styletts2/utils.py:12-21, losses.py:42-48); the
    reference's ``"mel"`` transform passes ``norm="slaney"`` (utils/heavy.py:57-68)."""
    if norm is not None and norm != "slaney":
        raise ValueError('norm must be one of None or "slaney"')
    if mel_scale not in ("htk", "slaney"):
        raise ValueError('mel_scale should be one of "htk" or "slaney".')
    if f_min > f_max:
        raise ValueError(f"Require f_min: {f_min} <= f_max: {f_max}")
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_pts = torch.linspace(_hz_to_mel(f_min, mel_scale), _hz_to_mel(f_max, mel_scale), n_mels + 2)
    if mel_scale == "htk":
        f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    else:
        f_sp = 200.0 / 3
        f_pts = f_sp * m_pts
        min_log_mel = 1000.0 / f_sp
        log_t = m_pts >= min_log_mel
        f_pts[log_t] = 1000.0 * torch.exp((math.log(6.4) / 27.0) * (m_pts[log_t] - min_log_mel))
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down_slopes = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up_slopes = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(torch.zeros(1), torch.min(down_slopes, up_slopes))
    if norm == "slaney":
        enorm = 2.0 / (f_pts[2 : n_mels + 2] - f_pts[:n_mels])
        fb = fb * enorm.unsqueeze(0)
    return fb.contiguous()


def melscale_fbanks_htk_slaney(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> torch.Tensor:
    """The bank of the reference's ``"mel"`` transform (htk scale, slaney norm)."""
    return melscale_fbanks(n_freqs, f_min, f_max, n_mels, sample_rate, "slaney", "htk")


def librosa_mel_basis(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: float | None) -> torch.Tensor:
    """Returns ``fb[n_freqs, n_mels]`` float32 (i.e. librosa's ``[n_mels, n_freqs]`` transposed)."""
    if fmax is None:
        fmax = float(sr) / 2
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0

    def hz_to_mel(f: float) -> float:
        if f >= min_log_hz:
            return min_log_mel + np.log(f / min_log_hz) / logstep
        return f / f_sp

    mels = np.linspace(hz_to_mel(float(fmin)), hz_to_mel(float(fmax)), n_mels + 2)
    mel_f = f_sp * mels
    log_t = mels >= min_log_mel
    mel_f[log_t] = min_log_hz * np.exp(logstep * (mels[log_t] - min_log_mel))

    # == numpy's rfftfreq(n_fft, 1 / sr), operation for operation (a frequency grid, not a transform)
    fftfreqs = np.arange(0, n_fft // 2 + 1, dtype=int) * (1.0 / (n_fft * (1.0 / sr)))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    weights = np.maximum(0, np.minimum(lower, upper))
    weights *= (2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels]))[:, None]
    return torch.from_numpy(np.ascontiguousarray(weights.astype(np.float32).T))
