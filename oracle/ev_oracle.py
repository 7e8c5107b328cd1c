"""CPU restatement of EveryVoice's preprocessing feature-extraction hot path.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``): the checker for the CUDA path and the
CPU baseline timed by ``bench.py``.  Never imported by ``everyvoice_b200``.

Every function cites the reference lines it follows (paths relative to
``/root/reference``).  The arithmetic of the path lives in third-party code that is not
under ``/root/reference``:

* torchaudio 2.7.1 (pinned, ``pyproject.toml:80``) -- ``transforms.MelSpectrogram`` /
  ``Spectrogram`` / ``MelScale``, ``functional.spectrogram`` / ``melscale_fbanks``;
* torch 2.7.1 -- ``torch.stft``, ``hann_window``, ``linalg.norm``, ``std``, ``nanmean``;
* librosa 0.11.0 (``pyproject.toml:58``) -- ``filters.mel``.

The restatement below calls the same torch primitives torchaudio calls (``torch.stft``
with ``center=True, pad_mode="reflect", onesided=True``; fp32 ``matmul``), in the same
order, so on CPU it reproduces the reference bit-for-bit for ``mel`` / ``linear`` / ``raw``
(checked against ``tests/golden`` which was generated from the live reference).  An
independent numpy fp64 evaluation (``truth_*``) is kept alongside to separate "kernel
error" from "reference fp32 noise".
"""

from __future__ import annotations

import math

import numpy as np
import torch

SPEC_TYPES = ("mel", "mel-librosa", "linear", "raw")


# --------------------------------------------------------------------------------------
# Mel filterbanks
# --------------------------------------------------------------------------------------
def _hz_to_mel_htk(freq: float) -> float:
    # torchaudio.functional.functional._hz_to_mel, mel_scale="htk"
    return 2595.0 * math.log10(1.0 + (freq / 700.0))


def _hz_to_mel_slaney(freq: float) -> float:
    # torchaudio.functional.functional._hz_to_mel, mel_scale="slaney"
    f_sp = 200.0 / 3
    mels = freq / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    if freq >= min_log_hz:
        mels = min_log_mel + math.log(freq / min_log_hz) / logstep
    return mels


def torchaudio_melscale_fbanks(
    n_freqs: int,
    f_min: float,
    f_max: float,
    n_mels: int,
    sample_rate: int,
    norm: str | None = "slaney",
    mel_scale: str = "htk",
) -> torch.Tensor:
    """``torchaudio.functional.melscale_fbanks`` restated with the same fp32 torch ops.

    This is what ``T.MelSpectrogram(..., norm="slaney")`` builds at
    ``everyvoice/utils/heavy.py:57-68`` (mel_scale defaults to "htk").
    Returns ``fb[n_freqs, n_mels]`` float32.
    """
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    if mel_scale == "htk":
        m_min, m_max = _hz_to_mel_htk(f_min), _hz_to_mel_htk(f_max)
    else:
        m_min, m_max = _hz_to_mel_slaney(f_min), _hz_to_mel_slaney(f_max)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    if mel_scale == "htk":
        f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    else:
        f_sp = 200.0 / 3
        f_pts = f_sp * m_pts
        min_log_hz = 1000.0
        min_log_mel = min_log_hz / f_sp
        logstep = math.log(6.4) / 27.0
        log_t = m_pts >= min_log_mel
        f_pts[log_t] = min_log_hz * torch.exp(logstep * (m_pts[log_t] - min_log_mel))
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    zero = torch.zeros(1)
    down_slopes = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up_slopes = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(zero, torch.min(down_slopes, up_slopes))
    if norm == "slaney":
        enorm = 2.0 / (f_pts[2 : n_mels + 2] - f_pts[:n_mels])
        fb = fb * enorm.unsqueeze(0)
    return fb


def librosa_mel(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    """``librosa.filters.mel(sr=, n_fft=, n_mels=, fmin=, fmax=)`` (librosa 0.11.0 defaults:
    ``htk=False`` i.e. Slaney scale, ``norm="slaney"``, ``dtype=float32``), restated from
    its published algorithm; call site ``everyvoice/utils/heavy.py:70,84-91``.
    Returns ``w[n_mels, 1 + n_fft//2]`` float32 (math in float64, as librosa does).
    """
    if fmax is None:
        fmax = float(sr) / 2
    n_freqs = 1 + n_fft // 2
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)

    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0

    def hz_to_mel(f):
        f = np.asanyarray(f, dtype=np.float64)
        mels = f / f_sp
        if f.ndim:
            log_t = f >= min_log_hz
            mels[log_t] = min_log_mel + np.log(f[log_t] / min_log_hz) / logstep
        elif f >= min_log_hz:
            mels = min_log_mel + np.log(f / min_log_hz) / logstep
        return mels

    def mel_to_hz(m):
        m = np.asanyarray(m, dtype=np.float64)
        freqs = f_sp * m
        log_t = m >= min_log_mel
        freqs[log_t] = min_log_hz * np.exp(logstep * (m[log_t] - min_log_mel))
        return freqs

    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, n_freqs), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights.astype(np.float32)


# --------------------------------------------------------------------------------------
# Spectral transforms  (everyvoice/utils/heavy.py:39-119)
# --------------------------------------------------------------------------------------
def _spectrogram(x: torch.Tensor, n_fft: int, win_length: int, hop_length: int, power):
    """``torchaudio.functional.spectrogram(pad=0, window=hann_window(win_length),
    normalized=False, center=True, pad_mode="reflect", onesided=True)``: what
    ``T.Spectrogram`` / ``T.MelSpectrogram.spectrogram`` run for heavy.py:58,72,102,108.
    """
    window = torch.hann_window(win_length)
    shape = x.size()
    x2 = x.reshape(-1, shape[-1])
    spec_f = torch.stft(
        input=x2,
        n_fft=n_fft,
        hop_length=hop_length,
        win_length=win_length,
        window=window,
        center=True,
        pad_mode="reflect",
        normalized=False,
        onesided=True,
        return_complex=True,
    )
    spec_f = spec_f.reshape(shape[:-1] + spec_f.shape[-2:])
    if power is None:
        return spec_f
    if power == 1.0:
        return spec_f.abs()
    return spec_f.abs().pow(power)


def get_spectral_transform(
    spec_type,
    n_fft,
    win_length,
    hop_length,
    sample_rate=None,
    n_mels=None,
    f_min=0,
    f_max=8000,
):
    """Restatement of ``everyvoice/utils/heavy.py:47-119``: returns a callable
    ``x[..., L] -> [..., F, 1 + L//hop]`` or ``None`` for an unsupported type
    ("istft" is out of scope for this path and also returns ``None`` here)."""
    if spec_type == "mel":  # heavy.py:57-68
        fb = torchaudio_melscale_fbanks(
            n_fft // 2 + 1, float(f_min), float(f_max), n_mels, sample_rate, "slaney", "htk"
        )

        def mel_tf(x):
            spec = _spectrogram(x, n_fft, win_length, hop_length, 2.0)
            # torchaudio.transforms.MelScale.forward
            return torch.matmul(spec.transpose(-1, -2), fb).transpose(-1, -2)

        return mel_tf
    if spec_type == "mel-librosa":  # heavy.py:69-100
        mel_basis = torch.from_numpy(librosa_mel(sample_rate, n_fft, n_mels, f_min, f_max)).float()

        def mel_librosa_tf(x):
            spec = _spectrogram(x, n_fft, win_length, hop_length, 2.0)
            sine_windowed_spec = torch.sqrt(spec + 1e-9)
            return torch.matmul(mel_basis, sine_windowed_spec)

        return mel_librosa_tf
    if spec_type == "linear":  # heavy.py:101-106 (power=2 by torchaudio default)
        return lambda x: _spectrogram(x, n_fft, win_length, hop_length, 2.0)
    if spec_type == "raw":  # heavy.py:107-113
        return lambda x: _spectrogram(x, n_fft, win_length, hop_length, None)
    return None


def dynamic_range_compression_torch(x, C=1, clip_val=1e-5):
    """``everyvoice/utils/heavy.py:39-40``."""
    return torch.log(torch.clamp(x, min=clip_val) * C)


def extract_spectral_features(audio_tensor, transform, normalize=True):
    """``Preprocessor.extract_spectral_features``, preprocessor/preprocessor.py:220-233."""
    mel = transform(audio_tensor)
    if normalize:
        mel = dynamic_range_compression_torch(mel)
    return mel


def process_spec(audio: torch.Tensor, transform, hop_size: int):
    """The in-memory core of ``Preprocessor.process_spec``, preprocessor.py:917-928:
    ``max_frames = L // hop``; ``extract_spectral_features(...)[:, :max_frames]``."""
    audio = audio.squeeze()
    max_frames = audio.size(0) // hop_size
    spec = extract_spectral_features(audio, transform)[:, :max_frames]
    assert max_frames == spec.size(1)
    return spec


def extract_energy(spectral_feature_tensor: torch.Tensor):
    """``Preprocessor.extract_energy``, preprocessor.py:302-309: L2 norm over the
    frequency axis of the (log-compressed) spectrogram."""
    return torch.linalg.norm(spectral_feature_tensor, dim=0)


def average_data_by_durations(data: torch.Tensor, durations: torch.Tensor):
    """``Preprocessor.average_data_by_durations``, preprocessor.py:287-300 -- the same
    Python loop (slices clip, ``d <= 0`` gives 1e-7, empty positive slice gives NaN)."""
    current_frame_position = 0
    new_data = []
    for duration in durations.numpy().tolist():
        if duration > 0:
            new_data.append(
                torch.mean(data[current_frame_position : current_frame_position + duration])
            )
        else:
            new_data.append(1e-7)
        current_frame_position += duration
    return torch.tensor(new_data)


class Scaler:
    """``everyvoice/preprocessor/helpers.py:47-106``."""

    def __init__(self):
        self._data = []
        self._tensor_data = None
        self.min = self.max = self.std = self.mean = None
        self.norm_min = self.norm_max = None

    def __len__(self):
        return len(self._data)

    @property
    def data(self):
        return self._data

    def append(self, value):
        self._data.append(value)

    def clear_data(self):
        self.__init__()

    def normalize(self, data):
        return (data - self.mean) / self.std

    def denormalize(self, data):
        return (data * self.std) + self.mean

    def calculate_stats(self):
        if not len(self):
            return
        if self._tensor_data is None:
            self._tensor_data = torch.cat(self._data)
        non_nan_data = self._tensor_data[~torch.isnan(self._tensor_data)]
        self.min = torch.min(non_nan_data)
        self.max = torch.max(non_nan_data)
        self.mean = torch.nanmean(self._tensor_data)
        self.std = torch.std(non_nan_data)
        self.norm_max = self.normalize(self.max)
        self.norm_min = self.normalize(self.min)
        return {
            "sample_size": len(self),
            "norm_min": float(self.norm_min),
            "norm_max": float(self.norm_max),
            "min": float(self.min),
            "max": float(self.max),
            "mean": float(self.mean),
            "std": float(self.std),
        }


# --------------------------------------------------------------------------------------
# Whole-path helpers used by the parity tests and the CPU baseline
# --------------------------------------------------------------------------------------
def features_one(audio: torch.Tensor, transform, hop: int, durations: torch.Tensor | None = None):
    """``process_spec`` (preprocessor.py:917-928) then ``process_energy``
    (preprocessor.py:641-650) for one in-memory utterance.
    Returns ``(spec[F, T], energy[T], phone_energy[P] | None)``."""
    spec = process_spec(audio, transform, hop)
    energy = extract_energy(spec)
    phone = average_data_by_durations(energy, durations) if durations is not None else None
    return spec, energy, phone


def features_ragged(audios, transform, hop: int, durations=None):
    """The reference's per-utterance loop (what the loky workers do one item at a time,
    preprocessor.py:1197-1209) over a list of 1-D tensors."""
    out = []
    for i, a in enumerate(audios):
        out.append(features_one(a, transform, hop, None if durations is None else durations[i]))
    return out


# --------------------------------------------------------------------------------------
# fp64 "truth": independent numpy evaluation of the published formulas
# --------------------------------------------------------------------------------------
def truth_power_spectrogram(x: np.ndarray, n_fft: int, win_length: int, hop: int) -> np.ndarray:
    """|STFT|^2 in float64: periodic Hann of ``win_length`` zero-padded (centred) to
    ``n_fft``, reflect padding of ``n_fft//2``, frames ``t = 0 .. L//hop`` centred at
    ``t*hop``.  Returns ``[n_fft//2+1, T]`` with torch.stft's frame count ``T = 1 + (L + 2*(n_fft//2) - n_fft)//hop``."""
    x = np.asarray(x, dtype=np.float64)
    L = x.shape[-1]
    n = np.arange(win_length, dtype=np.float64)
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)
    left = (n_fft - win_length) // 2
    win = np.zeros(n_fft)
    win[left : left + win_length] = w
    xp = np.pad(x, (n_fft // 2, n_fft // 2), mode="reflect")
    T1 = 1 + (L + 2 * (n_fft // 2) - n_fft) // hop  # == 1 + L // hop for even n_fft
    idx = np.arange(n_fft)[None, :] + hop * np.arange(T1)[:, None]
    frames = xp[idx] * win[None, :]
    X = np.fft.rfft(frames, axis=-1)
    return (X.real**2 + X.imag**2).T


def truth_features(x, spec_type, n_fft, win_length, hop, sample_rate, n_mels, f_min, f_max):
    """float64 log-spectrogram ``[F, L//hop]`` and energy ``[L//hop]`` for the non-complex types."""
    P = truth_power_spectrogram(x, n_fft, win_length, hop)
    if spec_type == "mel":
        fb = torchaudio_melscale_fbanks(
            n_fft // 2 + 1, float(f_min), float(f_max), n_mels, sample_rate
        ).double().numpy()
        S = fb.T @ P
    elif spec_type == "mel-librosa":
        S = librosa_mel(sample_rate, n_fft, n_mels, f_min, f_max).astype(np.float64) @ np.sqrt(P + 1e-9)
    elif spec_type == "linear":
        S = P
    else:
        raise ValueError(spec_type)
    T = x.shape[-1] // hop
    logS = np.log(np.maximum(S, 1e-5))[:, :T]
    return logS, np.sqrt((logS**2).sum(axis=0))


# ------------------------------------------------------------------------------------------------
# audio front-end: the numerics of Preprocessor.process_audio
# everyvoice/preprocessor/preprocessor.py:131-218.  Third-party arithmetic restated from
# torchaudio 2.7.1 (pinned in pyproject.toml:78-80): functional.resample (_get_sinc_resample_kernel,
# _apply_sinc_resample_kernel), functional.loudness, functional.{treble,highpass}_biquad, lfilter.
# Pinned against the live reference by oracle/make_golden_frontend.py -> tests/golden/frontend.npz.
# ------------------------------------------------------------------------------------------------
def sinc_resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """torchaudio _get_sinc_resample_kernel (sinc_interp_hann, dtype=None: float64 math, float32 result).
    Returns ``(kernel[new, taps] float32, width, orig, new)`` with the frequencies reduced by their gcd."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base_freq = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base_freq)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    t = np.arange(0, -new, -1, dtype=np.float64)[:, None] / new + idx
    t = t * base_freq
    t = np.clip(t, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    scale = base_freq / orig
    with np.errstate(invalid="ignore", divide="ignore"):
        kern = np.where(t == 0, 1.0, np.sin(t) / t)
    kern = kern * window * scale
    return kern.astype(np.float32), width, orig, new


def resample(x: np.ndarray, orig_freq: int, new_freq: int) -> np.ndarray:
    """torchaudio.functional.resample(x[L], orig_freq, new_freq) (preprocessor.py:196-198), float32
    accumulation through a strided sliding-window matmul."""
    if orig_freq == new_freq:
        return np.asarray(x, dtype=np.float32)
    kern, width, orig, new = sinc_resample_kernel(orig_freq, new_freq)
    x = np.asarray(x, dtype=np.float32)
    L = len(x)
    xp = np.concatenate([np.zeros(width, np.float32), x, np.zeros(width + orig, np.float32)])
    taps = kern.shape[1]
    n_blocks = (len(xp) - taps) // orig + 1
    win = np.lib.stride_tricks.sliding_window_view(xp, taps)[::orig][:n_blocks]  # [blocks, taps]
    y = (win @ kern.T).reshape(-1)  # block-major, phase-minor == conv1d(...).transpose(1, 2).reshape
    target = -((-new * L) // orig)  # ceil(new * L / orig)
    return np.ascontiguousarray(y[:target], dtype=np.float32)


def _biquad_f32(x: np.ndarray, b, a) -> np.ndarray:
    """``torchaudio.functional.biquad`` -> ``lfilter(clamp=True)`` in float32, operation for operation (torchaudio
    2.7 - 2.11 ``_lfilter``): the coefficients are divided by a0 FIRST; the feed-forward part is ``conv1d`` over
    ``[b2, b1, b0] / a0`` whose accumulation is ``fma(b0, x[t], fma(b1, x[t-1], b2 * x[t-2]))``; the recursion
    (``_lfilter_core_loop``, C++) subtracts ``a2 * y[t-2]`` and then ``a1 * y[t-1]`` with separate roundings, on the
    UNclamped outputs; the result is clamped to [-1, 1].  Bit-exact against this container's torchaudio
    (tests/test_frontend_oracle.py): the 38 Hz high-pass has a double pole at 1 - 2 pi 38 / sr, so any other
    rounding order moves its output by 1e-4 and the loudness by up to 1e-3 LKFS."""
    f = np.float32
    b0, b1, b2 = (f(v) for v in b)
    a0, a1, a2 = (f(v) for v in a)
    b0, b1, b2, a1, a2 = f(b0 / a0), f(b1 / a0), f(b2 / a0), f(a1 / a0), f(a2 / a0)
    xp = np.concatenate([np.zeros(2, f), x.astype(f)]).astype(np.float64)
    t2 = (f(b2) * xp[:-2].astype(f)).astype(f).astype(np.float64)            # b2 * x[t-2], rounded
    t1 = (np.float64(b1) * xp[1:-1] + t2).astype(f).astype(np.float64)       # fma(b1, x[t-1], .): exact product in float64
    ff = (np.float64(b0) * xp[2:] + t1).astype(f)                            # fma(b0, x[t], .)
    y = np.zeros(len(x) + 2, f)
    for t in range(len(x)):
        y[t + 2] = f(f(ff[t] - f(a2 * y[t])) - f(a1 * y[t + 1]))
    return np.clip(y[2:], -1.0, 1.0)


def k_weighting_coeffs(sample_rate: int):
    """Coefficients of torchaudio's treble_biquad(4 dB, 1500 Hz, Q=1/sqrt(2)) and highpass_biquad(38 Hz, Q=0.5),
    evaluated with the SAME float32 torch tensor operations torchaudio runs (torch.sin / cos / exp / sqrt differ from
    libm's by an ulp for some arguments -- at 48 kHz that alone moves the loudness by 2e-3 LKFS)."""
    t = lambda v: torch.as_tensor(v, dtype=torch.float32)  # noqa: E731

    def treble(gain, central_freq, Q):  # torchaudio.functional.treble_biquad
        central_freq, Q, gain = t(central_freq), t(Q), t(gain)
        w0 = 2 * math.pi * central_freq / sample_rate
        alpha = torch.sin(w0) / 2 / Q
        A = torch.exp(gain / 40 * math.log(10))
        temp1 = 2 * torch.sqrt(A) * alpha
        temp2 = (A - 1) * torch.cos(w0)
        temp3 = (A + 1) * torch.cos(w0)
        b0 = A * ((A + 1) + temp2 + temp1)
        b1 = -2 * A * ((A - 1) + temp3)
        b2 = A * ((A + 1) + temp2 - temp1)
        a0 = (A + 1) - temp2 + temp1
        a1 = 2 * ((A - 1) - temp3)
        a2 = (A + 1) - temp2 - temp1
        return tuple(np.float32(float(v)) for v in (b0, b1, b2)), tuple(np.float32(float(v)) for v in (a0, a1, a2))

    def highpass(cutoff_freq, Q):  # torchaudio.functional.highpass_biquad
        cutoff_freq, Q = t(cutoff_freq), t(Q)
        w0 = 2 * math.pi * cutoff_freq / sample_rate
        alpha = torch.sin(w0) / 2.0 / Q
        b0 = (1 + torch.cos(w0)) / 2
        b1 = -1 - torch.cos(w0)
        b2 = b0
        a0 = 1 + alpha
        a1 = -2 * torch.cos(w0)
        a2 = 1 - alpha
        return tuple(np.float32(float(v)) for v in (b0, b1, b2)), tuple(np.float32(float(v)) for v in (a0, a1, a2))

    return treble(4.0, 1500.0, 1 / math.sqrt(2)), highpass(38.0, 0.5)


def loudness(x: np.ndarray, sample_rate: int) -> float:
    """torchaudio.functional.loudness for a mono waveform (ITU-R BS.1770-4), preprocessor.py:177-179."""
    gate = int(round(0.4 * sample_rate))
    step = int(round(gate * (1 - 0.75)))
    (tb, ta), (hb, ha) = k_weighting_coeffs(sample_rate)
    z = _biquad_f32(_biquad_f32(np.asarray(x, np.float32), tb, ta), hb, ha)
    if len(z) < gate:
        return float("nan")
    sq = np.square(z).astype(np.float32)
    blocks = np.lib.stride_tricks.sliding_window_view(sq, gate)[::step]
    energy = blocks.mean(axis=-1, dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        lk = np.float32(-0.691) + np.float32(10) * np.log10(energy)
        g1 = lk > -70.0
        e1 = np.float32(energy[g1].sum(dtype=np.float32)) / np.float32(g1.sum())
        gamma_rel = np.float32(-0.691) + np.float32(10) * np.log10(e1) - np.float32(10)
        g2 = g1 & (lk > gamma_rel)
        e2 = np.float32(energy[g2].sum(dtype=np.float32)) / np.float32(g2.sum())
        return float(np.float32(-0.691) + np.float32(10) * np.log10(e2))


def pcm16(x: np.ndarray) -> np.ndarray:
    """float -> PCM_S 16 bit as torchaudio.save(encoding="PCM_S", bits_per_sample=16) stores it
    (helpers.py:31-44).  PARITY UNPINNED: torchaudio.save needs torchcodec, which this container lacks; the rule is
    restated from ffmpeg swresample's flt -> s16 conversion (lrintf(x * 32768), clipped)."""
    return np.clip(np.rint(np.asarray(x, np.float32) * np.float32(32768.0)), -32768, 32767).astype(np.int16)


def process_audio_tensor(audio: np.ndarray, sr: int, *, normalize=True, resample_rate=None, hop_size=None,
                         min_audio_length=0.4, max_audio_length=11.0):
    """Preprocessor.process_audio after load_audio (preprocessor.py:148-218) for a [C, L] float32 waveform.
    Returns ``(audio[L'], sr)`` or ``(None, reason)`` when a gate skips the file."""
    audio = np.asarray(audio, np.float32)
    if audio.ndim == 1:
        audio = audio[None]
    if audio.shape[0] > 2:
        return None, "multichannel_files"
    seconds = audio.shape[1] / sr
    if seconds > max_audio_length:
        return None, "audio_too_long"
    if seconds < min_audio_length:
        return None, "audio_too_short"
    if audio.shape[0] != 1:
        raise NotImplementedError("the restatement covers mono input")
    lk = loudness(audio[0], sr)
    if math.isnan(lk) or lk < -36:
        return None, "audio_empty"
    x = audio[0]
    if resample_rate is not None and resample_rate != sr:
        x = resample(x, sr, resample_rate)
        sr = resample_rate
    if normalize:
        x = (x / np.max(np.abs(x))).astype(np.float32)
        x = (x * np.float32(0.95)).astype(np.float32)
    if hop_size is None:
        raise ValueError("We must know the hop size for processing audio because EveryVoice enforces that the "
                         "number of samples is evenly divisible by the hop size")
    return x[: (len(x) // hop_size) * hop_size], sr


# ------------------------------------------------------------------------------------------------
# pitch post-processing: everything in Preprocessor.extract_pitch after pyworld's dio / stonemask
# (everyvoice/preprocessor/preprocessor.py:236-285).  Pinned against the live reference (with a stub pyworld that
# returns seeded tracks) by oracle/make_golden_pitch.py -> tests/golden/pitch.npz.
# ------------------------------------------------------------------------------------------------
def postprocess_pitch(pitch: np.ndarray) -> np.ndarray:
    """``pitch[pitch == 0] = nan``; NaNs filled by ``np.interp`` over the frame index (preprocessor.py:236-242);
    nothing voiced (``ValueError`` from np.interp on an empty array) -> zeros; ``torch.tensor(pitch).float()``."""
    x = np.array(pitch, dtype=np.float64, copy=True)
    x[x == 0] = np.nan
    nans = np.isnan(x)
    try:
        x[nans] = np.interp(nans.nonzero()[0], (~nans).nonzero()[0], x[~nans])
    except ValueError:
        x[np.isnan(x)] = 0
    return x.astype(np.float32)
