"""Drop-in for the hot-path half of ``everyvoice/utils/heavy.py`` on a B200.

Same names, same arguments, same shapes as the reference:

* ``get_spectral_transform(spec_type, n_fft, win_length, hop_length, sample_rate=None,
  n_mels=None, f_min=0, f_max=8000)``  (reference: utils/heavy.py:47-119) returns a
  callable ``y = t(x)``, ``x[..., L] -> y[..., F, 1 + L//hop]`` or ``None`` for an
  unsupported type (so ``Preprocessor.__init__`` raises its ``ConfigError``);
* ``dynamic_range_compression_torch(x, C=1, clip_val=1e-5)``  (utils/heavy.py:39-40).

Everything numeric runs in ``libevfeat.so`` (hand-written sm_100a kernels) through the C
ABI of ``include/evfeat.h``.  There is no CPU / PyTorch fallback: without the library or a
CUDA device these calls raise.  CPU tensors are accepted for drop-in compatibility with
the reference's call sites (they are copied to the GPU and the result is copied back).

The batched, ragged entry points (``SpectralTransform.make_batch`` / ``run`` /
``features_ragged``) are the throughput path used by ``Preprocessor`` and ``bench.py``:
one launch computes log-spectrogram + energy for a whole packed batch of utterances.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib, filterbanks
from .config import AudioSpecTypeEnum


def _require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "everyvoice_b200 needs a CUDA sm_100 (B200) device; it has no CPU or PyTorch fallback"
        )
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"everyvoice_b200 only runs on CUDA devices, not {device}")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def _stream_ptr(device: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t: torch.Tensor | None) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


class _Plan:
    """Owns one ``evf_plan`` (immutable after creation)."""

    def __init__(self, tf: "SpectralTransform", device: torch.device, apply_log: bool,
                 keep_last: bool, sample_format: int):
        lib = _lib.load()
        cfg = _lib.evf_config(
            spec_type=_lib.SPEC_TYPES[tf.spec_type],
            sample_rate=int(tf.sample_rate or 0),
            n_fft=tf.n_fft,
            win_length=tf.win_length,
            hop_length=tf.hop_length,
            n_mels=int(tf.n_mels or 0),
            apply_log=int(apply_log),
            keep_last_frame=int(keep_last),
            sample_format=sample_format,
            log_clip=1e-5,
            fft_path=_lib.FFT_GENERIC if tf.fft_path == "generic" else _lib.FFT_AUTO,
        )
        win = tf.window.contiguous()
        fb = tf.mel_fb.contiguous() if tf.mel_fb is not None else None
        handle = C.c_void_p()
        _lib.check(lib.evf_plan_create(C.byref(cfg), _ptr(win), _ptr(fb), device.index, C.byref(handle)))
        self._lib = lib
        self.handle = handle
        self.device = device
        self.apply_log = apply_log
        self.keep_last = keep_last
        self.sample_format = sample_format
        rf = C.c_int32()
        _lib.check(lib.evf_plan_row_floats(handle, C.byref(rf)))
        self.row_floats = rf.value

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                self._lib.evf_plan_destroy(h)
            except Exception:
                pass


class RaggedBatch:
    """Owns one ``evf_batch``: per-utterance frame counts / offsets and the tile work list
    for a packed ragged batch (utterance ``b`` = samples ``[offsets[b], offsets[b+1])``)."""

    def __init__(self, plan: _Plan, sample_offsets):
        off = np.ascontiguousarray(np.asarray(sample_offsets, dtype=np.int64))
        if off.ndim != 1 or off.size < 1:
            raise ValueError("sample_offsets must be a 1-D array of B+1 offsets")
        lib = plan._lib
        self._lib = lib
        self.plan = plan
        self.n_utts = int(off.size - 1)
        self.sample_offsets = off
        handle = C.c_void_p()
        _lib.check(lib.evf_batch_create(plan.handle, off.ctypes.data_as(C.c_void_p), self.n_utts, C.byref(handle)))
        self.handle = handle
        tf = C.c_int64()
        _lib.check(lib.evf_batch_total_frames(handle, C.byref(tf)))
        self.total_frames = int(tf.value)
        fo = np.zeros(self.n_utts + 1, dtype=np.int64)
        _lib.check(lib.evf_batch_frame_offsets(handle, fo.ctypes.data_as(C.c_void_p)))
        self.frame_offsets = fo
        dev = C.c_void_p()
        _lib.check(lib.evf_batch_frame_offsets_dev(handle, C.byref(dev)))
        self.frame_offsets_dev_ptr = dev.value  # int64[n_utts+1] on the device, owned by the batch

    @property
    def total_samples(self) -> int:
        return int(self.sample_offsets[-1] - self.sample_offsets[0])

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                self._lib.evf_batch_destroy(h)
            except Exception:
                pass


@dataclass
class RaggedFeatures:
    spec: torch.Tensor            # [total_frames, row_floats] time-major, packed
    energy: torch.Tensor | None   # [total_frames]
    frame_offsets: np.ndarray     # int64[B+1]
    n_rows: int                   # F (n_mels or n_fft//2+1)
    is_complex: bool = False

    def utterance(self, b: int) -> torch.Tensor:
        """``[F, T_b]`` view of utterance ``b`` -- the layout the reference returns."""
        s = self.spec[self.frame_offsets[b] : self.frame_offsets[b + 1]]
        if self.is_complex:
            s = torch.view_as_complex(s.view(s.shape[0], self.n_rows, 2))
        return s.transpose(0, 1)

    def utterance_energy(self, b: int) -> torch.Tensor:
        return self.energy[self.frame_offsets[b] : self.frame_offsets[b + 1]]


class _SpectralFn(torch.autograd.Function):
    """Differentiable linear-domain transform for training through the spectrogram (HiFiGAN's mel loss on generated
    audio, hfgl/model.py:581-590, 719-721).  Forward: the fused feature kernel; backward: ``evf_features_backward``
    (transposed mel projection, inverse FFT through the forward FFT code, windowed overlap-add with the reflect
    padding folded back).

    The output is the ``[B, F, T]`` VIEW of the kernel's time-major result; the gradient autograd hands back is
    read in whichever of the two layouts it arrives (``[B, F, T]`` contiguous -- bin-major per utterance -- or the
    view's own), so neither direction copies.  With ``fused_log`` the kernel's epilogue applies
    ``log(clamp(., 1e-5))`` (utils/heavy.py:39-40) and the backward multiplies by its derivative from the saved
    output: no separate log kernels, no linear-domain intermediate."""

    @staticmethod
    def forward(ctx, x2d: torch.Tensor, tf: "SpectralTransform", keep_last: bool, fused_log: bool = False):
        B, L = x2d.shape
        batch = tf.uniform_batch(B, L, x2d.device, apply_log=fused_log, keep_last=keep_last)
        spec, _ = tf.run(batch, x2d.reshape(-1), want_energy=False)
        T = tf.num_frames(L, keep_last)
        ctx.save_for_backward(x2d, spec) if fused_log else ctx.save_for_backward(x2d)
        ctx.batch, ctx.dims, ctx.fused_log = batch, (B, L, T), fused_log
        return spec.view(B, T, tf.n_rows).transpose(1, 2)

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        x2d = ctx.saved_tensors[0]
        log_spec = ctx.saved_tensors[1] if ctx.fused_log else None
        batch, (B, L, T) = ctx.batch, ctx.dims
        plan = batch.plan
        g = grad_out.to(torch.float32)
        if g.transpose(1, 2).is_contiguous():      # the layout of the forward output
            layout = _lib.GRAD_FRAME_MAJOR
        else:                                      # [B, F, T] contiguous: per utterance [F][T]
            g, layout = g.contiguous(), _lib.GRAD_BIN_MAJOR
        lib = plan._lib
        n = int(lib.evf_features_backward_scratch_floats(plan.handle, batch.handle))
        scratch = torch.empty(max(n, 1), dtype=torch.float32, device=x2d.device)
        gx = torch.empty_like(x2d)
        with torch.cuda.device(x2d.device):
            _lib.check(lib.evf_features_backward_ex(plan.handle, batch.handle, _ptr(x2d), _ptr(g), layout,
                                                    _ptr(log_spec), _ptr(scratch), _ptr(gx), _stream_ptr(x2d.device)))
        return gx, None, None, None


def _is_dense(x: torch.Tensor) -> bool:
    """True if ``x`` covers its storage without gaps or overlaps in SOME dimension order (a contiguous tensor, or a
    permuted view of one such as the ``[B, F, T]`` view of the time-major spectrogram): an elementwise kernel may then
    walk the storage in memory order and ``torch.empty_like`` reproduces the strides."""
    if x.is_contiguous():
        return True
    order = sorted(range(x.dim()), key=lambda d: (-x.stride(d), -x.shape[d]))
    return x.permute(order).is_contiguous()


def _as_dense_f32(x: torch.Tensor) -> torch.Tensor:
    x = x.to(torch.float32)
    return x if _is_dense(x) else x.contiguous()


class _LogCompressFn(torch.autograd.Function):
    """``log(clamp(x, min=clip) * C)`` with its gradient ``1 / x`` where the clamp passes (utils/heavy.py:39-40).
    Elementwise, so any dense layout is processed in memory order and kept (no copy of the transposed view the
    transform returns)."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, C: float, clip_val: float):
        out = torch.empty_like(x)   # preserve_format: the strides of a dense x
        lib = _lib.load()
        with torch.cuda.device(x.device):
            _lib.check(lib.evf_log_compress(_ptr(x), _ptr(out), x.numel(), float(C), float(clip_val), _stream_ptr(x.device)))
        ctx.save_for_backward(x)
        ctx.clip = float(clip_val)
        return out

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        (x,) = ctx.saved_tensors
        g = grad_out.to(torch.float32)
        if g.stride() != x.stride():   # the kernel pairs x and g element by element in memory order
            g = torch.empty_like(x).copy_(g)
        gi = torch.empty_like(x)
        lib = _lib.load()
        with torch.cuda.device(x.device):
            _lib.check(lib.evf_log_compress_backward(_ptr(x), _ptr(g), _ptr(gi), x.numel(), ctx.clip, _stream_ptr(x.device)))
        return gi, None, None


class SpectralTransform:
    """The object ``get_spectral_transform`` returns.  ``t(x)`` mirrors the torchaudio
    transform the reference builds (linear-domain output, all ``1 + L//hop`` frames);
    ``t.features(...)`` / ``t.run(...)`` are the fused log + energy paths."""

    def __init__(self, spec_type, n_fft, win_length, hop_length, sample_rate=None, n_mels=None,
                 f_min=0, f_max=8000, device=None, *, norm="slaney", mel_scale="htk", mel_fb=None, window=None,
                 fft_path="auto"):
        """The keyword-only arguments go beyond ``get_spectral_transform``'s signature (training-side mels, SURVEY
        8f N4): ``norm`` / ``mel_scale`` are ``T.MelSpectrogram``'s (StyleTTS2 uses its defaults, None / "htk");
        ``mel_fb`` is any ``[n_fft // 2 + 1, n_mels]`` basis (triangular banks take the sparse projection, anything
        else a dense one); ``window`` replaces the Hann window (``[win_length]`` or ``[n_fft]``); ``fft_path="generic"``
        forces the any-size FFT kernel (evf_fft_path)."""
        spec_type = getattr(spec_type, "value", spec_type)
        if spec_type not in _lib.SPEC_TYPES:
            raise ValueError(f"unsupported spec_type {spec_type!r}")
        if fft_path not in ("auto", "generic"):
            raise ValueError("fft_path must be 'auto' or 'generic'")
        self.spec_type = spec_type
        self.fft_path = fft_path
        self.n_fft, self.win_length, self.hop_length = int(n_fft), int(win_length), int(hop_length)
        self.sample_rate, self.n_mels, self.f_min, self.f_max = sample_rate, n_mels, f_min, f_max
        self.n_freqs = self.n_fft // 2 + 1
        if window is None:
            self.window = filterbanks.hann_window_padded(self.win_length, self.n_fft)
        else:
            w = torch.as_tensor(window, dtype=torch.float32).detach().cpu().reshape(-1)
            if w.numel() == self.win_length and self.win_length < self.n_fft:
                left = (self.n_fft - self.win_length) // 2
                w = torch.nn.functional.pad(w, (left, self.n_fft - self.win_length - left))
            if w.numel() != self.n_fft:
                raise ValueError("window must have win_length or n_fft entries")
            self.window = w.contiguous()
        if mel_fb is not None and spec_type in ("mel", "mel-librosa"):
            fb = torch.as_tensor(mel_fb, dtype=torch.float32).detach().cpu()
            if fb.dim() != 2 or fb.shape[0] != self.n_freqs:
                raise ValueError(f"mel_fb must be [n_fft // 2 + 1 = {self.n_freqs}, n_mels]")
            self.mel_fb = fb.contiguous()
            self.n_mels = n_mels = int(fb.shape[1])
        elif spec_type == "mel":  # heavy.py:57-68
            self.mel_fb = filterbanks.melscale_fbanks(
                self.n_freqs, float(f_min), float(f_max if f_max is not None else sample_rate // 2),
                int(n_mels), int(sample_rate), norm, mel_scale)
        elif spec_type == "mel-librosa":  # heavy.py:69-100
            self.mel_fb = filterbanks.librosa_mel_basis(int(sample_rate), self.n_fft, int(n_mels), f_min, f_max)
        else:
            self.mel_fb = None
        self.is_complex = spec_type == "raw"
        self.n_rows = int(n_mels) if self.mel_fb is not None else self.n_freqs
        self._device = None if device is None else _require_cuda(device)
        self._plans: dict = {}

    # -- nn.Module-ish surface the reference relies on -----------------------------------
    def to(self, device):
        self._device = _require_cuda(device)
        return self

    def __repr__(self):
        return (f"SpectralTransform(spec_type={self.spec_type!r}, n_fft={self.n_fft}, win_length={self.win_length}, "
                f"hop_length={self.hop_length}, sample_rate={self.sample_rate}, n_mels={self.n_mels}, "
                f"f_min={self.f_min}, f_max={self.f_max})")

    # -- plans / batches --------------------------------------------------------------------
    def _resolve_device(self, x: torch.Tensor | None = None) -> torch.device:
        if x is not None and x.is_cuda:
            return x.device
        return self._device if self._device is not None else _require_cuda(None)

    def plan(self, device=None, apply_log=True, keep_last=False, sample_format=_lib.SAMPLES_F32) -> _Plan:
        device = _require_cuda(device if device is not None else self._device)
        key = (device.index, bool(apply_log), bool(keep_last), int(sample_format))
        p = self._plans.get(key)
        if p is None:
            p = self._plans[key] = _Plan(self, device, bool(apply_log), bool(keep_last), int(sample_format))
        return p

    def num_frames(self, n_samples: int, keep_last=False) -> int:
        """``L // hop`` (what ``process_spec`` keeps, preprocessor.py:921) or what ``torch.stft(center=True)`` yields:
        ``1 + (L + 2 * (n_fft // 2) - n_fft) // hop`` (= ``L // hop + 1`` for even ``n_fft``)."""
        if keep_last:
            return 1 + (int(n_samples) + 2 * (self.n_fft // 2) - self.n_fft) // self.hop_length
        return int(n_samples) // self.hop_length

    def make_batch(self, sample_offsets, device=None, apply_log=True, keep_last=False,
                   sample_dtype=torch.float32) -> RaggedBatch:
        fmt = _lib.SAMPLES_S16 if sample_dtype == torch.int16 else _lib.SAMPLES_F32
        return RaggedBatch(self.plan(device, apply_log, keep_last, fmt), sample_offsets)

    def uniform_batch(self, B: int, L: int, device=None, apply_log=True, keep_last=False) -> RaggedBatch:
        """The batch of ``B`` equal-length utterances (training segments).  Cached: a training loop calls the
        transform with the same shape every step, and building a batch costs a host pass over the tile list, a
        device allocation and a copy."""
        device = _require_cuda(device if device is not None else self._device)
        key = (device.index, int(B), int(L), bool(apply_log), bool(keep_last))
        cache = self.__dict__.setdefault("_uniform_batches", {})
        batch = cache.get(key)
        if batch is None:
            if len(cache) >= 8:
                cache.pop(next(iter(cache)))
            batch = cache[key] = self.make_batch(np.arange(B + 1, dtype=np.int64) * int(L), device, apply_log, keep_last)
        return batch

    def run(self, batch: RaggedBatch, samples: torch.Tensor, spec_out: torch.Tensor | None = None,
            energy_out: torch.Tensor | None = None, want_energy=True):
        """One asynchronous launch on the current stream.  ``samples`` is the packed device
        buffer the batch's offsets index (float32, or int16 if the batch was made for it)."""
        plan = batch.plan
        want = torch.int16 if plan.sample_format == _lib.SAMPLES_S16 else torch.float32
        if not samples.is_cuda or samples.device != plan.device:
            raise ValueError(f"samples must live on {plan.device}")
        if samples.dtype != want or not samples.is_contiguous():
            raise ValueError(f"samples must be a contiguous {want} tensor")
        if samples.numel() < int(batch.sample_offsets[-1]):
            raise ValueError("samples buffer is shorter than the last sample offset")
        if spec_out is None:
            spec_out = torch.empty((batch.total_frames, plan.row_floats), dtype=torch.float32, device=plan.device)
        elif (spec_out.dtype != torch.float32 or not spec_out.is_contiguous()
              or spec_out.numel() < batch.total_frames * plan.row_floats or spec_out.device != plan.device):
            raise ValueError("spec_out must be a contiguous float32 device tensor of total_frames * row_floats")
        if self.is_complex:
            want_energy = False
        if want_energy and energy_out is None:
            energy_out = torch.empty((batch.total_frames,), dtype=torch.float32, device=plan.device)
        with torch.cuda.device(plan.device):
            _lib.check(plan._lib.evf_features_run(plan.handle, batch.handle, _ptr(samples), _ptr(spec_out),
                                                  _ptr(energy_out if want_energy else None),
                                                  _stream_ptr(plan.device)))
        return spec_out, (energy_out if want_energy else None)

    def features_ragged(self, samples: torch.Tensor, sample_offsets, apply_log=True, keep_last=False,
                        want_energy=True) -> RaggedFeatures:
        """Log-spectrogram (+ energy) of a packed ragged batch: for every utterance what
        ``process_spec`` (preprocessor.py:917-928) + ``extract_energy`` (:302-309) produce."""
        device = self._resolve_device(samples)
        if not samples.is_cuda:
            samples = samples.to(device, non_blocking=True)
        if samples.dtype not in (torch.float32, torch.int16):
            samples = samples.float()
        batch = self.make_batch(sample_offsets, device, apply_log, keep_last, samples.dtype)
        spec, energy = self.run(batch, samples.contiguous(), want_energy=want_energy)
        return RaggedFeatures(spec, energy, batch.frame_offsets, self.n_rows, self.is_complex)

    # -- dense [..., L] entry points ------------------------------------------------------------
    def features(self, x: torch.Tensor, normalize=True, keep_last=True) -> torch.Tensor:
        """``x[..., L] -> [..., F, T]`` with ``T = L//hop (+1)``; ``normalize`` applies the
        fused ``log(clamp(., 1e-5))``.  Result lives on ``x``'s device."""
        if x.dim() < 1:
            raise ValueError("expected a tensor with a trailing time dimension")
        src_device = x.device
        device = self._resolve_device(x)
        xd = x.to(device=device, dtype=torch.float32).contiguous()
        L = x.shape[-1]
        lead = tuple(x.shape[:-1])
        B = int(np.prod(lead)) if lead else 1
        if x.requires_grad and torch.is_grad_enabled():
            # training through the transform: custom autograd functions over the same kernels
            if self.is_complex:
                raise NotImplementedError("the complex ('raw') transform has no backward")
            # [B, F, T]; the log (normalize) is fused into the kernel's epilogue and into its backward
            out = _SpectralFn.apply(xd.view(B, L), self, bool(keep_last), bool(normalize))
            out = out.reshape(*lead, self.n_rows, out.shape[-1])
            return out if src_device == device else out.to(src_device)
        offsets = np.arange(B + 1, dtype=np.int64) * L
        batch = self.make_batch(offsets, device, apply_log=normalize and not self.is_complex, keep_last=keep_last)
        spec, _ = self.run(batch, xd.view(-1), want_energy=False)
        T = self.num_frames(L, keep_last)
        if self.is_complex:
            out = torch.view_as_complex(spec.view(*lead, T, self.n_freqs, 2))
        else:
            out = spec.view(*lead, T, self.n_rows)
        out = out.transpose(-1, -2)
        return out if src_device == device else out.to(src_device)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """The bare transform of the reference: linear-domain, ``1 + L//hop`` frames
        (``T.MelSpectrogram`` / ``T.Spectrogram`` / the mel-librosa closure)."""
        return self.features(x, normalize=False, keep_last=True)

    def training_mel(self, generated_wav: torch.Tensor, drop_first: bool = True) -> torch.Tensor:
        """HiFiGAN's training-time mel of the generated audio (hfgl/model.py:719-721, 812-814):
        ``dynamic_range_compression_torch(transform(wav).squeeze(1)[:, :, 1:])`` for ``wav[B, 1, L]`` -- the log is
        fused into the kernel, the first frame is dropped like the reference does (its target mels come from
        ``process_spec``'s ``[:, :L // hop]`` of a segment that starts one hop earlier).  Differentiable."""
        y = self.features(generated_wav, normalize=True, keep_last=True)
        if y.dim() >= 4 and y.shape[-3] == 1:
            y = y.squeeze(-3)
        return y[..., 1:] if drop_first else y


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
    """Reference: ``everyvoice/utils/heavy.py:47-119``.  ``"istft"`` (synthesis-side inverse
    transform) is outside this path and, like any unknown type, yields ``None``."""
    st = getattr(spec_type, "value", spec_type)
    if st in (AudioSpecTypeEnum.mel.value, AudioSpecTypeEnum.mel_librosa.value,
              AudioSpecTypeEnum.linear.value, AudioSpecTypeEnum.raw.value):
        return SpectralTransform(st, n_fft, win_length, hop_length, sample_rate, n_mels, f_min, f_max)
    return None


def dynamic_range_compression_torch(x: torch.Tensor, C=1, clip_val=1e-5) -> torch.Tensor:
    """Reference: ``everyvoice/utils/heavy.py:39-40`` -- ``log(clamp(x, min=clip_val) * C)``
    as a stand-alone operator (the fused kernels do this in their epilogue)."""
    device = x.device if x.is_cuda else _require_cuda(None)
    xd = _as_dense_f32(x.to(device=device))
    if x.requires_grad and torch.is_grad_enabled():
        out = _LogCompressFn.apply(xd, float(C), float(clip_val))
        return out if x.is_cuda else out.to(x.device)
    out = torch.empty_like(xd)
    lib = _lib.load()
    with torch.cuda.device(device):
        _lib.check(lib.evf_log_compress(_ptr(xd), _ptr(out), xd.numel(), float(C), float(clip_val), _stream_ptr(device)))
    return out if x.is_cuda else out.to(x.device)
