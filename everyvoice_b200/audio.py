"""Audio front-end on the B200: the numerics of ``Preprocessor.process_audio``
(everyvoice/preprocessor/preprocessor.py:131-218) for a whole list of loaded waveforms at once.

What the reference does per file after ``load_audio`` -- channel / length gates, the BS.1770
loudness gate (``torchaudio.transforms.Loudness`` < -36 or NaN => skip, :177-186), sinc resampling
(``torchaudio.functional.resample``, :196-198), peak normalisation ``x / max|x| * 0.95``
(:199-201) and truncation to a multiple of the hop size (:216-218) -- runs here as four kernels
over a packed ragged batch (``evf_audio_loudness`` / ``_resample`` / ``_absmax`` / ``_finalize``
in ``include/evfeat.h``).  The result stays on the device, packed, as float32 or as the PCM16 the
reference would write with ``save_wav`` (helpers.py:31-44), i.e. exactly the two sample formats
the feature kernel stages.  sox effects and file decoding stay with the reference.
"""

from __future__ import annotations

import ctypes as C
from collections import Counter
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib
from .heavy import _ptr, _require_cuda, _stream_ptr

LOUDNESS_GATE_LKFS = -36.0  # preprocessor.py:180


@dataclass
class ProcessedAudio:
    """Packed result of ``process_audio_batch``.

    ``samples``: device tensor (float32 or int16), utterance ``i`` of ``kept`` owns
    ``samples[offsets[i]:offsets[i+1]]`` (length a multiple of ``hop_size``); ``kept``: indices into
    the input list; ``skipped``: input index -> the reference's counter name (``audio_too_long``,
    ``audio_too_short``, ``audio_empty``, ``multichannel_files``); ``loudness``: LKFS of every
    input that reached the loudness gate (NaN elsewhere)."""

    samples: torch.Tensor
    offsets: np.ndarray
    sr: int
    kept: list = field(default_factory=list)
    skipped: dict = field(default_factory=dict)
    loudness: np.ndarray | None = None

    def utterance(self, i: int) -> torch.Tensor:
        return self.samples[int(self.offsets[i]) : int(self.offsets[i + 1])]


class Resampler:
    """``torchaudio.functional.resample(x, orig_freq, new_freq)`` with torchaudio's defaults
    (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99) for packed ragged batches."""

    def __init__(self, orig_freq: int, new_freq: int, device=None, lowpass_filter_width: int = 6, rolloff: float = 0.99):
        self.device = _require_cuda(device)
        self.orig_freq, self.new_freq = int(orig_freq), int(new_freq)
        self._lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self._lib.evf_resampler_create(self.orig_freq, self.new_freq, int(lowpass_filter_width),
                                                  float(rolloff), self.device.index, C.byref(h)))
        self.handle = h

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                self._lib.evf_resampler_destroy(h)
            except Exception:
                pass

    def out_length(self, n: int) -> int:
        return int(self._lib.evf_resampler_out_length(self.handle, int(n)))

    def out_lengths(self, lens: np.ndarray) -> np.ndarray:
        """``ceil(new * n / orig)`` for an array of lengths (the integer form of evf_resampler_out_length)."""
        g = int(np.gcd(self.orig_freq, self.new_freq))
        o, n = self.orig_freq // g, self.new_freq // g
        return (np.asarray(lens, dtype=np.int64) * n + o - 1) // o

    def launch(self, samples, d_in_off, d_out_off, n_utts: int, max_out_len: int, out: torch.Tensor):
        """The raw asynchronous call with offsets already on the device."""
        fmt = _lib.SAMPLES_S16 if samples.dtype == torch.int16 else _lib.SAMPLES_F32
        with torch.cuda.device(self.device):
            _lib.check(self._lib.evf_audio_resample(self.handle, _ptr(samples), fmt, _ptr(d_in_off), _ptr(d_out_off),
                                                    int(n_utts), int(max_out_len), _ptr(out), _stream_ptr(self.device)))

    def __call__(self, samples: torch.Tensor, offsets) -> tuple[torch.Tensor, np.ndarray]:
        """``samples``: packed float32 / int16 device tensor; returns ``(packed float32, out_offsets)``."""
        offsets = np.ascontiguousarray(np.asarray(offsets, dtype=np.int64))
        lens = np.diff(offsets)
        out_lens = self.out_lengths(lens)
        out_off = np.concatenate([[0], np.cumsum(out_lens)]).astype(np.int64)
        out = torch.empty(int(out_off[-1]), dtype=torch.float32, device=self.device)
        if samples.dtype not in (torch.float32, torch.int16) or not samples.is_contiguous():
            raise ValueError("samples must be a contiguous float32 or int16 tensor")
        d_in, d_out = torch.from_numpy(offsets).to(self.device), torch.from_numpy(out_off).to(self.device)
        self.launch(samples, d_in, d_out, len(lens), int(out_lens.max()) if len(lens) else 0, out)
        return out, out_off


def _fmt(t: torch.Tensor) -> int:
    if t.dtype == torch.int16:
        return _lib.SAMPLES_S16
    if t.dtype == torch.float32:
        return _lib.SAMPLES_F32
    raise ValueError("samples must be float32 or int16 PCM")


_KW_CACHE: dict = {}


def k_weighting_coefficients(sample_rate: int) -> np.ndarray:
    """The ten float32 numbers of BS.1770's K-weighting as ``torchaudio.functional.loudness`` gets them:
    ``treble_biquad(4 dB, 1500 Hz, Q = 1/sqrt(2))`` and ``highpass_biquad(38 Hz, Q = 0.5)``, each as
    ``{b0, b1, b2, a1, a2} / a0`` (``_lfilter`` normalises first).  Evaluated with float32 torch tensor operations,
    like torchaudio does, by whatever torch build is installed: torch's sin / cos / exp differ from libm's by an ulp
    for some arguments, and the 38 Hz high-pass (double pole at 1 - 2 pi 38 / sr) turns that ulp into 1e-3 LKFS.
    Host-side constants (ten scalars), no signal processing."""
    sr = int(sample_rate)
    if sr not in _KW_CACHE:
        import math

        t = lambda v: torch.as_tensor(v, dtype=torch.float32)  # noqa: E731
        out = []
        # treble_biquad
        central_freq, Q, gain = t(1500.0), t(1 / math.sqrt(2)), t(4.0)
        w0 = 2 * math.pi * central_freq / sr
        alpha = torch.sin(w0) / 2 / Q
        A = torch.exp(gain / 40 * math.log(10))
        temp1 = 2 * torch.sqrt(A) * alpha
        temp2 = (A - 1) * torch.cos(w0)
        temp3 = (A + 1) * torch.cos(w0)
        b = torch.stack([A * ((A + 1) + temp2 + temp1), -2 * A * ((A - 1) + temp3), A * ((A + 1) + temp2 - temp1)])
        a = torch.stack([(A + 1) - temp2 + temp1, 2 * ((A - 1) - temp3), (A + 1) - temp2 - temp1])
        out += (b / a[0]).tolist() + (a / a[0])[1:].tolist()
        # highpass_biquad
        cutoff_freq, Q = t(38.0), t(0.5)
        w0 = 2 * math.pi * cutoff_freq / sr
        alpha = torch.sin(w0) / 2.0 / Q
        b0 = (1 + torch.cos(w0)) / 2
        b = torch.stack([b0, -1 - torch.cos(w0), b0])
        a = torch.stack([1 + alpha, -2 * torch.cos(w0), 1 - alpha])
        out += (b / a[0]).tolist() + (a / a[0])[1:].tolist()
        _KW_CACHE[sr] = np.asarray(out, dtype=np.float32)
    return _KW_CACHE[sr]


LOUDNESS_REFINE_BAND = 0.05  # LKFS: 25x the fast pass' error; inside it the exact pass decides


def loudness_batch(samples: torch.Tensor, offsets, sample_rate: int, refine_band: float = LOUDNESS_REFINE_BAND,
                   return_refined: bool = False):
    """``torchaudio.transforms.Loudness(sr)`` of every mono utterance of a packed float32 / int16 PCM device batch
    (int16 is ``s / 32768``, what ``torchaudio.load`` returns for a PCM16 wav).  Utterances within ``refine_band`` LKFS
    of the -36 LKFS gate (or with a block that close to a block-gating threshold) are re-evaluated with the
    reference's exact float32 arithmetic, so that the keep / skip decision is the reference's."""
    lib = _lib.load()
    device = samples.device
    offsets = np.ascontiguousarray(np.asarray(offsets, dtype=np.int64))
    lens = np.diff(offsets)
    step = int(lib.evf_audio_loudness_step(int(sample_rate)))   # the 100 ms step in samples, as the library rounds it
    if step < 1:
        raise ValueError("unsupported sampling rate for the loudness measurement")
    per = 5 * (lens // step + 5)          # == evf_audio_loudness_scratch_floats(sr, n) for every n
    s_off = np.concatenate([[0], np.cumsum(per)]).astype(np.int64)
    scratch = torch.empty(int(s_off[-1]), dtype=torch.float32, device=device)
    out = torch.empty(len(lens), dtype=torch.float32, device=device)
    flags = torch.zeros(max(len(lens), 1), dtype=torch.int32, device=device)
    coeffs = k_weighting_coefficients(sample_rate)
    d_off, d_soff = torch.from_numpy(offsets).to(device), torch.from_numpy(s_off).to(device)
    with torch.cuda.device(device):
        _lib.check(lib.evf_audio_loudness(_ptr(samples), _fmt(samples), _ptr(d_off), len(lens),
                                          int(lens.max()) if len(lens) else 0, int(sample_rate),
                                          coeffs.ctypes.data_as(C.c_void_p), float(refine_band), LOUDNESS_GATE_LKFS,
                                          _ptr(scratch), _ptr(d_soff), _ptr(flags), _ptr(out), _stream_ptr(device)))
    return (out, flags[: len(lens)]) if return_refined else out


def _consecutive_views(waves) -> bool:
    """True if the 1-D CPU tensors (one dtype) are contiguous views that follow one another in ONE storage."""
    if not waves or any(w.device.type != "cpu" or w.dtype != waves[0].dtype or not w.is_contiguous() for w in waves):
        return False
    base = waves[0].untyped_storage().data_ptr()
    nxt = waves[0].storage_offset()
    for w in waves:
        if w.untyped_storage().data_ptr() != base or w.storage_offset() != nxt:
            return False
        nxt += w.numel()
    return True


class AudioFrontEnd:
    """Batched ``process_audio``.  One instance per (device, audio config); resamplers are cached per rate pair."""

    def __init__(self, audio_config, device=None):
        self.audio_config = audio_config
        self.device = _require_cuda(device)
        self.counters: Counter = Counter()
        self.multichannel_files_list: list = []
        self._resamplers: dict = {}

    def resampler(self, orig: int, new: int) -> Resampler:
        key = (int(orig), int(new))
        if key not in self._resamplers:
            self._resamplers[key] = Resampler(orig, new, self.device)
        return self._resamplers[key]

    def process_audio_batch(self, audios, sr: int, normalize=True, resample_rate=None, hop_size=None,
                            out_dtype=torch.float32, update_counters=True, names=None) -> ProcessedAudio:
        """``audios``: list of ``[L]`` / ``[C, L]`` float32 tensors or arrays as ``load_audio`` returns them, all at
        sampling rate ``sr`` -- or, all of them, the int16 PCM samples of the wav files themselves (converted on the
        device as ``s / 32768``, bit-identical to loading them as float; half the host-to-device bytes).  Mirrors
        process_audio's order of gates and operations (preprocessor.py:148-218)."""
        if hop_size is None:
            raise ValueError(
                "We must know the hop size for processing audio because EveryVoice enforces that the number of "
                "samples is evenly divisible by the hop size"
            )
        if out_dtype not in (torch.float32, torch.int16):
            raise ValueError("out_dtype must be torch.float32 or torch.int16")
        lib = _lib.load()
        dev = self.device
        ac = self.audio_config
        skipped: dict = {}
        cand, waves = [], []
        for i, a in enumerate(audios):
            t = a if torch.is_tensor(a) else torch.from_numpy(np.asarray(a))
            if t.dim() == 1:
                t = t[None]
            if t.shape[0] > 2:  # :151-161
                skipped[i] = "multichannel_files"
                self.multichannel_files_list.append(str(names[i]) if names else str(i))
                continue
            seconds = t.shape[1] / sr
            if seconds > ac.max_audio_length:  # :163-169
                skipped[i] = "audio_too_long"
                continue
            if seconds < ac.min_audio_length:  # :170-176
                skipped[i] = "audio_too_short"
                continue
            if t.shape[0] != 1:
                raise NotImplementedError("process_audio_batch handles mono input; downmix stereo first "
                                          "(the reference does it through its sox effects)")
            cand.append(i)
            waves.append(t[0] if t.dtype == torch.int16 else t[0].to(torch.float32))
        loud = np.full(len(audios), np.nan, dtype=np.float32)
        if not cand:
            self._count(skipped, 0.0, 0, update_counters)
            return ProcessedAudio(torch.empty(0, dtype=out_dtype, device=dev), np.zeros(1, np.int64),
                                  int(resample_rate or sr), [], skipped, loud)
        lens = np.array([w.numel() for w in waves], dtype=np.int64)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        # one device buffer, one asynchronous copy per utterance (pinned host tensors overlap; no host-side concat)
        in_dtype = torch.int16 if all(w.dtype == torch.int16 for w in waves) else torch.float32
        if in_dtype == torch.float32:  # a mixed batch: PCM promoted to float is s / 32768 (torchaudio.load), not the integer
            waves = [w.to(torch.float32) * (1.0 / 32768.0) if w.dtype == torch.int16 else w for w in waves]
        x = torch.empty(int(off[-1]), dtype=in_dtype, device=dev)
        if _consecutive_views(waves):
            # the utterances already sit back to back in one host buffer (e.g. a memory-mapped corpus): one copy
            whole = torch.empty(0, dtype=in_dtype).set_(waves[0].untyped_storage(), waves[0].storage_offset(),
                                                         (int(off[-1]),))
            x.copy_(whole, non_blocking=True)
        else:
            for j, w in enumerate(waves):
                x[int(off[j]) : int(off[j + 1])].copy_(w, non_blocking=True)
        # ---- loudness gate (:177-186) -----------------------------------------------------------
        lk = loudness_batch(x, off, sr).cpu().numpy()
        loud[cand] = lk
        ok = ~(np.isnan(lk) | (lk < LOUDNESS_GATE_LKFS))
        for j, i in enumerate(cand):
            if not ok[j]:
                skipped[i] = "audio_empty"
        # ---- resample (:196-198) ----------------------------------------------------------------
        out_sr = sr
        if resample_rate is not None and resample_rate != sr:
            x, off = self.resampler(sr, resample_rate)(x, off)
            out_sr = int(resample_rate)
        full = np.diff(off)
        d_off = torch.from_numpy(off).to(dev)
        # ---- peak (:199-201), truncation (:216-218), output format ------------------------------
        absmax = None
        if normalize:
            absmax = torch.empty(len(cand), dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                _lib.check(lib.evf_audio_absmax(_ptr(x), _fmt(x), _ptr(d_off), len(cand), int(full.max()), _ptr(absmax),
                                                _stream_ptr(dev)))
        kept_len = np.where(ok, (full // int(hop_size)) * int(hop_size), 0).astype(np.int64)
        dst = np.concatenate([[0], np.cumsum(kept_len)]).astype(np.int64)
        d_dst = torch.from_numpy(dst).to(dev)
        out = torch.empty(int(dst[-1]), dtype=out_dtype, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.evf_audio_finalize(
                _ptr(x), _fmt(x), _ptr(d_off), _ptr(d_dst), len(cand), int(kept_len.max()), _ptr(absmax),
                _ptr(out) if out_dtype == torch.float32 else None, _ptr(out) if out_dtype == torch.int16 else None,
                _stream_ptr(dev)))
        kept = [i for j, i in enumerate(cand) if ok[j]]
        offsets = np.concatenate([[0], np.cumsum(kept_len[ok])]).astype(np.int64)
        self._count(skipped, float(sum(lens[j] for j in range(len(cand)) if ok[j])) / sr, len(kept), update_counters)
        return ProcessedAudio(out, offsets, out_sr, kept, skipped, loud)

    def _count(self, skipped, seconds, n_ok, update):
        if not update:
            return
        for reason in skipped.values():
            self.counters[reason] += 1
        self.counters["processed_files"] += n_ok
        self.counters["duration"] += seconds
