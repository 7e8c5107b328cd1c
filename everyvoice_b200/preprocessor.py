"""Drop-in for the numeric methods of ``everyvoice/preprocessor/preprocessor.py`` on a B200.

Mirrors, with the reference's names and argument meaning:

* ``Preprocessor.__init__``'s transform construction            (preprocessor.py:94-129)
* ``Preprocessor.extract_spectral_features(audio, transform, normalize=True)``  (:220-233)
* ``Preprocessor.extract_energy(spec[F, T]) -> [T]``                            (:302-309)
* ``Preprocessor.average_data_by_durations(data[T], durations[P]) -> [P]``      (:287-300)
* the in-memory cores of ``process_spec`` (:917-928) and ``process_energy`` (:641-650) as
  batched, ragged calls (one kernel launch for a whole list of utterances), and
  ``compute_stats`` / ``normalize_stats`` (:378-490) over in-memory shards.

* ``process_audio``'s numerics (:131-218: length / loudness gates, resampling, peak normalisation,
  truncation to a multiple of the hop size) as ``process_audio_batch`` / ``process_audio``
  (``everyvoice_b200.audio``).

File decoding beyond PCM wav, sox effects, text processing, config locks and the CLI are out of
scope (they stay with the reference; see INTEGRATION.md for where these calls slot in).
"""

from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .config import AudioConfig, ConfigError
from .heavy import (RaggedFeatures, SpectralTransform, _ptr, _require_cuda, _stream_ptr,
                    get_spectral_transform)
from .helpers import Scaler


def _to_sample_dtype(a: torch.Tensor, dt) -> torch.Tensor:
    """int16 PCM promoted to float32 is ``s / 32768`` (what torchaudio.load yields), never the raw integer value."""
    if a.dtype == dt:
        return a
    if a.dtype == torch.int16 and dt == torch.float32:
        return a.to(torch.float32) * (1.0 / 32768.0)
    return a.to(dt)


def _as_i64_dev(x, device) -> torch.Tensor:
    if torch.is_tensor(x):
        return x.to(device=device, dtype=torch.int64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.int64))).to(device)


class Preprocessor:
    def __init__(self, config=None, device=None):
        """``config``: an ``AudioConfig`` (ours or the reference's) or any object exposing
        ``.preprocessing.audio`` like the reference's FeaturePrediction/Vocoder configs."""
        if config is None:
            config = AudioConfig()
        self.config = config
        self.audio_config = getattr(getattr(config, "preprocessing", None), "audio", config)
        self.device = device
        self.pitch_scaler = Scaler(device)
        self.energy_scaler = Scaler(device)
        self.sep = "--"
        ac = self.audio_config
        self.input_sampling_rate = ac.input_sampling_rate
        self.output_sampling_rate = ac.output_sampling_rate
        self.sampling_rate_change = self.output_sampling_rate // self.input_sampling_rate
        self.output_hop_size = ac.fft_hop_size * self.sampling_rate_change
        spec_type = getattr(ac.spec_type, "value", ac.spec_type)
        # preprocessor.py:102-121 (note: the output transform's mel basis also uses the
        # *input* sampling rate, exactly as the reference does)
        self.input_spectral_transform = get_spectral_transform(
            spec_type, ac.n_fft, ac.fft_window_size, ac.fft_hop_size,
            sample_rate=self.input_sampling_rate, n_mels=ac.n_mels, f_min=ac.f_min, f_max=ac.f_max,
        )
        self.output_spectral_transform = get_spectral_transform(
            spec_type, ac.n_fft * self.sampling_rate_change, ac.fft_window_size * self.sampling_rate_change,
            self.output_hop_size,
            sample_rate=self.input_sampling_rate, n_mels=ac.n_mels, f_min=ac.f_min, f_max=ac.f_max,
        )
        if self.input_spectral_transform is None or self.output_spectral_transform is None:
            raise ConfigError(
                f"Spectral feature specification '{spec_type}' is not supported. Please edit your config file."
            )
        if device is not None:
            self.input_spectral_transform.to(device)
            self.output_spectral_transform.to(device)

    # ------------------------------------------------------------------------------------------
    # Audio front-end (preprocessor.py:131-218)
    # ------------------------------------------------------------------------------------------
    @property
    def audio_front_end(self):
        from .audio import AudioFrontEnd

        if getattr(self, "_audio_front_end", None) is None:
            self._audio_front_end = AudioFrontEnd(self.audio_config, _require_cuda(self.device))
            self.counters = self._audio_front_end.counters
            self.multichannel_files_list = self._audio_front_end.multichannel_files_list
        return self._audio_front_end

    def process_audio_batch(self, audios, sr, normalize=True, resample_rate=None, hop_size=None,
                            out_dtype=torch.float32, update_counters=True, names=None):
        """``process_audio`` for a list of loaded waveforms (what ``load_audio`` returns) at rate ``sr``: gates,
        loudness, resampling, peak normalisation and truncation on the device; see ``audio.ProcessedAudio``."""
        return self.audio_front_end.process_audio_batch(audios, sr, normalize, resample_rate, hop_size, out_dtype,
                                                        update_counters, names)

    def process_audio(self, wav_path, normalize=True, resample_rate=None, sox_effects=None, hop_size=None,
                      update_counters=True):
        """Reference: preprocessor.py:131-218, same arguments and return value ``(audio[L] float32, sr)`` or
        ``(None, None)`` when a gate skips the file.  Reads wav files (PCM 8 / 16 / 24 / 32 bit, IEEE float 32 / 64 bit:
        ``wavio.read_wav``, scaled like ``torchaudio.load``);
        ``sox_effects`` must be empty (sox stays with the reference)."""
        from .wavio import read_wav

        if sox_effects:
            raise NotImplementedError("sox effects are applied by the reference before this call")
        a, sr = read_wav(wav_path)   # [C, L]: int16 for 16-bit PCM (converted on the device), float32 otherwise
        audio = torch.from_numpy(a)
        res = self.process_audio_batch([audio], sr, normalize, resample_rate, hop_size, torch.float32,
                                       update_counters, [wav_path])
        if not res.kept:
            return None, None
        return res.utterance(0).cpu(), res.sr

    # ------------------------------------------------------------------------------------------
    # Pitch (preprocessor.py:236-285): DIO + StoneMask and everything after them run on the device.  pyworld is a
    # third-party dependency that is not available offline: the kernels restate WORLD's published algorithm and are
    # PARITY UNPINNED (include/evfeat.h, oracle/world_pitch.py).
    # ------------------------------------------------------------------------------------------
    def postprocess_pitch_batch(self, tracks):
        """``tracks``: list of float64 pitch tracks as pyworld returns them (0 = unvoiced).  Unvoiced frames are
        filled by linear interpolation over the frame index (``_interpolate`` / np.interp, :236-242, :278-283), a
        track without a voiced frame becomes zeros, float32 out.  Returns ``(packed device tensor, offsets)``."""
        device = _require_cuda(self.device)
        arrs = [np.ascontiguousarray(t.detach().cpu().numpy() if torch.is_tensor(t) else t, dtype=np.float64).reshape(-1)
                for t in tracks]
        lens = np.array([len(a) for a in arrs], dtype=np.int64)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        packed = torch.from_numpy(np.concatenate(arrs) if arrs else np.zeros(0)).to(device)
        return self._fill_unvoiced(packed, off), off

    def _fill_unvoiced(self, packed_f64: torch.Tensor, off: np.ndarray) -> torch.Tensor:
        device = packed_f64.device
        out = torch.empty(int(off[-1]), dtype=torch.float32, device=device)
        d_off = torch.from_numpy(np.ascontiguousarray(off, dtype=np.int64)).to(device)
        lib = _lib.load()
        with torch.cuda.device(device):
            _lib.check(lib.evf_pitch_fill_unvoiced(_ptr(packed_f64), _ptr(d_off), len(off) - 1, _ptr(out), _stream_ptr(device)))
        return out

    def track_pitch_batch(self, samples: torch.Tensor, sample_offsets, speed: int = 4, f0_floor: float = 71.0,
                          f0_ceil: float = 800.0, channels_in_octave: float = 2.0, allowed_range: float = 0.1):
        """``pw.dio(x, sr, frame_period=hop / sr * 1000, speed=4)`` then ``pw.stonemask(x, f0, t, sr)`` (:257-277) for a
        packed ragged batch (float32 or int16 PCM, on the device or not).  Returns ``(f0 packed float64 device tensor,
        frame offsets)``: WORLD's ``f0_length = int(1000 * L / sr / frame_period) + 1`` values per utterance, 0 where
        unvoiced."""
        import ctypes as C

        device = samples.device if samples.is_cuda else _require_cuda(self.device)
        x = samples.to(device)
        if x.dtype not in (torch.float32, torch.int16):
            x = x.to(torch.float32)
        x = x.contiguous()
        off = np.ascontiguousarray(np.asarray(sample_offsets, dtype=np.int64))
        n = len(off) - 1
        sr = int(self.input_sampling_rate)
        frame_period = self.audio_config.fft_hop_size / self.input_sampling_rate * 1000     # the reference's expression
        lib = _lib.load()
        flen = np.array([lib.evf_pitch_num_frames(sr, frame_period, int(L)) for L in np.diff(off)], dtype=np.int64)
        f_off = np.concatenate([[0], np.cumsum(flen)]).astype(np.int64)
        nbytes = int(lib.evf_pitch_scratch_bytes(off.ctypes.data_as(C.c_void_p), n, sr, frame_period, int(speed)))
        if nbytes < 0:
            raise ValueError("invalid arguments for the pitch tracker")
        scratch = torch.empty((nbytes + 7) // 8, dtype=torch.float64, device=device)
        f0 = torch.empty(int(f_off[-1]), dtype=torch.float64, device=device)
        fmt = _lib.SAMPLES_S16 if x.dtype == torch.int16 else _lib.SAMPLES_F32
        with torch.cuda.device(device):
            _lib.check(lib.evf_pitch_dio_stonemask(_ptr(x), fmt, off.ctypes.data_as(C.c_void_p), n, sr, frame_period,
                                                   int(speed), float(f0_floor), float(f0_ceil), float(channels_in_octave),
                                                   float(allowed_range), _ptr(scratch), nbytes, _ptr(f0), _stream_ptr(device)))
        return f0, f_off

    def extract_pitch_batch(self, samples: torch.Tensor, sample_offsets):
        """``extract_pitch`` for a packed ragged batch: tracker, then the unvoiced frames filled by interpolation.
        Returns ``(pitch packed float32 device tensor, frame offsets)``."""
        f0, f_off = self.track_pitch_batch(samples, sample_offsets)
        return self._fill_unvoiced(f0, f_off), f_off

    def extract_pitch(self, audio_tensor: torch.Tensor):
        """Reference: preprocessor.py:244-285, same argument and result (``[T]`` float32 CPU tensor): DIO (``speed=4``,
        ``frame_period = hop / sr * 1000``), StoneMask, zeros -> interpolated -- all on the device."""
        x = audio_tensor.squeeze(0).detach().reshape(-1)
        out, _ = self.extract_pitch_batch(x, np.array([0, x.numel()], dtype=np.int64))
        return out.cpu()

    # ------------------------------------------------------------------------------------------
    # The reference's per-utterance operators
    # ------------------------------------------------------------------------------------------
    def extract_spectral_features(self, audio_tensor: torch.Tensor, transform, normalize=True):
        """Reference: preprocessor.py:220-233.  ``transform`` must be one returned by
        ``everyvoice_b200.get_spectral_transform`` (window, FFT, mel and the log run as one
        fused kernel); there is no fallback for foreign callables."""
        if not isinstance(transform, SpectralTransform):
            raise TypeError(
                "transform must come from everyvoice_b200.get_spectral_transform; "
                "everyvoice_b200 has no CPU / torchaudio fallback"
            )
        return transform.features(audio_tensor, normalize=normalize, keep_last=True)

    def extract_energy(self, spectral_feature_tensor: torch.Tensor):
        """Reference: preprocessor.py:302-309 -- ``torch.linalg.norm(spec, dim=0)`` of a
        ``[F, T]`` (log-)spectrogram."""
        spec = spectral_feature_tensor
        if spec.dim() != 2:
            raise ValueError("extract_energy expects a [F, T] spectrogram")
        device = spec.device if spec.is_cuda else _require_cuda(self.device)
        tm = spec.to(device=device, dtype=torch.float32).transpose(0, 1).contiguous()  # [T, F]; free for our views
        T, F = tm.shape
        out = torch.empty(T, dtype=torch.float32, device=device)
        lib = _lib.load()
        with torch.cuda.device(device):
            _lib.check(lib.evf_energy_from_spec(_ptr(tm), T, F, _ptr(out), _stream_ptr(device)))
        return out if spec.is_cuda else out.to(spec.device)

    def average_data_by_durations(self, data: torch.Tensor, durations: torch.Tensor):
        """Reference: preprocessor.py:287-300 (one utterance)."""
        device = data.device if data.is_cuda else _require_cuda(self.device)
        vals = data.to(device=device, dtype=torch.float32).contiguous().view(-1)
        durs = _as_i64_dev(durations, device).view(-1)
        out = self.average_data_by_durations_ragged(
            vals, np.array([0, vals.numel()], dtype=np.int64), durs, np.array([0, durs.numel()], dtype=np.int64)
        )
        return out if data.is_cuda else out.to(data.device)

    # ------------------------------------------------------------------------------------------
    # Batched / ragged drivers (one launch per batch instead of one Python call per file)
    # ------------------------------------------------------------------------------------------
    def average_data_by_durations_ragged(self, values: torch.Tensor, value_offsets, durations, phone_offsets):
        """``average_data_by_durations`` for a packed batch: utterance ``b`` owns
        ``values[value_offsets[b]:value_offsets[b+1]]`` and
        ``durations[phone_offsets[b]:phone_offsets[b+1]]``.  Returns packed ``[sum P_b]`` float32."""
        device = values.device if values.is_cuda else _require_cuda(self.device)
        vals = values.to(device=device, dtype=torch.float32).contiguous()
        v_off = _as_i64_dev(value_offsets, device)
        durs = _as_i64_dev(durations, device)
        p_off = _as_i64_dev(phone_offsets, device)
        n_utts = v_off.numel() - 1
        if p_off.numel() != n_utts + 1:
            raise ValueError("value_offsets and phone_offsets must describe the same number of utterances")
        if n_utts > 0 and not (torch.is_tensor(value_offsets) and value_offsets.is_cuda):
            # host-side offsets: check them against the buffers before a kernel indexes with them
            vo, po = np.asarray(value_offsets, dtype=np.int64), np.asarray(
                phone_offsets.cpu() if torch.is_tensor(phone_offsets) else phone_offsets, dtype=np.int64)
            if (np.diff(vo) < 0).any() or vo[0] < 0 or vo[-1] > vals.numel():
                raise ValueError("value_offsets must be non-decreasing and end inside `values`")
            if (np.diff(po) < 0).any() or po[0] < 0 or po[-1] > durs.numel():
                raise ValueError("phone_offsets must be non-decreasing and end inside `durations`")
        out = torch.empty(durs.numel(), dtype=torch.float32, device=device)
        lib = _lib.load()
        with torch.cuda.device(device):
            _lib.check(lib.evf_segment_mean(_ptr(vals), _ptr(v_off), _ptr(durs), _ptr(p_off), n_utts, _ptr(out),
                                            _stream_ptr(device)))
        return out

    def process_spec_batch(self, audios, sample_offsets=None, output=False, want_energy=True) -> RaggedFeatures:
        """The in-memory core of ``process_spec`` (preprocessor.py:917-928) for many items at
        once: ``extract_spectral_features(audio, transform)[:, :L // hop]`` for each utterance
        (+ ``extract_energy`` of it, fused).  ``audios`` is a list of 1-D tensors, or a packed
        1-D tensor (float32 / int16 PCM) together with ``sample_offsets``."""
        transform = self.output_spectral_transform if output else self.input_spectral_transform
        device = _require_cuda(self.device)
        if sample_offsets is None:
            lens = np.array([a.numel() for a in audios], dtype=np.int64)
            sample_offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
            dt = torch.int16 if all(a.dtype == torch.int16 for a in audios) else torch.float32
            packed = torch.cat([_to_sample_dtype(a.reshape(-1), dt) for a in audios]) if len(audios) \
                else torch.zeros(0, dtype=dt)
        else:
            packed = audios
        if not packed.is_cuda:
            packed = packed.pin_memory().to(device, non_blocking=True) if packed.numel() else packed.to(device)
        return transform.features_ragged(packed, sample_offsets, apply_log=True, keep_last=False,
                                         want_energy=want_energy)

    def make_corpus_pipeline(self, sample_offsets, sample_dtype=torch.float32, durations=None, phone_offsets=None,
                             output=False, chunk_bytes: int = 16 << 20, head_chunk_bytes: int | None = None):
        """Host-buffer front door for a whole shard (``pipeline.CorpusPipeline``): H2D of the
        next chunk, the kernels of this one and D2H of the previous one overlap on three streams."""
        from .pipeline import CorpusPipeline, PipelineResources

        transform = self.output_spectral_transform if output else self.input_spectral_transform
        device = _require_cuda(self.device)
        res = getattr(self, "_pipeline_resources", None)
        if res is None or res.device != device:
            res = self._pipeline_resources = PipelineResources(device)   # buffers / streams live across batches
        return CorpusPipeline(transform, sample_offsets, device, sample_dtype, durations, phone_offsets, chunk_bytes,
                              resources=res, head_chunk_bytes=head_chunk_bytes)

    def make_flow_pipeline(self, raw_offsets, sr: int, sample_dtype=torch.int16, durations=None, phone_offsets=None,
                           normalize=True, chunk_bytes: int = 16 << 20, head_chunk_bytes: int | None = None):
        """The whole ``process_audio -> process_spec -> process_energy -> statistics`` flow for one batch of loaded
        wavs as one chunked pipeline with the loudness gate consumed on the device (``pipeline.FlowPipeline``)."""
        from .pipeline import FlowPipeline, PipelineResources

        device = _require_cuda(self.device)
        res = getattr(self, "_pipeline_resources", None)
        if res is None or res.device != device:
            res = self._pipeline_resources = PipelineResources(device)
        return FlowPipeline(self.input_spectral_transform, raw_offsets, sr, self.audio_config.fft_hop_size, device,
                            sample_dtype, durations, phone_offsets, chunk_bytes, normalize, resources=res,
                            head_chunk_bytes=head_chunk_bytes)

    def process_energy_batch(self, feats: RaggedFeatures, durations=None, phone_offsets=None):
        """The in-memory core of ``process_energy`` (preprocessor.py:641-650): frame energy, and
        phone-level averages when ``durations`` (packed int64 + ``phone_offsets``, or a list of
        per-utterance tensors) is given.  Returns ``(values_packed, offsets)``."""
        energy = feats.energy
        if durations is None:
            return energy, feats.frame_offsets
        if phone_offsets is None:
            lens = np.array([int(d.numel()) for d in durations], dtype=np.int64)
            phone_offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
            durations = torch.cat([d.reshape(-1).to(torch.int64) for d in durations])
        out = self.average_data_by_durations_ragged(energy, feats.frame_offsets, durations, phone_offsets)
        return out, np.asarray(phone_offsets, dtype=np.int64)

    # ------------------------------------------------------------------------------------------
    # Corpus statistics (preprocessor.py:378-490 driven by fs2/cli/preprocess.py:44-77)
    # ------------------------------------------------------------------------------------------
    def compute_stats(self, energy=None, pitch=None, n_energy_files=None, n_pitch_files=None):
        """``energy`` / ``pitch``: a list of per-utterance tensors, or one packed device tensor
        holding this rank's shard (then ``n_*_files`` gives the number of utterances, the
        reference's ``sample_size``).  Returns ``(energy_scaler, pitch_scaler)``."""
        scalers = []
        for data, n_files in ((energy, n_energy_files), (pitch, n_pitch_files)):
            if data is None:
                scalers.append(None)
                continue
            s = Scaler(self.device)
            if torch.is_tensor(data):
                s.append(data)
                s._n_files = int(n_files) if n_files is not None else 1
            else:
                for t in data:
                    s.append(t)
            scalers.append(s)
        self.energy_scaler, self.pitch_scaler = scalers[0] or self.energy_scaler, scalers[1] or self.pitch_scaler
        return scalers[0], scalers[1]

    def normalize_stats(self, energy_scaler: Scaler | None, pitch_scaler: Scaler | None, group=None,
                        distributed=None) -> dict:
        """Reference: preprocessor.py:453-490 -- statistics, then ``(x - mean) / std`` over every
        stored value (here: in place over the scaler's device tensors, one launch each)."""
        stats = {}
        for name, scaler in (("energy", energy_scaler), ("pitch", pitch_scaler)):
            if scaler is None:   # an EMPTY scaler still joins the collectives (a rank whose shard was gated out)
                continue
            n_files = getattr(scaler, "_n_files", None)
            if n_files is not None:
                st = _packed_stats(scaler, n_files, group, distributed)
            else:
                st = scaler.calculate_stats(group=group, distributed=distributed)
            for i, t in enumerate(scaler.data):
                if t.is_cuda and t.dtype == torch.float32 and t.is_contiguous():
                    scaler.normalize_(t)
                else:
                    scaler.data[i] = scaler.normalize(t)
            stats[name] = st
        return stats


def _packed_stats(scaler: Scaler, n_files: int, group, distributed):
    """calculate_stats for a scaler that holds one packed shard: ``sample_size`` must count
    utterances (files), not tensors."""
    import torch.distributed as dist

    from .distributed import allreduce_stats, finalize_stats

    if distributed is None:
        distributed = dist.is_available() and dist.is_initialized()
    stats5, sample_size = scaler.partial_stats(), n_files
    if distributed:
        stats5, sample_size = allreduce_stats(stats5, sample_size, group)
    st = finalize_stats(stats5.cpu().tolist(), sample_size)
    for k in ("min", "max", "mean", "std", "norm_min", "norm_max"):
        setattr(scaler, k, torch.tensor(st[k], dtype=torch.float32, device=stats5.device))
    return st
