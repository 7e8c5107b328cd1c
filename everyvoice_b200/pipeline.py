"""Host-buffer front door of the hot path: a chunked, triple-stream pipeline.

What ``everyvoice preprocess`` does per file -- ``process_spec`` (preprocessor.py:870-929),
``process_energy`` (:632-651) and the statistics / normalisation pass (:378-490) -- for a
whole shard whose audio lives in (pinned) HOST memory and whose results are wanted back in
HOST memory (to be written to disk by the caller):

    copy-in stream :  H2D samples of chunk i+1 ........ overlaps
    compute stream :  fused features kernel + phone averaging of chunk i ........ overlaps
    copy-out stream:  D2H log-spectrogram + energy of chunk i-1

PCIe is full duplex, so the step costs max(H2D, D2H) instead of their sum, and the kernels
hide entirely behind the copies.  Phone-level values stay resident until the last chunk, are
reduced to the five-number summary, (all-gathered across ranks,) normalised in place and
copied out once.  int16 PCM input (what ``process_audio`` writes, preprocessor.py:196-218)
halves the H2D bytes; it is converted on load inside the kernel, bit-identically to float input.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .distributed import allgather_stats
from .heavy import RaggedBatch, SpectralTransform, _require_cuda


@dataclass
class _Chunk:
    u0: int           # first utterance
    u1: int           # one past the last
    s0: int           # first sample in the packed host buffer
    s1: int
    f0: int           # first frame in the packed output
    f1: int
    pad: int          # elements of lead-in in the device sample buffer (keeps the 16-byte alignment class)


class PipelineResources:
    """The reusable half of a pipeline: streams, events and the double-buffered device chunks.
    Grows on demand, never shrinks; shared by every ``CorpusPipeline`` planned on one device."""

    def __init__(self, device):
        self.device = _require_cuda(device)
        dev = self.device
        self.s_in = torch.cuda.Stream(dev)
        self.s_out = torch.cuda.Stream(dev)
        self.s_compute = torch.cuda.Stream(dev)
        self.ev_in = [torch.cuda.Event() for _ in range(2)]        # H2D of buffer b done
        self.ev_comp = [torch.cuda.Event() for _ in range(2)]      # kernels on buffer b done
        self.ev_out = [torch.cuda.Event() for _ in range(2)]       # D2H of buffer b done
        self._bufs: dict = {}

    def buffer(self, name: str, numel: int, dtype) -> torch.Tensor:
        t = self._bufs.get(name)
        if t is None or t.dtype != dtype or t.numel() < numel:
            t = self._bufs[name] = torch.empty(max(int(numel), 1), dtype=dtype, device=self.device)
        return t

    def streams(self, name: str, count: int) -> list:
        key = "streams:" + name
        have = self._bufs.setdefault(key, [])
        while len(have) < count:
            have.append(torch.cuda.Stream(self.device))
        return have[:count]

    def events(self, name: str, count: int) -> list:
        key = "events:" + name
        have = self._bufs.setdefault(key, [])
        while len(have) < count:
            have.append(torch.cuda.Event())
        return have[:count]

    def host_buffer(self, name: str, numel: int, dtype) -> torch.Tensor:
        """Pinned host staging (page-locking memory costs about a millisecond: never per batch)."""
        t = self._bufs.get("host:" + name)
        if t is None or t.dtype != dtype or t.numel() < numel:
            t = self._bufs["host:" + name] = torch.empty(max(int(numel), 1), dtype=dtype).pin_memory()
        return t


class CorpusPipeline:
    """The plan for one shard layout (``sample_offsets`` [+ durations]): chunk boundaries and the
    per-chunk batch descriptors.  Cheap to build (host index arithmetic + one small upload per
    chunk), so a new one is planned for every new batch of files; the device buffers, streams
    and events come from a shared ``PipelineResources``.  ``run`` may be called repeatedly."""

    def __init__(self, transform: SpectralTransform, sample_offsets, device=None, sample_dtype=torch.float32,
                 durations=None, phone_offsets=None, chunk_bytes: int = 16 << 20, want_energy: bool = True,
                 resources: PipelineResources | None = None, head_chunk_bytes: int | None = None):
        self.tf = transform
        self.device = _require_cuda(device if device is not None else transform._device)
        self.res = resources if resources is not None else PipelineResources(self.device)
        self.sample_dtype = sample_dtype
        self.lib = _lib.load()
        off = np.ascontiguousarray(np.asarray(sample_offsets, dtype=np.int64))
        self.sample_offsets = off
        self.n_utts = int(off.size - 1)
        hop = transform.hop_length
        frames = (np.diff(off) // hop).astype(np.int64)
        self.frame_offsets = np.concatenate([[0], np.cumsum(frames)]).astype(np.int64)
        self.total_frames = int(self.frame_offsets[-1])
        self.total_samples = int(off[-1] - off[0])
        esize = 2 if sample_dtype == torch.int16 else 4
        # ---- chunk plan: consecutive utterances, about chunk_bytes of samples each ---------------
        # Small chunks at the END only (... / 2, / 4, / 8): the last D2H waits for the last chunk's kernel, so the drain
        # shrinks with them.  The FIRST chunks are full-sized (head_chunk_bytes): the step is bound by the copy-in
        # engine, and the first two copies must outlast the host's batch planning, or the engine idles until the
        # host is back to enqueue chunk 2 (measured: 0.55 ms of a 5.7 ms step with a 2 / 4 / 8 MB ramp-up,
        # profiles/r02z_e2e_timeline.txt).
        # ONE batch descriptor for the shard (built lazily by run(), after the first copies are in flight); chunks
        # are utterance ranges of it (evf_features_run_range)
        self._batch = None
        self.chunks: list[_Chunk] = []
        full = max(1, chunk_bytes // esize)
        align = 16 // esize
        total = int(off[-1] - off[0])
        u = 0
        ramp = max(1, (head_chunk_bytes if head_chunk_bytes is not None else chunk_bytes) // esize)
        while u < self.n_utts:
            left = total - int(off[u] - off[0])
            limit = max(1, min(ramp, full, max(left // 2, full // 8)))
            ramp *= 2
            v = int(np.searchsorted(off, off[u] + limit, side="right")) - 1   # last v with off[v] - off[u] <= limit
            v = min(max(v, u + 1), self.n_utts)
            s0 = int(off[u] - off[0])
            self.chunks.append(_Chunk(u, v, s0, int(off[v] - off[0]), int(self.frame_offsets[u]),
                                      int(self.frame_offsets[v]), s0 % align))
            u = v
        self.row_floats = transform.plan(self.device, True, False,
                                         _lib.SAMPLES_S16 if sample_dtype == torch.int16 else _lib.SAMPLES_F32).row_floats
        max_s = max((c.s1 - c.s0 + c.pad for c in self.chunks), default=0)
        max_f = max((c.f1 - c.f0 for c in self.chunks), default=0)
        self.want_energy = want_energy and not transform.is_complex
        res = self.res
        tag = "s16" if sample_dtype == torch.int16 else "f32"
        self._d_samples = [res.buffer(f"samples{b}_{tag}", max_s, sample_dtype) for b in range(2)]
        self._d_spec = [res.buffer(f"spec{b}", max_f * self.row_floats, torch.float32) for b in range(2)]
        self._d_energy = [res.buffer(f"energy{b}", max_f, torch.float32) for b in range(2)]
        # ---- phone-level averaging (optional) ------------------------------------------------------
        self.has_phones = durations is not None
        self.n_phones = 0
        if self.has_phones:
            p_off = np.ascontiguousarray(np.asarray(phone_offsets, dtype=np.int64))
            if p_off.size != self.n_utts + 1:
                raise ValueError("phone_offsets must have one entry per utterance plus one")
            self.phone_offsets = p_off
            self.n_phones = int(p_off[-1])
            d = durations if torch.is_tensor(durations) else torch.from_numpy(np.ascontiguousarray(durations))
            if d.dtype != torch.int64 or d.numel() != self.n_phones:
                raise ValueError("durations must be an int64 tensor with phone_offsets[-1] entries")
            self._h_durations = d.contiguous()      # uploaded by run() on the copy-in stream (counted as H2D)
            self._h_phone_off = torch.from_numpy(p_off)
            self._d_durations = res.buffer("durations", self.n_phones, torch.int64)
            self._d_phone_off = res.buffer("phone_off", self.n_utts + 1, torch.int64)
            self._d_phone = res.buffer("phone", self.n_phones, torch.float32)
            self._d_stats = res.buffer("stats5", 5, torch.float64)
            self._d_gathered = None
        self.h2d_bytes = self.total_samples * esize + (self.n_phones * 8 + (self.n_utts + 1) * 8 if self.has_phones else 0)
        self.d2h_bytes = self.total_frames * self.row_floats * 4 + (self.total_frames * 4 if self.want_energy else 0) \
            + (self.n_phones * 4 if self.has_phones else 0)
        self.kernel_launches_per_run = len(self.chunks) * (2 if self.has_phones else 1) + (3 if self.has_phones else 0)
        self.esize = esize

    @property
    def batch(self) -> RaggedBatch:
        """The shard's batch descriptor (tile work list, frame offsets): host index arithmetic plus one upload."""
        if self._batch is None:
            off = self.sample_offsets
            self._batch = self.tf.make_batch(off - off[0], self.device, apply_log=True, keep_last=False,
                                             sample_dtype=self.sample_dtype)
            assert np.array_equal(self._batch.frame_offsets, self.frame_offsets)
        return self._batch

    # ------------------------------------------------------------------------------------------
    def run(self, host_samples: torch.Tensor, host_spec: torch.Tensor, host_energy: torch.Tensor | None = None,
            host_phone: torch.Tensor | None = None, normalize_phones: bool = True, group=None):
        """``host_samples``: packed 1-D (pinned) host tensor of the pipeline's sample dtype;
        ``host_spec [total_frames, F]``, ``host_energy [total_frames]``, ``host_phone [n_phones]``:
        (pinned) host outputs.  Returns the float64[5] device summary of the phone values (or None)."""
        lib, dev = self.lib, self.device
        if host_samples.dtype != self.sample_dtype or host_samples.numel() < int(self.sample_offsets[-1]):
            raise ValueError("host_samples does not match the pipeline's dtype / layout")
        if tuple(host_spec.shape) != (self.total_frames, self.row_floats) or host_spec.dtype != torch.float32:
            raise ValueError(f"host_spec must be float32 [{self.total_frames}, {self.row_floats}]")
        res = self.res
        esize = 2 if self.sample_dtype == torch.int16 else 4
        h0 = int(self.sample_offsets[0])
        s_in, s_compute, s_out = res.s_in, res.s_compute, res.s_out
        ev_in, ev_comp, ev_out = res.ev_in, res.ev_comp, res.ev_out
        cur = torch.cuda.current_stream(dev)
        for s in (s_in, s_compute, s_out):
            s.wait_stream(cur)
        with torch.cuda.device(dev):
            if self.has_phones:
                with torch.cuda.stream(s_in):
                    self._d_durations[: self.n_phones].copy_(self._h_durations, non_blocking=True)
                    self._d_phone_off[: self.n_utts + 1].copy_(self._h_phone_off, non_blocking=True)
                s_compute.wait_stream(s_in)
            def copy_in(i):
                c, b = self.chunks[i], i & 1
                # buffer b is free once the kernels of chunk i-2 have consumed it
                if i >= 2:
                    s_in.wait_event(ev_comp[b])
                with torch.cuda.stream(s_in):
                    self._d_samples[b][c.pad : c.pad + c.s1 - c.s0].copy_(host_samples[h0 + c.s0 : h0 + c.s1],
                                                                         non_blocking=True)
                    ev_in[b].record(s_in)

            # the first two copies fly while the host builds the batch descriptor (tile list + one upload)
            for i in range(min(2, len(self.chunks))):
                copy_in(i)
            batch = self.batch
            for i, c in enumerate(self.chunks):
                b = i & 1
                ns, nf = c.s1 - c.s0, c.f1 - c.f0
                # -- kernels (output buffer b is free once chunk i-2 has been copied out)
                s_compute.wait_event(ev_in[b])
                if i >= 2:
                    s_compute.wait_event(ev_out[b])
                st = C.c_void_p(s_compute.cuda_stream)
                # addresses sample 0 / frame 0 of the whole shard would have if the chunk sat in place
                samples_base = self._d_samples[b].data_ptr() + (c.pad - c.s0) * esize
                spec_base = self._d_spec[b].data_ptr() - c.f0 * self.row_floats * 4
                energy_base = self._d_energy[b].data_ptr() - c.f0 * 4
                need_energy = self.want_energy or self.has_phones
                _lib.check(lib.evf_features_run_range(batch.plan.handle, batch.handle, c.u0, c.u1,
                                                      C.c_void_p(samples_base), C.c_void_p(spec_base),
                                                      C.c_void_p(energy_base if need_energy else 0), st))
                if self.has_phones:
                    _lib.check(lib.evf_segment_mean(C.c_void_p(energy_base),
                                                    C.c_void_p(batch.frame_offsets_dev_ptr + 8 * c.u0),
                                                    C.c_void_p(self._d_durations.data_ptr()),
                                                    C.c_void_p(self._d_phone_off.data_ptr() + 8 * c.u0),
                                                    c.u1 - c.u0, C.c_void_p(self._d_phone.data_ptr()), st))
                ev_comp[b].record(s_compute)
                # -- the copy-in engine bounds the step: hand it chunk i+2 (into the buffer these kernels free) before
                #    anything else is enqueued
                if i + 2 < len(self.chunks):
                    copy_in(i + 2)
                # -- D2H
                s_out.wait_event(ev_comp[b])
                with torch.cuda.stream(s_out):
                    host_spec[c.f0 : c.f1].copy_(self._d_spec[b][: nf * self.row_floats].view(nf, self.row_floats), non_blocking=True)
                    if self.want_energy and host_energy is not None:
                        host_energy[c.f0 : c.f1].copy_(self._d_energy[b][:nf], non_blocking=True)
                    ev_out[b].record(s_out)
            stats = None
            if self.has_phones:
                st = C.c_void_p(s_compute.cuda_stream)
                with torch.cuda.stream(s_compute):
                    _lib.check(lib.evf_stats_partial(C.c_void_p(self._d_phone.data_ptr()), self.n_phones,
                                                     C.c_void_p(self._d_stats.data_ptr()), 0, st))
                    stats = self._d_stats
                    if normalize_phones:
                        parts = allgather_stats(self._d_stats, group, out=self._d_gathered)
                        if parts.is_cuda and parts.dim() == 2 and parts.shape[0] > 1:
                            self._d_gathered = parts       # NCCL: reuse the gather buffer next run
                        # a gloo group gathers through host memory; the kernel below takes DEVICE pointers
                        parts = parts.to(self.device).contiguous()
                        _lib.check(lib.evf_normalize_by_gathered_stats(
                            C.c_void_p(self._d_phone.data_ptr()), self.n_phones, C.c_void_p(parts.data_ptr()),
                            parts.shape[0], parts.shape[1], st))
                        stats = parts
                    if host_phone is not None:
                        host_phone.copy_(self._d_phone[: self.n_phones], non_blocking=True)
        for s in (s_in, s_compute, s_out):
            cur.wait_stream(s)
        return stats


@dataclass
class FlowResult:
    """What ``FlowPipeline.run`` leaves behind once the streams have drained (``sync()``)."""

    keep: np.ndarray        # bool[n]: the loudness gate's decision per candidate utterance
    loudness: np.ndarray    # float32[n] LKFS
    stats5: torch.Tensor    # float64 device summary of the kept phone values ({n, sum, sumsq, min, max}; [world, 5] gathered)


class FlowPipeline:
    """The whole numeric flow behind ``everyvoice preprocess`` for one batch of loaded wavs, host to host, as ONE
    chunked three-stream pipeline: ``process_audio`` (loudness gate, peak normalisation, truncation, PCM16;
    preprocessor.py:148-218) -> ``process_spec`` (:917-928) -> ``process_energy`` (:632-651) -> statistics and
    normalisation (:378-490).

        copy-in stream :  H2D raw samples of chunk i+1
        compute stream :  loudness (fast + exact pass) -> absmax -> finalize -> features -> phone averages -> gate mask
        copy-out stream:  D2H processed PCM16 audio, log-spectrogram, energy of chunk i-1

    The loudness gate is CONSUMED ON THE DEVICE: every candidate (an utterance that passed the host-side channel /
    length gates) is processed and laid out as if it were kept -- all offsets are host-known integers -- and a
    skipped one only has its phone values turned into NaN (left out of the statistics) and its flag cleared; the
    caller drops it after the batch (``FlowResult.keep``).  No host synchronisation inside the batch.  Same-rate
    input (``resample_rate == sr``); resampled input takes ``Preprocessor.process_audio_batch``."""

    def __init__(self, transform: SpectralTransform, raw_offsets, sr: int, hop_size: int, device=None,
                 sample_dtype=torch.int16, durations=None, phone_offsets=None, chunk_bytes: int = 16 << 20,
                 normalize: bool = True, resources: PipelineResources | None = None,
                 head_chunk_bytes: int | None = None):
        from .audio import LOUDNESS_GATE_LKFS, LOUDNESS_REFINE_BAND, k_weighting_coefficients

        self.tf = transform
        self.device = _require_cuda(device if device is not None else transform._device)
        self.res = resources if resources is not None else PipelineResources(self.device)
        self.lib = _lib.load()
        self.sr, self.hop, self.normalize = int(sr), int(hop_size), bool(normalize)
        if self.hop != transform.hop_length:
            raise ValueError("hop_size must be the transform's hop length")
        self.sample_dtype = sample_dtype
        self.gate, self.band = LOUDNESS_GATE_LKFS, LOUDNESS_REFINE_BAND
        self.kw = k_weighting_coefficients(self.sr)
        raw = np.ascontiguousarray(np.asarray(raw_offsets, dtype=np.int64))
        self.raw_offsets = raw
        self.n_utts = n = int(raw.size - 1)
        lens = np.diff(raw)
        klen = (lens // self.hop) * self.hop                       # preprocessor.py:216-218
        self.kept_offsets = np.concatenate([[0], np.cumsum(klen)]).astype(np.int64)
        self.frame_offsets = self.kept_offsets // self.hop
        self.total_frames = int(self.frame_offsets[-1])
        step = int(self.lib.evf_audio_loudness_step(self.sr))
        if step < 1:
            raise ValueError("unsupported sampling rate for the loudness measurement")
        per = 5 * (lens // step + 5)
        self.scratch_offsets = np.concatenate([[0], np.cumsum(per)]).astype(np.int64)
        esize = 2 if sample_dtype == torch.int16 else 4
        self.esize = esize
        # ---- chunk plan over the RAW layout (what is copied in): full-sized head, ramped-down tail like CorpusPipeline's
        self.chunks: list[_Chunk] = []
        full = max(1, chunk_bytes // esize)
        total = int(raw[-1] - raw[0])
        u, ramp = 0, max(1, (head_chunk_bytes if head_chunk_bytes is not None else chunk_bytes) // esize)
        while u < n:
            left = total - int(raw[u] - raw[0])
            limit = max(1, min(ramp, full, max(left // 2, full // 8)))
            ramp *= 2
            v = int(np.searchsorted(raw, raw[u] + limit, side="right")) - 1
            v = min(max(v, u + 1), n)
            self.chunks.append(_Chunk(u, v, int(raw[u] - raw[0]), int(raw[v] - raw[0]), int(self.frame_offsets[u]),
                                      int(self.frame_offsets[v]), int(self.kept_offsets[u]) % (16 // 2)))
            u = v
        self._batch = None
        self._chunk_max_len = [int(lens[c.u0 : c.u1].max()) for c in self.chunks]
        self.row_floats = transform.plan(self.device, True, False, _lib.SAMPLES_S16).row_floats
        res = self.res
        max_raw = max((c.s1 - c.s0 for c in self.chunks), default=0)
        max_kept = max((int(self.kept_offsets[c.u1] - self.kept_offsets[c.u0]) + c.pad for c in self.chunks), default=0)
        max_f = max((c.f1 - c.f0 for c in self.chunks), default=0)
        max_scr = max((int(self.scratch_offsets[c.u1] - self.scratch_offsets[c.u0]) for c in self.chunks), default=0)
        tag = "s16" if sample_dtype == torch.int16 else "f32"
        self.NB = NB = 4   # chunk slots in flight (the loudness of a chunk outlives its kernels)
        self._d_raw = [res.buffer(f"flow_raw{b}_{tag}", max_raw, sample_dtype) for b in range(NB)]
        self._d_pcm = [res.buffer(f"flow_pcm{b}", max_kept, torch.int16) for b in range(NB)]
        self._d_spec = [res.buffer(f"spec{b}", max_f * self.row_floats, torch.float32) for b in range(NB)]
        self._d_energy = [res.buffer(f"energy{b}", max_f, torch.float32) for b in range(NB)]
        self._d_scratch = [res.buffer(f"flow_scratch{b}", max_scr, torch.float32) for b in range(NB)]
        self._d_lkfs = res.buffer("flow_lkfs", n, torch.float32)
        self._d_flags = res.buffer("flow_flags", n, torch.int32)
        self._d_keep = res.buffer("flow_keep", n, torch.int32)
        self._d_absmax = res.buffer("flow_absmax", n, torch.float32)
        # one upload for the three offset tables (raw, kept, scratch)
        self._h_tables = res.host_buffer("flow_tables", 3 * (n + 1), torch.int64)[: 3 * (n + 1)]
        self._h_tables.copy_(torch.from_numpy(np.concatenate([raw - raw[0], self.kept_offsets, self.scratch_offsets])))
        self._d_tables = res.buffer("flow_tables", 3 * (n + 1), torch.int64)
        self.has_phones = durations is not None
        self.n_phones = 0
        if self.has_phones:
            p_off = np.ascontiguousarray(np.asarray(phone_offsets, dtype=np.int64))
            if p_off.size != n + 1:
                raise ValueError("phone_offsets must have one entry per utterance plus one")
            self.phone_offsets = p_off
            self.n_phones = int(p_off[-1])
            d = durations if torch.is_tensor(durations) else torch.from_numpy(np.ascontiguousarray(durations))
            if d.dtype != torch.int64 or d.numel() != self.n_phones:
                raise ValueError("durations must be an int64 tensor with phone_offsets[-1] entries")
            self._h_durations = d.contiguous()
            self._h_phone_off = torch.from_numpy(p_off)
            self._d_durations = res.buffer("durations", self.n_phones, torch.int64)
            self._d_phone_off = res.buffer("phone_off", n + 1, torch.int64)
            self._d_phone = res.buffer("phone", self.n_phones, torch.float32)
            self._d_stats = res.buffer("stats5", 5, torch.float64)
        self._h_lkfs = res.host_buffer("flow_lkfs", n, torch.float32)[:n]
        self._h_keep = res.host_buffer("flow_keep", n, torch.int32)[:n]
        self.h2d_bytes = total * esize + 3 * (n + 1) * 8 + (self.n_phones * 8 + (n + 1) * 8 if self.has_phones else 0)
        self.d2h_bytes = self.total_frames * self.row_floats * 4 + self.total_frames * 4 + self.n_phones * 4 + 8 * n
        # per chunk: loudness 4 (fast, gate, exact, gate), absmax, finalize, features (+ phone averages); then mask + stats
        self.kernel_launches_per_run = len(self.chunks) * (7 + (1 if self.has_phones else 0)) + 1 + (3 if self.has_phones else 0)

    @property
    def batch(self) -> RaggedBatch:
        if self._batch is None:
            self._batch = self.tf.make_batch(self.kept_offsets, self.device, apply_log=True, keep_last=False,
                                             sample_dtype=torch.int16)
            assert np.array_equal(self._batch.frame_offsets, self.frame_offsets)
        return self._batch

    def run(self, host_samples: torch.Tensor, host_spec: torch.Tensor, host_energy: torch.Tensor | None = None,
            host_phone: torch.Tensor | None = None, host_audio: torch.Tensor | None = None, group=None):
        """``host_samples``: packed (pinned) raw samples, int16 PCM or float32; outputs (pinned): ``host_spec
        [total_frames, F]``, ``host_energy [total_frames]``, ``host_phone [n_phones]`` (normalised over the kept
        utterances), ``host_audio`` int16 ``[kept_offsets[-1]]`` (the processed wav samples, optional; its bytes are
        not in ``d2h_bytes``).  Returns a function that synchronises and yields the ``FlowResult``.

        Stream plan: nothing but the final mask depends on the loudness, and its kernel is latency bound (a 100 ms
        recursion per thread, ~0.5 ms per launch whatever the chunk size), so the loudness of chunk i runs on its own
        pair of streams under the copies and kernels of the following chunks; the gate is applied to the phone values
        once, after the last chunk."""
        lib, dev, n = self.lib, self.device, self.n_utts
        if host_samples.dtype != self.sample_dtype or host_samples.numel() < int(self.raw_offsets[-1]):
            raise ValueError("host_samples does not match the pipeline's dtype / layout")
        if tuple(host_spec.shape) != (self.total_frames, self.row_floats) or host_spec.dtype != torch.float32:
            raise ValueError(f"host_spec must be float32 [{self.total_frames}, {self.row_floats}]")
        res, esize, NB = self.res, self.esize, self.NB
        fmt = _lib.SAMPLES_S16 if self.sample_dtype == torch.int16 else _lib.SAMPLES_F32
        h0 = int(self.raw_offsets[0])
        s_in, s_compute, s_out = res.s_in, res.s_compute, res.s_out
        s_loud = res.streams("loud", 2)
        ev_in, ev_comp, ev_out, ev_loud = (res.events(k, NB) for k in ("in", "comp", "out", "loud"))
        cur = torch.cuda.current_stream(dev)
        for s in (s_in, s_compute, s_out, *s_loud):
            s.wait_stream(cur)
        P = C.c_void_p
        with torch.cuda.device(dev):
            with torch.cuda.stream(s_in):
                self._d_tables[: 3 * (n + 1)].copy_(self._h_tables, non_blocking=True)
                if self.has_phones:
                    self._d_durations[: self.n_phones].copy_(self._h_durations, non_blocking=True)
                    self._d_phone_off[: n + 1].copy_(self._h_phone_off, non_blocking=True)
            for s in (s_compute, *s_loud):
                s.wait_stream(s_in)
            t_raw = self._d_tables.data_ptr()
            t_kept = t_raw + 8 * (n + 1)
            t_scr = t_kept + 8 * (n + 1)

            def copy_in(i):
                c, b = self.chunks[i], i % NB
                if i >= NB:  # slot b is free once the kernels AND the loudness of chunk i - NB have consumed it
                    s_in.wait_event(ev_comp[b])
                    s_in.wait_event(ev_loud[b])
                with torch.cuda.stream(s_in):
                    self._d_raw[b][: c.s1 - c.s0].copy_(host_samples[h0 + c.s0 : h0 + c.s1], non_blocking=True)
                    ev_in[b].record(s_in)

            for i in range(min(NB, len(self.chunks))):
                copy_in(i)
            batch = self.batch
            st = P(s_compute.cuda_stream)
            for i, c in enumerate(self.chunks):
                b = i % NB
                nu, nf = c.u1 - c.u0, c.f1 - c.f0
                k0, k1 = int(self.kept_offsets[c.u0]), int(self.kept_offsets[c.u1])
                max_len = self._chunk_max_len[i]
                # bases such that (base + table offset) lands inside this chunk's buffers
                raw_base = self._d_raw[b].data_ptr() - c.s0 * esize
                scr_base = self._d_scratch[b].data_ptr() - int(self.scratch_offsets[c.u0]) * 4
                pcm_base = self._d_pcm[b].data_ptr() + (c.pad - k0) * 2
                spec_base = self._d_spec[b].data_ptr() - c.f0 * self.row_floats * 4
                energy_base = self._d_energy[b].data_ptr() - c.f0 * 4
                # -- loudness of the chunk, on its own stream
                sl = s_loud[i & 1]
                sl.wait_event(ev_in[b])
                _lib.check(lib.evf_audio_loudness(P(raw_base), fmt, P(t_raw + 8 * c.u0), nu, max_len, self.sr,
                                                  self.kw.ctypes.data_as(P), float(self.band), float(self.gate),
                                                  P(scr_base), P(t_scr + 8 * c.u0), P(self._d_flags.data_ptr() + 4 * c.u0),
                                                  P(self._d_lkfs.data_ptr() + 4 * c.u0), P(sl.cuda_stream)))
                ev_loud[b].record(sl)
                # -- peak, normalise + truncate + PCM16, features, phone averages
                s_compute.wait_event(ev_in[b])
                if i >= NB:
                    s_compute.wait_event(ev_out[b])
                if self.normalize:
                    _lib.check(lib.evf_audio_absmax(P(raw_base), fmt, P(t_raw + 8 * c.u0), nu, max_len,
                                                    P(self._d_absmax.data_ptr() + 4 * c.u0), st))
                _lib.check(lib.evf_audio_finalize(P(raw_base), fmt, P(t_raw + 8 * c.u0), P(t_kept + 8 * c.u0), nu,
                                                  (max_len // self.hop) * self.hop,
                                                  P(self._d_absmax.data_ptr() + 4 * c.u0) if self.normalize else P(0),
                                                  P(0), P(pcm_base), st))
                _lib.check(lib.evf_features_run_range(batch.plan.handle, batch.handle, c.u0, c.u1, P(pcm_base),
                                                      P(spec_base), P(energy_base), st))
                if self.has_phones:
                    _lib.check(lib.evf_segment_mean(P(energy_base), P(batch.frame_offsets_dev_ptr + 8 * c.u0),
                                                    P(self._d_durations.data_ptr()), P(self._d_phone_off.data_ptr() + 8 * c.u0),
                                                    nu, P(self._d_phone.data_ptr()), st))
                ev_comp[b].record(s_compute)
                # -- D2H
                s_out.wait_event(ev_comp[b])
                with torch.cuda.stream(s_out):
                    host_spec[c.f0 : c.f1].copy_(self._d_spec[b][: nf * self.row_floats].view(nf, self.row_floats), non_blocking=True)
                    if host_energy is not None:
                        host_energy[c.f0 : c.f1].copy_(self._d_energy[b][:nf], non_blocking=True)
                    if host_audio is not None:
                        host_audio[k0:k1].copy_(self._d_pcm[b][c.pad : c.pad + k1 - k0], non_blocking=True)
                    ev_out[b].record(s_out)
                if i + NB < len(self.chunks):
                    copy_in(i + NB)
            stats = None
            for sl in s_loud:
                s_compute.wait_stream(sl)  # every loudness is known
            with torch.cuda.stream(s_compute):
                # the gate, consumed on the device: skipped utterances' phone values -> NaN, keep flags
                _lib.check(lib.evf_audio_gate_mask(P(self._d_lkfs.data_ptr()), float(self.gate),
                                                   P(self._d_phone.data_ptr()) if self.has_phones else P(0),
                                                   P(self._d_phone_off.data_ptr()) if self.has_phones else P(0),
                                                   n, P(self._d_keep.data_ptr()), st))
                if self.has_phones:
                    _lib.check(lib.evf_stats_partial(P(self._d_phone.data_ptr()), self.n_phones, P(self._d_stats.data_ptr()), 0, st))
                    parts = allgather_stats(self._d_stats, group).to(dev).contiguous()
                    _lib.check(lib.evf_normalize_by_gathered_stats(P(self._d_phone.data_ptr()), self.n_phones,
                                                                   P(parts.data_ptr()), parts.shape[0], parts.shape[1], st))
                    stats = parts
                    if host_phone is not None:
                        host_phone.copy_(self._d_phone[: self.n_phones], non_blocking=True)
                self._h_lkfs.copy_(self._d_lkfs[:n], non_blocking=True)
                self._h_keep.copy_(self._d_keep[:n], non_blocking=True)
        for s in (s_in, s_compute, s_out, *s_loud):
            cur.wait_stream(s)

        def result() -> FlowResult:
            torch.cuda.current_stream(dev).synchronize()
            return FlowResult(self._h_keep.numpy().astype(bool), self._h_lkfs.numpy().copy(), stats)  # copies: the staging is reused

        return result
