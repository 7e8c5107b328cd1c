#!/usr/bin/env python
"""Benchmark of the preprocessing feature-extraction hot path (BASELINE.json metric:
audio-seconds per second, log-mel + energy + phone-level averaging, at N B200s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
                    [--data white|speech|lj_tiled] [--no-extras] [--no-parity]

One "step" = one pass of the hot path over one batch of synthetic utterances per rank:
fused log-spectrogram + energy kernel -> phone-level averaging of energy by durations ->
{count,sum,sumsq,min,max} reduction -> (N>1: one NCCL all-gather of those five numbers) ->
in-place normalisation.  Ranks hold independent shards (weak scaling); `value` is the
whole-job audio-seconds processed per second, timed on the device, max over ranks.

After the timed loop the buffers that were timed (log-spectrogram, energy, normalised phone values) are compared
with the CPU oracle on EVERY utterance of rank 0's shard (`parity` key; tests/parity_pool.py, all host cores).

`e2e` is the same work through the public API from pinned HOST buffers holding the wav files' own int16 PCM
(float32 input is reported next to it), against the box's measured host<->device copy bound.

`extra` holds the device-timed lines of the other BASELINE configs (44.1 kHz / n_fft 2048, linear, the 100 h corpus,
and two more transform sizes: n_fft 512 in the warp kernel with two packed jobs per warp, 4096 as four phase-stream
transforms of 1024 points plus a combine).

`--impl reference` times the CPU implementation of the same path (the oracle port of the
reference, which on CPU is bit-identical to it; /root/reference itself cannot travel to
the GPU box) on all host cores, on the same workload and config.
"""

from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# name -> (spec_type, sample_rate, n_fft, win, hop, n_mels, f_min, f_max, n_utts, min_s, max_s)
WORKLOADS = {
    # BASELINE.json configs[1]: 1k synthetic 1-10 s utterances at 22.05 kHz, 80-mel log-mel + energy
    "mel80_22k_1k_ragged": ("mel", 22050, 1024, 1024, 256, 80, 0, 8000, 1000, 1.0, 10.0),
    # configs[2]: 44.1 kHz n_fft 2048 hop 512 128 mels
    "mel128_44k_1k_ragged": ("mel", 44100, 2048, 2048, 512, 128, 0, 8000, 1000, 1.0, 10.0),
    # configs[3]: linear spectrogram + energy + phone averaging
    "linear_22k_1k_ragged": ("linear", 22050, 1024, 1024, 256, 80, 0, 8000, 1000, 1.0, 10.0),
    # configs[4]: a 100 h corpus (65 455 utterances, seed 1238) sharded by utterance over the N ranks with the greedy
    # longest-first partition: STRONG scaling (the corpus is fixed, a rank holds 1/N of it)
    "mel80_22k_100h_corpus": ("mel", 22050, 1024, 1024, 256, 80, 0, 8000, 65455, 1.0, 10.0),
    # the rest of the config-field domain: 16 kHz corpora (n_fft 512: warp kernel) and 4096-point vocoder targets (phase streams)
    "mel80_16k_nfft512": ("mel", 16000, 512, 512, 128, 80, 0, 8000, 1000, 1.0, 10.0),
    "mel128_44k_nfft4096": ("mel", 44100, 4096, 4096, 1024, 128, 0, 8000, 1000, 1.0, 10.0),
    "mel80_48k_nfft3072": ("mel", 48000, 3072, 3072, 768, 80, 0, 8000, 1000, 1.0, 10.0),   # the sizes a 16 -> 48 kHz configuration gives its output transform
}
CORPUS_WORKLOADS = {"mel80_22k_100h_corpus": 1238}  # name -> seed of the global utterance list
DEFAULT_WORKLOAD = "mel80_22k_1k_ragged"
EXTRA_WORKLOADS = ["mel128_44k_1k_ragged", "linear_22k_1k_ragged", "mel80_22k_100h_corpus", "mel80_16k_nfft512",
                   "mel128_44k_nfft4096", "mel80_48k_nfft3072"]
METRIC = "audio-sec/sec (log-mel+energy+phone-avg)"
UNIT = "audio-s/s"
STEP_DESC = ("features(log-spec+energy) -> phone averaging -> stats -> all-gather of the 5-number summaries (N>1) "
             "-> normalise")


def algorithmic_bytes_per_frame(spec_type, hop, n_mels, n_fft, sample_bytes=4):
    """every input sample read once, every output written once."""
    n_out = n_mels if spec_type in ("mel", "mel-librosa") else n_fft // 2 + 1
    return hop * sample_bytes + n_out * 4 + 4


def make_lengths(w, seed):
    from everyvoice_b200 import synth

    spec_type, sr, n_fft, win, hop, n_mels, f_min, f_max, n_utts, min_s, max_s = w
    return synth.utterance_lengths(n_utts, sr, hop, seed, min_s, max_s)


def make_durations(lengths, hop, seed):
    from everyvoice_b200 import synth

    durs = [synth.synthetic_durations(int(L) // hop, seed=seed + i) for i, L in enumerate(lengths)]
    return synth.pack_ragged(durs)


def shard_lengths(w, wname, rank, world):
    """Utterance lengths of this rank's shard and the size of the whole corpus (strong-scaling workloads)."""
    if wname in CORPUS_WORKLOADS:
        from everyvoice_b200.distributed import shard_utterances

        all_lengths = make_lengths(w, CORPUS_WORKLOADS[wname])
        mine = np.asarray(shard_utterances(all_lengths, world)[rank], dtype=np.int64)
        return all_lengths[mine], len(all_lengths)
    return make_lengths(w, 1234 + rank), None


def workload_config(w, wname, world, lengths, n_corpus, data):
    """The `config` object of the JSON line -- ONE function for both arms, so that they describe the same job."""
    spec_type, sr, n_fft, win, hop, n_mels, *_ = w
    total_samples = int(np.sum(lengths))
    frames = int(np.sum(np.asarray(lengths) // hop))
    n_out = n_mels if spec_type in ("mel", "mel-librosa") else n_fft // 2 + 1
    return {
        "workload": wname, "spec_type": spec_type, "sample_rate": sr, "n_fft": n_fft, "win": win, "hop": hop,
        "n_mels": n_mels, "utterances_per_gpu": int(len(lengths)), "audio_s_per_gpu": total_samples / sr,
        "frames_per_gpu": frames, "sample_dtype": "f32", "data": data, "step": STEP_DESC,
        "l2_policy": f"inputs larger than L2 ({total_samples * 4 / 1e6:.0f} MB read + {frames * n_out * 4 / 1e6:.0f} MB written per step)",
        "parallelism": (f"one corpus of {n_corpus} utterances sharded x{world} (greedy longest-first)"
                        if n_corpus else f"utterance shards x{world}") + ", stats all-gather only",
    }


def source_sha() -> str:
    """Hash of the kernel sources: ties profiles/traffic.json (an ncu capture) to the build being timed."""
    h = hashlib.sha256()
    for f in sorted((ROOT / "everyvoice_b200" / "csrc").glob("*")) + [ROOT / "include" / "evfeat.h"]:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()[:16]


# ------------------------------------------------------------------------------------------------
# synthetic data (white, speech-like, the bundled LJ wavs tiled)
# ------------------------------------------------------------------------------------------------
def make_samples(kind, lengths, sr, seed, device):
    """Packed float32 samples on the device."""
    import torch

    offsets = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    total = int(offsets[-1])
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    if kind == "white":
        return (torch.rand(total, device=device, generator=gen) * 1.9 - 0.95).contiguous()
    if kind == "speech":
        # harmonic stacks (5-20 harmonics of f0 in [80, 300] Hz, 1/k decay) under a slow AM envelope plus -60 dB noise:
        # the device-side twin of everyvoice_b200.synth.speech_like (wide dynamic range across the spectrum)
        x = torch.empty(total, device=device)
        rng = np.random.default_rng(seed)
        for b, L in enumerate(lengths):
            L = int(L)
            t = torch.arange(L, device=device, dtype=torch.float64) / sr
            f0, n_h = rng.uniform(80.0, 300.0), int(rng.integers(5, 21))
            vib = 1.0 + 0.02 * torch.sin(2 * np.pi * rng.uniform(3.0, 7.0) * t)
            phase = 2 * np.pi * f0 * torch.cumsum(vib, 0) / sr
            y = torch.zeros(L, device=device, dtype=torch.float64)
            for k in range(1, n_h + 1):
                if k * f0 * 1.02 >= sr / 2:
                    break
                y += torch.sin(k * phase + rng.uniform(0, 2 * np.pi)) / k
            y *= 0.55 + 0.45 * torch.sin(2 * np.pi * rng.uniform(0.5, 3.0) * t + rng.uniform(0, 2 * np.pi))
            y /= y.abs().max().clamp(min=1e-12)
            x[int(offsets[b]):int(offsets[b + 1])] = (0.95 * y).float()
        x += 1e-3 * torch.randn(total, device=device, generator=gen)
        return x.clamp_(-1.0, 1.0).contiguous()
    if kind == "lj_tiled":
        gold = np.load(ROOT / "tests" / "golden" / "lj_config1.npz")
        wavs = [gold[f"LJ050-02{n}/pcm16"].astype(np.float32) / np.float32(32768.0) for n in (69, 70, 71, 72, 73)]
        lj = torch.from_numpy(np.concatenate(wavs)).to(device)
        idx = (torch.arange(total, device=device) + seed * 7919) % lj.numel()
        return lj[idx].contiguous()
    raise ValueError(kind)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region from a background thread
    (NVML through pynvml, ~every 5 ms; the timed region can be a few tens of ms long, far
    shorter than nvidia-smi's start-up)."""

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = None
        self._thread = None
        self.source = None

    def _resolve_nvml_index(self, pynvml):
        # CUDA_VISIBLE_DEVICES may remap ordinals; match by PCI bus id through torch
        try:
            import torch

            pr = torch.cuda.get_device_properties(self.gpu_index)
            return pynvml.nvmlDeviceGetHandleByPciBusId(
                f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0".encode())
        except Exception:
            return pynvml.nvmlDeviceGetHandleByIndex(self.gpu_index)

    def start(self):
        import threading

        try:
            import pynvml

            pynvml.nvmlInit()
            h = self._resolve_nvml_index(pynvml)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            names = {
                "hw_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(pynvml, "nvmlClocksEventReasonSwPowerCap", 0x4),
            }
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        except Exception:
            return
        self.source = "nvml"
        self._stop = threading.Event()

        def loop():
            while not self._stop.is_set():
                try:
                    self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                    mask = int(get_reasons(h))
                    for n, bit in names.items():
                        if mask & bit:
                            self.reasons.add(n)
                except Exception:
                    pass
                self._stop.wait(0.005)

        self._thread = threading.Thread(target=loop, daemon=True)
        self._thread.start()

    def stop(self) -> dict:
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
               "samples": len(self.samples), "source": self.source}
        if self.samples:
            out["sm_mhz"] = float(np.median(self.samples))
        return out


# ------------------------------------------------------------------------------------------------
# CPU side (oracle port of the reference): cpu_baseline leg and --impl reference
# ------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_init(w, lengths, seed):
    """Build the CPU sample (shared by fork with the workers): white noise like the GPU arm's default data."""
    import torch

    from everyvoice_b200 import synth
    from oracle import ev_oracle as O

    spec_type, sr, n_fft, win, hop, n_mels, f_min, f_max, *_ = w
    rng = np.random.default_rng(seed)
    _CPU["audio"] = [torch.from_numpy(rng.uniform(-0.95, 0.95, size=int(L)).astype(np.float32)) for L in lengths]
    _CPU["durs"] = [torch.from_numpy(synth.synthetic_durations(int(L) // hop, seed=seed + i)) for i, L in enumerate(lengths)]
    _CPU["tf"] = O.get_spectral_transform(spec_type, n_fft, win, hop, sr, n_mels, f_min, f_max)
    _CPU["hop"] = hop
    return float(sum(int(L) for L in lengths)) / sr


def _cpu_chunk(idx_range):
    """What one loky worker of the reference does for its items (process_spec + process_energy,
    preprocessor.py:917-928, 641-650), in memory, one intra-op thread."""
    import torch

    from oracle import ev_oracle as O

    torch.set_num_threads(1)
    acc = 0.0
    vals = []
    for i in range(*idx_range):
        spec, energy, phone = O.features_one(_CPU["audio"][i], _CPU["tf"], _CPU["hop"], _CPU["durs"][i])
        acc += float(spec[0, 0])
        vals.append(phone)
    return acc, vals


def _cpu_step(pool, n_items, cpus):
    from oracle import ev_oracle as O

    if pool is None:
        results = [_cpu_chunk((0, n_items))]
    else:
        bs = min(100, 1 + n_items // (cpus * 2))  # the reference's batch rule, preprocessor.py:1198
        chunks = [(a, min(a + bs, n_items)) for a in range(0, n_items, bs)]
        results = pool.map(_cpu_chunk, chunks)
    # stats + normalise (Scaler, helpers.py:86-106; normalize_stats preprocessor.py:453-490)
    s = O.Scaler()
    for _, vals in results:
        for v in vals:
            s.append(v)
    s.calculate_stats()
    return [s.normalize(v) for v in s.data]


def cpu_baseline_single_thread(w, seed, n_sample=250):
    """Oracle port, 1 thread, in process, on a bounded sample of the same workload."""
    import torch

    lengths = make_lengths(w, seed)[:n_sample]
    audio_s = _cpu_init(w, lengths, seed)
    nthreads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        _cpu_step(None, min(8, n_sample), 1)  # warm-up
        t0 = time.perf_counter()
        _cpu_step(None, n_sample, 1)
        dt = time.perf_counter() - t0
    finally:
        torch.set_num_threads(nthreads)
    return {
        "value": audio_s / dt, "unit": UNIT, "cores": 1, "kind": "port",
        "sample": f"first {n_sample} utterances of the workload ({audio_s:.0f} audio-s), oracle port of the reference, "
                  f"torch {torch.__version__} CPU, 1 thread, in memory (no file I/O)",
        "seconds": dt,
    }


def run_reference_arm(args, w, wname):
    """--impl reference: the CPU path on all host cores (fork pool standing in for loky), on the SAME workload and
    `config` as our arm (rank 0's shard: the whole N = 1 job)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0
    import multiprocessing as mp

    import torch

    lengths, n_corpus = shard_lengths(w, wname, 0, world)
    n_items = len(lengths)
    audio_s = _cpu_init(w, lengths, seed=1234)
    cpus = os.cpu_count() or 1
    try:
        cpus = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(1)
    pool = mp.get_context("fork").Pool(cpus) if cpus > 1 else None
    try:
        for _ in range(args.warmup):
            _cpu_step(pool, n_items, cpus)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            _cpu_step(pool, n_items, cpus)
        dt = time.perf_counter() - t0
    finally:
        if pool is not None:
            pool.terminate()
    value = audio_s * args.steps / dt
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "strong" if n_corpus else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "impl": "reference",
        "config": workload_config(w, wname, world, lengths, n_corpus, args.data),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpus, "kind": "port",
                         "sample": f"the whole step: {n_items} utterances ({audio_s:.0f} audio-s); oracle port of the "
                                   f"reference (bit-identical to it on CPU), fork pool of {cpus} workers with the "
                                   "reference's batch rule, 1 intra-op thread each, in memory (no per-file torch.save/load)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class DeviceJob:
    """One rank's shard resident in HBM and the step over it.  Everything runs on ONE stream: overlapping the
    statistics exchange of step i with the feature kernel of step i + 1 on a side stream was measured and is slower
    (N = 2: 0.766 vs 0.720 ms per step, profiles/r02g_bench_n2_side_stream.json) -- the persistent feature kernel owns
    every SM (148 CTAs x 64 K registers), so the NCCL kernel cannot co-reside and delays one CTA of a statically
    partitioned grid instead."""

    def __init__(self, w, wname, rank, world, device, data):
        import torch

        import everyvoice_b200 as ev

        spec_type, sr, n_fft, win, hop, n_mels, f_min, f_max, *_ = w
        self.w, self.wname, self.device, self.world, self.rank = w, wname, device, world, rank
        self.lengths, self.n_corpus = shard_lengths(w, wname, rank, world)
        self.sample_offsets = np.concatenate([[0], np.cumsum(self.lengths)]).astype(np.int64)
        self.total_samples = int(self.sample_offsets[-1])
        self.audio_s = self.total_samples / sr
        self.seed = 1234 + rank
        self.samples = make_samples(data, self.lengths, sr, self.seed, device)
        self.d_packed, self.phone_offsets = make_durations(self.lengths, hop, self.seed)
        self.durations_dev = torch.from_numpy(self.d_packed.astype(np.int64)).to(device)
        self.phone_offsets_dev = torch.from_numpy(self.phone_offsets).to(device)
        self.pre = ev.Preprocessor(ev.AudioConfig(input_sampling_rate=sr, output_sampling_rate=sr, n_fft=n_fft,
                                                  fft_window_size=win, fft_hop_size=hop, n_mels=n_mels, f_min=f_min,
                                                  f_max=f_max, spec_type=spec_type), device=device)
        self.tf = self.pre.input_spectral_transform
        self.batch = self.tf.make_batch(self.sample_offsets, device, apply_log=True, keep_last=False)
        self.frame_offsets_dev = torch.from_numpy(self.batch.frame_offsets).to(device)
        self.total_frames = self.batch.total_frames
        self.spec = torch.empty((self.total_frames, self.batch.plan.row_floats), dtype=torch.float32, device=device)
        self.energy = torch.empty(self.total_frames, dtype=torch.float32, device=device)
        n_ph = int(self.phone_offsets[-1])
        self.phone = [torch.empty(n_ph, dtype=torch.float32, device=device) for _ in range(2)]
        self.stats = [torch.empty(5, dtype=torch.float64, device=device) for _ in range(2)]
        self.gathered = [torch.empty((world, 5), dtype=torch.float64, device=device) for _ in range(2)] if world > 1 else None
        self.main = torch.cuda.current_stream(device)
        self.side = None
        self.ev_stats = [torch.cuda.Event() for _ in range(2)]
        self.ev_norm = [torch.cuda.Event() for _ in range(2)]
        self.launches = 0
        self.feat_events = []
        self.it = 0

    def step(self, timed: bool):
        import ctypes as C

        import torch

        from everyvoice_b200 import _lib
        from everyvoice_b200.distributed import allgather_stats

        lib, dev = _lib.load(), self.device
        b = self.it & 1
        self.it += 1
        st = C.c_void_p(self.main.cuda_stream)
        if self.side is not None and self.it > 2:
            self.main.wait_event(self.ev_norm[b])          # phone[b] / stats[b] of step i - 2 have been normalised
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.main)
        self.tf.run(self.batch, self.samples, self.spec, self.energy)                              # 1 launch
        if timed:
            e1.record(self.main)
            self.feat_events.append((e0, e1))
        phone, stats = self.phone[b], self.stats[b]
        _lib.check(lib.evf_segment_mean(C.c_void_p(self.energy.data_ptr()), C.c_void_p(self.frame_offsets_dev.data_ptr()),
                                        C.c_void_p(self.durations_dev.data_ptr()),
                                        C.c_void_p(self.phone_offsets_dev.data_ptr()), len(self.lengths),
                                        C.c_void_p(phone.data_ptr()), st))                         # 1
        _lib.check(lib.evf_stats_partial(C.c_void_p(phone.data_ptr()), phone.numel(), C.c_void_p(stats.data_ptr()), 0, st))  # 2
        if self.world == 1:
            _lib.check(lib.evf_normalize_by_stats(C.c_void_p(phone.data_ptr()), phone.numel(),
                                                  C.c_void_p(stats.data_ptr()), st))               # 1
        elif self.side is None:
            parts = allgather_stats(stats, out=self.gathered[b])                                   # ONE NCCL collective
            _lib.check(lib.evf_normalize_by_gathered_stats(
                C.c_void_p(phone.data_ptr()), phone.numel(), C.c_void_p(parts.data_ptr()), parts.shape[0],
                parts.shape[1], st))                                                               # 1
        else:
            self.ev_stats[b].record(self.main)
            self.side.wait_event(self.ev_stats[b])
            with torch.cuda.stream(self.side):
                parts = allgather_stats(stats, out=self.gathered[b])                               # ONE NCCL collective
                _lib.check(lib.evf_normalize_by_gathered_stats(
                    C.c_void_p(phone.data_ptr()), phone.numel(), C.c_void_p(parts.data_ptr()), parts.shape[0],
                    parts.shape[1], C.c_void_p(self.side.cuda_stream)))                            # 1
                self.ev_norm[b].record(self.side)
        self.launches += 5
        return phone

    def finish(self):
        """Joins the side stream into the main one (the timed region ends after the last normalisation)."""
        if self.side is not None:
            self.main.wait_stream(self.side)

    def last_phone(self):
        return self.phone[(self.it - 1) & 1]

    def last_stats(self):
        b = (self.it - 1) & 1
        return self.gathered[b] if self.gathered is not None else self.stats[b]


def measure_device(job, steps, warmup, world, clocks=None):
    import torch
    import torch.distributed as dist

    dev = job.device

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(warmup):
        job.step(False)
    job.finish()
    job.launches = 0
    sync_all()
    if clocks is not None:
        clocks.samples.clear()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record(job.main)
    for _ in range(steps):
        job.step(True)
    job.finish()
    t_end.record(job.main)
    sync_all()
    elapsed_ms = t_start.elapsed_time(t_end)
    feat_ms = float(np.mean([a.elapsed_time(b) for a, b in job.feat_events]))
    audio_s_all, frames_all = job.audio_s, float(job.total_frames)
    if world > 1:
        t = torch.tensor([elapsed_ms, feat_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, feat_ms = float(t[0]), float(t[1])
        tot = torch.tensor([audio_s_all, frames_all], dtype=torch.float64, device=dev)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        audio_s_all = float(tot[0])
    return elapsed_ms, feat_ms, audio_s_all


def roofline_block(job, feat_ms, step_ms):
    spec_type, sr, n_fft, win, hop, n_mels, *_ = job.w
    bpf = algorithmic_bytes_per_frame(spec_type, hop, n_mels, n_fft)
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = job.total_frames * bpf / (feat_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        try:
            rec = json.load(open(tp)).get(job.wname)
            if isinstance(rec, dict):
                traffic = rec.get("bytes_per_launch")
                traffic_src = {k: rec.get(k) for k in ("source_sha", "git", "capture")}
                traffic_src["matches_this_build"] = rec.get("source_sha") == source_sha()
            elif rec is not None:
                traffic = rec
        except Exception:
            traffic = None
    if n_fft in (256, 512, 1024, 2048):
        kernel = "features_kernel"
    elif n_fft in (3072, 4096) and hop % (n_fft // 1024) == 0:
        kernel = "deinterleave_kernel + features_kernel (raw, per phase stream) + combine_kernel"
    else:
        kernel = "features_generic_kernel"
    return {
        "bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_frame": bpf,
        "frames_per_launch": int(job.total_frames), "kernel_ms": feat_ms, "kernel_share_of_step": feat_ms / step_ms,
    }


def check_parity(job, max_utts=None):
    """The buffers the timed loop left behind against the oracle, every utterance of this rank's shard (or the first
    `max_utts`): log-spectrogram, energy, and the phone values (de-normalised with the statistics the step itself
    derived, then compared with the oracle's averages; at N = 1 the statistics are compared as well)."""
    import torch

    sys.path.insert(0, str(ROOT / "tests"))
    from parity_pool import compare_all

    from everyvoice_b200.distributed import finalize_stats, merge_stats

    spec_type, sr, n_fft, win, hop, n_mels, f_min, f_max, *_ = job.w
    n = len(job.lengths) if max_utts is None else min(max_utts, len(job.lengths))
    off, f_off, p_off = job.sample_offsets[: n + 1], job.batch.frame_offsets[: n + 1], job.phone_offsets[: n + 1]
    t0 = time.perf_counter()
    x = job.samples[: int(off[-1])].cpu().numpy()
    spec = job.spec[: int(f_off[-1])].cpu().numpy()
    energy = job.energy[: int(f_off[-1])].cpu().numpy()
    st5 = merge_stats(job.last_stats().reshape(-1, 5)).cpu().tolist()
    ours = finalize_stats(st5, len(job.lengths))
    phone_norm = job.last_phone()[: int(p_off[-1])].cpu().numpy()
    phone_raw = phone_norm * np.float32(ours["std"]) + np.float32(ours["mean"])      # undo the step's normalisation
    res = compare_all(x, off, spec, energy, f_off, (sr, n_fft, win, hop, n_mels, f_min, f_max), spec_type,
                      job.d_packed[: int(p_off[-1])], p_off, phone_raw, return_oracle_phone=True)
    out = {
        "utterances_checked": res["utterances"], "frames_checked": res["frames"], "failures": len(res["failures"]),
        "max_abs_logspec": res["max_spec"], "max_abs_energy": res["max_energy"],
        "max_abs_phone_after_denormalising": res["max_phone"], "frames_exact": bool(res["frames"] == int(f_off[-1])),
        "nan_positions_equal": not any("NaN" in why for _, why in res["failures"]),
        "criterion": "tests/parity_pool.py: log-mel / energy / phone averages within 1e-3 of the oracle (linear: the "
                     "weak-bin criterion), frame counts and NaN positions exact",
        "oracle_workers": res["workers"], "seconds": None,
    }
    if res["failures"]:
        out["first_failures"] = [f"utterance {b}: {why}" for b, why in res["failures"][:3]]
    if job.world == 1 and n == len(job.lengths):
        from oracle import ev_oracle as O

        s = O.Scaler()
        for v in res["oracle_phone"]:
            s.append(torch.from_numpy(v))
        ref = s.calculate_stats()
        out["stats_max_rel_diff"] = max(abs(ours[k] - ref[k]) / max(1.0, abs(ref[k])) for k in ("mean", "std", "min", "max"))
        ref_norm = torch.cat([s.normalize(torch.from_numpy(v)) for v in res["oracle_phone"]]).numpy()
        m = ~np.isnan(ref_norm)
        out["max_abs_phone_normalised"] = float(np.abs(phone_norm[m] - ref_norm[m]).max())
    out["seconds"] = time.perf_counter() - t0
    return out


def measure_link(device, h2d_bytes, d2h_bytes, chunk=16 << 20):
    """What the box's host <-> device link does with the byte counts of ONE end-to-end step and no kernels at all:
    pinned H2D alone, D2H alone, and both directions at once on two streams (in chunks, like the pipeline).  The
    last one is the bound of the end-to-end path; max(alone) is what an ideal full-duplex link would allow."""
    import torch

    h_a, h_b = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory(), torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(h2d_bytes, dtype=torch.uint8, device=device)
    d_b = torch.zeros(d2h_bytes, dtype=torch.uint8, device=device)
    s1, s2 = torch.cuda.Stream(device), torch.cuda.Stream(device)

    def run(h2d, d2h, reps=5):
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    for a in range(0, h2d_bytes, chunk):
                        d_a[a:a + chunk].copy_(h_a[a:a + chunk], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    for a in range(0, d2h_bytes, chunk):
                        h_b[a:a + chunk].copy_(d_b[a:a + chunk], non_blocking=True)
        torch.cuda.synchronize(device)
        return (time.perf_counter() - t0) / reps * 1e3

    run(True, True, 1)
    t_in, t_out, t_both = run(True, False), run(False, True), run(True, True)
    return {"h2d_alone_ms": t_in, "d2h_alone_ms": t_out, "both_directions_ms": t_both,
            "h2d_gbs": h2d_bytes / t_in / 1e6, "d2h_gbs": d2h_bytes / t_out / 1e6,
            "ideal_full_duplex_ms": max(t_in, t_out)}


def measure_e2e(job, args, world, sample_dtype):
    """Every step: plan the batch (chunking + tile descriptors), H2D of samples + durations from pinned host memory,
    kernels, D2H of log-spectrogram + energy + normalised phone values into pinned host memory -- what
    `everyvoice preprocess` would hand to its file writers."""
    import torch
    import torch.distributed as dist

    dev = job.device
    if sample_dtype == torch.int16:
        host_in = torch.empty(job.total_samples, dtype=torch.int16).pin_memory()
        host_in.copy_((job.samples * 32767.0).round().to(torch.int16))
    else:
        host_in = torch.empty(job.total_samples, dtype=torch.float32).pin_memory()
        host_in.copy_(job.samples)
    host_spec = torch.empty((job.total_frames, job.batch.plan.row_floats), dtype=torch.float32).pin_memory()
    host_energy = torch.empty(job.total_frames, dtype=torch.float32).pin_memory()
    host_phone = torch.empty(int(job.phone_offsets[-1]), dtype=torch.float32).pin_memory()
    host_durs = torch.from_numpy(job.d_packed.astype(np.int64)).pin_memory()
    info = {}

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def e2e_step():
        pipe = job.pre.make_corpus_pipeline(job.sample_offsets, sample_dtype, host_durs, job.phone_offsets)
        pipe.run(host_in, host_spec, host_energy, host_phone)
        torch.cuda.synchronize(dev)
        info.update(h2d=pipe.h2d_bytes, d2h=pipe.d2h_bytes, chunks=len(pipe.chunks), launches=pipe.kernel_launches_per_run)
        return float(host_phone[0])

    steps = max(3, min(args.steps, 10))
    for _ in range(2):
        e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    sync_all()
    dt = time.perf_counter() - t0
    audio_s_all = job.audio_s
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        a = torch.tensor([job.audio_s], dtype=torch.float64, device=dev)
        dist.all_reduce(a, op=dist.ReduceOp.SUM)
        dt, audio_s_all = float(t[0]), float(a[0])
    del host_in, host_spec, host_energy, host_phone
    return audio_s_all * steps / dt, steps, info, dt / steps


def measure_preprocess_flow(job, steps=6, pcm16=True):
    """The whole numeric flow behind `everyvoice preprocess` for one batch, host to host, every step: the wav files'
    samples (pinned host) -> process_audio (loudness gate with its exact pass, peak normalisation, truncation, PCM16) ->
    process_spec -> process_energy (phone-level) -> statistics / normalisation -> processed PCM16 audio,
    log-spectrogram, energy and normalised phone values back in pinned host memory, as ONE chunked three-stream
    pipeline with the gate consumed on the device (Preprocessor.make_flow_pipeline).  Planning included, every step.
    An extra, informative key of the bench line (N = 1 only)."""
    import torch

    spec_type, sr, n_fft, win, hop, *_ = job.w
    pre, device, off = job.pre, job.device, job.sample_offsets
    if pcm16:  # the wav files' own samples
        host_all = torch.empty(int(off[-1]), dtype=torch.int16).pin_memory()
        host_all.copy_((job.samples * 32767.0).round().to(torch.int16))
    else:      # what load_audio returns
        host_all = torch.empty(int(off[-1]), dtype=torch.float32).pin_memory()
        host_all.copy_(job.samples)
    host_durs = torch.from_numpy(job.d_packed.astype(np.int64)).pin_memory()
    h_spec = torch.empty((job.total_frames, job.batch.plan.row_floats), dtype=torch.float32).pin_memory()
    h_energy = torch.empty(job.total_frames, dtype=torch.float32).pin_memory()
    h_phone = torch.empty(int(job.phone_offsets[-1]), dtype=torch.float32).pin_memory()
    h_audio = torch.empty(int(off[-1]), dtype=torch.int16).pin_memory()
    out = {}

    def step():
        flow = pre.make_flow_pipeline(off, sr, host_all.dtype, host_durs, job.phone_offsets)
        res = flow.run(host_all, h_spec, h_energy, h_phone, h_audio)()
        out.update(kept=int(res.keep.sum()), h2d=flow.h2d_bytes, d2h=flow.d2h_bytes + h_audio.numel() * 2,
                   chunks=len(flow.chunks))

    for _ in range(2):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": job.audio_s / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "utterances_kept": out["kept"],
            "input_format": "int16 PCM (the wav files' samples), pinned host" if pcm16 else "float32, pinned host",
            "h2d_bytes_per_step": out["h2d"], "d2h_bytes_per_step": out["d2h"], "chunks": out["chunks"],
            "api": "Preprocessor.make_flow_pipeline(...).run(host buffers): process_audio (gates on the device) -> "
                   "process_spec -> process_energy -> statistics -> normalise; processed PCM16 audio, log-mel, energy "
                   "and phone values back on the host"}


def run_extra(wname, args, rank, local_rank, world, device):
    """Device-timed line of another BASELINE config (no e2e / CPU legs), with parity on a 250-utterance sample."""
    import torch

    w = WORKLOADS[wname]
    job = DeviceJob(w, wname, rank, world, device, "white")
    steps = max(3, min(args.steps, 10))
    elapsed_ms, feat_ms, audio_s_all = measure_device(job, steps, 3, world)
    out = None
    if rank == 0:
        step_ms = elapsed_ms / steps
        out = {"metric": METRIC, "value": audio_s_all * steps / (elapsed_ms * 1e-3), "unit": UNIT, "n_gpus": world,
               "steps": steps, "ms_per_step": step_ms, "scaling": "strong" if job.n_corpus else "weak",
               "config": workload_config(w, wname, world, job.lengths, job.n_corpus, "white"),
               "roofline": roofline_block(job, feat_ms, step_ms)}
        if not args.no_parity:
            try:
                out["parity"] = check_parity(job, max_utts=250)
            except Exception as e:
                out["parity"] = {"error": repr(e)[:300]}
    del job
    torch.cuda.empty_cache()
    return out


def run_ours(args, w, wname):
    import torch
    import torch.distributed as dist

    from everyvoice_b200 import _lib

    spec_type, sr, n_fft, win, hop, n_mels, f_min, f_max, n_utts, min_s, max_s = w
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the CPU arm")
    _lib.load()
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    # staging buffers on the GPU's own NUMA node (matters for the end-to-end leg at N > 2)
    from everyvoice_b200.distributed import bind_to_gpu_numa_node

    numa_cpus = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on the C-level stdout when the communicator is created; stdout carries
        # exactly ONE JSON line (bench contract), so file descriptor 1 points at stderr while NCCL comes up
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier(device_ids=[local_rank])
            torch.cuda.synchronize(device)
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    # CPU baseline first (rank 0, N == 1 only), before the GPU gets busy
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_baseline_single_thread(w, seed=1234)

    corpus = wname in CORPUS_WORKLOADS
    job = DeviceJob(w, wname, rank, world, device, args.data)
    # NVML init takes tens of ms: start the sampler BEFORE the barrier, or rank 0 enters the timed
    # region late and every other rank waits for it in the first exchange
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    elapsed_ms, feat_ms, audio_s_all = measure_device(job, args.steps, args.warmup, world, clocks if rank == 0 else None)
    clock_info = clocks.stop() if rank == 0 else None
    value = audio_s_all * args.steps / (elapsed_ms * 1e-3)
    launches = job.launches

    parity = None
    if rank == 0 and not args.no_parity:
        try:
            parity = check_parity(job)
        except Exception as e:  # never costs the bench line; a failure is visible in the line
            parity = {"error": repr(e)[:300]}

    # ---- end to end through the public API with HOST buffers --------------------------------
    # pinned host copies of a whole shard: only while they stay below 8 GB of float32 input per rank
    e2e_ok = job.total_samples * 4 <= 8 * 2**30
    if e2e_ok:
        e2e_s16, e2e_steps, info_s16, sec_s16 = measure_e2e(job, args, world, torch.int16)
        e2e_f32, _, info_f32, sec_f32 = measure_e2e(job, args, world, torch.float32)
        # the copies alone (no kernels), same byte counts: rank 0 only, the other ranks idle (N > 1: the single-rank
        # link rate; tools/pcie_probe.py measures all ranks copying at once)
        link = link32 = None
        if rank == 0:
            link = measure_link(device, int(info_s16["h2d"]), int(info_s16["d2h"]))
            link32 = measure_link(device, int(info_f32["h2d"]), int(info_f32["d2h"]))
        if world > 1:
            dist.barrier()

    flow = None
    if world == 1 and e2e_ok and spec_type in ("mel", "mel-librosa") and not corpus:
        try:
            flow = measure_preprocess_flow(job, pcm16=True)
            f32 = measure_preprocess_flow(job, pcm16=False)
            flow["float32_input"] = {k: f32[k] for k in ("value", "unit", "ms_per_step", "h2d_bytes_per_step")}
        except Exception as e:  # informative extra: never costs the bench line
            flow = {"error": repr(e)[:300]}

    line = None
    if rank == 0:
        step_ms = elapsed_ms / args.steps
        e2e = None
        if e2e_ok:
            e2e = {"value": e2e_s16, "unit": UNIT, "h2d_bytes_per_step": int(info_s16["h2d"]),
                   "d2h_bytes_per_step": int(info_s16["d2h"]), "steps": e2e_steps, "ms_per_step": sec_s16 * 1e3,
                   "input_format": "int16 PCM (the on-disk format process_audio writes), pinned host; converted in the "
                                   "kernel, bit-identical to float input",
                   "chunks": info_s16["chunks"], "gpu_launches_per_step": info_s16["launches"],
                   "copies_only": link,
                   "frac_of_copies_only": link["both_directions_ms"] / (sec_s16 * 1e3),
                   "frac_of_ideal_full_duplex": link["ideal_full_duplex_ms"] / (sec_s16 * 1e3),
                   "host_affinity": (f"rank 0 bound to the {len(numa_cpus)} cores NVML reports local to its GPU"
                                     if numa_cpus else "not bound"),
                   "api": "Preprocessor.make_corpus_pipeline(...).run(host buffers): batch planning + chunked "
                          "H2D / kernels / D2H on three streams, every step",
                   "float32_input": {"value": e2e_f32, "unit": UNIT, "h2d_bytes_per_step": int(info_f32["h2d"]),
                                     "d2h_bytes_per_step": int(info_f32["d2h"]), "ms_per_step": sec_f32 * 1e3,
                                     "copies_only": link32,
                                     "frac_of_copies_only": link32["both_directions_ms"] / (sec_f32 * 1e3),
                                     "note": "same call fed float32 samples (what torchaudio.load returns): twice the H2D bytes"}}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "strong" if corpus else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(w, wname, world, job.lengths, job.n_corpus, args.data),
            "realtime_factor_per_gpu": value / world,
            "roofline": roofline_block(job, feat_ms, step_ms),
            "parity": parity,
            "cpu_baseline": cpu_base,
            "e2e": e2e,
            "preprocess_flow": flow,
            "gpu_launches": launches,
            "clocks": clock_info,
            "exchange": ("one all-gather of the 5-number summaries per step, on the compute stream" if world > 1
                         else "none (N = 1)"),
        }
    del job
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs, device-timed (every rank takes part; rank 0 reports) ---------------------
    extras = {}
    if not args.no_extras and wname == DEFAULT_WORKLOAD:
        for name in EXTRA_WORKLOADS:
            try:
                r = run_extra(name, args, rank, local_rank, world, device)
            except Exception as e:
                r = {"error": repr(e)[:300]}
                if world > 1:
                    raise   # a rank that leaves the collectives would hang the others
            if rank == 0:
                extras[name] = r
    if rank == 0:
        if extras:
            line["extra"] = extras
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=list(WORKLOADS), default=DEFAULT_WORKLOAD)
    ap.add_argument("--data", choices=["white", "speech", "lj_tiled"], default="white",
                    help="synthetic signal: U(-0.95, 0.95) noise, harmonic speech-like stacks, or the bundled LJ wavs tiled")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the `extra` block (other BASELINE configs)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the timed buffers")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference_arm(args, w, args.workload)
    return run_ours(args, w, args.workload)


if __name__ == "__main__":
    sys.exit(main())
