#!/usr/bin/env python
"""Timing of the device pitch tracker (DIO speed 4 + StoneMask + unvoiced fill; evfeat_pitch.cu) on the
1 000-utterance workload, next to the CPU restatement (oracle/world_pitch.py, numpy, one process per host core) on a
bounded sample.  PARITY UNPINNED (pyworld is not available offline).
    python tools/pitch_bench.py > profiles/rNN_pitch_bench.json"""
import json, os, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import everyvoice_b200 as ev
from everyvoice_b200 import synth


def _cpu(args):
    from oracle import world_pitch as W
    x, sr, hop = args
    return len(W.extract_pitch(x, sr, hop))


def main():
    dev = torch.device("cuda", 0)
    sr, hop, n = 22050, 256, 1000
    lens = synth.utterance_lengths(n, sr, hop, 1234)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    xs = [synth.speech_like(int(L), sr, seed=i) for i, L in enumerate(lens[:64])]
    x = torch.from_numpy(np.concatenate([xs[i % 64][: int(L)] if len(xs[i % 64]) >= L else np.resize(xs[i % 64], int(L))
                                         for i, L in enumerate(lens)])).to(dev)
    pre = ev.Preprocessor(ev.AudioConfig(spec_type="mel"), device=dev)
    for _ in range(2):
        pre.extract_pitch_batch(x, off)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        p, f_off = pre.extract_pitch_batch(x, off)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    audio_s = float(off[-1]) / sr
    import multiprocessing as mp
    cpus = len(os.sched_getaffinity(0))
    sample = [(xs[i].astype(np.float32), sr, hop) for i in range(min(64, 4 * cpus))]
    with mp.get_context("spawn").Pool(cpus) as pool:
        pool.map(_cpu, sample[:cpus])
        t0 = time.perf_counter()
        pool.map(_cpu, sample)
        dt = time.perf_counter() - t0
    cpu_audio_s = sum(len(s[0]) for s in sample) / sr
    print(json.dumps({"workload": f"{n} utterances 1-10 s at {sr} Hz, hop {hop}: dio(speed=4) + stonemask + unvoiced fill",
                      "audio_s": audio_s, "frames": int(f_off[-1]), "device_ms": float(np.mean(ts)), "device_ms_min": float(np.min(ts)),
                      "device_audio_s_per_s": audio_s / (np.mean(ts) * 1e-3),
                      "cpu_restatement": {"audio_s_per_s": cpu_audio_s / dt, "cores": cpus, "sample": f"{len(sample)} utterances ({cpu_audio_s:.0f} audio-s)",
                                          "note": "numpy restatement of WORLD (not pyworld's C++, which is unavailable here)"},
                      "parity": "unpinned (pyworld absent); kernels == restatement: tests/test_gpu_frontend.py::test_pitch_tracker_matches_the_world_restatement"}))


if __name__ == "__main__":
    main()
