"""Timing of the audio front-end kernels (evfeat_audio.cu) on one B200.

    python tools/frontend_bench.py [n_utts] [in_rate] [out_rate] > profiles/rNN_frontend_bench.json

Workload: `n_utts` synthetic 1-10 s mono utterances at `in_rate` (default 1000 at 44.1 kHz, seed 1234), resampled to
`out_rate` (22.05 kHz), peak-normalised, truncated to a multiple of hop 256 and written as PCM16.  Reports, per
kernel, the CUDA-event time with device-resident input, the algorithmic bytes (every input read once, every output
written once) and the achieved fraction of the measured HBM peak; the end-to-end rate of
`Preprocessor.process_audio_batch` from host tensors; and the same operations through torchaudio on the host cores
(bounded sample) as the CPU baseline."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import everyvoice_b200 as ev  # noqa: E402
from everyvoice_b200 import _lib, synth  # noqa: E402
from everyvoice_b200.heavy import _ptr, _stream_ptr  # noqa: E402


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts)), float(np.min(ts))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    sr_in = int(sys.argv[2]) if len(sys.argv) > 2 else 44100
    sr_out = int(sys.argv[3]) if len(sys.argv) > 3 else 22050
    hop = 256
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    peak = 6561.0
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peak = float(json.loads(pk.read_text()).get("hbm_gbs", peak))
    lens = synth.utterance_lengths(n, sr_in, 1, 1234)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    x = (torch.rand(int(off[-1]), device=dev, generator=g) * 1.9 - 0.95) * 0.5
    audio_s = float(off[-1]) / sr_in
    out = {"workload": f"{n} utterances 1-10 s, {sr_in} -> {sr_out} Hz, hop {hop}, PCM16 out", "audio_s": audio_s,
           "hbm_peak_gbs": peak, "kernels": {}}

    def rec(name, ms, nbytes):
        out["kernels"][name] = {"ms": ms[0], "ms_min": ms[1], "algorithmic_bytes": int(nbytes),
                                "achieved_gbs": nbytes / ms[0] / 1e6, "frac_of_hbm_peak": nbytes / ms[0] / 1e6 / peak}

    st = _stream_ptr(dev)
    d_off = torch.from_numpy(off).to(dev)
    step = int(lib.evf_audio_loudness_step(sr_in))
    s_off = np.concatenate([[0], np.cumsum(5 * (lens // step + 5))]).astype(np.int64)
    d_soff = torch.from_numpy(s_off).to(dev)
    scratch = torch.empty(int(s_off[-1]), dtype=torch.float32, device=dev)
    lk = torch.empty(n, dtype=torch.float32, device=dev)
    flags = torch.zeros(n, dtype=torch.int32, device=dev)
    from everyvoice_b200.audio import k_weighting_coefficients
    import ctypes as C
    kw = k_weighting_coefficients(sr_in)
    rec("loudness (partial + gate + exact pass for utterances at the gate)",
        timed(lambda: _lib.check(lib.evf_audio_loudness(_ptr(x), _lib.SAMPLES_F32, _ptr(d_off), n, int(lens.max()), sr_in,
                                                        kw.ctypes.data_as(C.c_void_p), 0.05, -36.0, _ptr(scratch),
                                                        _ptr(d_soff), _ptr(flags), _ptr(lk), st))), 4 * int(off[-1]))
    rs = ev.Resampler(sr_in, sr_out, dev)
    y, y_off = rs(x, off)
    d_yoff = torch.from_numpy(y_off).to(dev)
    rec("resample", timed(lambda: rs.launch(x, d_off, d_yoff, n, int(np.diff(y_off).max()), y)),
        4 * int(off[-1]) + 4 * int(y_off[-1]))
    full = np.diff(y_off)
    absmax = torch.empty(n, dtype=torch.float32, device=dev)
    rec("absmax", timed(lambda: _lib.check(lib.evf_audio_absmax(_ptr(y), _lib.SAMPLES_F32, _ptr(d_yoff), n, int(full.max()), _ptr(absmax), st))),
        4 * int(y_off[-1]))
    kept = (full // hop) * hop
    dst = np.concatenate([[0], np.cumsum(kept)]).astype(np.int64)
    d_dst = torch.from_numpy(dst).to(dev)
    o16 = torch.empty(int(dst[-1]), dtype=torch.int16, device=dev)
    rec("finalize (normalise + truncate + PCM16)",
        timed(lambda: _lib.check(lib.evf_audio_finalize(_ptr(y), _lib.SAMPLES_F32, _ptr(d_yoff), _ptr(d_dst), n, int(kept.max()), _ptr(absmax),
                                                        None, _ptr(o16), st))), 6 * int(dst[-1]))
    total_ms = sum(k["ms"] for k in out["kernels"].values())
    out["device_total_ms"] = total_ms
    out["device_audio_s_per_s"] = audio_s / (total_ms / 1e3)

    # end to end through the operator surface, host tensors in (pinned), packed PCM16 on the device out
    host = [x[int(off[i]):int(off[i + 1])].cpu().pin_memory() for i in range(n)]
    pre = ev.Preprocessor(ev.AudioConfig(spec_type="mel"), device=dev)
    for _ in range(2):
        res = pre.process_audio_batch(host, sr_in, resample_rate=sr_out, hop_size=hop, out_dtype=torch.int16)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        res = pre.process_audio_batch(host, sr_in, resample_rate=sr_out, hop_size=hop, out_dtype=torch.int16)
    torch.cuda.synchronize()
    e2e = (time.perf_counter() - t0) / 3
    out["e2e"] = {"audio_s_per_s": audio_s / e2e, "ms": e2e * 1e3, "h2d_bytes": 4 * int(off[-1]), "kept": len(res.kept),
                  "api": "Preprocessor.process_audio_batch(list of pinned host tensors) -> packed int16 on the device"}

    # CPU baseline: what process_audio calls (torchaudio on the host), first `m` utterances, all intra-op threads
    import torchaudio

    m = min(n, 100)
    t0 = time.perf_counter()
    for i in range(m):
        a = host[i][None].clone()
        torchaudio.functional.loudness(a, sr_in)
        a = torchaudio.functional.resample(a, sr_in, sr_out)
        a /= torch.max(torch.abs(a))
        a *= 0.95
        a = a.squeeze()
        a = a[: (a.size(0) // hop) * hop]
    cpu = time.perf_counter() - t0
    out["cpu_baseline"] = {"audio_s_per_s": float(off[m]) / sr_in / cpu, "kind": "torchaudio on the host (what process_audio calls)",
                           "threads": torch.get_num_threads(), "sample": f"first {m} utterances"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
