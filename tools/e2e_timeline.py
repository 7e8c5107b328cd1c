#!/usr/bin/env python
"""Per-chunk timeline of the end-to-end step (bench.py `e2e`): when each chunk's H2D copy, kernels and D2H copy finish,
in ms from the start of `CorpusPipeline.run`, from CUDA events recorded in place of the pipeline's own.  One B200.
    python tools/e2e_timeline.py [--chunk-mb 16] [--head-mb 2] [--dtype s16|f32]"""
import argparse, sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import everyvoice_b200 as ev
from everyvoice_b200 import synth


class LoggedEvent:
    """Stands in for one of the pipeline's reusable events: every record() is a fresh timing event kept in `log`."""

    def __init__(self, log, kind):
        self.log, self.kind, self.cur = log, kind, None

    def record(self, stream=None):
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream)
        self.cur = e
        self.log.append((self.kind, e))

    def wait(self, stream=None):
        self.cur.wait(stream)


ap = argparse.ArgumentParser()
ap.add_argument("--chunk-mb", type=float, default=16.0)
ap.add_argument("--dtype", default="s16")
ap.add_argument("--flow", action="store_true", help="the FlowPipeline (front-end included, PCM16 audio coming back)")
ap.add_argument("--head-mb", type=float, default=None, help="size of the first chunks (default: the chunk size)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
sr, hop = 22050, 256
lens = synth.utterance_lengths(1000, sr, hop, 1234)
off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
durs = [synth.synthetic_durations(int(L) // hop, seed=1234 + i) for i, L in enumerate(lens)]
d_packed, p_off = synth.pack_ragged(durs)
dt = torch.int16 if a.dtype == "s16" else torch.float32
n = int(off[-1])
host_in = torch.empty(n, dtype=dt).pin_memory()
host_in.copy_((torch.rand(n) * 1.9 - 0.95) * (32767 if dt == torch.int16 else 1))
pre = ev.Preprocessor(ev.AudioConfig(spec_type="mel"), device=dev)
T = int((lens // hop).sum())
host_spec = torch.empty((T, 80), dtype=torch.float32).pin_memory()
host_energy = torch.empty(T, dtype=torch.float32).pin_memory()
host_phone = torch.empty(int(p_off[-1]), dtype=torch.float32).pin_memory()
host_durs = torch.from_numpy(d_packed.astype(np.int64)).pin_memory()
h_audio = torch.empty(n, dtype=torch.int16).pin_memory() if a.flow else None
for rep in range(4):
    head = None if a.head_mb is None else int(a.head_mb * 2**20)
    if a.flow:
        pipe = pre.make_flow_pipeline(off, sr, dt, host_durs, p_off, chunk_bytes=int(a.chunk_mb * 2**20), head_chunk_bytes=head)
    else:
        pipe = pre.make_corpus_pipeline(off, dt, host_durs, p_off, chunk_bytes=int(a.chunk_mb * 2**20), head_chunk_bytes=head)
    res = pipe.res
    log = []
    if rep == 3:
        if a.flow:
            for k in ("in", "comp", "out", "loud"):
                res._bufs["events:" + k] = [LoggedEvent(log, k) for _ in range(pipe.NB)]
        else:
            res.ev_in = [LoggedEvent(log, "in") for _ in range(2)]
            res.ev_comp = [LoggedEvent(log, "comp") for _ in range(2)]
            res.ev_out = [LoggedEvent(log, "out") for _ in range(2)]
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    start.record()
    if a.flow:
        fin = pipe.run(host_in, host_spec, host_energy, host_phone, h_audio)
    else:
        pipe.run(host_in, host_spec, host_energy, host_phone)
    t1 = time.perf_counter()
    end.record()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"rep {rep}: enqueue {1e3 * (t1 - t0):.2f} ms, total {1e3 * (t2 - t0):.2f} ms, device {start.elapsed_time(end):.2f} ms")
per = {"in": [], "comp": [], "out": [], "loud": []}
for kind, e in log:
    per[kind].append(start.elapsed_time(e))
print("chunk  MB_in   h2d_done  comp_done  d2h_done  loudness_done   (ms from the start of run)")
for i, c in enumerate(pipe.chunks):
    mb = (c.s1 - c.s0) * pipe.esize / 2**20
    ld = f"{per['loud'][i]:9.3f}" if per["loud"] else ""
    print(f"{i:5d} {mb:6.2f} {per['in'][i]:9.3f} {per['comp'][i]:10.3f} {per['out'][i]:9.3f} {ld}")
