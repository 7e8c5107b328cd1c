#!/usr/bin/env python
"""Where the end-to-end step (bench.py `e2e`) spends its time: planning, enqueueing, device.  One B200.
    python tools/e2e_probe.py [--chunk-mb 16] [--dtype s16|f32]"""
import argparse, sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import everyvoice_b200 as ev
from everyvoice_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--chunk-mb", type=float, default=16.0)
ap.add_argument("--dtype", default="s16")
ap.add_argument("--reps", type=int, default=10)
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
rows = []
for i in range(a.reps + 2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pipe = pre.make_corpus_pipeline(off, dt, host_durs, p_off, chunk_bytes=int(a.chunk_mb * 2**20))
    t1 = time.perf_counter()
    pipe.run(host_in, host_spec, host_energy, host_phone)
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    if i >= 2:
        rows.append((t1 - t0, t2 - t1, t3 - t2, t3 - t0))
r = np.array(rows) * 1e3
print(f"dtype {a.dtype} chunk {a.chunk_mb} MB chunks {len(pipe.chunks)}: plan {r[:,0].mean():.2f} ms, enqueue {r[:,1].mean():.2f} ms, "
      f"drain {r[:,2].mean():.2f} ms, total {r[:,3].mean():.2f} ms (min {r[:,3].min():.2f}); h2d {pipe.h2d_bytes/1e6:.0f} MB d2h {pipe.d2h_bytes/1e6:.0f} MB")
