#!/usr/bin/env python
"""Where the pipelined preprocess flow (bench.py `preprocess_flow`) spends its time: planning, enqueueing, device."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import everyvoice_b200 as ev
from everyvoice_b200 import synth

dev = torch.device("cuda", 0)
sr, hop = 22050, 256
with_audio = "--no-audio" not in sys.argv
head_mb = float(sys.argv[sys.argv.index("--head-mb") + 1]) if "--head-mb" in sys.argv else None
lens = synth.utterance_lengths(1000, sr, hop, 1234)
off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
durs = [synth.synthetic_durations(int(L) // hop, seed=1234 + i) for i, L in enumerate(lens)]
d_packed, p_off = synth.pack_ragged(durs)
n = int(off[-1])
host_in = torch.empty(n, dtype=torch.int16).pin_memory()
host_in.copy_(((torch.rand(n) * 1.9 - 0.95) * 32767).to(torch.int16))
pre = ev.Preprocessor(ev.AudioConfig(spec_type="mel"), device=dev)
T = int((lens // hop).sum())
h_spec = torch.empty((T, 80), dtype=torch.float32).pin_memory()
h_energy = torch.empty(T, dtype=torch.float32).pin_memory()
h_phone = torch.empty(int(p_off[-1]), dtype=torch.float32).pin_memory()
h_audio = torch.empty(n, dtype=torch.int16).pin_memory() if with_audio else None
h_durs = torch.from_numpy(d_packed.astype(np.int64)).pin_memory()
rows = []
for i in range(10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    flow = pre.make_flow_pipeline(off, sr, torch.int16, h_durs, p_off,
                                   head_chunk_bytes=None if head_mb is None else int(head_mb * 2**20))
    t1 = time.perf_counter()
    fin = flow.run(host_in, h_spec, h_energy, h_phone, h_audio)
    t2 = time.perf_counter()
    res = fin()
    t3 = time.perf_counter()
    if i >= 2:
        rows.append((t1 - t0, t2 - t1, t3 - t2, t3 - t0))
r = np.array(rows) * 1e3
print(f"audio back {with_audio} head {head_mb} MB: chunks {len(flow.chunks)}: plan {r[:,0].mean():.2f} ms, enqueue {r[:,1].mean():.2f} ms, drain {r[:,2].mean():.2f} ms, "
      f"total {r[:,3].mean():.2f} ms (min {r[:,3].min():.2f}); kept {int(res.keep.sum())}; h2d {flow.h2d_bytes/1e6:.0f} MB d2h {(flow.d2h_bytes + (n*2 if with_audio else 0))/1e6:.0f} MB")
