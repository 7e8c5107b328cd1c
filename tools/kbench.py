"""Kernel-only timing of features_kernel for A/B runs on the GPU box.

    [EVF_LIB=path/to/variant.so] [EVF_WARPS=8] python tools/kbench.py [reps]

Prints one line per workload: mean / min kernel ms over `reps` launches (CUDA events, inputs
resident and larger than L2) and the max |diff| against the default library's output when
EVF_REF_OUT points at a saved tensor (so a variant's numerics are checked in the same call)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import everyvoice_b200 as ev  # noqa: E402
from everyvoice_b200 import synth  # noqa: E402

WORK = {
    "melA": ("mel", 22050, 1024, 1024, 256, 80, 0, 8000),
    "linA": ("linear", 22050, 1024, 1024, 256, 80, 0, 8000),
    "melB": ("mel", 44100, 2048, 2048, 512, 128, 0, 8000),
    "librosaA": ("mel-librosa", 22050, 1024, 1024, 256, 80, 0, 8000),
    "melA_s16": ("mel", 22050, 1024, 1024, 256, 80, 0, 8000),
    "mel512": ("mel", 16000, 512, 512, 128, 80, 0, 8000),        # warp kernel, two packed jobs per warp
    "lin512": ("linear", 16000, 512, 512, 128, 80, 0, 8000),
    "mel256": ("mel", 8000, 256, 256, 64, 40, 0, 4000),          # warp kernel, four packed jobs per warp
    "mel4096": ("mel", 44100, 4096, 4096, 1024, 128, 0, 8000),   # four phase-stream transforms of 1024 points + combine
    "mel3072": ("mel", 48000, 3072, 3072, 768, 80, 0, 8000),     # three phase streams (16 -> 48 kHz vocoder output transform)
}


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    names = sys.argv[2].split(",") if len(sys.argv) > 2 else [n for n in WORK if "512" not in n and "256" not in n and "4096" not in n and "3072" not in n]
    dev = torch.device("cuda", 0)
    tag = os.environ.get("EVF_TAG", Path(os.environ.get("EVF_LIB", "default")).stem)
    for name in names:
        st, sr, n_fft, win, hop, n_mels, f_min, f_max = WORK[name]
        lens = synth.utterance_lengths(1000, sr, hop, 1234)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        g = torch.Generator(device=dev)
        g.manual_seed(1234)
        x = torch.rand(int(off[-1]), device=dev, generator=g) * 1.9 - 0.95
        tf = ev.get_spectral_transform(st, n_fft, win, hop, sr, n_mels, f_min, f_max).to(dev)
        if name.endswith("_s16"):
            x = (x * 32767.0).round().to(torch.int16)
        batch = tf.make_batch(off, dev, sample_dtype=x.dtype)
        spec = torch.empty((batch.total_frames, batch.plan.row_floats), device=dev)
        en = torch.empty(batch.total_frames, device=dev)
        for _ in range(3):
            tf.run(batch, x, spec, en)
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            tf.run(batch, x, spec, en)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        chk = ""
        refdir = os.environ.get("EVF_REF_OUT")
        if refdir:
            f = Path(refdir) / f"{name}.pt"
            if f.exists():
                r = torch.load(f)
                chk = f" max|d spec|={float((spec.cpu() - r['spec']).abs().max()):.2e} max|d energy|={float((en.cpu() - r['energy']).abs().max()):.2e}"
            else:
                f.parent.mkdir(parents=True, exist_ok=True)
                torch.save({"spec": spec.cpu(), "energy": en.cpu()}, f)
                chk = " (saved as reference)"
        print(f"{tag:18s} {name:9s} frames={batch.total_frames} mean={np.mean(ts):.4f} ms min={np.min(ts):.4f} ms{chk}", flush=True)


if __name__ == "__main__":
    main()
