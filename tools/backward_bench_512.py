"""Forward + backward of the log-mel at n_fft 512 (16 kHz / hop 128 and StyleTTS2's 24 kHz loss transform 512 / 50 / 240):
the warp kernels with two packed jobs per warp against the any-size kernels (fft_path="generic").

    python tools/backward_bench_512.py > profiles/rNN_backward_bench_512.json"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import everyvoice_b200 as ev  # noqa: E402


def timed(fn, reps=20):
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
    return float(np.median(ts))


def main():
    dev = torch.device("cuda", 0)
    out = []
    for name, sr, win, hop, n_mels, B, L in (("16k_512_128_1000x5s", 16000, 512, 128, 80, 1000, 80000),
                                             ("styletts2_loss_512_50_240_16x24000", 24000, 240, 50, 128, 16, 24000),
                                             ("styletts2_loss_512_50_240_256x24000", 24000, 240, 50, 128, 256, 24000)):
        x = (torch.rand(B, L, device=dev) * 1.9 - 0.95) * 0.5
        rec = {"case": name, "n_fft": 512, "win": win, "hop": hop, "n_mels": n_mels, "batch": B, "samples": L}
        grads = {}
        for path in ("auto", "generic"):
            tf = ev.SpectralTransform("mel", 512, win, hop, sr, n_mels, 0, 8000, fft_path=path)
            target = None

            def step():
                nonlocal target
                xg = x.detach().requires_grad_(True)
                y = tf.features(xg, normalize=True, keep_last=True)
                if target is None:   # the same target for both paths
                    target = torch.randn(y.shape, device=y.device, generator=torch.Generator(device=y.device).manual_seed(7)).to(y.dtype)
                    target = torch.empty_like(y).copy_(target)
                ((y - target) ** 2).mean().backward()
                return xg.grad

            grads[path] = step()
            rec[("warp_kernels_ms" if path == "auto" else "any_size_kernels_ms")] = timed(step)
        rec["max_rel_grad_diff"] = float((grads["auto"] - grads["generic"]).abs().max() / grads["generic"].abs().max())
        out.append(rec)
    print(json.dumps({"cases": out}))


if __name__ == "__main__":
    main()
