"""Timing of the differentiable transform (forward + backward) on one B200, next to the
same graph in plain torch on the same GPU (torch.stft through cuFFT + mel-basis matmul + log: what the reference's
HiFiGAN training executes, hfgl/model.py:581-590, 719-721), as a measured baseline only.

    python tools/backward_bench.py > profiles/rNN_backward_bench.json"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import everyvoice_b200 as ev  # noqa: E402
from everyvoice_b200 import filterbanks  # noqa: E402


def timed(fn, reps):
    for _ in range(5):
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
    return float(np.median(ts)), float(np.min(ts))


def main():
    dev = torch.device("cuda", 0)
    tf = ev.get_spectral_transform("mel", 1024, 1024, 256, 22050, 80, 0, 8000).to(dev)
    fb = filterbanks.melscale_fbanks_htk_slaney(513, 0.0, 8000.0, 80, 22050).to(dev, torch.float32)
    win = torch.hann_window(1024, device=dev)

    def otf(x):  # T.MelSpectrogram's graph on the GPU: torch.stft (cuFFT) -> |.|^2 -> (spec^T @ fb)^T
        spec = torch.stft(x, 1024, 256, 1024, window=win, center=True, pad_mode="reflect", normalized=False,
                          onesided=True, return_complex=True)
        power = spec.real.square() + spec.imag.square()
        return torch.matmul(power.transpose(-1, -2), fb).transpose(-1, -2)

    out = {"spec_type": "mel", "n_fft": 1024, "hop": 256, "n_mels": 80, "cases": []}
    for name, B, L, reps in (("hifigan_step_16x8192", 16, 8192, 200), ("hifigan_step_64x8192", 64, 8192, 200),
                             ("batch_1000x5s", 1000, 110336, 20)):
        x = (torch.rand(B, L, device=dev) * 1.9 - 0.95) * 0.5
        target = None

        def ours():
            xg = x.detach().requires_grad_(True)
            y = ev.dynamic_range_compression_torch(tf(xg))
            nonlocal target
            if target is None:
                target = torch.randn_like(y)
            (torch.nn.functional.l1_loss(y, target) * 45).backward()
            return xg.grad

        def ours_fused():   # log fused into the forward epilogue and into the backward (SpectralTransform.training_mel)
            xg = x.detach().requires_grad_(True)
            y = tf.features(xg, normalize=True, keep_last=True)
            (torch.nn.functional.l1_loss(y, target) * 45).backward()
            return xg.grad

        def torch_gpu():
            xg = x.detach().requires_grad_(True)
            y = torch.log(torch.clamp(otf(xg), min=1e-5))
            (torch.nn.functional.l1_loss(y, target) * 45).backward()
            return xg.grad

        g1 = ours()
        try:
            g2 = torch_gpu()
            rel = float((g1 - g2).abs().max() / g2.abs().max())
            t_ref = timed(torch_gpu, reps)
        except Exception as e:
            rel, t_ref = None, (None, None)
            out.setdefault("torch_gpu_error", repr(e)[:200])
        t_ours = timed(ours, reps)
        g3 = ours_fused()
        rel_fused = float((g3 - g1).abs().max() / g1.abs().max())
        t_fused = timed(ours_fused, reps)
        audio_s = B * L / 22050
        out["cases"].append({"case": name, "batch": B, "samples": L, "ours_ms": t_ours[0], "ours_ms_min": t_ours[1],
                             "ours_fused_log_ms": t_fused[0], "ours_fused_log_ms_min": t_fused[1],
                             "fused_vs_two_step_max_rel_grad_diff": rel_fused,
                             "torch_gpu_ms": t_ref[0], "torch_gpu_ms_min": t_ref[1], "max_rel_grad_diff": rel,
                             "ours_audio_s_per_s": audio_s / (t_ours[0] / 1e3)})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
