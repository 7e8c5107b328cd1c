"""One forward + backward of the differentiable transform on the 1 000 x 5 s batch, for an ncu launch list:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/launches_backward.csv python tools/backward_launches.py"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import everyvoice_b200 as ev  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    fused = "fused" in sys.argv   # log folded into the forward epilogue and the backward (features(normalize=True))
    nums = [int(a) for a in sys.argv[1:] if a.isdigit()]
    n_fft, hop = (nums[0], nums[1]) if len(nums) > 1 else (1024, 256)
    tf = ev.get_spectral_transform("mel", n_fft, n_fft, hop, 22050, 80, 0, 8000).to(dev)
    x = (torch.rand(1000, 110336, device=dev) * 1.9 - 0.95) * 0.5
    target = None
    for _ in range(2):
        xg = x.detach().requires_grad_(True)
        y = tf.features(xg, normalize=True, keep_last=True) if fused else ev.dynamic_range_compression_torch(tf(xg))
        if target is None:
            target = torch.randn_like(y)
        (torch.nn.functional.l1_loss(y, target) * 45).backward()
    torch.cuda.synchronize()
    print("ok", float(xg.grad.abs().max()))


if __name__ == "__main__":
    main()
