"""Error statistics of the CUDA path and of the CPU reference port against the fp64 truth,
bucketed by the exact log-power (run on the GPU box)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import numpy as np, torch
import everyvoice_b200 as ev
from everyvoice_b200 import synth
from oracle import ev_oracle as O

dev = torch.device("cuda", 0)
CASES = {"A": (22050, 1024, 1024, 256, 80), "B": (44100, 2048, 2048, 512, 128)}
if "--wide" in sys.argv:   # the any-size kernel on the largest / smallest transforms
    CASES = {"H4096": (44100, 4096, 4096, 1024, 128), "W4096": (44100, 4096, 3000, 750, 128), "S512": (16000, 512, 512, 128, 80),
             "R3": (16000, 3072, 3072, 768, 80)}
for cfg_name, (sr, n_fft, win, hop, n_mels) in CASES.items():
    for st in ("linear", "mel"):
        tf = ev.get_spectral_transform(st, n_fft, win, hop, sr, n_mels, 0, 8000).to(dev)
        otf = O.get_spectral_transform(st, n_fft, win, hop, sr, n_mels, 0, 8000)
        for kind in ("speech", "white"):
            errs_o, errs_r, dif, tr = [], [], [], []
            for seed in range(6):
                L = 120 * hop
                x = synth.speech_like(L, sr, 900 + seed) if kind == "speech" else synth.white_noise(L, 900 + seed)
                ours = tf.features_ragged(torch.from_numpy(x).to(dev), np.array([0, L])).utterance(0).cpu().double().numpy()
                ref = O.process_spec(torch.from_numpy(x), otf, hop).double().numpy()
                truth, _ = O.truth_features(x, st, n_fft, win, hop, sr, n_mels, 0, 8000)
                errs_o.append(np.abs(ours - truth).ravel()); errs_r.append(np.abs(ref - truth).ravel())
                dif.append(np.abs(ours - ref).ravel()); tr.append(truth.ravel())
            eo, er, d, t = map(np.concatenate, (errs_o, errs_r, dif, tr))
            print(f"== {cfg_name} {st} {kind}: n={t.size}")
            for lo, hi in ((-12, -11), (-11, -10), (-10, -9), (-9, -7), (-7, -3), (-3, 20)):
                m = (t >= lo) & (t < hi)
                if m.sum() == 0:
                    continue
                print(f"  truth in [{lo:4d},{hi:4d}): n={m.sum():7d} ours-truth max {eo[m].max():.2e} rms {np.sqrt((eo[m]**2).mean()):.2e} | "
                      f"ref-truth max {er[m].max():.2e} rms {np.sqrt((er[m]**2).mean()):.2e} | ours-ref max {d[m].max():.2e} frac>1e-3 {(d[m]>1e-3).mean():.2e}")
