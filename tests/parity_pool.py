"""Compare EVERY utterance of a large batch with the CPU oracle, on all host cores.

The oracle runs ~1.8 k audio-seconds per second and core, so the 1 000-utterance headline batch (5 459 audio-s) is a
few seconds of CPU work and one rank's shard of the 100 h corpus (45 000 audio-s) well under a minute on the GPU
box's cores.  Workers are SPAWNED (never forked: the parent holds a CUDA context and OpenMP threads) and map the
inputs / outputs from .npy files in a temporary directory.  Test infrastructure: imports ``oracle``.
"""

from __future__ import annotations

import multiprocessing as mp
import os
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ATOL_LOG = 1e-3  # north_star: log-mel / energy within max-abs 1e-3 of the reference

_W = {}


def log_spec_mismatch(ours, ref, truth, spec_type: str):
    """None if ``ours`` meets the parity bar against ``ref`` (both torch [F, T]), else a description.
    mel / mel-librosa: max-abs <= 1e-3.  ``linear``: the weak-bin criterion documented at
    tests/test_gpu_parity.py::assert_log_spec_close (strong bins 1e-3; weak bins no worse than the reference
    is against the float64 truth, or within 16 float32 eps of the frame's RMS amplitude)."""
    import torch

    if tuple(ours.shape) != tuple(ref.shape):
        return f"shape {tuple(ours.shape)} != {tuple(ref.shape)}"
    d = (ours - ref).abs()
    if spec_type != "linear" or truth is None:
        m = float(d.max()) if d.numel() else 0.0
        return None if m <= ATOL_LOG else f"max |d| {m:.3e}"
    t = torch.from_numpy(truth).to(torch.float64)
    strong = t >= -3.0
    if bool(strong.any()) and float(d[strong].max()) > ATOL_LOG:
        return f"strong bins: max |d| {float(d[strong].max()):.3e}"
    ours_err = (ours.double() - t).abs()
    ref_err = (ref.double() - t).abs()
    if float(d.max()) > max(1e-2, 2.5 * float(ref_err.max())):
        return f"weak bins: max |d| {float(d.max()):.3e} (reference vs truth {float(ref_err.max()):.3e})"
    if float((d > ATOL_LOG).float().mean()) >= 1e-2:
        return f"{float((d > ATOL_LOG).float().mean()):.2%} of the bins beyond 1e-3"
    if float(ours_err.pow(2).mean().sqrt()) > 1.5 * float(ref_err.pow(2).mean().sqrt()) + 1e-5:
        return "RMS error against the float64 truth worse than 1.5x the reference's"
    if float(ours_err.max()) > max(3.0 * float(ref_err.max()), ATOL_LOG):
        # a max over ~1e5 bins is an extreme-value statistic: before calling it a failure, look at the offending bins in
        # the amplitude domain.  A float32 FFT leaves an absolute amplitude error of a few eps of the frame's RMS
        # amplitude in EVERY bin; on a bin far below the frame's level that alone moves the log-power by
        # 2 * error / amplitude.  Accept bins whose error is below 16 eps of the frame RMS amplitude.
        P = torch.exp(t)
        eps_units = ours_err * (P / P.mean(dim=0, keepdim=True)).sqrt() / (2 * 1.19e-7)
        bad = (ours_err > max(3.0 * float(ref_err.max()), ATOL_LOG)) & (eps_units > 16.0)
        if bool(bad.any()):
            k = int((ours_err * bad).argmax())
            return (f"max error against the float64 truth {float(ours_err.flatten()[k]):.3e} vs the reference's "
                    f"{float(ref_err.max()):.3e} (log-power {float(t.flatten()[k]):.2f}, "
                    f"{float(eps_units.flatten()[k]):.1f} eps of the frame RMS amplitude)")
    return None


def _init(tmp, cfg, spec_type, with_phones, want_phone=False):
    import torch

    for p in (str(ROOT), str(ROOT / "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from oracle import ev_oracle as O

    if mp.current_process().name != "MainProcess":
        torch.set_num_threads(1)
    sr, n_fft, win, hop, n_mels, f_min, f_max = cfg
    _W.clear()
    _W.update(O=O, cfg=cfg, st=spec_type, hop=hop,
              tf=O.get_spectral_transform(spec_type, n_fft, win, hop, sr, n_mels, f_min, f_max),
              x=np.load(os.path.join(tmp, "x.npy"), mmap_mode="r"), off=np.load(os.path.join(tmp, "off.npy")),
              spec=np.load(os.path.join(tmp, "spec.npy"), mmap_mode="r"),
              energy=np.load(os.path.join(tmp, "energy.npy"), mmap_mode="r"), f_off=np.load(os.path.join(tmp, "f_off.npy")))
    _W["want_phone"] = want_phone
    if with_phones:
        _W.update(dur=np.load(os.path.join(tmp, "dur.npy")), p_off=np.load(os.path.join(tmp, "p_off.npy")),
                  phone=np.load(os.path.join(tmp, "phone.npy")))


def _check(rng):
    import torch

    O, hop, st = _W["O"], _W["hop"], _W["st"]
    sr, n_fft, win, _, n_mels, f_min, f_max = _W["cfg"]
    out = []
    for b in range(*rng):
        xb = np.array(_W["x"][_W["off"][b]:_W["off"][b + 1]])
        if xb.dtype == np.int16:
            xb = xb.astype(np.float32) / np.float32(32768.0)
        dur = torch.from_numpy(np.array(_W["dur"][_W["p_off"][b]:_W["p_off"][b + 1]])) if "dur" in _W else None
        o_spec, o_energy, o_phone = O.features_one(torch.from_numpy(xb), _W["tf"], hop, dur)
        f0, f1 = int(_W["f_off"][b]), int(_W["f_off"][b + 1])
        ours = torch.from_numpy(np.array(_W["spec"][f0:f1])).transpose(0, 1)
        truth = O.truth_features(xb, st, n_fft, win, hop, sr, n_mels, f_min, f_max)[0] if st == "linear" else None
        err = log_spec_mismatch(ours, o_spec, truth, st)
        e = torch.from_numpy(np.array(_W["energy"][f0:f1]))
        d_e = float((e - o_energy).abs().max()) if f1 > f0 else 0.0
        d_s = float((ours - o_spec).abs().max()) if err is None and f1 > f0 else float("nan")
        bar_e = ATOL_LOG if st != "linear" else 2e-2   # linear: 513+ weak-bin differences add up in the norm
        if err is None and d_e > bar_e:
            err = f"energy max |d| {d_e:.3e}"
        d_p, nan_ok = 0.0, True
        if dur is not None:
            got = _W["phone"][_W["p_off"][b]:_W["p_off"][b + 1]]
            want = o_phone.numpy()
            nan_ok = bool(np.array_equal(np.isnan(got), np.isnan(want)))
            m = ~np.isnan(want)
            d_p = float(np.abs(got[m] - want[m]).max()) if m.any() else 0.0
            if err is None and (not nan_ok or d_p > bar_e):
                err = f"phone averages: NaN positions equal {nan_ok}, max |d| {d_p:.3e}"
        out.append((b, err, d_s, d_e, d_p, f1 - f0, o_phone.numpy() if (dur is not None and _W.get('want_phone')) else None))
    return out


def compare_all(x, off, spec, energy, f_off, cfg, spec_type, durations=None, phone_off=None, phone=None, workers=None,
                utterances=None, return_oracle_phone=False):
    """``x`` packed samples (float32 / int16 numpy), ``spec [frames, F]``, ``energy [frames]`` (numpy, ours).  Runs the
    oracle over every utterance (or ``utterances``) and returns ``dict(failures=[(b, why)], max_spec=, max_energy=,
    max_phone=, frames=, utterances=)``."""
    n = len(off) - 1
    idx = list(range(n)) if utterances is None else list(utterances)
    workers = workers or max(1, min(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1), 32))
    with tempfile.TemporaryDirectory(prefix="evf_parity_") as tmp:
        for name, a in (("x", x), ("off", np.asarray(off, np.int64)), ("spec", spec), ("energy", energy),
                        ("f_off", np.asarray(f_off, np.int64))):
            np.save(os.path.join(tmp, name + ".npy"), np.ascontiguousarray(a))
        with_phones = durations is not None
        if with_phones:
            np.save(os.path.join(tmp, "dur.npy"), np.asarray(durations, np.int64))
            np.save(os.path.join(tmp, "p_off.npy"), np.asarray(phone_off, np.int64))
            np.save(os.path.join(tmp, "phone.npy"), np.asarray(phone, np.float32))
        if utterances is None:
            per = max(1, n // (workers * 8))
            chunks = [(a, min(a + per, n)) for a in range(0, n, per)]
        else:
            chunks = [(b, b + 1) for b in idx]
        if workers == 1:
            _init(tmp, cfg, spec_type, with_phones, return_oracle_phone)
            rows = [r for c in chunks for r in _check(c)]
        else:
            with mp.get_context("spawn").Pool(workers, initializer=_init, initargs=(tmp, cfg, spec_type, with_phones, return_oracle_phone)) as pool:
                rows = [r for part in pool.map(_check, chunks) for r in part]
    ok = [r for r in rows if r[1] is None]
    rows.sort(key=lambda r: r[0])
    return {
        "oracle_phone": [r[6] for r in rows] if return_oracle_phone else None,
        "failures": [(r[0], r[1]) for r in rows if r[1] is not None],
        "max_spec": max((r[2] for r in ok), default=0.0), "max_energy": max((r[3] for r in rows), default=0.0),
        "max_phone": max((r[4] for r in rows), default=0.0), "frames": int(sum(r[5] for r in rows)),
        "utterances": len(rows), "workers": workers,
    }
