"""CPU property tests of the host-side integer arithmetic the device code relies on: shard balance and coverage,
resampled / kept lengths against the reference's formulas, loudness scratch sizing, frame counts."""
import math

import numpy as np
import pytest

from everyvoice_b200 import _lib, synth
from everyvoice_b200.distributed import shard_utterances


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_shards_cover_every_utterance_once_and_balance(world):
    rng = np.random.default_rng(world)
    for n in (0, 1, world - 1, world, 7 * world + 3, 5000):
        lens = (rng.integers(1, 900, size=max(n, 0)) * 256).astype(np.int64)
        shards = shard_utterances(lens, world)
        assert len(shards) == world
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(n))
        if n >= 50 * world:
            loads = np.array([lens[s].sum() for s in shards], dtype=np.float64)
            assert loads.max() - loads.min() <= lens.max()          # greedy longest-first: within one utterance
    assert shard_utterances(np.array([5, 5, 5]), 8).count([]) == 5   # more ranks than utterances: empty shards


def test_resampled_and_kept_lengths_match_the_reference_formulas():
    """torchaudio crops to ceil(new * L / orig) computed in floating point (functional.py), process_audio keeps
    (L' // hop) * hop (preprocessor.py:216-218); the library does both in integers."""
    lib = _lib.load()
    rng = np.random.default_rng(3)
    import ctypes as C

    for orig, new in ((44100, 22050), (48000, 22050), (16000, 22050), (22050, 44100), (8000, 48000), (44100, 16000)):
        g = math.gcd(orig, new)
        o, n = orig // g, new // g
        for L in [0, 1, 2, o - 1, o, o + 1, 12345, 2**31 - 7] + rng.integers(1, 10**7, size=200).tolist():
            want = int(math.ceil(n * L / o))                         # the reference's float formula
            got = (L * n + o - 1) // o                               # evf_resampler_out_length / Resampler.out_lengths
            assert got == want, (orig, new, L)
    # the C entry point itself needs a device to create a resampler; its formula is the line above (evfeat_audio.cu)
    assert lib.evf_resampler_out_length(C.c_void_p(0), 100) == -1


@pytest.mark.parametrize("sr", [8000, 11025, 16000, 22050, 24000, 32000, 44099, 44100, 48000])
def test_loudness_block_arithmetic(sr):
    """torchaudio.functional.loudness: gate = round(0.4 sr), step = round(0.25 gate) (Python round-half-even); the
    kernels keep one 5-float record per step (sum, first two, last two squares), a block is four steps plus / minus
    gate - 4 * step in [-2, 2] samples, and a block reads up to 4 records past its index."""
    lib = _lib.load()
    gate = int(round(0.4 * sr))
    step = int(round(gate * (1 - 0.75)))
    assert -2 <= gate - 4 * step <= 2 and lib.evf_audio_loudness_step(sr) == step
    for n in (0, 1, step - 1, step, gate - 1, gate, 10 * sr + 17):
        assert lib.evf_audio_loudness_scratch_floats(sr, n) == 5 * (n // step + 5)
        n_blk = (n - gate) // step + 1 if n >= gate else 0          # unfold(-1, gate, step)
        assert n_blk + 4 <= n // step + 5
    assert lib.evf_audio_loudness_scratch_floats(0, 10) == -1


def test_frame_counts_and_synthetic_workload_shapes():
    """T = L // hop (process_spec, preprocessor.py:921) for the bench workload; the durations generator produces
    P = max(1, T // 7) phones whose sum is T or T +- {1, 2}."""
    lens = synth.utterance_lengths(1000, 22050, 256, 1234)
    assert (lens % 256 == 0).all() and lens.min() >= 22050 // 256 * 256 - 256 and lens.max() <= 10 * 22050
    assert int((lens // 256).sum()) == 470190                       # the number every bench line quotes
    for i, L in enumerate(lens[:50]):
        T = int(L) // 256
        d = synth.synthetic_durations(T, seed=i)
        assert len(d) == max(1, T // 7) and (d >= 0).all() and abs(int(d.sum()) - T) <= 2
