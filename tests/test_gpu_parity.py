"""Parity of the CUDA path (through the C ABI) against the reference.

* golden vectors produced by the LIVE reference (tests/golden, oracle/make_golden.py);
* the CPU oracle (oracle/ev_oracle.py, bit-identical to the reference on CPU) on seeded
  ragged batches at sizes it finishes in seconds.

Tolerances (BASELINE.json north_star): frame counts / offsets / indexing bit-exact;
log-spectrogram and energy max-abs <= 1e-3.
"""

import numpy as np
import pytest
import torch

from conftest import CONFIGS, SPEC_TYPES, golden_inputs

pytestmark = pytest.mark.gpu

ATOL_LOG = 1e-3  # north_star: log-mel / energy within max-abs 1e-3 of the reference


def assert_log_spec_close(ours: torch.Tensor, ref: torch.Tensor, truth: np.ndarray | None, spec_type: str):
    """max-abs <= 1e-3 against the reference (north_star tolerance for log-mel / energy).

    For the 513/1025-bin ``linear`` type only: weak bins (log-power below about -3) of a frame
    whose strongest bins are ~1e9 stronger are at the fp32 FFT round-off floor.  Measured on the
    B200 (tests/tools/err_stats.py): against the exact float64 value the
    reference's own CPU fp32 result is off by up to 7.4e-3 there and the CUDA path by up to
    6.8e-3, with equal RMS error -- two correct fp32 FFTs cannot agree to 1e-3 on those bins.
    The bar for ``linear`` is therefore:
      * every bin with exact log-power >= -3: within 1e-3 of the reference;
      * weaker bins: within 1e-2 of the reference -- or, for the larger transforms whose round-off floor is higher
        (n_fft 3072 / 4096: the reference itself is 1.0e-2 from the truth there, profiles/r02c_err_stats_wide.txt),
        within 2.5x the reference's own worst error -- and fewer than 1 % beyond 1e-3;
      * accuracy no worse than the reference's: RMS error against the float64 truth
        <= 1.5x the reference's + 1e-5 (1 % of the tolerance: when every bin is strong both
        implementations sit at 1e-6 and the ratio of two round-off levels means nothing),
        max error <= max(3x the reference's, 1e-3), or -- the max over ~1e5 bins being an extreme-value statistic -- the
        offending bin's error in the amplitude domain below 16 float32 eps of the frame's RMS amplitude."""
    from parity_pool import log_spec_mismatch   # one statement of the criterion, shared with the every-utterance checker

    why = log_spec_mismatch(ours, ref, truth, spec_type)
    assert why is None, why


def _transform(config, spec_type):
    import everyvoice_b200 as ev

    sr, n_fft, win, hop, n_mels, f_min, f_max = CONFIGS[config]
    return ev.get_spectral_transform(spec_type, n_fft, win, hop, sr, n_mels, f_min, f_max), hop


def _oracle_transform(config, spec_type):
    from oracle import ev_oracle as O

    sr, n_fft, win, hop, n_mels, f_min, f_max = CONFIGS[config]
    return O.get_spectral_transform(spec_type, n_fft, win, hop, sr, n_mels, f_min, f_max), hop


# ------------------------------------------------------------------------------------------
# golden vectors from the live reference
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("config", list(CONFIGS))
@pytest.mark.parametrize("spec_type", ["mel", "mel-librosa", "linear"])
def test_golden_log_spec_and_energy(cuda_device, golden_dir, config, spec_type):
    import everyvoice_b200 as ev
    from oracle import ev_oracle as O

    gold = np.load(golden_dir / f"spectral_{config}.npz")
    tf, hop = _transform(config, spec_type)
    pre = ev.Preprocessor(device=cuda_device)
    n_checked = 0
    for name, x in golden_inputs(config).items():
        key = f"{spec_type}/{name}/spec"
        if key not in gold:
            continue
        xt = torch.from_numpy(x)
        T = len(x) // hop
        # exactly what process_spec does (preprocessor.py:921-927), on a CPU tensor like the reference
        spec = pre.extract_spectral_features(xt, tf)[:, :T]
        assert spec.device.type == "cpu"
        ref = torch.from_numpy(gold[key])
        assert tuple(spec.shape) == tuple(ref.shape)  # bit-exact frame count
        sr, n_fft, win, _, n_mels, f_min, f_max = CONFIGS[config]
        truth = O.truth_features(x, spec_type, n_fft, win, hop, sr, n_mels, f_min, f_max)[0] if spec_type == "linear" else None
        assert_log_spec_close(spec, ref, truth, spec_type)
        energy = pre.extract_energy(spec)
        ref_e = torch.from_numpy(gold[f"{spec_type}/{name}/energy"])
        assert tuple(energy.shape) == tuple(ref_e.shape)
        assert float((energy - ref_e).abs().max()) <= ATOL_LOG
        # the bare transform returns T + 1 linear-domain frames; check the one process_spec drops
        lin = tf(xt)
        assert lin.shape[-1] == 1 + (len(x) + 2 * (n_fft // 2) - n_fft) // hop   # T + 1 for even n_fft
        ref_last = torch.from_numpy(gold[f"{spec_type}/{name}/lin_last"])
        log_last = torch.log(torch.clamp(lin[:, -1], min=1e-5))
        truth_last = None
        if spec_type == "linear":
            # same criterion as the kept frames: weak bins of a linear spectrogram sit in the reference's own fp32 noise
            truth_last = np.log(np.maximum(O.truth_power_spectrogram(x, n_fft, win, hop)[:, lin.shape[-1] - 1], 1e-5))
        assert_log_spec_close(log_last, torch.log(torch.clamp(ref_last, min=1e-5)), truth_last, spec_type)
        n_checked += 1
    assert n_checked >= 1 or (config == "Bfull" and spec_type == "linear")


@pytest.mark.parametrize("config", [c for c in CONFIGS if c != "Bfull"])
def test_golden_raw_complex(cuda_device, golden_dir, config):
    gold = np.load(golden_dir / f"spectral_{config}.npz")
    tf, hop = _transform(config, "raw")
    n = 0
    for name, x in golden_inputs(config).items():
        if f"raw/{name}/re" not in gold:
            continue
        out = tf(torch.from_numpy(x).to(cuda_device)).cpu()
        ref = torch.complex(torch.from_numpy(gold[f"raw/{name}/re"]), torch.from_numpy(gold[f"raw/{name}/im"]))
        assert out.dtype == torch.complex64 and tuple(out.shape) == tuple(ref.shape)
        scale = float(ref.abs().max())
        # complex STFT: absolute error relative to the largest bin (fp32 FFT round-off)
        assert float((out - ref).abs().max()) <= 2e-6 * scale + 1e-5
        n += 1
    assert n >= 1


# ------------------------------------------------------------------------------------------
# oracle on seeded ragged batches
# ------------------------------------------------------------------------------------------
def _ragged_inputs(sr, hop, n_utts, seed, max_s=3.0):
    from everyvoice_b200 import synth

    lens = synth.utterance_lengths(n_utts, sr, hop, seed, 0.3, max_s)
    lens[0] += 13  # one length that is not a multiple of hop
    xs = []
    for i, L in enumerate(lens):
        xs.append(synth.white_noise(int(L), seed * 1000 + i) if i % 2 == 0 else synth.speech_like(int(L), sr, seed * 1000 + i))
    return xs


@pytest.mark.parametrize("config", list(CONFIGS))
@pytest.mark.parametrize("spec_type", ["mel", "mel-librosa", "linear"])
def test_ragged_batch_matches_oracle(cuda_device, config, spec_type):
    import everyvoice_b200 as ev
    from everyvoice_b200 import synth
    from oracle import ev_oracle as O

    sr, n_fft, win, hop, n_mels, f_min, f_max = CONFIGS[config]
    xs = _ragged_inputs(sr, hop, 14, seed=7)
    packed, offsets = synth.pack_ragged(xs)
    tf, _ = _transform(config, spec_type)
    feats = tf.to(cuda_device).features_ragged(torch.from_numpy(packed).to(cuda_device), offsets)
    otf, _ = _oracle_transform(config, spec_type)
    # frame counts / offsets: bit-exact
    exp_T = np.array([len(x) // hop for x in xs], dtype=np.int64)
    assert np.array_equal(np.diff(feats.frame_offsets), exp_T)
    worst_e = 0.0
    for b, x in enumerate(xs):
        o_spec, o_energy, _ = O.features_one(torch.from_numpy(x), otf, hop)
        spec = feats.utterance(b).cpu()
        energy = feats.utterance_energy(b).cpu()
        truth = O.truth_features(x, spec_type, n_fft, win, hop, sr, n_mels, f_min, f_max)[0] if spec_type == "linear" else None
        assert_log_spec_close(spec, o_spec, truth, spec_type)
        worst_e = max(worst_e, float((energy - o_energy).abs().max()))
    assert worst_e <= ATOL_LOG, worst_e


@pytest.mark.parametrize("config", ["A", "B", "W", "ST2", "S512", "W512", "S256", "W256", "R3", "W3072", "H4096"])
@pytest.mark.parametrize("spec_type", SPEC_TYPES)
def test_warp_kernel_and_any_size_kernel_agree(cuda_device, config, spec_type):
    """Two independent FFT implementations behind one plan interface (evf_fft_path): the warp-per-FFT kernel and the
    shared-memory mixed-radix kernel compute the same transform on the sizes both cover."""
    import everyvoice_b200 as ev
    from everyvoice_b200 import synth

    sr, n_fft, win, hop, n_mels, f_min, f_max = CONFIGS[config]
    xs = _ragged_inputs(sr, hop, 9, seed=19, max_s=1.5)
    packed, offsets = synth.pack_ragged(xs)
    x = torch.from_numpy(packed).to(cuda_device)
    fast = ev.SpectralTransform(spec_type, n_fft, win, hop, sr, n_mels, f_min, f_max)
    slow = ev.SpectralTransform(spec_type, n_fft, win, hop, sr, n_mels, f_min, f_max, fft_path="generic")
    a, b = fast.features_ragged(x, offsets), slow.features_ragged(x, offsets)
    assert np.array_equal(a.frame_offsets, b.frame_offsets)
    if spec_type == "raw":
        scale = float(a.spec.abs().max())
        assert float((a.spec - b.spec).abs().max()) <= 4e-6 * scale + 1e-5
    elif spec_type == "linear":   # weak bins sit at the float32 round-off floor of either FFT: compare powers
        pa, pb = torch.exp(a.spec), torch.exp(b.spec)
        assert float((pa - pb).abs().max()) <= 2e-5 * float(pa.max())
        assert float((a.energy - b.energy).abs().max()) <= 2e-2
    else:
        assert float((a.spec - b.spec).abs().max()) <= ATOL_LOG
        assert float((a.energy - b.energy).abs().max()) <= ATOL_LOG


def test_dense_projection_for_a_bank_that_is_not_triangular(cuda_device):
    """Any [n_freq, n_mels] basis is accepted: overlapping / negative / random filters take the dense projection of
    the any-size kernel (a triangular bank takes the sparse one); both against a float64 matmul of the power
    spectrogram."""
    import everyvoice_b200 as ev
    from everyvoice_b200 import synth
    from oracle import ev_oracle as O

    sr, n_fft, win, hop = 22050, 1024, 1024, 256
    rng = np.random.default_rng(77)
    fb = rng.uniform(-0.2, 1.0, size=(n_fft // 2 + 1, 24)).astype(np.float32)
    x = synth.speech_like(60 * hop, sr, seed=5)
    P = O.truth_power_spectrogram(x, n_fft, win, hop)                      # [513, T + 1] float64
    want = fb.astype(np.float64).T @ P
    tf = ev.SpectralTransform("mel", n_fft, win, hop, sr, mel_fb=fb)
    got = tf(torch.from_numpy(x).to(cuda_device)).cpu().double().numpy()   # linear domain, T + 1 frames
    assert got.shape == want.shape == (24, 61)
    assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max()
    # and a triangular bank handed over explicitly equals the built-in one bit for bit
    tri = ev.SpectralTransform("mel", n_fft, win, hop, sr, 80, 0, 8000)
    same = ev.SpectralTransform("mel", n_fft, win, hop, sr, mel_fb=tri.mel_fb)
    xd = torch.from_numpy(x).to(cuda_device)
    assert torch.equal(tri(xd), same(xd))


def test_misaligned_sample_buffer_view(cuda_device):
    """A contiguous view that does not start on a 16-byte boundary (wav[1:], pcm[3:]) keeps its odd data_ptr: the
    bulk-copy staging must notice and fall back to element loads (advisor finding, round 1)."""
    from everyvoice_b200 import synth

    tf, hop = _transform("A", "mel")
    x = torch.from_numpy(synth.speech_like(40 * hop + 8, 22050, seed=9)).to(cuda_device)
    pcm = (x * 30000).round().to(torch.int16)
    for buf, shifts in ((x, (1, 2, 3)), (pcm, (1, 3, 5, 7))):
        for s in shifts:
            view = buf[s:]
            assert view.data_ptr() % 16 != 0
            L = view.numel()
            got = tf.features_ragged(view, np.array([0, L]))
            ref = tf.features_ragged(view.clone(), np.array([0, L]))     # a fresh, aligned allocation
            assert torch.equal(got.spec, ref.spec) and torch.equal(got.energy, ref.energy)


def test_int16_input_equals_float_input(cuda_device):
    from everyvoice_b200 import synth

    tf, hop = _transform("A", "mel")
    rng = np.random.default_rng(5)
    lens = synth.utterance_lengths(6, 22050, hop, 11, 0.5, 2.0)
    pcm = [rng.integers(-30000, 30000, size=int(L)).astype(np.int16) for L in lens]
    packed, offsets = synth.pack_ragged(pcm)
    f_i16 = tf.features_ragged(torch.from_numpy(packed).to(cuda_device), offsets)
    f_f32 = tf.features_ragged(torch.from_numpy(packed.astype(np.float32) / 32768.0).to(cuda_device), offsets)
    assert torch.equal(f_i16.spec, f_f32.spec) and torch.equal(f_i16.energy, f_f32.energy)


def test_batched_equals_single_bitwise(cuda_device):
    """An utterance's features do not depend on what else is in the batch."""
    from everyvoice_b200 import synth

    tf, hop = _transform("A", "mel")
    xs = _ragged_inputs(22050, hop, 9, seed=3)
    packed, offsets = synth.pack_ragged(xs)
    feats = tf.features_ragged(torch.from_numpy(packed).to(cuda_device), offsets)
    for b in (0, 4, 8):
        single = tf.features_ragged(torch.from_numpy(xs[b]).to(cuda_device), np.array([0, len(xs[b])]))
        assert torch.equal(single.spec, feats.spec[feats.frame_offsets[b] : feats.frame_offsets[b + 1]])
        assert torch.equal(single.energy, feats.utterance_energy(b))


def test_leading_batch_dims_and_device_round_trip(cuda_device):
    from oracle import ev_oracle as O

    tf, hop = _transform("A", "mel")
    otf, _ = _oracle_transform("A", "mel")
    x = torch.from_numpy(np.random.default_rng(0).uniform(-0.9, 0.9, size=(2, 3, 5000)).astype(np.float32))
    y = tf(x)  # CPU in -> CPU out, like the reference's transform
    ref = otf(x)
    assert y.device.type == "cpu" and tuple(y.shape) == tuple(ref.shape) == (2, 3, 80, 5000 // hop + 1)
    assert float((torch.log(y.clamp(min=1e-5)) - torch.log(ref.clamp(min=1e-5))).abs().max()) <= ATOL_LOG
    yc = tf(x.to(cuda_device))
    assert yc.is_cuda and torch.equal(yc.cpu(), y)


def test_dynamic_range_compression_operator(cuda_device):
    import everyvoice_b200 as ev
    from oracle import ev_oracle as O

    x = torch.from_numpy(np.random.default_rng(1).uniform(0, 50, size=(80, 77)).astype(np.float32))
    x[3, 5] = 0.0
    x[4, 6] = 1e-7
    out = ev.dynamic_range_compression_torch(x)
    ref = O.dynamic_range_compression_torch(x)
    assert float((out - ref).abs().max()) <= 1e-6
    assert float(out[3, 5]) == pytest.approx(np.log(1e-5), abs=1e-6)


# ------------------------------------------------------------------------------------------
# phone-level averaging
# ------------------------------------------------------------------------------------------
def _same_with_nans(a: torch.Tensor, b: torch.Tensor, atol):
    assert tuple(a.shape) == tuple(b.shape)
    na, nb = torch.isnan(a), torch.isnan(b)
    assert torch.equal(na, nb), "NaN positions differ"
    assert float((a[~na] - b[~nb]).abs().max()) <= atol if (~na).any() else True


def test_average_by_durations_golden(cuda_device, golden_dir):
    import everyvoice_b200 as ev

    gold = np.load(golden_dir / "average_by_durations.npz")
    pre = ev.Preprocessor(device=cuda_device)
    names = sorted({k.split("/")[0] for k in gold.files})
    assert len(names) >= 9
    for name in names:
        vals = torch.from_numpy(gold[f"{name}/values"])
        durs = torch.from_numpy(gold[f"{name}/durations"])
        out = pre.average_data_by_durations(vals, durs)
        ref = torch.from_numpy(gold[f"{name}/out"])
        assert out.dtype == torch.float32 and out.device.type == "cpu"
        scale = max(1.0, float(vals.abs().max()))
        _same_with_nans(out, ref, atol=2e-6 * scale)  # fp32 summation-order noise only
        # d <= 0 entries are exactly the reference's 1e-7
        assert torch.equal(out[durs <= 0], torch.full((int((durs <= 0).sum()),), 1e-7))


def test_average_by_durations_ragged_matches_oracle(cuda_device):
    import everyvoice_b200 as ev
    from everyvoice_b200 import synth
    from oracle import ev_oracle as O

    rng = np.random.default_rng(17)
    Ts = [int(t) for t in rng.integers(30, 900, size=40)] + [1, 2, 700]
    vals = [rng.uniform(80, 300, size=T).astype(np.float32) for T in Ts]
    durs = [synth.synthetic_durations(T, seed=100 + i) for i, T in enumerate(Ts)]
    durs[-1] = np.array([700], dtype=np.int64)  # one phone spanning the whole utterance (cooperative path)
    durs[5] = np.concatenate([durs[5], [9, 9]])  # overrun: clipped, then empty -> NaN
    v_packed, v_off = synth.pack_ragged(vals)
    d_packed, p_off = synth.pack_ragged(durs)
    pre = ev.Preprocessor(device=cuda_device)
    out = pre.average_data_by_durations_ragged(torch.from_numpy(v_packed).to(cuda_device), v_off,
                                               torch.from_numpy(d_packed.astype(np.int64)), p_off).cpu()
    assert out.numel() == int(p_off[-1])
    for b in range(len(Ts)):
        ref = O.average_data_by_durations(torch.from_numpy(vals[b]), torch.from_numpy(durs[b].astype(np.int64)))
        _same_with_nans(out[p_off[b] : p_off[b + 1]], ref, atol=1e-3)


def test_energy_phone_pipeline_matches_oracle(cuda_device):
    """process_spec -> process_energy with phone-level averaging, batched, vs the oracle loop."""
    import everyvoice_b200 as ev
    from everyvoice_b200 import synth
    from oracle import ev_oracle as O

    cfg = ev.AudioConfig(spec_type="linear")
    pre = ev.Preprocessor(cfg, device=cuda_device)
    hop, sr = cfg.fft_hop_size, cfg.input_sampling_rate
    xs = _ragged_inputs(sr, hop, 10, seed=23, max_s=2.0)
    durs = [synth.synthetic_durations(len(x) // hop, seed=300 + i) for i, x in enumerate(xs)]
    feats = pre.process_spec_batch([torch.from_numpy(x) for x in xs])
    phone, p_off = pre.process_energy_batch(feats, [torch.from_numpy(d) for d in durs])
    otf = O.get_spectral_transform("linear", cfg.n_fft, cfg.fft_window_size, hop, sr, cfg.n_mels, cfg.f_min, cfg.f_max)
    for b, x in enumerate(xs):
        _, _, o_phone = O.features_one(torch.from_numpy(x), otf, hop, torch.from_numpy(durs[b]))
        _same_with_nans(phone[p_off[b] : p_off[b + 1]].cpu(), o_phone, atol=ATOL_LOG)


# ------------------------------------------------------------------------------------------
# statistics / normalisation
# ------------------------------------------------------------------------------------------
def test_scaler_golden(cuda_device, golden_dir):
    import everyvoice_b200 as ev

    gold = np.load(golden_dir / "scaler.npz")
    s = ev.Scaler(device=cuda_device)
    chunks = [torch.from_numpy(gold[f"chunk{i}"]) for i in range(7)]
    for c in chunks:
        s.append(c)
    stats = s.calculate_stats(distributed=False)
    assert stats["sample_size"] == int(gold["stats/sample_size"])
    for k in ("min", "max"):
        assert stats[k] == float(gold[f"stats/{k}"])  # exact: min / max are selections
    for k in ("mean", "std", "norm_min", "norm_max"):
        assert stats[k] == pytest.approx(float(gold[f"stats/{k}"]), rel=2e-6, abs=1e-6)
    for i, c in enumerate(chunks):
        out = s.normalize(c)
        ref = torch.from_numpy(gold[f"norm{i}"])
        _same_with_nans(out, ref, atol=1e-5)


def test_scaler_matches_oracle_on_device_shard(cuda_device):
    import everyvoice_b200 as ev
    from oracle import ev_oracle as O

    rng = np.random.default_rng(9)
    x = rng.normal(180.0, 40.0, size=200_003).astype(np.float32)
    x[rng.integers(0, x.size, size=50)] = np.nan
    pre = ev.Preprocessor(device=cuda_device)
    xd = torch.from_numpy(x).to(cuda_device)
    e_scaler, _ = pre.compute_stats(energy=xd, n_energy_files=123)
    stats = pre.normalize_stats(e_scaler, None, distributed=False)["energy"]
    o = O.Scaler()
    o.append(torch.from_numpy(x))
    ref = o.calculate_stats()
    assert stats["sample_size"] == 123
    for k in ("min", "max"):
        assert stats[k] == ref[k]
    for k in ("mean", "std", "norm_min", "norm_max"):
        assert stats[k] == pytest.approx(ref[k], rel=5e-6)
    _same_with_nans(xd.cpu(), o.normalize(torch.from_numpy(x)), atol=1e-4)  # normalised in place


def test_gathered_stats_merge_and_normalise_on_device(cuda_device):
    """The N > 1 exchange step as the ranks see it after the all-gather: [world, 5] summaries are
    merged and applied on the device; result == the reference's Scaler over the whole corpus."""
    import ctypes as C

    import everyvoice_b200 as ev
    from everyvoice_b200 import _lib
    from everyvoice_b200.distributed import finalize_stats, merge_stats
    from oracle import ev_oracle as O

    rng = np.random.default_rng(21)
    shards = [rng.normal(25.0, 6.0, size=n).astype(np.float32) for n in (5003, 7001, 1, 2999)]
    shards[1][10] = np.nan
    parts = []
    for sh in shards:
        s = ev.Scaler(cuda_device)
        s.append(torch.from_numpy(sh).to(cuda_device))
        parts.append(s.partial_stats().clone())
    gathered = torch.stack(parts).contiguous()  # what allgather_stats returns on every rank
    merged = torch.empty(5, dtype=torch.float64, device=cuda_device)
    lib = _lib.load()
    _lib.check(lib.evf_stats_merge(C.c_void_p(gathered.data_ptr()), 4, 5, C.c_void_p(merged.data_ptr()), None))
    torch.cuda.synchronize()
    assert torch.allclose(merged.cpu(), merge_stats(gathered).cpu(), rtol=1e-15, atol=0)
    o = O.Scaler()
    for sh in shards:
        o.append(torch.from_numpy(sh))
    ref = o.calculate_stats()
    st = finalize_stats(merged.cpu().tolist(), len(shards))
    assert st["min"] == ref["min"] and st["max"] == ref["max"] and st["sample_size"] == 4
    for k in ("mean", "std", "norm_min", "norm_max"):
        assert st[k] == pytest.approx(ref[k], rel=5e-6)
    mine = torch.from_numpy(shards[0]).to(cuda_device)
    ev.Scaler(cuda_device).normalize_by_device_stats_(mine, gathered)
    _same_with_nans(mine.cpu(), o.normalize(torch.from_numpy(shards[0])), atol=1e-5)
    single = torch.from_numpy(shards[0]).to(cuda_device)
    ev.Scaler(cuda_device).normalize_by_device_stats_(single, merged)
    assert torch.equal(single, mine)  # [5] and [world, 5] forms agree bit for bit


@pytest.mark.parametrize("head", [None, 25_000])
@pytest.mark.parametrize("dtype", ["f32", "s16"])
def test_corpus_pipeline_equals_single_batch(cuda_device, dtype, head):
    """The chunked three-stream host pipeline returns exactly what one resident batch does (first chunks full-sized
    -- the default -- or ramped up from ``head_chunk_bytes``)."""
    import everyvoice_b200 as ev
    from everyvoice_b200 import synth

    pre = ev.Preprocessor(ev.AudioConfig(spec_type="mel"), device=cuda_device)
    hop = 256
    lens = synth.utterance_lengths(37, 22050, hop, 5, 0.3, 2.5)
    lens[3] += 5   # every later chunk starts at an odd sample index: exercises the alignment lead-in
    lens[11] += 2
    rng = np.random.default_rng(8)
    if dtype == "s16":
        xs = [rng.integers(-30000, 30000, size=int(L)).astype(np.int16) for L in lens]
        tdt = torch.int16
    else:
        xs = [synth.white_noise(int(L), 50 + i) for i, L in enumerate(lens)]
        tdt = torch.float32
    packed, off = synth.pack_ragged(xs)
    durs = [synth.synthetic_durations(int(L) // hop, seed=70 + i) for i, L in enumerate(lens)]
    d_packed, p_off = synth.pack_ragged(durs)
    host = torch.from_numpy(packed).pin_memory()
    pipe = pre.make_corpus_pipeline(off, tdt, torch.from_numpy(d_packed.astype(np.int64)), p_off,
                                    chunk_bytes=200_000, head_chunk_bytes=head)  # forces ~10 chunks
    assert len(pipe.chunks) >= 5
    first = (pipe.chunks[0].s1 - pipe.chunks[0].s0) * pipe.esize
    assert first <= (head if head is not None else 200_000) or pipe.chunks[0].u1 == pipe.chunks[0].u0 + 1
    h_spec = torch.empty((pipe.total_frames, 80), dtype=torch.float32).pin_memory()
    h_energy = torch.empty(pipe.total_frames, dtype=torch.float32).pin_memory()
    h_phone = torch.empty(pipe.n_phones, dtype=torch.float32).pin_memory()
    for _ in range(2):  # twice: buffers and events are reused across runs
        h_spec.zero_(); h_energy.zero_(); h_phone.zero_()
        pipe.run(host, h_spec, h_energy, h_phone)
        torch.cuda.synchronize()
        feats = pre.process_spec_batch(host.to(cuda_device), off)
        phone, _ = pre.process_energy_batch(feats, torch.from_numpy(d_packed.astype(np.int64)), p_off)
        s = ev.Scaler(cuda_device)
        s.append(phone)
        s.normalize_by_device_stats_(phone, s.partial_stats())
        assert np.array_equal(pipe.frame_offsets, feats.frame_offsets)
        assert torch.equal(h_spec, feats.spec.cpu()) and torch.equal(h_energy, feats.energy.cpu())
        _same_with_nans(h_phone, phone.cpu(), atol=0.0)


@pytest.mark.parametrize("k", [3, 4])
@pytest.mark.parametrize("dtype", ["f32", "s16"])
def test_output_transform_with_a_sampling_rate_change(cuda_device, k, dtype):
    """preprocessor.py:94-121: the "output" transform of a vocoder configuration multiplies n_fft / win / hop by
    output_sr // input_sr (16 kHz input -> 48 / 64 kHz output: 3072 / 768 and 4096 / 1024) and keeps the input rate for
    the mel basis.  These sizes run as phase streams through the 1024-point kernel plus a combine
    (csrc/evfeat_decimated.cu): one resident batch against the oracle, and the chunked host pipeline (utterance
    ranges, scratch allocated per chunk on the pipeline's stream) against the resident batch, bit for bit."""
    import everyvoice_b200 as ev
    from everyvoice_b200 import synth
    from oracle import ev_oracle as O

    sr_in = 16000
    ac = ev.AudioConfig(spec_type="mel", input_sampling_rate=sr_in, output_sampling_rate=sr_in * k)
    pre = ev.Preprocessor(ac, device=cuda_device)
    tf = pre.output_spectral_transform
    assert (tf.n_fft, tf.win_length, tf.hop_length, tf.sample_rate) == (1024 * k, 1024 * k, 256 * k, sr_in)
    hop = 256 * k
    lens = synth.utterance_lengths(23, sr_in * k, hop, 9, 0.3, 1.5)
    lens[2] += 11
    if dtype == "s16":
        rng = np.random.default_rng(4)
        xs = [(synth.speech_like(int(L), sr_in * k, seed=400 + i) * 20000).astype(np.int16) for i, L in enumerate(lens)]
        tdt = torch.int16
        as_float = [x.astype(np.float32) / np.float32(32768.0) for x in xs]
    else:
        xs = [synth.speech_like(int(L), sr_in * k, seed=400 + i) for i, L in enumerate(lens)]
        tdt = torch.float32
        as_float = xs
    packed, off = synth.pack_ragged(xs)
    feats = pre.process_spec_batch(torch.from_numpy(packed).to(cuda_device), off, output=True)
    otf = O.get_spectral_transform("mel", 1024 * k, 1024 * k, hop, sr_in, ac.n_mels, ac.f_min, ac.f_max)
    assert np.array_equal(np.diff(feats.frame_offsets), lens // hop)
    for b in (0, 2, 7, 22):
        o_spec, o_energy, _ = O.features_one(torch.from_numpy(as_float[b]), otf, hop)
        assert float((feats.utterance(b).cpu() - o_spec).abs().max()) <= ATOL_LOG
        assert float((feats.utterance_energy(b).cpu() - o_energy).abs().max()) <= ATOL_LOG
    durs = [synth.synthetic_durations(int(L) // hop, seed=90 + i) for i, L in enumerate(lens)]
    d_packed, p_off = synth.pack_ragged(durs)
    host = torch.from_numpy(packed).pin_memory()
    pipe = pre.make_corpus_pipeline(off, tdt, torch.from_numpy(d_packed.astype(np.int64)), p_off, output=True,
                                    chunk_bytes=400_000)
    assert len(pipe.chunks) >= 3
    h_spec = torch.empty((pipe.total_frames, ac.n_mels), dtype=torch.float32).pin_memory()
    h_energy = torch.empty(pipe.total_frames, dtype=torch.float32).pin_memory()
    h_phone = torch.empty(pipe.n_phones, dtype=torch.float32).pin_memory()
    pipe.run(host, h_spec, h_energy, h_phone)
    torch.cuda.synchronize()
    assert torch.equal(h_spec, feats.spec.cpu()) and torch.equal(h_energy, feats.energy.cpu())


# ------------------------------------------------------------------------------------------
# reference test-suite invariants (everyvoice/tests/test_preprocessing.py:385-435, 496-568)
# ------------------------------------------------------------------------------------------
def test_reference_shape_invariants(cuda_device, golden_dir):
    import everyvoice_b200 as ev

    lj = torch.from_numpy(np.load(golden_dir / "lj_excerpt_int16.npy").astype(np.float32) / 32768.0)
    durs = np.load(golden_dir / "lj_durations.npz")
    frames = {}
    for st in SPEC_TYPES:
        pre = ev.Preprocessor(ev.AudioConfig(spec_type=st), device=cuda_device)
        feats = pre.extract_spectral_features(lj, pre.input_spectral_transform)
        frames[st] = feats.size(1)
        if st in ("mel", "mel-librosa"):
            assert feats.size(0) == pre.audio_config.n_mels
        else:
            assert feats.size(0) == pre.audio_config.n_fft // 2 + 1
        if st != "raw":
            energy = pre.extract_energy(feats)
            assert energy.size(0) == feats.size(1)
            d = torch.from_numpy(durs["LJ050-0269"])
            assert len(pre.average_data_by_durations(energy, d)) == len(d)
    assert len(set(frames.values())) == 1  # mel / linear / raw agree on the frame count
    with pytest.raises(ev.ConfigError):
        ev.Preprocessor(ev.AudioConfig(spec_type="bogus"), device=cuda_device)


def test_error_behaviour(cuda_device):
    import everyvoice_b200 as ev
    from everyvoice_b200 import _lib

    tf, hop = _transform("A", "mel")
    # torch.stft raises for L <= n_fft // 2 (reflect padding); so do we, with a status code
    with pytest.raises(_lib.EvfError) as ei:
        tf(torch.zeros(512, device=cuda_device))
    assert ei.value.status == _lib.EVF_ERR_SHORT_INPUT
    assert tf(torch.zeros(513, device=cuda_device)).shape == (80, 3)
    # every n_fft torch.stft accepts has a kernel; only absurd sizes (no audio config) are refused
    huge = ev.get_spectral_transform("linear", 20000, 20000, 5000)
    with pytest.raises(_lib.EvfError) as ei:
        huge(torch.zeros(40000, device=cuda_device))
    assert ei.value.status == _lib.EVF_ERR_UNSUPPORTED
    assert ev.get_spectral_transform("istft", 1024, 1024, 256) is None
    assert ev.get_spectral_transform("nope", 1024, 1024, 256) is None
    pre = ev.Preprocessor(device=cuda_device)
    with pytest.raises(TypeError):
        pre.extract_spectral_features(torch.zeros(4000), lambda x: x)
    # empty batch is fine and launches nothing
    f = tf.features_ragged(torch.zeros(0, device=cuda_device), np.array([0], dtype=np.int64))
    assert f.spec.shape == (0, 80) and f.energy.shape == (0,)


# ------------------------------------------------------------------------------------------
# BASELINE.json full sizes: every utterance against the oracle on all host cores, plus size-independent properties
# ------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def full_size_batch(cuda_device):
    """configs[1]: 1 000 utterances of 1-10 s at 22.05 kHz (470 190 frames), white + speech-like."""
    from everyvoice_b200 import synth

    sr, hop = 22050, 256
    lens = synth.utterance_lengths(1000, sr, hop, 1234)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    g = torch.Generator(device=cuda_device)
    g.manual_seed(99)
    x = torch.rand(int(off[-1]), device=cuda_device, generator=g) * 1.9 - 0.95
    for b in range(0, 1000, 50):  # every 50th utterance is a speech-like signal (wide dynamic range)
        x[off[b] : off[b + 1]] = torch.from_numpy(synth.speech_like(int(lens[b]), sr, seed=b)).to(cuda_device)
    return lens, off, x


def _assert_every_utterance_matches_oracle(x, off, feats, cfg, spec_type, durations=None, phone_off=None, phone=None):
    """All utterances, all frames, against the oracle on every host core (tests/parity_pool.py)."""
    from parity_pool import compare_all

    res = compare_all(x.cpu().numpy(), off, feats.spec.cpu().numpy(), feats.energy.cpu().numpy(), feats.frame_offsets,
                      cfg, spec_type, durations, phone_off, None if phone is None else phone.cpu().numpy())
    assert res["utterances"] == len(off) - 1 and res["frames"] == int(feats.frame_offsets[-1])
    assert not res["failures"], res["failures"][:5]
    return res


def test_full_size_every_utterance_matches_oracle(cuda_device, full_size_batch):
    """BASELINE configs[1] at full size: ALL 1 000 utterances / 470 190 frames of log-mel, energy and phone-level
    averages against the oracle (north_star: 1e-3; frame counts and NaN positions exact)."""
    import everyvoice_b200 as ev
    from everyvoice_b200 import synth

    lens, off, x = full_size_batch
    pre = ev.Preprocessor(ev.AudioConfig(spec_type="mel"), device=cuda_device)
    feats = pre.process_spec_batch(x, off)
    assert np.array_equal(np.diff(feats.frame_offsets), lens // 256)      # bit-exact T = L // hop for all 1 000
    assert feats.spec.shape == (int((lens // 256).sum()), 80) == (470190, 80)
    assert bool(torch.isfinite(feats.spec).all()) and float(feats.spec.min()) >= np.log(1e-5) - 1e-6
    d_packed, p_off = synth.pack_ragged([synth.synthetic_durations(int(L) // 256, seed=1234 + i) for i, L in enumerate(lens)])
    phone, _ = pre.process_energy_batch(feats, torch.from_numpy(d_packed.astype(np.int64)), p_off)
    res = _assert_every_utterance_matches_oracle(x, off, feats, CONFIGS["A"], "mel", d_packed, p_off, phone)
    assert res["max_spec"] <= ATOL_LOG and res["max_energy"] <= ATOL_LOG and res["max_phone"] <= ATOL_LOG


def test_full_size_energy_is_norm_of_log_spec_and_checksums(cuda_device, full_size_batch):
    """energy == ||log-mel||_2 per frame (preprocessor.py:302-309) on all 470 190 frames, recomputed
    by an independent kernel and by torch; running the batch twice gives bit-identical output."""
    import everyvoice_b200 as ev

    lens, off, x = full_size_batch
    pre = ev.Preprocessor(ev.AudioConfig(spec_type="mel"), device=cuda_device)
    f1 = pre.process_spec_batch(x, off)
    f2 = pre.process_spec_batch(x, off)
    assert torch.equal(f1.spec, f2.spec) and torch.equal(f1.energy, f2.energy)       # deterministic
    e_kernel = pre.extract_energy(f1.spec.transpose(0, 1))                            # the stand-alone operator
    e_torch = torch.linalg.norm(f1.spec, dim=1)
    assert float((f1.energy - e_kernel).abs().max()) <= 2e-5 * float(e_torch.max())
    assert float((f1.energy - e_torch).abs().max()) <= 2e-5 * float(e_torch.max())


def test_full_size_parseval_on_raw_stft(cuda_device, full_size_batch):
    """Parseval on every frame of a 100-utterance slice: sum_k c_k |X[k]|^2 == N * sum_n (w[n] x[n])^2
    (c_k = 1 for DC / Nyquist, 2 otherwise).  Checks FFT, window, framing and reflect padding at
    full ragged scale without the oracle."""
    import everyvoice_b200 as ev

    lens, off, x = full_size_batch
    n_fft, hop = 1024, 256
    tf = ev.get_spectral_transform("raw", n_fft, n_fft, hop).to(cuda_device)
    sl = slice(200, 300)
    o = off[sl.start : sl.stop + 1] - off[sl.start]
    xs = x[off[sl.start] : off[sl.stop]]
    feats = tf.features_ragged(xs, o, keep_last=True)
    win = torch.hann_window(n_fft, device=cuda_device, dtype=torch.float64)
    c = torch.full((n_fft // 2 + 1,), 2.0, device=cuda_device, dtype=torch.float64)
    c[0] = c[-1] = 1.0
    worst = 0.0
    for b in range(sl.stop - sl.start):
        X = feats.utterance(b)                                              # [513, T + 1] complex64
        lhs = (c[:, None] * (X.real.double() ** 2 + X.imag.double() ** 2)).sum(dim=0)
        u = xs[o[b] : o[b + 1]].double()
        padded = torch.nn.functional.pad(u[None, None], (n_fft // 2, n_fft // 2), mode="reflect")[0, 0]
        frames = padded.unfold(0, n_fft, hop)                               # [T + 1, 1024]
        rhs = n_fft * ((frames * win) ** 2).sum(dim=1)
        assert lhs.shape == rhs.shape
        worst = max(worst, float(((lhs - rhs).abs() / rhs.clamp(min=1e-12)).max()))
    assert worst <= 2e-5, worst


def test_full_size_phone_average_round_trip(cuda_device, full_size_batch):
    """Sum over phones of (mean * clipped length) == sum of the covered frames, for all 1 000
    utterances (durations with zeros, overruns and underruns), and normalisation is idempotent:
    normalising already normalised data changes nothing beyond float round-off."""
    import everyvoice_b200 as ev
    from everyvoice_b200 import synth

    lens, off, x = full_size_batch
    pre = ev.Preprocessor(ev.AudioConfig(spec_type="mel"), device=cuda_device)
    T = lens // 256
    f_off = np.concatenate([[0], np.cumsum(T)]).astype(np.int64)
    g = torch.Generator(device=cuda_device)
    g.manual_seed(5)
    vals = torch.rand(int(f_off[-1]), device=cuda_device, generator=g) * 220 + 80
    durs = [synth.synthetic_durations(int(t), seed=900 + i) for i, t in enumerate(T)]
    d_packed, p_off = synth.pack_ragged(durs)
    out = pre.average_data_by_durations_ragged(vals, f_off, torch.from_numpy(d_packed.astype(np.int64)), p_off).cpu().double()
    v_cpu = vals.cpu().double()
    for b in range(0, 1000, 7):
        d = durs[b]
        start = np.concatenate([[0], np.cumsum(d)[:-1]])
        lo = np.clip(start, 0, T[b])
        hi = np.clip(start + d, 0, T[b])
        n = np.where(d > 0, hi - lo, 0)
        o = out[p_off[b] : p_off[b + 1]]
        assert torch.equal(torch.isnan(o), torch.from_numpy((d > 0) & (n == 0)))
        assert bool((o[torch.from_numpy(d <= 0)] == float(np.float32(1e-7))).all())
        covered = float(v_cpu[f_off[b] + lo.min() : f_off[b] + hi.max()].sum())
        recon = float((torch.nan_to_num(o) * torch.from_numpy(n.astype(np.float64)))[torch.from_numpy(n > 0)].sum())
        assert recon == pytest.approx(covered, rel=1e-5)
    s = ev.Scaler(cuda_device)
    phone = out.float().to(cuda_device)
    s.append(phone)
    s.normalize_by_device_stats_(phone, s.partial_stats())
    s2 = ev.Scaler(cuda_device)
    s2.append(phone)
    st = s2.calculate_stats(distributed=False)
    assert abs(st["mean"]) < 1e-4 and st["std"] == pytest.approx(1.0, abs=1e-4)


@pytest.mark.parametrize("config,spec_type", [("B", "mel"), ("A", "linear"), ("A", "mel-librosa")])
def test_other_baseline_configs_every_utterance(cuda_device, config, spec_type):
    """configs[2] (44.1 kHz / 2048 / 512 / 128 mels) and configs[3] (linear + energy) at 200
    ragged utterances: exact frame counts and oracle parity on every utterance."""
    import everyvoice_b200 as ev
    from everyvoice_b200 import synth
    from oracle import ev_oracle as O

    sr, n_fft, win, hop, n_mels, f_min, f_max = CONFIGS[config]
    lens = synth.utterance_lengths(200, sr, hop, 1236, 1.0, 10.0)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    g = torch.Generator(device=cuda_device)
    g.manual_seed(1236)
    x = torch.rand(int(off[-1]), device=cuda_device, generator=g) * 1.9 - 0.95
    tf = ev.get_spectral_transform(spec_type, n_fft, win, hop, sr, n_mels, f_min, f_max).to(cuda_device)
    feats = tf.features_ragged(x, off)
    assert np.array_equal(np.diff(feats.frame_offsets), lens // hop)
    _assert_every_utterance_matches_oracle(x, off, feats, CONFIGS[config], spec_type)


def test_config5_rank_shard_of_100h_corpus(cuda_device):
    """BASELINE.json configs[4]: a 100 h corpus at 22.05 kHz sharded by utterance over 8 GPUs.  One rank's shard
    (1/8 of the corpus: ~8.2 k utterances, 12.5 h, 4 GB of float32 samples) runs here at full size as ONE ragged
    batch; the corpus statistics are exchanged between 8 LOGICAL shards of it exactly as the ranks do after the
    all-gather.  Size-independent properties: exact frame counts, energy == ||log-mel||, the greedy shard balance,
    merged == global statistics, normalise-then-denormalise round trip."""
    import everyvoice_b200 as ev
    from everyvoice_b200 import synth
    from everyvoice_b200.distributed import finalize_stats, merge_stats, shard_utterances
    from oracle import ev_oracle as O

    sr, hop, world = 22050, 256, 8
    n_corpus = 65455                                         # ~65 455 utterances of mean 5.5 s = 100 h
    lens_all = synth.utterance_lengths(n_corpus, sr, hop, 1238)
    assert abs(lens_all.sum() / sr / 3600 - 100.0) < 1.0     # the corpus IS about 100 hours
    shards = shard_utterances(lens_all, world)
    loads = np.array([lens_all[s].sum() for s in shards], dtype=np.float64)
    assert loads.max() / loads.mean() < 1.001                # greedy longest-first: < 0.1 % imbalance
    assert sorted(np.concatenate(shards).tolist()) == list(range(n_corpus))
    mine = np.asarray(shards[3])                              # this "rank"
    lens = lens_all[mine]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    g = torch.Generator(device=cuda_device)
    g.manual_seed(1238 + 3)
    x = torch.rand(int(off[-1]), device=cuda_device, generator=g) * 1.9 - 0.95
    pre = ev.Preprocessor(ev.AudioConfig(spec_type="mel"), device=cuda_device)
    feats = pre.process_spec_batch(x, off)
    T = lens // hop
    assert np.array_equal(np.diff(feats.frame_offsets), T) and feats.spec.shape == (int(T.sum()), 80)
    e_torch = torch.linalg.norm(feats.spec, dim=1)
    assert float((feats.energy - e_torch).abs().max()) <= 2e-5 * float(e_torch.max())
    res = _assert_every_utterance_matches_oracle(x, off, feats, CONFIGS["A"], "mel")     # all ~8.2 k utterances
    assert res["max_spec"] <= ATOL_LOG and res["max_energy"] <= ATOL_LOG
    # statistics of the frame-level energy: 8 logical shards -> [8, 5] summaries -> merged == one global pass
    f_off = feats.frame_offsets
    cuts = [int(f_off[i]) for i in np.linspace(0, len(lens), world + 1).astype(int)]
    parts = []
    for r in range(world):
        s = ev.Scaler(cuda_device)
        s.append(feats.energy[cuts[r] : cuts[r + 1]])
        parts.append(s.partial_stats().clone())
    gathered = torch.stack(parts).contiguous()
    whole = ev.Scaler(cuda_device)
    whole.append(feats.energy)
    m, w = merge_stats(gathered).cpu(), whole.partial_stats().cpu()
    assert m[0] == w[0] == float(T.sum()) and m[3] == w[3] and m[4] == w[4]
    assert torch.allclose(m[1:3], w[1:3], rtol=1e-12, atol=0)
    st = finalize_stats(m.tolist(), len(lens))
    e64 = feats.energy.double()
    # finalize_stats reports float32-rounded numbers like the reference's Scaler (helpers.py:86-106)
    assert st["mean"] == pytest.approx(float(e64.mean()), rel=2e-7)
    assert st["std"] == pytest.approx(float(e64.std(unbiased=True)), rel=1e-6)
    normed = feats.energy.clone()
    ev.Scaler(cuda_device).normalize_by_device_stats_(normed, gathered)
    assert abs(float(normed.double().mean())) < 1e-4 and abs(float(normed.double().std()) - 1.0) < 1e-4
    back = normed * st["std"] + st["mean"]
    assert float((back - feats.energy).abs().max()) <= 1e-4 * float(feats.energy.abs().max())


@pytest.mark.parametrize("config", ["S512", "W512", "S256"])
@pytest.mark.parametrize("spec_type", ["mel", "linear"])
def test_nan_samples_stay_local_with_packed_jobs(cuda_device, config, spec_type):
    """n_fft 512 / 256: two / four packed jobs share a warp (register file, transpose scratch, the zero-padded power
    columns of the mel walk).  A NaN sample may reach the partner frame of its own job (frames 2j, 2j + 1) and nothing
    else -- in particular not the other jobs of the same warp."""
    from everyvoice_b200 import synth
    from oracle import ev_oracle as O

    sr, n_fft, win, hop, n_mels, f_min, f_max = CONFIGS[config]
    tf, _ = _transform(config, spec_type)
    otf, _ = _oracle_transform(config, spec_type)
    xs = [synth.speech_like(hop * n + 3, sr, seed=700 + i) for i, n in enumerate((150, 37, 260))]
    xs[0][hop * 41 + 5] = np.nan           # frames around 41: jobs 20 / 21, a warp's second / third job
    xs[2][hop * 130 + 1] = np.inf          # a later tile
    xs[2][3] = np.nan                      # reflected left margin
    packed, off = synth.pack_ragged(xs)
    feats = tf.features_ragged(torch.from_numpy(packed).to(cuda_device), off)
    for b, x in enumerate(xs):
        o_spec, o_energy, _ = O.features_one(torch.from_numpy(x), otf, hop)
        spec, energy = feats.utterance(b).cpu(), feats.utterance_energy(b).cpu()
        bad_o = ~torch.isfinite(o_spec).all(dim=0)
        bad = ~torch.isfinite(spec).all(dim=0)
        allowed = bad_o.clone()
        T = len(bad_o)
        for t in bad_o.nonzero().flatten().tolist():
            allowed[min(t ^ 1, T - 1)] = True
        assert bool((bad >= bad_o).all()) and bool((bad <= allowed).all()), (b, bad.nonzero().flatten().tolist(),
                                                                             bad_o.nonzero().flatten().tolist())
        assert torch.equal(~torch.isfinite(energy), bad)
        good = ~bad
        truth = O.truth_features(x, spec_type, n_fft, win, hop, sr, n_mels, f_min, f_max)[0] if spec_type == "linear" else None
        if truth is not None:
            truth = truth[:, good.numpy()]
        assert_log_spec_close(spec[:, good], o_spec[:, good], truth, spec_type)
        assert float((energy[good] - o_energy[good]).abs().max()) <= (ATOL_LOG if spec_type == "mel" else 2e-2)


def test_nan_and_inf_samples_stay_local(cuda_device):
    """A NaN / Inf sample poisons the frames whose window covers it (like torch.stft does) plus, for n_fft 1024, the
    partner frame of the same FFT job (frames 2j and 2j + 1 ride as real and imaginary part of one complex FFT, so a
    non-finite value in one reaches the other in the real-FFT separation) -- and nothing else: no stale value leaks
    between jobs, tiles or warps.  Documented deviation (INTEGRATION.md
    section 5); PCM input cannot contain non-finite samples."""
    from everyvoice_b200 import synth
    from oracle import ev_oracle as O

    tf, hop = _transform("A", "mel")
    otf, _ = _oracle_transform("A", "mel")
    xs = [synth.speech_like(hop * n, 22050, seed=300 + i) for i, n in enumerate((70, 33, 95, 40))]
    xs[0][hop * 20 + 7] = np.nan
    xs[2][hop * 64 + 100] = np.inf       # second tile of the utterance
    xs[2][5] = -np.inf                    # inside the reflected left margin: mirrored into frame 0 twice
    packed, off = synth.pack_ragged(xs)
    feats = tf.features_ragged(torch.from_numpy(packed).to(cuda_device), off)
    n_bad = 0
    for b, x in enumerate(xs):
        o_spec, o_energy, _ = O.features_one(torch.from_numpy(x), otf, hop)
        spec, energy = feats.utterance(b).cpu(), feats.utterance_energy(b).cpu()
        bad_o = ~torch.isfinite(o_spec).all(dim=0)
        bad = ~torch.isfinite(spec).all(dim=0)
        allowed = bad_o.clone()           # the reference's frames, widened to whole (2j, 2j + 1) pairs
        T = len(bad_o)
        for t in bad_o.nonzero().flatten().tolist():
            allowed[min(t ^ 1, T - 1)] = True
        assert bool((bad >= bad_o).all()) and bool((bad <= allowed).all()), (b, bad.nonzero().flatten().tolist())
        assert torch.equal(~torch.isfinite(energy), bad)
        good = ~bad
        assert float((spec[:, good] - o_spec[:, good]).abs().max()) <= ATOL_LOG
        assert float((energy[good] - o_energy[good]).abs().max()) <= ATOL_LOG
        n_bad += int(bad.sum())
    assert 11 <= n_bad <= 16 and int((~torch.isfinite(feats.spec).all(dim=1)).sum()) == n_bad
    # n_fft 2048: one frame per FFT job, so the locality is exactly the reference's
    tfB, hopB = _transform("B", "mel")
    otfB, _ = _oracle_transform("B", "mel")
    y = synth.speech_like(hopB * 40, 44100, seed=310)
    y[hopB * 17 + 3] = np.nan
    fB = tfB.features_ragged(torch.from_numpy(y).to(cuda_device), np.array([0, len(y)]))
    o_spec, _, _ = O.features_one(torch.from_numpy(y), otfB, hopB)
    assert torch.equal(~torch.isfinite(fB.utterance(0).cpu()).all(dim=0), ~torch.isfinite(o_spec).all(dim=0))


def test_packed_buffer_beyond_2_31_samples(cuda_device):
    """Maximum sizes: a packed int16 batch of 2.3e9 samples (4.6 GB; 64-bit sample and output offsets).  The last
    utterances start beyond the 2^31-th sample and must equal the same utterances processed alone, bit for bit."""
    from everyvoice_b200 import synth

    tf, hop = _transform("A", "mel")
    filler, n_fill = 10_000_128, 230                     # 230 x 7.6 minutes
    tail = [synth.speech_like(hop * n, 22050, seed=400 + i) for i, n in enumerate((50, 123))]
    tail_pcm = [np.clip(np.rint(t * 32767.0), -32768, 32767).astype(np.int16) for t in tail]
    lens = np.array([filler] * n_fill + [len(t) for t in tail_pcm], dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    assert off[n_fill] > 2**31
    x = torch.empty(int(off[-1]), dtype=torch.int16, device=cuda_device)
    g = torch.Generator(device=cuda_device)
    g.manual_seed(7)
    chunk = 2**27
    for s in range(0, int(off[n_fill]), chunk):          # seeded filler, generated in place
        e = min(s + chunk, int(off[n_fill]))
        x[s:e] = torch.randint(-20000, 20000, (e - s,), generator=g, device=cuda_device, dtype=torch.int16)
    for i, t in enumerate(tail_pcm):
        x[int(off[n_fill + i]) : int(off[n_fill + i + 1])] = torch.from_numpy(t).to(cuda_device)
    feats = tf.features_ragged(x, off)
    assert np.array_equal(np.diff(feats.frame_offsets), lens // hop)
    assert feats.spec.shape[0] == int((lens // hop).sum()) and feats.spec.shape[0] * 80 > 2**29
    for i, t in enumerate(tail_pcm):
        single = tf.features_ragged(torch.from_numpy(t).to(cuda_device), np.array([0, len(t)]))
        assert torch.equal(single.spec, feats.utterance(n_fill + i).transpose(0, 1))
        assert torch.equal(single.energy, feats.utterance_energy(n_fill + i))
    mid = feats.utterance(117)                           # a filler utterance in the middle: finite, above the floor
    assert bool(torch.isfinite(mid).all()) and float(mid.min()) >= np.log(1e-5) - 1e-6
