"""The whole per-batch flow behind ``everyvoice preprocess`` on the GPU, through the files it leaves on disk:
loaded wavs -> process_audio (gates, peak normalisation, truncation, PCM16) -> audio/*.wav -> process_spec ->
spec/*.pt -> process_energy (phone-level) -> compute_stats / normalize_stats -> energy/*.pt + stats.json, compared
file by file with the CPU oracle running the reference's per-utterance flow (preprocessor.py:131-218, 632-651,
870-929, 378-490; fs2/cli/preprocess.py:44-77)."""
import json
import wave

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ATOL_LOG = 1e-3


def test_preprocess_flow_files_match_reference_flow(cuda_device, tmp_path):
    import everyvoice_b200 as ev
    from everyvoice_b200 import synth
    from everyvoice_b200.artefacts import ArtefactWriter, create_path
    from oracle import ev_oracle as O

    sr, hop = 22050, 256
    pre = ev.Preprocessor(ev.AudioConfig(spec_type="mel"), device=cuda_device)
    rng = np.random.default_rng(11)
    secs = (0.9, 2.3, 0.2, 1.4, 1.1)                      # the third is shorter than min_audio_length
    gains = (0.5, 0.8, 0.5, 0.003, 0.3)                   # the fourth is quieter than -36 LKFS
    raw = [O.pcm16(synth.speech_like(int(sr * s) + 13, sr, seed=500 + i) * np.float32(g))
           for i, (s, g) in enumerate(zip(secs, gains))]
    loaded = [torch.from_numpy(r.astype(np.float32) / 32768.0)[None] for r in raw]   # what load_audio returns
    items = [{"basename": f"utt{i:03d}", "speaker": "default", "language": "und"} for i in range(len(raw))]

    # ---- ours: four batched calls + one writer -------------------------------------------------
    audio = pre.process_audio_batch(loaded, sr, resample_rate=sr, hop_size=hop, out_dtype=torch.int16)
    assert audio.kept == [0, 1, 4] and audio.skipped == {2: "audio_too_short", 3: "audio_empty"}
    kept_items = [items[i] for i in audio.kept]
    feats = pre.process_spec_batch(audio.samples, audio.offsets)
    durs = [synth.synthetic_durations(int(t), seed=600 + j) for j, t in enumerate(np.diff(feats.frame_offsets))]
    phone, p_off = pre.process_energy_batch(feats, [torch.from_numpy(d) for d in durs])
    e_scaler, _ = pre.compute_stats(energy=phone, n_energy_files=len(kept_items))
    stats = pre.normalize_stats(e_scaler, None, distributed=False)   # normalises `phone` in place
    with ArtefactWriter(tmp_path, workers=4) as w:
        w.write_audio(items, audio)
        w.write_specs(kept_items, feats, sr, "mel")
        w.write_energy(kept_items, phone, p_off)
        w.write_stats(stats)

    # ---- the reference's flow, one utterance at a time (oracle) -------------------------------------
    otf = O.get_spectral_transform("mel", 1024, 1024, hop, sr, 80, 0, 8000)
    o_scaler, o_phone = O.Scaler(), {}
    for i in audio.kept:
        a, a_sr = O.process_audio_tensor(loaded[i].numpy(), sr, resample_rate=sr, hop_size=hop)
        pcm = O.pcm16(a)                                    # save_wav ...
        with wave.open(str(create_path(tmp_path, items[i], "audio", f"audio-{sr}.wav")), "rb") as f:
            assert f.getframerate() == a_sr == sr
            assert np.array_equal(np.frombuffer(f.readframes(f.getnframes()), dtype="<i2"), pcm)   # bit-exact wav
        x = torch.from_numpy(pcm.astype(np.float32) / 32768.0)   # ... and load_audio in process_spec
        j = audio.kept.index(i)
        o_spec, _, o_ph = O.features_one(x, otf, hop, torch.from_numpy(durs[j]))
        spec = torch.load(create_path(tmp_path, items[i], "spec", f"spec-{sr}-mel.pt"), weights_only=True)
        assert tuple(spec.shape) == tuple(o_spec.shape) == (80, len(pcm) // hop)
        assert float((spec - o_spec).abs().max()) <= ATOL_LOG
        o_phone[i] = o_ph
        o_scaler.append(o_ph)
    o_stats = o_scaler.calculate_stats()
    got = json.loads((tmp_path / "stats.json").read_text())["energy"]
    assert got["sample_size"] == o_stats["sample_size"] == 3
    for k in ("min", "max", "mean", "std", "norm_min", "norm_max"):
        assert got[k] == pytest.approx(o_stats[k], rel=1e-4, abs=1e-4), k
    for i in audio.kept:
        e = torch.load(create_path(tmp_path, items[i], "energy", "energy.pt"), weights_only=True)
        want = o_scaler.normalize(o_phone[i])
        assert e.shape == want.shape
        nan = torch.isnan(want)
        assert torch.equal(torch.isnan(e), nan)
        assert float((e[~nan] - want[~nan]).abs().max()) <= 2e-3   # (1e-3 log-energy error) / std, std ~ 10
    for i in (2, 3):                                        # skipped utterances leave no files
        assert not create_path(tmp_path, items[i], "audio", f"audio-{sr}.wav").exists()


@pytest.mark.parametrize("dtype", ["s16", "f32"])
def test_pipelined_flow_equals_the_stepwise_calls(cuda_device, dtype):
    """FlowPipeline (one chunked three-stream pipeline, loudness gate consumed on the device) against the same flow
    through the stepwise API (process_audio_batch -> process_spec_batch -> process_energy_batch -> statistics): same
    keep / skip decisions, and for the kept utterances bit-identical PCM16 audio, log-mel, energy and normalised
    phone values; a skipped utterance's phone values are NaN and stay out of the statistics."""
    import everyvoice_b200 as ev
    from everyvoice_b200 import synth

    sr, hop = 22050, 256
    pre = ev.Preprocessor(ev.AudioConfig(spec_type="mel"), device=cuda_device)
    rng = np.random.default_rng(12)
    xs = []
    for i in range(23):
        n = int(rng.integers(9000, 60000))
        x = synth.speech_like(n, sr, seed=700 + i) * np.float32(rng.uniform(0.05, 0.9))
        if i in (3, 11, 19):
            x = x * np.float32(0.003)     # below -36 LKFS: "audio_empty"
        if i == 7:
            x = np.zeros(n, np.float32)    # silence: loudness NaN / -inf
        xs.append(x.astype(np.float32))
    if dtype == "s16":
        xs = [np.clip(np.rint(x * 32768.0), -32768, 32767).astype(np.int16) for x in xs]
    packed, raw_off = synth.pack_ragged(xs)
    durs = [synth.synthetic_durations(len(x) // hop, seed=40 + i) for i, x in enumerate(xs)]
    d_packed, p_off = synth.pack_ragged(durs)
    tdt = torch.int16 if dtype == "s16" else torch.float32
    # ---- stepwise reference of our own API --------------------------------------------------------------------
    audio = pre.process_audio_batch([torch.from_numpy(x) for x in xs], sr, resample_rate=sr, hop_size=hop, out_dtype=torch.int16)
    assert {3, 7, 11, 19} <= set(audio.skipped) and len(audio.kept) >= 15
    feats = pre.process_spec_batch(audio.samples, audio.offsets)
    kd = [torch.from_numpy(durs[i]) for i in audio.kept]
    phone, kp_off = pre.process_energy_batch(feats, kd)
    sc, _ = pre.compute_stats(energy=phone, n_energy_files=len(audio.kept))
    stats = pre.normalize_stats(sc, None, distributed=False)["energy"]
    # ---- the pipeline (small chunks: several of them) ---------------------------------------------------------
    flow = pre.make_flow_pipeline(raw_off, sr, tdt, torch.from_numpy(d_packed.astype(np.int64)), p_off, chunk_bytes=1 << 18)
    assert len(flow.chunks) >= 4
    h_in = torch.from_numpy(packed).pin_memory()
    h_spec = torch.empty((flow.total_frames, 80), dtype=torch.float32).pin_memory()
    h_energy = torch.empty(flow.total_frames, dtype=torch.float32).pin_memory()
    h_phone = torch.empty(int(p_off[-1]), dtype=torch.float32).pin_memory()
    h_audio = torch.empty(int(flow.kept_offsets[-1]), dtype=torch.int16).pin_memory()
    for _ in range(2):  # twice: buffers / events are reused
        res = flow.run(h_in, h_spec, h_energy, h_phone, h_audio)()
    assert res.keep.tolist() == [i in audio.kept for i in range(len(xs))]
    assert np.array_equal(res.loudness, audio.loudness, equal_nan=True)
    for j, i in enumerate(audio.kept):
        k0, k1 = int(flow.kept_offsets[i]), int(flow.kept_offsets[i + 1])
        f0, f1 = int(flow.frame_offsets[i]), int(flow.frame_offsets[i + 1])
        assert np.array_equal(h_audio[k0:k1].numpy(), audio.utterance(j).cpu().numpy()), i
        assert torch.equal(h_spec[f0:f1], feats.spec[feats.frame_offsets[j]:feats.frame_offsets[j + 1]].cpu()), i
        assert torch.equal(h_energy[f0:f1], feats.utterance_energy(j).cpu()), i
        got = h_phone[int(p_off[i]):int(p_off[i + 1])].numpy()
        want = phone[int(kp_off[j]):int(kp_off[j + 1])].cpu().numpy()
        assert np.allclose(got, want, atol=2e-6, equal_nan=True), i
    for i in audio.skipped:
        assert np.isnan(h_phone[int(p_off[i]):int(p_off[i + 1])].numpy()).all()
    from everyvoice_b200.distributed import finalize_stats, merge_stats
    st = finalize_stats(merge_stats(res.stats5).cpu().tolist(), int(res.keep.sum()))
    for k in ("mean", "std", "min", "max", "sample_size"):
        assert st[k] == pytest.approx(stats[k], rel=1e-6), k
