"""CPU: the oracle's restatement of the pitch post-processing (oracle/ev_oracle.py: postprocess_pitch; reference:
everyvoice/preprocessor/preprocessor.py:236-285 after pyworld) against the outputs of the LIVE reference's
``extract_pitch`` stored in tests/golden/pitch.npz by oracle/make_golden_pitch.py."""
import numpy as np

from oracle import ev_oracle as O
from oracle.make_golden_pitch import pitch_tracks


def test_postprocess_pitch_equals_live_reference(golden_dir):
    golden = np.load(golden_dir / "pitch.npz")
    tracks = pitch_tracks()
    assert sorted(golden.files) == sorted(tracks)
    for name, f0 in tracks.items():
        out = O.postprocess_pitch(f0)
        assert out.dtype == np.float32 and np.array_equal(out, golden[name]), name
        assert not np.isnan(out).any()
        voiced = f0 != 0
        assert np.array_equal(out[voiced], f0[voiced].astype(np.float32))   # voiced frames pass through
    assert not O.postprocess_pitch(np.zeros(9)).any()


# ---- DIO + StoneMask restatement (oracle/world_pitch.py): PARITY UNPINNED, see its header -------------------------
def test_world_decimation_table_is_cheby1():
    """WORLD's FilterForDecimate hard-codes its coefficients; they are MATLAB's cheby1(3, 0.05, 0.8 / r).  The one
    published entry recalled with confidence (r = 11) and scipy's design agree to 1e-13, and the table the oracle and
    the kernels use is that design for every r."""
    signal = __import__("pytest").importorskip("scipy.signal")
    from oracle import world_pitch as W

    a11, b11 = W.DECIMATE_COEFFS[11]
    assert np.allclose(a11, (2.450743295230728, -2.06794904601978, 0.59574774438332101), rtol=1e-13)
    assert np.allclose(b11, (0.0026822508007163792, 0.0080467524021491377), rtol=1e-13)
    for r, (a, b) in W.DECIMATE_COEFFS.items():
        bb, aa = signal.cheby1(3, 0.05, 0.8 / r)
        assert np.allclose(a, (-aa[1], -aa[2], -aa[3]), rtol=1e-14) and np.allclose(b, (bb[0], bb[1]), rtol=1e-14)
        assert np.allclose(bb, bb[0] * np.array([1, 3, 3, 1]), rtol=1e-12)


def test_world_pitch_restatement_tracks_known_f0():
    """Sanity of the restatement on signals whose f0 is known: a harmonic stack with vibrato is tracked within 1 %,
    noise and silence come out unvoiced, f0_length is WORLD's GetSamplesForDIO, and decimate() is zero phase."""
    from everyvoice_b200 import synth
    from oracle import world_pitch as W

    sr, hop = 22050, 256
    fp = hop / sr * 1000
    rng = np.random.default_rng(3)
    f_true = rng.uniform(80.0, 300.0)                        # the f0 synth.speech_like(seed=3) draws first
    x = synth.speech_like(int(1.5 * sr) // hop * hop, sr, seed=3).astype(np.float64)
    f0, t = W.dio(x, sr, frame_period=fp, speed=4)
    assert len(f0) == len(t) == int(1000.0 * len(x) / sr / fp) + 1 == len(x) // hop + 1
    f0r = W.stonemask(x, f0, t, sr)
    voiced = f0r > 0
    assert voiced.mean() > 0.8
    assert np.abs(f0r[voiced] / f_true - 1.0).max() < 0.035   # 2 % vibrato + estimation error
    assert abs(np.median(f0r[voiced]) / f_true - 1.0) < 0.01
    noise = rng.uniform(-0.5, 0.5, size=sr).astype(np.float64)
    assert (W.dio(noise, sr, frame_period=fp, speed=4)[0] > 0).mean() < 0.2
    assert not W.dio(np.zeros(sr), sr, frame_period=fp, speed=4)[0].any()
    # decimate: a slow sine keeps its phase and amplitude
    n = np.arange(4000)
    s = np.sin(2 * np.pi * 50 * n / sr)
    d = W.decimate(s, 4)            # MATLAB's decimate keeps samples nbeg - 1 + r * k (here 3, 7, 11, ...)
    ref = s[3::4]
    assert np.abs(d[20 : len(ref) - 20] - ref[20:-20]).max() < 2e-3
    # the whole extract_pitch: unvoiced frames filled, float32
    p = W.extract_pitch(x.astype(np.float32), sr, hop)
    assert p.dtype == np.float32 and len(p) == len(f0) and (p > 0).all()
