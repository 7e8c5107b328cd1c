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
