"""Host-side wav reader (everyvoice_b200/wavio.py) against scipy.io.wavfile as an independent reader and against the
scaling torchaudio.load applies (reference: everyvoice/utils load_audio -> torchaudio.load, normalize=True)."""
import struct

import numpy as np
import pytest

from everyvoice_b200.wavio import read_wav


def _riff(fmt_tag, channels, sr, bits, payload, extensible=False):
    block = channels * bits // 8
    if extensible:
        guid_tail = bytes.fromhex("000000001000800000aa00389b71")
        fmt = struct.pack("<HHIIHHHHIH", 0xFFFE, channels, sr, sr * block, block, bits, 22, bits, 0, fmt_tag) + guid_tail
    else:
        fmt = struct.pack("<HHIIHH", fmt_tag, channels, sr, sr * block, block, bits)
    chunks = b"fmt " + struct.pack("<I", len(fmt)) + fmt
    chunks += b"LIST" + struct.pack("<I", 5) + b"INFOx" + b"\x00"          # an odd-sized chunk in front of the data
    chunks += b"data" + struct.pack("<I", len(payload)) + payload
    return b"RIFF" + struct.pack("<I", 4 + len(chunks)) + b"WAVE" + chunks


@pytest.mark.parametrize("channels", [1, 2])
@pytest.mark.parametrize("kind", ["u8", "s16", "s24", "s32", "f32", "f64", "s24_ext", "f32_ext"])
def test_read_wav_formats(tmp_path, kind, channels):
    from scipy.io import wavfile

    rng = np.random.default_rng(5)
    n, sr = 777, 22050
    x = rng.uniform(-1, 1, size=(n, channels))
    ext = kind.endswith("_ext")
    base = kind.split("_")[0]
    if base == "u8":
        q = np.clip(np.rint(x * 127 + 128), 0, 255).astype(np.uint8)
        payload, want, tag, bits = q.tobytes(), (q.astype(np.float32) - 128) / 128, 1, 8
    elif base == "s16":
        q = np.rint(x * 32767).astype("<i2")
        payload, want, tag, bits = q.tobytes(), q, 1, 16
    elif base == "s24":
        q = np.rint(x * 8388607).astype(np.int32)
        payload = b"".join(int(v).to_bytes(3, "little", signed=True) for v in q.reshape(-1))
        want, tag, bits = (q.astype(np.float64) / 8388608).astype(np.float32), 1, 24
    elif base == "s32":
        q = np.rint(x * 2147483647).astype("<i4")
        payload, want, tag, bits = q.tobytes(), (q.astype(np.float64) / 2147483648).astype(np.float32), 1, 32
    elif base == "f32":
        q = x.astype("<f4")
        payload, want, tag, bits = q.tobytes(), q, 3, 32
    else:
        q = x.astype("<f8")
        payload, want, tag, bits = q.tobytes(), q.astype(np.float32), 3, 64
    path = tmp_path / f"{kind}.wav"
    path.write_bytes(_riff(tag, channels, sr, bits, payload, extensible=ext))
    got, got_sr = read_wav(path)
    assert got_sr == sr and got.shape == (channels, n)
    assert got.dtype == (np.int16 if base == "s16" else np.float32)
    assert np.array_equal(got, np.ascontiguousarray(want.T))
    # an independent reader agrees on the integers / floats in the file
    s_sr, s = wavfile.read(path)
    s = s.reshape(n, channels)
    assert s_sr == sr
    if base == "s24":
        s = s >> 8 if s.dtype == np.int32 and np.abs(s).max() > (1 << 23) else s
        assert np.array_equal((s.astype(np.float64) / 8388608).astype(np.float32), want)
    elif base == "u8":
        assert np.array_equal((s.astype(np.float32) - 128) / 128, want)
    elif base == "s32":
        assert np.array_equal((s.astype(np.float64) / 2147483648).astype(np.float32), want)
    else:
        assert np.array_equal(s.astype(want.dtype), want)


def test_read_wav_rejects_other_files(tmp_path):
    p = tmp_path / "x.wav"
    p.write_bytes(b"not a wav file at all")
    with pytest.raises(ValueError):
        read_wav(p)
    p.write_bytes(_riff(2, 1, 8000, 4, b"\x00" * 16))   # ADPCM
    with pytest.raises(ValueError):
        read_wav(p)
