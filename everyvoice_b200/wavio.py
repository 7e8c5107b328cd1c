"""RIFF / WAVE reader for the single-file surface of ``Preprocessor.process_audio`` (reference:
``everyvoice/utils/__init__.py`` ``load_audio`` -> ``torchaudio.load(path)`` with its default ``normalize=True``).

Host code: the formats a wav file on disk can have -- integer PCM of 8 / 16 / 24 / 32 bits, IEEE float of 32 / 64
bits, plain or ``WAVE_FORMAT_EXTENSIBLE`` headers -- scaled the way ``torchaudio.load`` scales them (unsigned 8 bit:
``(s - 128) / 128``; signed ``b`` bit: ``s / 2**(b - 1)``; float: as stored, float64 rounded to float32).  16-bit PCM
stays int16: the device kernels convert it as ``s / 32768`` on load, which is bit-identical and halves the
host-to-device bytes.  The standard library's ``wave`` reads neither 24-bit nor float files."""

from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

_PCM, _FLOAT, _EXTENSIBLE = 1, 3, 0xFFFE


def read_wav(path) -> tuple[np.ndarray, int]:
    """``(samples[C, L], sampling_rate)``; ``samples`` is int16 for 16-bit PCM and float32 otherwise."""
    data = Path(path).read_bytes()
    if len(data) < 12 or data[:4] not in (b"RIFF", b"RF64") or data[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF / WAVE file")
    pos, fmt, body = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack_from("<I", data, pos + 4)[0]
        start = pos + 8
        if cid == b"fmt ":
            tag, ch, sr, _, _, bits = struct.unpack_from("<HHIIHH", data, start)
            if tag == _EXTENSIBLE and size >= 26:
                tag = struct.unpack_from("<H", data, start + 24)[0]   # first two bytes of the sub-format GUID
            fmt = (tag, ch, sr, bits)
        elif cid == b"data":
            end = len(data) if size == 0xFFFFFFFF else min(len(data), start + size)   # streamed files leave the size open
            body = data[start:end]
            break
        pos = start + size + (size & 1)   # chunks are word aligned
    if fmt is None or body is None:
        raise ValueError(f"{path}: missing fmt or data chunk")
    tag, ch, sr, bits = fmt
    if ch < 1:
        raise ValueError(f"{path}: no channels")
    if (tag, bits) not in ((_PCM, 8), (_PCM, 16), (_PCM, 24), (_PCM, 32), (_FLOAT, 32), (_FLOAT, 64)):
        raise ValueError(f"{path}: unsupported wav encoding (format tag {tag}, {bits} bits)")
    width = bits // 8
    n = len(body) // (width * ch) * ch
    if tag == _PCM and bits == 16:
        a = np.frombuffer(body, dtype="<i2", count=n).copy()
    elif tag == _PCM and bits == 8:
        a = ((np.frombuffer(body, dtype=np.uint8, count=n).astype(np.float32) - 128.0) / 128.0).astype(np.float32)
    elif tag == _PCM and bits == 24:
        b = np.frombuffer(body, dtype=np.uint8, count=3 * n).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v >= 1 << 23, v - (1 << 24), v)
        a = (v.astype(np.float64) / 8388608.0).astype(np.float32)
    elif tag == _PCM and bits == 32:
        a = (np.frombuffer(body, dtype="<i4", count=n).astype(np.float64) / 2147483648.0).astype(np.float32)
    elif tag == _FLOAT and bits == 32:
        a = np.frombuffer(body, dtype="<f4", count=n).copy()
    else:
        a = np.frombuffer(body, dtype="<f8", count=n).astype(np.float32)
    return np.ascontiguousarray(a.reshape(-1, ch).T), int(sr)
