"""ctypes binding of ``libevfeat.so`` (``include/evfeat.h``).

There is no other code path: if the library is missing or no sm_100 device is present,
everything here raises.  The library is built in-tree by ``everyvoice_b200.build``.
"""

from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path

# EVF_LIB selects an experimental build of the same library (kernel A/B runs); never a fallback
LIB_PATH = Path(os.environ.get("EVF_LIB") or Path(__file__).resolve().parent / "libevfeat.so")

# evf_status
EVF_OK = 0
EVF_ERR_INVALID_ARGUMENT = 1
EVF_ERR_UNSUPPORTED = 2
EVF_ERR_SHORT_INPUT = 3
EVF_ERR_FILTERBANK = 4
EVF_ERR_CUDA = 5
EVF_ERR_NO_DEVICE = 6
EVF_ERR_OUT_OF_MEMORY = 7

# evf_spec_type  (everyvoice/config/preprocessing_config.py:18-22)
SPEC_TYPES = {"mel": 0, "mel-librosa": 1, "linear": 2, "raw": 3}
SAMPLES_F32, SAMPLES_S16 = 0, 1
GRAD_FRAME_MAJOR, GRAD_BIN_MAJOR = 0, 1   # evf_features_backward_ex: layout of the incoming gradient
FFT_AUTO, FFT_GENERIC = 0, 1

ABI_VERSION = 11


class evf_config(C.Structure):
    _fields_ = [
        ("spec_type", C.c_int32),
        ("sample_rate", C.c_int32),
        ("n_fft", C.c_int32),
        ("win_length", C.c_int32),
        ("hop_length", C.c_int32),
        ("n_mels", C.c_int32),
        ("apply_log", C.c_int32),
        ("keep_last_frame", C.c_int32),
        ("sample_format", C.c_int32),
        ("log_clip", C.c_float),
        ("fft_path", C.c_int32),
    ]


# name -> (restype, argtypes); the single source of truth for the symbol table test
_P = C.c_void_p
_I64P = C.POINTER(C.c_int64)
PROTOTYPES = {
    "evf_abi_version": (C.c_int, []),
    "evf_last_error": (C.c_char_p, []),
    "evf_plan_create": (C.c_int, [C.POINTER(evf_config), _P, _P, C.c_int, C.POINTER(_P)]),
    "evf_plan_destroy": (C.c_int, [_P]),
    "evf_plan_row_floats": (C.c_int, [_P, C.POINTER(C.c_int32)]),
    "evf_plan_num_frames": (C.c_int64, [_P, C.c_int64]),
    "evf_batch_create": (C.c_int, [_P, _P, C.c_int32, C.POINTER(_P)]),
    "evf_batch_destroy": (C.c_int, [_P]),
    "evf_batch_total_frames": (C.c_int, [_P, _I64P]),
    "evf_batch_frame_offsets": (C.c_int, [_P, _P]),
    "evf_batch_frame_offsets_dev": (C.c_int, [_P, C.POINTER(_P)]),
    "evf_features_run": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "evf_features_run_range": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P, _P, _P]),
    "evf_features_ragged": (C.c_int, [_P, _P, _P, C.c_int32, _P, _P, _P, _P]),
    "evf_energy_from_spec": (C.c_int, [_P, C.c_int64, C.c_int32, _P, _P]),
    "evf_log_compress": (C.c_int, [_P, _P, C.c_int64, C.c_float, C.c_float, _P]),
    "evf_segment_mean": (C.c_int, [_P, _P, _P, _P, C.c_int32, _P, _P]),
    "evf_stats_partial": (C.c_int, [_P, C.c_int64, _P, C.c_int32, _P]),
    "evf_normalize_inplace": (C.c_int, [_P, C.c_int64, C.c_float, C.c_float, _P]),
    "evf_normalize_by_stats": (C.c_int, [_P, C.c_int64, _P, _P]),
    "evf_stats_merge": (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P]),
    "evf_normalize_by_gathered_stats": (C.c_int, [_P, C.c_int64, _P, C.c_int32, C.c_int32, _P]),
    # backward of the transform / of the log compression (training through the mel, hfgl/model.py:581-590)
    "evf_features_backward_scratch_floats": (C.c_int64, [_P, _P]),
    "evf_features_backward": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "evf_features_backward_ex": (C.c_int, [_P, _P, _P, _P, C.c_int32, _P, _P, _P, _P]),
    "evf_log_compress_backward": (C.c_int, [_P, _P, _P, C.c_int64, C.c_float, _P]),
    "evf_pitch_fill_unvoiced": (C.c_int, [_P, _P, C.c_int32, _P, _P]),
    # pitch tracking (DIO + StoneMask, restated from WORLD; parity unpinned)
    "evf_pitch_num_frames": (C.c_int64, [C.c_int32, C.c_double, C.c_int64]),
    "evf_pitch_scratch_bytes": (C.c_int64, [_P, C.c_int32, C.c_int32, C.c_double, C.c_int32]),
    "evf_pitch_dio_stonemask": (C.c_int, [_P, C.c_int32, _P, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_double,
                                          C.c_double, C.c_double, C.c_double, _P, C.c_int64, _P, _P]),
    # audio front-end (process_audio numerics)
    "evf_resampler_create": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.POINTER(_P)]),
    "evf_resampler_destroy": (C.c_int, [_P]),
    "evf_resampler_out_length": (C.c_int64, [_P, C.c_int64]),
    "evf_audio_resample": (C.c_int, [_P, _P, C.c_int32, _P, _P, C.c_int32, C.c_int64, _P, _P]),
    "evf_audio_absmax": (C.c_int, [_P, C.c_int32, _P, C.c_int32, C.c_int64, _P, _P]),
    "evf_audio_finalize": (C.c_int, [_P, C.c_int32, _P, _P, C.c_int32, C.c_int64, _P, _P, _P, _P]),
    "evf_audio_loudness_scratch_floats": (C.c_int64, [C.c_int32, C.c_int64]),
    "evf_audio_loudness_step": (C.c_int32, [C.c_int32]),
    "evf_audio_gate_mask": (C.c_int, [_P, C.c_float, _P, _P, C.c_int32, _P, _P]),
    "evf_audio_loudness": (C.c_int, [_P, C.c_int32, _P, C.c_int32, C.c_int64, C.c_int32, _P, C.c_float, C.c_float, _P, _P,
                                     _P, _P, _P]),
}


class EvfError(RuntimeError):
    """A non-zero evf_status.  ``status`` carries the code, the message is evf_last_error()."""

    def __init__(self, status: int, message: str):
        super().__init__(f"libevfeat status {status}: {message}")
        self.status = status


class EvfLibraryMissing(ImportError):
    pass


_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """Load libevfeat.so (once).  Raises ``EvfLibraryMissing`` if it has not been built --
    there is deliberately no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            raise EvfLibraryMissing(
                f"{LIB_PATH} not found. Build it with `python -m everyvoice_b200.build` "
                "(needs nvcc; sm_100a only). everyvoice_b200 has no CPU or PyTorch fallback."
            )
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        v = lib.evf_abi_version()
        if v != ABI_VERSION:
            raise EvfLibraryMissing(f"libevfeat ABI version {v}, expected {ABI_VERSION}: rebuild")
        _lib = lib
    return _lib


def check(status: int) -> None:
    if status != EVF_OK:
        msg = load().evf_last_error()
        raise EvfError(status, msg.decode("utf-8", "replace") if msg else "")
