"""The drop-in boundary: libevfeat.so builds for sm_100a, loads, and exports exactly the
symbols include/evfeat.h declares.  No compute calls (there is no GPU in this pass); the
only calls made are the ones that must work -- or fail loudly -- without a device."""

import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "evfeat.h"


@pytest.fixture(scope="module")
def lib():
    from everyvoice_b200 import _lib, build

    build.build(force=False)  # nvcc cross-compiles without a GPU; no-op when up to date
    return _lib.load()


def _declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return re.findall(r"EVF_API\s+[\w\s\*]+?\b(evf_\w+)\s*\(", text)


def test_header_declares_the_operator_surface():
    names = _declared_symbols()
    assert len(names) == len(set(names)) >= 22
    for required in ("evf_plan_create", "evf_features_run", "evf_features_ragged", "evf_segment_mean",
                     "evf_stats_partial", "evf_normalize_inplace", "evf_energy_from_spec", "evf_log_compress",
                     "evf_last_error"):
        assert required in names


def test_library_exports_every_declared_symbol_and_nothing_else(lib):
    from everyvoice_b200 import _lib

    declared = set(_declared_symbols())
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    for name in declared:
        assert getattr(lib, name) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert declared <= exported
    # -fvisibility=hidden: the ABI is the header, no C++ or torch symbols leak
    assert {s for s in exported if s.startswith("evf_")} == declared
    assert not [s for s in exported if "torch" in s.lower() or "at::" in s]


def test_config_struct_layout_matches_header(lib):
    from everyvoice_b200 import _lib

    text = HEADER.read_text()
    body = re.search(r"typedef struct evf_config \{(.*?)\} evf_config;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"(int32_t|float)\s+(\w+)\s*;", body)
    assert [n for _, n in fields] == [n for n, _ in _lib.evf_config._fields_]
    for (ctype, name), (_, pytype) in zip(fields, _lib.evf_config._fields_):
        assert pytype is (C.c_int32 if ctype == "int32_t" else C.c_float), name
    assert C.sizeof(_lib.evf_config) == 4 * len(fields)
    status = dict(re.findall(r"(EVF_(?:OK|ERR_\w+))\s*=\s*(\d+)", text))
    for k, v in status.items():
        assert getattr(_lib, k) == int(v)
    for k, v in re.findall(r"(EVF_SPEC_\w+)\s*=\s*(\d+)", text):
        pyname = k[len("EVF_SPEC_"):].lower().replace("_", "-")
        assert _lib.SPEC_TYPES[pyname] == int(v)


def test_abi_version_and_pure_host_entry_points(lib):
    from everyvoice_b200 import _lib

    assert lib.evf_abi_version() == _lib.ABI_VERSION == int(re.search(r"#define EVF_ABI_VERSION (\d+)", HEADER.read_text()).group(1))
    assert lib.evf_plan_num_frames(None, 1000) == -1
    assert lib.evf_plan_destroy(None) == _lib.EVF_OK and lib.evf_batch_destroy(None) == _lib.EVF_OK
    rf = C.c_int32()
    assert lib.evf_plan_row_floats(None, C.byref(rf)) == _lib.EVF_ERR_INVALID_ARGUMENT
    assert b"null" in lib.evf_last_error()


def test_argument_validation_needs_no_device(lib):
    from everyvoice_b200 import _lib

    win = np.ones(1024, dtype=np.float32)
    handle = C.c_void_p()

    def create(**kw):
        cfg = dict(spec_type=2, sample_rate=22050, n_fft=1024, win_length=1024, hop_length=256, n_mels=80,
                   apply_log=1, keep_last_frame=0, sample_format=0, log_clip=1e-5)
        cfg.update(kw)
        c = _lib.evf_config(**cfg)
        return lib.evf_plan_create(C.byref(c), win.ctypes.data_as(C.c_void_p), None, 0, C.byref(handle))

    assert create(n_fft=0) == _lib.EVF_ERR_INVALID_ARGUMENT and b"n_fft" in lib.evf_last_error()
    assert create(n_fft=1000, win_length=1024) == _lib.EVF_ERR_INVALID_ARGUMENT   # win_length > n_fft
    assert create(fft_path=7) == _lib.EVF_ERR_INVALID_ARGUMENT
    assert create(spec_type=9) == _lib.EVF_ERR_UNSUPPORTED
    assert create(hop_length=0) == _lib.EVF_ERR_INVALID_ARGUMENT
    assert create(spec_type=0) == _lib.EVF_ERR_INVALID_ARGUMENT  # mel without a filterbank
    assert create(sample_format=5) == _lib.EVF_ERR_INVALID_ARGUMENT
    assert handle.value is None


def test_no_cpu_fallback_without_a_device(lib):
    """On a box without a GPU a valid plan request must FAIL (EVF_ERR_NO_DEVICE), never compute."""
    import torch

    from everyvoice_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU-only pass")
    cfg = _lib.evf_config(spec_type=2, sample_rate=22050, n_fft=1024, win_length=1024, hop_length=256, n_mels=80,
                          apply_log=1, keep_last_frame=0, sample_format=0, log_clip=1e-5)
    win = np.ones(1024, dtype=np.float32)
    handle = C.c_void_p()
    rc = lib.evf_plan_create(C.byref(cfg), win.ctypes.data_as(C.c_void_p), None, 0, C.byref(handle))
    assert rc == _lib.EVF_ERR_NO_DEVICE and handle.value is None
    assert b"no CPU path" in lib.evf_last_error()


def test_sass_is_sm100a_with_bulk_copy(lib):
    """The shipped binary holds sm_100a code only, and the feature kernel stages its input with
    the bulk-copy engine (UBLKCP = cp.async.bulk) and mbarriers (SYNCS)."""
    from everyvoice_b200 import _lib

    r = subprocess.run(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", r.stdout))
    assert archs == {"sm_100a"}, archs
    sass = subprocess.run(["cuobjdump", "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "features_kernel" in sass and "features_generic_kernel" in sass and "UBLKCP" in sass and "SYNCS" in sass
