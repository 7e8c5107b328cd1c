"""Builds ``libevfeat.so`` (hand-written sm_100a CUDA + the C ABI of ``include/evfeat.h``)
in-tree with nvcc.  ``python -m everyvoice_b200.build`` or ``__graft_entry__.build()``."""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "libevfeat.so"
SOURCES = ["evfeat_api.cu", "evfeat_features.cu", "evfeat_generic.cu", "evfeat_decimated.cu", "evfeat_aux.cu", "evfeat_audio.cu", "evfeat_backward.cu", "evfeat_pitch.cu"]
HEADERS = ["evfeat_internal.h", "evfeat_fft.cuh", "evfeat_device.cuh"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--cudart", "static",
    "-Xptxas", "-v",
]


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        raise RuntimeError("nvcc not found: libevfeat.so cannot be built (there is no other code path)")
    return nvcc


def is_stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES + HEADERS] + [ROOT / "include" / "evfeat.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines: list[str] | None = None, out: Path | None = None) -> Path:
    """``defines`` / ``out`` build an experimental variant next to the product library
    (kernel A/B runs: ``EVF_LIB=<path>`` makes ``_lib.load`` pick it up)."""
    global LIB
    if out is None and not force and not is_stale():
        return LIB
    nvcc = find_nvcc()
    objs = []
    build_dir = PKG / ("build" if out is None else "build_" + Path(out).stem)
    build_dir.mkdir(exist_ok=True)
    lib_out = LIB if out is None else Path(out)
    extra = [f"-D{d}" for d in (defines or [])]
    procs = []
    for s in SOURCES:
        obj = build_dir / (Path(s).stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-I", str(ROOT / "include"), "-I", str(CSRC), "-c", str(CSRC / s), "-o", str(obj)]
        procs.append((s, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    log = []
    for s, cmd, p in procs:
        out, _ = p.communicate()
        log.append(f"$ {' '.join(cmd)}\n{out}")
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {s}")
    cmd = [nvcc, "-shared", "--cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
           "-o", str(lib_out), *objs]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.append(f"$ {' '.join(cmd)}\n{r.stdout}")
    (build_dir / "build.log").write_text("\n".join(log))
    if r.returncode != 0:
        sys.stderr.write("\n".join(log))
        raise RuntimeError("link of libevfeat.so failed")
    if verbose:
        print("\n".join(log))
    return lib_out


if __name__ == "__main__":
    # python -m everyvoice_b200.build [-DNAME[=V] ...] [-o path/to/libvariant.so]
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outp = Path(sys.argv[sys.argv.index("-o") + 1]).resolve() if "-o" in sys.argv else None
    lib = build(force=True, verbose="-q" not in sys.argv, defines=defs, out=outp)
    print(f"built {lib} ({os.path.getsize(lib)} bytes)")
