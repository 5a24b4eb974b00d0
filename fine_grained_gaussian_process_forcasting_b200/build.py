"""In-tree build of libgpblur.so (the C-ABI CUDA library) for sm_100a.

``python -m fine_grained_gaussian_process_forcasting_b200.build`` compiles every ``csrc/*.cu`` with
``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` (object files in parallel) and links
``fine_grained_gaussian_process_forcasting_b200/libgpblur.so``.  The library has no torch dependency.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "csrc" / "_obj"
LIB = PKG / "libgpblur.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _flags():
    """GPBLUR_TRACE=1 compiles the clock64 event trace of the tensor-core point kernels in (scripts/tc2_trace.py)."""
    return NVCC_FLAGS + (["-DGPBLUR_TRACE=1"] if os.environ.get("GPBLUR_TRACE") == "1" else [])


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(_flags()).encode())
    return h.hexdigest()


def sources():
    return sorted(CSRC.glob("*.cu"))


def headers():
    return sorted(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "gpblur.h"]


def build(force: bool = False, verbose: bool = False) -> Path:
    """Build libgpblur.so if sources changed.  Returns the library path."""
    srcs = sources()
    stamp = OBJ / "stamp.sha256"
    want = _digest(srcs + headers())
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == want:
        return LIB
    OBJ.mkdir(parents=True, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src: Path):
        obj = OBJ / (src.stem + ".o")
        cmd = [nvcc, *_flags(), "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (OBJ / (src.stem + ".ptxas.log")).write_text(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stderr[-4000:]}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stderr[-4000:]}")
    stamp.write_text(want)
    if verbose:
        print(f"built {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
