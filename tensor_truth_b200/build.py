"""Builds ``libtt_b200.so`` (the C-ABI CUDA library, ``include/tt_b200.h``) in-tree with nvcc for sm_100a.

Cross-compiles without a GPU.  The .so is git-ignored but travels with the tree to the GPU box.
Every translation unit is compiled on its own (in parallel) and the objects are linked once; no relocatable
device code (each kernel lives in one unit).
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libtt_b200.so")
STAMP = os.path.join(HERE, "build", "sources.sha256")
SOURCES = ["api.cu", "scan_simt.cu", "scan_tc.cu", "scan_tc2.cu", "scan_gemm.cu", "linear.cu", "attention.cu", "rescore.cu",
           "automerge.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set $NVCC)")


def _deps():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(HERE, "..", "include", "tt_b200.h")]


def source_digest() -> str:
    """sha256 over every file the library is built from (what ``smoke()`` compares the loaded .so against)."""
    h = hashlib.sha256()
    for d in _deps():
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def is_current() -> bool:
    """True when the .so on disk was built from exactly the sources on disk."""
    try:
        with open(STAMP) as f:
            return os.path.exists(LIB) and f.read().strip() == source_digest()
    except OSError:
        return False


def needs_build() -> bool:
    return not is_current()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and is_current():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    extra = ["-Xptxas", "-v"] if verbose else []
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        r = subprocess.run([nvcc] + NVCC_FLAGS + extra + ["-c", src, "-o", obj], cwd=CSRC, capture_output=True, text=True)
        return src, obj, r

    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, srcs))
    for src, _obj, r in results:
        if verbose or r.returncode:
            print(f"--- {src}\n{r.stdout}{r.stderr}")
        if r.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + [o for _s, o, _r in results])
    with open(STAMP, "w") as f:
        f.write(source_digest())
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
