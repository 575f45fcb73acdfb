"""Builds libtokred_sm100a.so in-tree with nvcc (sm_100a only).  No torch headers, no pybind: the library is a
plain C-ABI shared object (include/tokred.h) that _lib.py loads with ctypes.

    python -m tokenreduction_b200.build [--force]
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(PKG_DIR, "build")
LIB_PATH = os.path.join(PKG_DIR, "libtokred_sm100a.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; libtokred_sm100a.so cannot be built")
    return cand


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    files = _sources() + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    files.append(os.path.join(os.path.dirname(PKG_DIR), "include", "tokred.h"))
    for f in files:
        h.update(f.encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False, stamps: bool = False) -> str:
    """stamps=True builds the bring-up variant libtokred_sm100a_dbg.so (-DTOKRED_STAMPS: clock64 phase stamps, see
    csrc/common.cuh and tools/diag/); the release library contains no stamp code."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    lib_path = LIB_PATH.replace(".so", "_dbg.so") if stamps else LIB_PATH
    extra = (["-DTOKRED_STAMPS"] + os.environ.get("TOKRED_DBG_FLAGS", "").split()) if stamps else []
    stamp = os.path.join(OBJ_DIR, "stamp_dbg.txt" if stamps else "stamp.txt")
    digest = _digest()
    if not force and os.path.exists(lib_path) and os.path.exists(stamp) and open(stamp).read() == digest:
        return lib_path
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ("_dbg.o" if stamps else ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", src, "-o", obj] + (["-Xptxas", "-v"] if verbose else [])
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [nvcc, "-shared", "-o", lib_path, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as fh:
        fh.write(digest)
    return lib_path


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv, stamps="--stamps" in sys.argv))
