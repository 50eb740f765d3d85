"""Build libsqsv.so (the sm_100a CUDA library behind include/sqsv.h) in-tree with nvcc.

The library is compiled for sm_100a only (B200); nvcc cross-compiles without a GPU.  The .so is
git-ignored but travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libsqsv.so")
STAMP = os.path.join(CSRC, ".libsqsv.stamp")
SOURCES = ["sqsv_space.cu", "sqsv_kernels.cu", "sqsv_api.cu", "sqsv_hamiltonian.cu", "sqsv_quad.cu", "sqsv_win.cu", "sqsv_win3.cu", "sqsv_reshard.cu", "sqsv_dmma.cu"]
HEADERS = [os.path.join(CSRC, "sqsv_internal.h"), os.path.join(ROOT, "include", "sqsv.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _digest() -> str:
    h = hashlib.sha256()
    for p in [os.path.join(CSRC, s) for s in SOURCES] + HEADERS:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources into csrc/libsqsv.so; returns the library path."""
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == digest:
                return LIB
    cmd = [
        _nvcc(),
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-O3", "-lineinfo", "-std=c++17",
        "-shared", "-Xcompiler", "-fPIC",
        "-I", os.path.join(ROOT, "include"), "-I", CSRC,
        "-o", LIB,
    ] + [os.path.join(CSRC, s) for s in SOURCES] + ["-lcudart", "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libsqsv.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    with open(STAMP, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
