"""In-tree build of libaad_b200.so (hand-written sm_100a CUDA behind a C ABI).

`python -m audioanalysisdetector_b200.build [--force]` or `build_library()`.
nvcc cross-compiles for sm_100a without a GPU; the .so is git-ignored but travels
with the tree to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libaad_b200.so")
SOURCES = [os.path.join(CSRC, "aad_api.cu"), os.path.join(CSRC, "aad_detector.cu")]
HEADERS = [os.path.join(CSRC, "aad_kernels.cuh"), os.path.join(CSRC, "aad_fft.cuh"),
           os.path.join(os.path.dirname(PKG_DIR), "include", "aad.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "--expt-relaxed-constexpr", "-O3", "-lineinfo",
    "-shared", "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libaad_b200.so")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        sys.stderr.write(proc.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
