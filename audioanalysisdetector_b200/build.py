"""In-tree build of libaad_b200.so (hand-written sm_100a CUDA behind a C ABI).

`python -m audioanalysisdetector_b200.build [--force] [-v] [--out PATH] [--flags "..."]` or `build_library()`.
nvcc cross-compiles for sm_100a without a GPU; the .so is git-ignored but travels with the tree to the GPU box.
The translation units are compiled in parallel (k_stft_fb is instantiated per transform size in its own
unit: aad_stft_inst.cu with -DAAD_INST_L=4|8|16|32) and linked into one shared library.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libaad_b200.so")
# (source, extra defines)
UNITS = [("aad_api.cu", []), ("aad_detector.cu", []), ("aad_flac.cpp", []), ("aad_cqcc.cu", [])] + [("aad_stft_inst.cu", [f"-DAAD_INST_L={L}"]) for L in (32, 8, 16, 4)]
SOURCES = sorted({os.path.join(CSRC, u) for u, _ in UNITS})
HEADERS = [os.path.join(CSRC, "aad_kernels.cuh"), os.path.join(CSRC, "aad_fft.cuh"), os.path.join(CSRC, "aad_stft_inst.h"),
           os.path.join(os.path.dirname(PKG_DIR), "include", "aad.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "--expt-relaxed-constexpr", "-O3", "-lineinfo",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libaad_b200.so")


def needs_build(lib_path: str = LIB_PATH) -> bool:
    if not os.path.exists(lib_path):
        return True
    t = os.path.getmtime(lib_path)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS)


def build_library(force: bool = False, verbose: bool = False, extra_flags=(), out: str = LIB_PATH) -> str:
    if not force and not needs_build(out):
        return out
    nvcc = _nvcc()
    common = NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else [])
    with tempfile.TemporaryDirectory(prefix="aad_build_") as tmp:
        jobs = []
        for i, (src, defs) in enumerate(UNITS):
            obj = os.path.join(tmp, f"u{i}.o")
            jobs.append((obj, [nvcc] + common + defs + ["-c", "-o", obj, os.path.join(CSRC, src)]))

        def run(job):
            return subprocess.run(job[1], capture_output=True, text=True)

        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            results = list(ex.map(run, jobs))
        log = "".join(r.stdout + r.stderr for r in results)
        if any(r.returncode != 0 for r in results):
            raise RuntimeError("nvcc failed:\n" + log)
        link = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + [j[0] for j in jobs],
                              capture_output=True, text=True)
        if link.returncode != 0:
            raise RuntimeError("link failed:\n" + link.stdout + link.stderr)
        if verbose:
            sys.stderr.write(log)
    return out


if __name__ == "__main__":
    argv = sys.argv[1:]
    out = argv[argv.index("--out") + 1] if "--out" in argv else LIB_PATH
    flags = argv[argv.index("--flags") + 1].split() if "--flags" in argv else []
    print(build_library(force="--force" in argv or "--out" in argv, verbose="-v" in argv, extra_flags=flags, out=out))
