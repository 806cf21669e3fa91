"""Build libpvqt.so (the C-ABI library, include/pvqt.h) for sm_100a, in-tree.

    python -m pitchvis_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The output, pitchvis_b200/lib/libpvqt.so, is
git-ignored but travels with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OUT_DIR = os.path.join(PKG, "lib")
BUILD_DIR = os.path.join(PKG, "_build")
LIB = os.path.join(OUT_DIR, "libpvqt.so")

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
              "--expt-relaxed-constexpr"]
# the analysis epilogue mirrors un-contracted f32 expressions of the reference (see the file header)
PER_FILE_FLAGS = {"analysis_kernels.cu": ["-fmad=false"], "agc_kernels.cu": ["-fmad=false"],
                  "chroma_kernels.cu": ["-fmad=false"], "spectrogram_kernels.cu": ["-fmad=false"]}
# the kernel builder mirrors the reference's f32 op order: no FMA contraction on the host
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-Wall", "-Wextra"]


def _sources():
    cu = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    cpp = sorted(f for f in os.listdir(CSRC) if f.endswith(".cpp"))
    return cu, cpp


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    os.makedirs(BUILD_DIR, exist_ok=True)
    inc = ["-I", os.path.join(ROOT, "include"), "-I", CSRC]
    cu, cpp = _sources()
    objs = []
    log = []
    for f in cpp:
        o = os.path.join(BUILD_DIR, f + ".o")
        cmd = ["g++", *CXX_FLAGS, *inc, "-c", os.path.join(CSRC, f), "-o", o]
        log.append(subprocess.run(cmd, check=True, capture_output=True, text=True).stderr)
        objs.append(o)
    for f in cu:
        o = os.path.join(BUILD_DIR, f + ".o")
        cmd = [NVCC, *ARCH, *NVCC_FLAGS, *PER_FILE_FLAGS.get(f, []), *inc, "-c", os.path.join(CSRC, f), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(r.stdout + r.stderr)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise subprocess.CalledProcessError(r.returncode, cmd)
        objs.append(o)
    cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs]
    subprocess.run(cmd, check=True)
    with open(os.path.join(BUILD_DIR, "build.log"), "w") as fh:
        fh.write("\n".join(log))
    if verbose:
        sys.stderr.write("\n".join(log))
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
