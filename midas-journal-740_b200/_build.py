"""Build recipe of libcuberille_cuda.so (nvcc, sm_100a only, in-tree)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcuberille_cuda.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",            # the projection / position arithmetic must not be contracted
    "-Xcompiler", "-fPIC", "-shared",
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC))


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    inc = os.path.join(os.path.dirname(HERE), "include", "cuberille_c.h")
    return any(os.path.getmtime(s) > t for s in _sources() + [inc])


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library.  nvcc cross-compiles without a GPU."""
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB, os.path.join(CSRC, "cuberille_capi.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB
