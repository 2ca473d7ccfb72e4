"""Builds pgvectorscale_b200/libdiskann_b200.so (sm_100a only) with nvcc, in-tree."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libdiskann_b200.so")
SOURCES = ["diskann_b200.cu"]


def _dependencies():
    """Every source the library is compiled from: all of csrc/ plus the public header."""
    import glob
    deps = glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh"))
    deps += glob.glob(os.path.join(CSRC, "*.h"))    # dann_plan.h decides every launch shape, dann_coalescer.h is compiled in
    deps += glob.glob(os.path.join(_HERE, "..", "include", "*.h"))
    deps.append(os.path.abspath(__file__))      # flag changes rebuild too
    return deps


NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",   # Blackwell B200 only, no PTX fallback for other parts
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",          # never contract a*b+c: rerank sums must round like the reference's AVX2 code
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libdiskann_b200.so")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in _dependencies())


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library if it is missing or older than its sources."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS]
    # experiment knob: extra -D... for this build only, e.g. DANN_NVCC_DEFINES="-DDANN_HV_NO_FUSED" with force=True
    cmd += [d for d in os.environ.get("DANN_NVCC_DEFINES", "").split() if d.startswith("-D")]
    # development loops only: e.g. DANN_NVCC_EXTRA="-split-compile=8" builds in ~40 s instead of ~90 s, but ptxas then
    # schedules differently - the result is NOT the measured binary (tools/sass_signature.py shows it)
    cmd += os.environ.get("DANN_NVCC_EXTRA", "").split()
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", LIB_PATH, *[os.path.join(CSRC, s) for s in SOURCES]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
