"""Builds the CPU SIMT-emulation harness (tests/simt/*.cpp + the kernel sources) with g++.  Test infrastructure."""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "pgvectorscale_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libemu_search.so")


def sources():
    return [os.path.join(HERE, "emu_search.cpp"), os.path.join(HERE, "emu_kernels.cpp"), os.path.join(HERE, "simt_emu.cpp")]


def is_stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = sources() + glob.glob(os.path.join(HERE, "*.h")) + glob.glob(os.path.join(HERE, "shim", "*.h")) + \
        glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(ROOT, "include", "*.h")) + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False):
    if not force and not is_stale():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-DDANN_SIMT_EMU", "-Wall", "-Wno-unknown-pragmas",
           "-Wno-unused-function", "-Wno-unused-variable", "-Wno-sign-compare", "-fno-omit-frame-pointer", "-ffp-contract=off",
           "-I", os.path.join(HERE, "shim"), "-I", CSRC, "-x", "c++"] + sources() + ["-o", OUT]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
