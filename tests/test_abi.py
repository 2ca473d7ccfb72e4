"""The C-ABI library loads and exports every symbol include/diskann_b200.h declares, and the
product path fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "diskann_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(dann_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_symbols_are_exported(lib_built):
    from pgvectorscale_b200 import diskann
    lib = C.CDLL(lib_built)
    declared = _declared_functions()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/diskann_b200.h but not exported"
    assert sorted(diskann.EXPORTS) == declared, "diskann.EXPORTS out of sync with the header"


def test_library_is_sm100a_only(lib_built):
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", lib_built], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_no_cpu_fallback_without_device(lib_built):
    from pgvectorscale_b200 import diskann
    if diskann.device_count() > 0:
        pytest.skip("a CUDA device is visible; the no-device behaviour is checked on the CPU box")
    from pgvectorscale_b200.snapshot import Snapshot, make_heap_tids
    n, dim = 4, 8
    s = Snapshot(n=n, dim=dim, dim_index=dim, bits=2, words=1, R=4, distance_type=1, has_labels=False,
                 count=n, mean=np.zeros(dim, np.float32), m2=np.ones(dim, np.float32),
                 codes=np.zeros((n, 1), np.uint64), nbrs=np.full((n, 4), 0xFFFFFFFF, np.uint32),
                 heap_tid=make_heap_tids(n), vectors=np.zeros((n, dim), np.float32), start_default=0)
    with pytest.raises(diskann.DiskAnnError) as e:
        diskann.DiskAnnIndex(s)
    assert e.value.code == -3          # DANN_ERR_NO_DEVICE
    assert "no CPU path" in str(e.value)


def _tiny_snapshot(**over):
    from pgvectorscale_b200.snapshot import Snapshot, make_heap_tids
    n, dim = 4, 8
    kw = dict(n=n, dim=dim, dim_index=dim, bits=2, words=1, R=4, distance_type=1, has_labels=False,
              count=n, mean=np.zeros(dim, np.float32), m2=np.ones(dim, np.float32),
              codes=np.zeros((n, 1), np.uint64), nbrs=np.full((n, 4), 0xFFFFFFFF, np.uint32),
              heap_tid=make_heap_tids(n), vectors=np.zeros((n, dim), np.float32), start_default=0)
    kw.update(over)
    return Snapshot(**kw)


@pytest.mark.parametrize("over,msg", [
    (dict(nbrs=np.array([[1, 2, 0xFFFFFFFF, 9]] * 4, np.uint32)), None),            # ids after the sentinel are ignored
    (dict(nbrs=np.array([[1, 4, 0xFFFFFFFF, 0]] * 4, np.uint32)), "outside the index"),
    (dict(start_default=4), "start_default"),
    (dict(has_labels=True, label_off=np.array([0, 2, 1, 3, 3], np.uint32), labels=np.array([1, 2, 3], np.int16),
          start_labels=np.array([1], np.int16), start_label_nodes=np.array([0], np.uint32)), "monotone"),
    (dict(has_labels=True, label_off=np.array([0, 1, 2, 3, 3], np.uint32), labels=np.array([1, 2, 3], np.int16),
          start_labels=np.array([1], np.int16), start_label_nodes=np.array([7], np.uint32)), "start node"),
])
def test_snapshot_is_validated_on_the_host_before_any_copy(lib_built, over, msg):
    """Out-of-range ids/offsets would become out-of-bounds gathers in the kernels: dann_index_load rejects them
    on the host (so this runs without a GPU: a well-formed snapshot gets as far as DANN_ERR_NO_DEVICE)."""
    from pgvectorscale_b200 import diskann
    if diskann.device_count() > 0 and msg is None:
        diskann.DiskAnnIndex(_tiny_snapshot(**over)).close()
        return
    with pytest.raises(diskann.DiskAnnError) as e:
        diskann.DiskAnnIndex(_tiny_snapshot(**over))
    if msg is None:
        assert e.value.code == -3
    else:
        assert e.value.code == -1 and msg in str(e.value)


def test_missing_library_raises(tmp_path):
    from pgvectorscale_b200 import diskann
    with pytest.raises(diskann.DiskAnnError):
        diskann.load_library(str(tmp_path / "nope.so"))


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under pgvectorscale_b200/ may reference it."""
    pkg = os.path.join(ROOT, "pgvectorscale_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "oracle.h" not in text, f
                # ... nor the CPU SIMT emulator / the emulated ABI build (tests/simt): the only trace allowed in the
                # product sources is the DANN_SIMT_EMU preprocessor guard around PTX and shared-memory declarations
                assert "simt_emu.h" not in text and "libdiskann_b200_emu" not in text and "fake_cuda" not in text, f
                assert not re.search(r"^\s*(from|import)\s+(emu|build_emu|cu2cpp)\b", text, flags=re.M), f


def test_plain_c_program_links_and_fails_loudly_without_a_device(lib_built, tmp_path):
    """include/diskann_b200.h is a self-contained C99 header; harness/executor_harness.c drives the scan operator
    the way the pgrx shim would.  Without a GPU it must report DANN_ERR_NO_DEVICE (exit code 3), not crash."""
    import shutil
    import subprocess
    from pgvectorscale_b200 import diskann
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    exe = str(tmp_path / "executor_harness")
    libdir = os.path.dirname(lib_built)
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "harness", "executor_harness.c"), "-L" + libdir, "-ldiskann_b200",
                    "-Wl,-rpath," + libdir, "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    if diskann.device_count() > 0:
        assert r.returncode == 0, r.stderr
        assert "rows" in r.stdout
    else:
        assert r.returncode == 3
        assert "no CPU path" in r.stderr
