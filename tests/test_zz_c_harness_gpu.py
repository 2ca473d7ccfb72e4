"""Runs last in the GPU suite (file name sorts after the other tests): the plain-C executor harness."""
import numpy as np  # noqa: F401
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib(lib_built):
    from pgvectorscale_b200 import diskann
    if diskann.device_count() < 1:
        pytest.fail("no CUDA device visible: -m gpu tests need the B200 box")
    return diskann


def test_plain_c_executor_harness_runs_a_scan(lib, lib_built, tmp_path):
    """The C99 harness (the pgrx shim's call sequence) loads an index and streams every row of a scan."""
    import os
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    exe = str(tmp_path / "executor_harness")
    libdir = os.path.dirname(lib_built)
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(root, "include"),
                    os.path.join(root, "harness", "executor_harness.c"), "-L" + libdir, "-l:" + os.path.basename(lib_built),
                    "-Wl,-rpath," + libdir, "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rows; visits=" in r.stdout
