"""The C ABI used from plain C (examples/ozl_msm_demo.c), with no Python between the caller and
libozl_b200.so -- the position the Rust shim's FFI is in.  CPU: it builds against include/ozl.h and
fails loudly (exit 3, OZL_ERR_NO_DEVICE) because there is no CPU fallback.  GPU: its two
self-checks (MSM against a fixed-base multiple, NTT round trip) pass."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "openzl_b200")


def _build(tmp_path):
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    exe = str(tmp_path / "ozl_msm_demo")
    subprocess.check_call([cc, "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "ozl_msm_demo.c"), "-L", LIBDIR, "-lozl_b200",
                           f"-Wl,-rpath,{LIBDIR}", "-o", exe])
    return exe


def test_c_demo_builds_and_fails_loudly_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3
    assert "no usable CUDA device" in r.stderr


@pytest.mark.gpu
def test_c_demo_on_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, "200000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "MISMATCH" not in r.stdout and r.stdout.count(": ok") == 2
