"""The C-ABI shared library loads and exports every symbol include/ozl.h declares (no compute
calls: this runs without a GPU).  Also checks the product never reaches into oracle/."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ozl.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ozl_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from openzl_b200 import _lib
    assert sorted(_lib.EXPORTS) == header_symbols()


def test_library_exports_all_symbols():
    from openzl_b200 import _lib
    lib = _lib.load()
    for sym in header_symbols():
        assert hasattr(lib, sym), sym
    assert lib.ozl_version() >= 100
    assert lib.ozl_strerror(0) == b"ok" and lib.ozl_strerror(6).startswith(b"domain")
    assert [lib.ozl_curve_coord_limbs(c) for c in range(4)] == [6, 12, 4, 8]
    assert lib.ozl_curve_coord_limbs(9) == 0


def test_sass_is_sm100a_and_uses_wide_imad():
    so = os.path.join(ROOT, "openzl_b200", "libozl_b200.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "openzl_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn
                assert "oracle/" not in text or fn == "__init__.py", fn


def test_no_device_fails_loudly():
    """Without a CUDA device the context constructor raises (no silent CPU path)."""
    import openzl_b200
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present: covered by -m gpu tests")
    with pytest.raises(openzl_b200.OzlError):
        openzl_b200.Context(0)


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): exactly one JSON line on
    stdout with the contract's keys, whatever native code writes to file descriptor 1 meanwhile."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-log-n", "12"],
                       capture_output=True, text=True, cwd=root, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["gpu_launches"] == 0
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "config", "cpu_baseline", "e2e"):
        assert k in d
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0
