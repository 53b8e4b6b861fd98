"""CPU-side checks: the C-ABI library loads, exports every symbol include/jmb200.h declares, and fails
loudly (no CPU path) when no GPU is present.  No compute calls."""
import ctypes
import os
import re

import pytest

from jm_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "jmb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jmb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(api.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(api.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/jmb200.h but not exported"
    assert lib.jmb_abi_version() == 2


def test_struct_layouts_match_header():
    assert api.ME_REQ.itemsize == 40 and api.ME_RES.itemsize == 24
    assert api.MB_PRED.itemsize == 72 and api.QUANT_DESC.itemsize == 980
    assert api.ME_REQ.fields["min_mcost"][1] == 32 and api.ME_REQ.fields["lambda"][1] == 16


def test_no_cpu_fallback():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(api.JMBError, match="no CUDA device"):
        api.Context(0)


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ (checker only)."""
    for dp, _, files in os.walk(os.path.join(ROOT, "jm_b200")):
        if "_build" in dp or "/lib" in dp:
            continue
        for f in files:
            if f.endswith((".py", ".c", ".h", ".cu", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "pyoracle" not in txt and "jm_oracle" not in txt and "libjmref" not in txt, os.path.join(dp, f)


def test_canonical_partition_order():
    parts = api.mb_partitions()
    assert len(parts) == 41 and parts[0] == (1, 0, 0) and parts[-1] == (7, 12, 12)


def test_c_example_builds_against_the_header_and_library(tmp_path):
    """examples/picture_form.c is the picture form of the ABI from plain C (no Python, no torch): it must compile with a C
    compiler against include/jmb200.h alone and link against libjmb200.so; without a GPU it stops at jmb_create."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "picture_form")
    libdir = os.path.dirname(api.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "picture_form.c"),
                        "-L", libdir, "-ljmb200", f"-Wl,-rpath,{libdir}", "-lm", "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1500:]
    try:
        import torch
        if torch.cuda.is_available():
            return
    except ImportError:
        pass
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU path" in r.stderr, (r.returncode, r.stderr[-300:])
