"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/moped_cuda.h
declares; without a device it fails loudly instead of falling back to anything."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "moped_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from moped_b200 import build, capi
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in moped_cuda.h but not exported"
    assert sorted(capi.SIGNATURES) == declared, "capi.SIGNATURES and moped_cuda.h disagree"


def test_product_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under moped_b200/ (nor the C ABI) may reference it."""
    for base, _, files in os.walk(os.path.join(ROOT, "moped_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "moped_oracle.h" not in txt and "libmoped_ref" not in txt, f


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from moped_b200 import capi
    with pytest.raises(capi.MopedCudaError):
        capi.Context(0)
    lib = capi.load()
    assert b"sm_100a" in lib.mc_version()


def test_sass_has_blackwell_tensor_and_tma_instructions():
    """UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (B200_PROFILING.md)."""
    import shutil
    import subprocess
    from moped_b200 import build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", build.build()], capture_output=True, text=True).stdout
    for mnem in ("UTCHMMA", "LDTM", "UBLKCP"):
        assert mnem in sass, mnem


def test_every_option_key_is_documented_in_the_header():
    """mc_set_option keys accepted by api.cu == keys described in include/moped_cuda.h."""
    api = open(os.path.join(ROOT, "moped_b200", "csrc", "api.cu")).read()
    body = api[api.index("mc_status mc_set_option("):]
    body = body[:body.index("return MC_OK;")]
    accepted = set(re.findall(r'k == "([a-z0-9_]+)"', body))
    header = open(os.path.join(ROOT, "include", "moped_cuda.h")).read()
    documented = set(re.findall(r'^\s*\*\s+"([a-z0-9_]+)"\s', header, flags=re.M))
    assert accepted and accepted <= documented, sorted(accepted - documented)
