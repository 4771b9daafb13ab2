"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/sister_b200.h declares.
No compute call is made here (CPU-only container)."""
import ctypes
import os
import re

import pytest

import sister_b200
from sister_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    sister_b200.build_library()
    return ctypes.CDLL(sister_b200.library_path())


def declared_functions():
    text = open(os.path.join(ROOT, "include", "sister_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sister_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_functions() == sorted(api.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_functions():
        assert hasattr(lib, name), name


def test_version_and_strerror(lib):
    lib.sister_version.restype = ctypes.c_int
    assert lib.sister_version() == 100
    lib.sister_strerror.restype = ctypes.c_char_p
    assert b"shape" in lib.sister_strerror(-2)


def test_no_silent_fallback_without_gpu(lib):
    """On a box without a CUDA device the product must fail loudly, not compute on the CPU."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(sister_b200.SisterError) as e:
        sister_b200.Engine(64, 48, 16)
    assert e.value.code == -5


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under sister_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sister_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src and "libsister_oracle" not in src \
                    and "libsister_ref" not in src, f
