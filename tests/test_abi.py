"""CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/skirtgpu.h declares; without a CUDA device the engine refuses to start (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import common
import skirt_b200 as sk

HEADER = os.path.join(common.ROOT, "include", "skirtgpu.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(skg_[a-z_0-9]+)\s*\(", text)))


def test_library_is_built():
    assert sk.lib_available(), "skirt_b200/libskirtgpu.so missing: run `make` or __graft_entry__.build()"


def test_every_declared_symbol_is_exported():
    lib = ctypes.CDLL(sk.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in skirtgpu.h but not exported: {missing}"


def test_version_and_error_string():
    lib = sk.load_library()
    assert lib.skg_version() >= 1
    assert isinstance(lib.skg_last_error(), bytes)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(sk.EngineError, match="no CPU fallback|CUDA"):
        sk.Engine(0)


def test_product_never_touches_the_oracle():
    """the product path (skirt_b200/, include/) must not import, link or call anything under oracle/"""
    bad = []
    for base, _, files in os.walk(os.path.join(common.ROOT, "skirt_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"(from|import)\s+oracle|oracle/|liboracle|libskirtref", text):
                    bad.append(f)
    assert not bad, bad
    ldd = os.popen(f"ldd {sk.LIB_PATH}").read()
    assert "oracle" not in ldd and "skirtref" not in ldd
