"""The C-ABI library builds, loads and exports exactly what include/mcg_b200.h declares.

No compute calls here: this file runs without a GPU.
"""
import ctypes
import os
import re

import pytest

from metacoag_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mcg_b200.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mcg_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib_path():
    return build.build_library()


def test_header_and_binding_agree():
    assert _declared() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in _declared():
        assert hasattr(lib, name), name


def test_kmer_table_is_host_arithmetic(lib_path):
    # mcg_kmer_table is the one entry point that needs no device
    import hashlib
    table = _lib.kmer_table()
    assert hashlib.sha256(table.tobytes()).hexdigest() == \
        "9a57f94a839828cfdd42650170aa4db6e761932f7724c8056ab4ea818d0bc72f"


def test_no_cpu_fallback(lib_path):
    lib = _lib.load_library()
    if lib.mcg_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.McgError):
        _lib.Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "metacoag_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "libmcg_oracle" not in text, f
