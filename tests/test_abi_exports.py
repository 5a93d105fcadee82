"""The C-ABI library loads and exports every symbol include/gunship_collision.h declares; the Python
prototypes cover the same set; without a GPU the product path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from gunship_rs_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gunship_collision.h")


def _header_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gsc_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_python_binds():
    assert _header_functions() == sorted(N.EXPORTS)


def test_library_exports_every_header_symbol():
    if not os.path.exists(N.LIB_PATH):
        pytest.fail(f"{N.LIB_PATH} not built: run __graft_entry__.build()")
    lib = C.CDLL(N.LIB_PATH)
    for name in _header_functions():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert N.load().gsc_abi_version() == 1


def test_struct_layouts_match_header():
    # sizes as the C compiler lays them out (x86-64): see include/gunship_collision.h
    assert C.sizeof(N.GscConfig) == 40
    assert C.sizeof(N.GscStepOut) == 8 * 5 + 4 * 2 + 12 + 12 + 4 * 2 + 4 * 6 + 4 + 4 + 8   # 120: d_pairs is 8-aligned
    assert C.sizeof(N.GscBounds) == 36


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from gunship_rs_b200 import context as G
    with pytest.raises(N.GscError) as ei:
        G.CollisionContext(16)
    assert ei.value.code == N.GSC_ERR_CUDA and "no CPU fallback" in str(ei.value)
    with pytest.raises(N.GscError):
        G.test_sphere_sphere([[0, 0, 0, 1]], [[0, 0, 0, 1]])


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gunship_rs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in text and "gunship_oracle" not in text, f
