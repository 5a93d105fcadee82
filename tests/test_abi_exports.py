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
    assert N.load().gsc_abi_version() == 2


def test_struct_layouts_match_header(tmp_path):
    """sizeof / offsetof as the C compiler sees the header vs the ctypes mirrors"""
    import subprocess
    src = tmp_path / "layout.c"
    src.write_text('''#include <stdio.h>
#include <stddef.h>
#include "gunship_collision.h"
int main(void) {
    printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(gsc_config), sizeof(gsc_step_out), sizeof(gsc_bounds),
           offsetof(gsc_step_out, cell_size), offsetof(gsc_step_out, stage_ms), offsetof(gsc_step_out, d_pairs),
           offsetof(gsc_bounds, home_x_max));
    return 0;
}''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(N.GscConfig), C.sizeof(N.GscStepOut), C.sizeof(N.GscBounds), N.GscStepOut.cell_size.offset,
            N.GscStepOut.stage_ms.offset, N.GscStepOut.d_pairs.offset, N.GscBounds.home_x_max.offset]
    assert got == want


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


def test_rust_ffi_declares_every_header_function():
    """rust/gunship-collision/src/ffi.rs cannot be compiled here (no rustc); at least keep its extern "C" block in
    step with the header"""
    ffi = open(os.path.join(ROOT, "rust", "gunship-collision", "src", "ffi.rs")).read()
    declared = set(re.findall(r"pub fn (gsc_[a-z0-9_]+)\s*\(", ffi))
    assert sorted(declared) == _header_functions()
