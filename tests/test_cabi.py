"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/genima_b200.h declares; the Python
binding table (genima_b200/_cabi.py) lists exactly the same set.  No compute call is made here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    with open(os.path.join(ROOT, "include", "genima_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gn_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    from genima_b200 import build

    return build.build()


def test_header_and_binding_table_agree():
    from genima_b200 import _cabi

    assert _header_symbols() == sorted(_cabi.SIGNATURES)


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in _header_symbols():
        assert hasattr(lib, name), f"{name} is declared in include/genima_b200.h but not exported"


def test_version_and_no_driver_error_code(lib_path):
    import torch

    from genima_b200 import _cabi

    lib = _cabi.load_library(lib_path)
    assert b"sm_100a" in lib.gn_version()
    if not torch.cuda.is_available():
        h = ctypes.c_void_p()
        assert lib.gn_create(0, ctypes.byref(h)) == _cabi.GN_ERR_NODRIVER
        assert not h.value


def test_epilogue_struct_layout_matches_header():
    from genima_b200._cabi import GnEpilogue

    # 4 pointers, int64, 3 x int32, 2 x float, 2 x int32, pad, 2 pointers, int32, float, pointer, 2 x int32, pointer,
    # 2 x int32
    assert ctypes.sizeof(GnEpilogue) == 4 * 8 + 8 + 3 * 4 + 2 * 4 + 2 * 4 + 4 + 2 * 8 + 4 + 4 + 8 + 2 * 4 + 8 + 2 * 4
    assert GnEpilogue.ldr.offset == 32 and GnEpilogue.ln_stats.offset == 72 and GnEpilogue.rowstats_out.offset == 96
    assert GnEpilogue.gn_bucket.offset == 108 and GnEpilogue.gnstats_out.offset == 112


def test_sass_uses_blackwell_tensor_and_tma_paths(lib_path):
    """The GEMM / attention kernels must be tcgen05 + TMA code (UTC*MMA / UTMALDG in SASS), not legacy HMMA."""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", lib_path], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass
    assert "HMMA." not in sass.replace("UTCHMMA", "")
