"""The C-ABI library loads without a GPU and exports every symbol include/bopy_b200.h declares.
No compute entry is called here."""
import ctypes
import os
import re

import pytest

from bopy_b200 import _native, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    return build.build()          # compiles with nvcc if missing/stale (cross-compiles without a GPU)


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "bopy_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bopy_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_native.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_abi_version_and_error_string(lib_path):
    lib = _native.load()
    assert lib.bopy_abi_version() == _native.ABI_VERSION
    h = ctypes.c_void_p()
    # argument checking happens before any CUDA call, so this is safe without a device
    assert lib.bopy_gp_create(ctypes.byref(h), 0, 0, 0, 0, 1) == _native.ERR_BAD_ARG
    assert b"n must be >= 1" in lib.bopy_last_error()
    assert lib.bopy_gp_create(None, 0, 0, 0, 1, 1) == _native.ERR_BAD_ARG


def test_built_for_sm_100a_with_tma_bulk_copies(lib_path):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    elf = subprocess.run([cuobjdump, "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in elf
    sass = subprocess.run([cuobjdump, "-sass", lib_path], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "DFMA" in sass and "SYNCS" in sass   # bulk async copies + mbarriers + fp64 FMA
    assert "DMMA" in sass                                            # fp64 engine: mma.sync.m8n8k4.f64
    # fp32 engine: tcgen05.mma kind::tf32 (UTCHMMA), its commits (UTCBAR) and the TMEM read-back (LDTM)
    assert "UTCHMMA" in sass and "UTCBAR" in sass and "LDTM" in sass
