"""The C-ABI libraries load and export every symbol include/alis_b200.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

from alis_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "alis_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(alisb_[A-Za-z_0-9]+)\s*\(", src)))


def test_header_declares_the_bound_symbols():
    names = declared_symbols()
    assert len(names) >= 30
    capi.load_oracle()
    assert set(capi.ABI_SYMBOLS) == set(names), set(capi.ABI_SYMBOLS) ^ set(names)


def test_cuda_engine_exports_every_declared_symbol():
    assert os.path.exists(capi.ENGINE_SO), "CUDA engine not built (run __graft_entry__.build())"
    lib = ctypes.CDLL(capi.ENGINE_SO)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.alisb_abi_version() == 1
    lib.alisb_backend.restype = ctypes.c_char_p
    assert lib.alisb_backend() == b"cuda-sm_100a"


def test_oracle_exports_every_declared_symbol():
    lib = ctypes.CDLL(capi.ORACLE_SO)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    lib.alisb_backend.restype = ctypes.c_char_p
    assert lib.alisb_backend() == b"oracle-cpu"


def test_engine_so_contains_sm100a_code_only():
    """The product library carries sm_100a SASS (checked with cuobjdump when available)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        return
    out = subprocess.run([cuobjdump, "-lelf", capi.ENGINE_SO], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out
