"""The C-ABI library builds, loads and exports every symbol include/vmas_b200.h declares."""
import ctypes
import os
import re

from vectorizedmultiagentsimulator_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vmas_b200.h")).read()
    return sorted(set(re.findall(r"\b(vmas_b200_\w+)\s*\(", text)))


def test_header_symbols_are_exported():
    _native.build()
    lib = ctypes.CDLL(_native.LIB_PATH)
    symbols = declared_symbols()
    assert len(symbols) >= 8
    for name in symbols:
        assert hasattr(lib, name), f"{name} declared in vmas_b200.h but not exported"
    assert sorted(_native.EXPORTS) == symbols


def test_abi_version_and_struct_sizes():
    lib = _native.load()
    assert lib.vmas_b200_abi_version() == 1
    # 10 int32 + 9 float
    assert ctypes.sizeof(_native.WorldConfig) == 19 * 4
    assert ctypes.sizeof(_native.PlanTablesC) == 10 * 8 + 4 * 4
    assert ctypes.sizeof(_native.StateC) == 6 * 8


def test_argument_validation_without_gpu():
    """Error paths that return before any CUDA call."""
    lib = _native.load()
    assert lib.vmas_b200_world_step(None, None, None, None, 1, None) < 0
    assert b"null" in lib.vmas_b200_last_error()
