"""CPU-only checks of the boundary: the library builds/loads and exports every declared symbol."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    from gordo_b200 import build_native, _native
    build_native.build()
    lib = _native.lib()
    header = open(os.path.join(ROOT, "include", "gordo_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(gb200_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"libgordo_b200.so does not export {name}"
    assert set(_native.SIGNATURES) == declared, (set(_native.SIGNATURES) ^ declared)
    assert lib.gb200_abi_version() == 1


def test_arch_helpers_and_param_counts():
    from gordo_b200 import _native as N
    lib = N.lib()
    a = N.make_ff_arch([50, 42, 33, 25, 25, 33, 42, 50], ["tanh"] * 6 + ["linear"])
    assert lib.gb200_ff_param_count(C.byref(a)) == 9497            # T=50 hourglass
    assert lib.gb200_ff_packed_bytes(C.byref(a)) % 16 == 0 and lib.gb200_ff_packed_bytes(C.byref(a)) > 0
    la = N.make_lstm_arch(200, 200, [167, 133, 100, 100, 133, 167], ["tanh"] * 6, "linear", 128, 0)
    assert lib.gb200_lstm_param_count(C.byref(la)) == 939112
    assert lib.gb200_lstm_out_rows(C.byref(la), 100000) == 99873   # §8 a8
    with pytest.raises(ValueError):
        N.make_ff_arch([3] * 19, ["tanh"] * 18)


def test_no_product_import_of_oracle():
    """The product path must never import the oracle (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "gordo_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports oracle"
