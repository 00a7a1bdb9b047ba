"""CPU: the C-ABI library loads, exports every symbol include/eo_b200.h declares, binds them all in
the ctypes table, and fails loudly (no CPU fallback) when there is no GPU."""

import ctypes
import os
import re

import pytest

import dolfinx_external_operator_b200 as eo
from dolfinx_external_operator_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "eo_b200.h")) as fh:
        src = fh.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(eo_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 25
    for nm in names:
        assert hasattr(lib, nm), f"{nm} declared in eo_b200.h but not exported"
        assert nm in _lib.PROTOTYPES, f"{nm} has no ctypes prototype"
    assert set(_lib.PROTOTYPES) == set(names)


def test_version_and_struct_layout():
    lib = _lib.load()
    assert lib.eo_version() == 100
    assert ctypes.sizeof(_lib.Stats) == 8 * (4 + _lib.EO_NITER_BINS + 4)
    assert ctypes.sizeof(_lib.VmParams) == 32


def test_no_cpu_fallback_without_gpu():
    lib = _lib.load()
    if lib.eo_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(eo.EOError) as ei:
        eo.Context(0)
    assert ei.value.code == -5
    with pytest.raises(eo.EOError):
        eo.VonMises()  # the model needs a context -> must raise, not compute on the CPU


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dolfinx-external-operator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                with open(os.path.join(dirpath, f)) as fh:
                    txt = fh.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports oracle"
                assert "liboracle" not in txt
