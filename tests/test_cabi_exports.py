"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/qtt_b200.h declares, and fails loudly (no CPU fallback) when no GPU is present."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "qtt_b200.h")).read()
    return sorted(set(re.findall(r"QTT_API\s+[\w\s\*]+?\b(qtt_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    import itensorqtt_jl_b200 as q
    L = ctypes.CDLL(q.lib_path)
    decl = _declared_symbols()
    assert len(decl) >= 30
    for name in decl:
        assert hasattr(L, name), f"{name} declared in include/qtt_b200.h but not exported"
    assert set(decl) == set(q.EXPORTED_SYMBOLS), "ctypes table and header disagree"
    assert q.lib.qtt_version() == 100


def test_struct_layouts_match_header():
    import itensorqtt_jl_b200 as q
    # qtt_sweep_params: int32, 3 pointers, int32, 3 doubles, ... (natural alignment, 64-bit)
    assert ctypes.sizeof(q.SweepParams) == 104
    assert ctypes.sizeof(q.SweepInfo) == 96


def test_no_cpu_fallback():
    import torch
    import itensorqtt_jl_b200 as q
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(q.QttError):
        q.Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "itensorqtt.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
