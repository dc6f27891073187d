"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/pcc/search.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

from pointcloudcomparator_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "pcc", "search.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pcc_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_listed_in_binding():
    assert _declared() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.SO_PATH):
        pytest.fail(f"{_lib.SO_PATH} missing -- run __graft_entry__.build()")
    L = ctypes.CDLL(_lib.SO_PATH)
    for name in _declared():
        assert hasattr(L, name), name
    assert L.pcc_version() >= 100


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pointcloudcomparator_b200.search import GridSearch
    with pytest.raises(_lib.PccError) as e:
        GridSearch(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pointcloudcomparator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in src and "from oracle" not in src and "pcc_oracle" not in src, f
    for f in os.listdir(os.path.join(ROOT, "include", "pcc")):
        assert "oracle" not in open(os.path.join(ROOT, "include", "pcc", f)).read().lower(), f
