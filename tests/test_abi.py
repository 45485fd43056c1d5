"""CPU-only: the C-ABI library loads and exports every symbol include/binius_b200.h declares; the
product path fails loudly without a GPU (no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "binius_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    import binius_b200

    lib = binius_b200._lib.load()
    syms = _header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/binius_b200.h but not exported"
    assert sorted(binius_b200._lib.SYMBOLS) == syms


def test_no_cpu_fallback():
    import torch

    import binius_b200

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(binius_b200.DeviceError):
        binius_b200.B200Layer()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "binius_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|#include\s+[\"<].*oracle/", txt, re.M), f
