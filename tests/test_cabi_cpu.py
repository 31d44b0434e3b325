"""CPU-only checks of the drop-in boundary: the shared library loads and exports every symbol the header declares,
and the product path fails loudly (no fallback) when no CUDA device is present."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "ttn_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ttn_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import ctypes
    import ttn_b200
    from ttn_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 45
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/ttn_b200.h but not exported"
    assert lib.ttn_version() == 100
    assert isinstance(lib, ctypes.CDLL)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    import ttn_b200 as t
    x = t.TTvector(2, [np.ones((2, 1, 1)), np.ones((2, 1, 1))], (2, 2), [1, 1, 1])
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        t.norm(x)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "tensortrainnumerics.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", ".jl")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "ttn_oracle" not in src, f"{f} references the oracle"
