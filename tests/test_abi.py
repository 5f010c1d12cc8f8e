"""The C-ABI library loads on a CPU-only box, exports every symbol include/scribe_b200.h declares, and the
product path fails loudly (no CPU fallback) when no CUDA device exists."""
import ctypes
import os
import re

import numpy as np
import pytest

from scribe_py_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "scribe_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(scribe_[a-z_]+)\s*\(", src)))


def test_header_symbols_exported():
    lib = _lib.load_library()
    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_lib.EXPORTS)
    assert lib.scribe_abi_version() == 1


def test_struct_layouts():
    assert ctypes.sizeof(_lib.Opts) == 32
    assert _lib.JOB_DTYPE.itemsize == 48
    assert ctypes.sizeof(_lib.Stats) == 88


def _has_gpu():
    try:
        _lib.Handle(-1).close()
        return True
    except _lib.ScribeCudaError:
        return False


@pytest.mark.skipif(_has_gpu(), reason="CUDA device present")
def test_no_cpu_fallback():
    import scribe_py_b200 as s
    x = np.random.default_rng(0).random((50, 1)).tolist()
    with pytest.raises(_lib.ScribeCudaError):
        s.cmi(x, x, x)
    with pytest.raises(_lib.ScribeCudaError):
        s.mi(x, x)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "scribe_py_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no CPU fallback", "").lower() or f == "_lib.py" and "imports `oracle/`" in text, f
