"""CPU: the C-ABI library loads, exports every symbol include/b2k.h declares, fails loudly without a GPU,
and the product never touches the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "b2k.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2k_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree(b2k):
    assert header_symbols() == sorted(b2k.SYMBOLS)


def test_library_exports_every_declared_symbol(b2k):
    lib = ctypes.CDLL(b2k.LIB_PATH)
    for name in header_symbols():
        assert hasattr(lib, name), "libb2k.so does not export %s" % name
    assert lib.b2k_version() >= 100


def test_no_cpu_fallback(b2k):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        b2k.assign(np.zeros((4, 2), np.float32), np.zeros((2, 2), np.float32))
    import pyemma_b200 as p
    with pytest.raises(Exception):
        p.cluster_kmeans(np.random.rand(100, 2), k=3)


def test_product_never_references_oracle():
    bad = []
    for base, _dirs, files in os.walk(os.path.join(ROOT, "pyemma_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                txt = open(os.path.join(base, f)).read()
                if re.search(r"(^|\n)\s*(from|import)\s+oracle|liboracle|#include\s+[<\"][^\n]*oracle|CDLL\([^\n]*oracle", txt):
                    bad.append(os.path.join(base, f))
    assert not bad, "product files reference the oracle: %s" % bad


def test_argument_validation_without_gpu(b2k):
    with pytest.raises(ValueError):
        b2k.metric_id("manhattan")
    with pytest.raises(ValueError):
        b2k.assign(np.zeros((4, 3), np.float32), np.zeros((2, 2), np.float32))  # dim mismatch (test_assign.py:183-197)
    with pytest.raises(ValueError):
        b2k.assign(np.zeros((4, 3), np.float32), np.zeros(3, np.float32))
