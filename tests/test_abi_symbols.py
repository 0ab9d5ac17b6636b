"""CPU: the C-ABI library builds, loads and exports every symbol include/colbert_b200.h declares;
compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import colbert_jl_b200 as cb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "colbert_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cb_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    syms = _header_symbols()
    assert len(syms) >= 15
    assert sorted(cb.SIGNATURES) == syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(cb.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(cb.LIB_PATH)
    for s in _header_symbols():
        assert hasattr(lib, s), s
    assert b"sm_100a" in cb.load().cb_version()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "colbert.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


@pytest.mark.skipif(cb.load().cb_device_count() > 0, reason="only meaningful on a machine without a GPU")
def test_no_cpu_fallback():
    cfg = cb.ColBERTConfig(dim=8, nbits=1)
    with pytest.raises(cb.CudaError):
        cb.Searcher(cfg, np.zeros((8, 2), np.float32), None, np.zeros(2, np.float32), [1], [1, 0], [1],
                    np.ones(1, np.uint32), np.zeros((1, 1), np.uint8))
    with pytest.raises(cb.CudaError):
        cb.decompress(8, 1, np.zeros((8, 2), np.float32), np.zeros(2, np.float32), np.ones(1, np.uint32),
                      np.zeros((1, 1), np.uint8))
    with pytest.raises(cb.CudaError):
        cb.maxsim(np.zeros((8, 2), np.float32), np.zeros((8, 1), np.float32), [1], [1])


def test_argument_validation_happens_before_the_device():
    # shape errors mirror the reference's exception types and need no GPU
    with pytest.raises(cb.DomainError):
        cb.decompress(7, 1, np.zeros((7, 2), np.float32), np.zeros(2, np.float32), np.ones(1, np.uint32),
                      np.zeros((1, 1), np.uint8))
    with pytest.raises(cb.DomainError):   # bucket_weights length (residual.jl:705)
        cb.decompress(8, 2, np.zeros((8, 2), np.float32), np.zeros(3, np.float32), np.ones(1, np.uint32),
                      np.zeros((2, 1), np.uint8))
    with pytest.raises(cb.DimensionMismatch):  # ranking.jl:71-74
        cb.maxsim(np.zeros((8, 2), np.float32), np.zeros((8, 2), np.float32), [1, 2], [1, 2])
