"""The C-ABI library loads on a machine without a GPU and exports every symbol that
include/progan_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "progan_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import progan_b200
    from progan_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "missing symbol " + n
    assert set(_lib.SIGNATURES) | {"pg_last_error"} == set(names)
    lib.pg_abi_version.restype = ctypes.c_int
    assert lib.pg_abi_version() == 1


def test_product_path_fails_loudly_without_cuda():
    """No CPU fallback: CPU tensors must raise, not silently compute."""
    import torch
    import progan_b200
    from progan_b200.kernels import CudaKernels
    prev = progan_b200.set_kernels(None)
    try:
        D = progan_b200.Discriminator(feat_dim=32, precision="fp32")
        with pytest.raises(RuntimeError, match="no .*CPU fallback|CUDA-only"):
            D(torch.zeros(2, 3, 8, 8), step=1, alpha=1.0)
    finally:
        progan_b200.set_kernels(prev)
