"""The C-ABI shared library loads and exports every symbol include/lbm_b200.h declares
(no compute calls: this file runs without a GPU)."""
import ctypes as C

import pytest


def test_exports_every_declared_symbol(L):
    from lbm_b200._lib import SIGNATURES, header_symbols
    lib = L.lib()
    declared = header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/lbm_b200.h but not exported"
    assert set(declared) == set(SIGNATURES), "python binding and header out of sync"


def test_version_and_structure_helper(L):
    lib = L.lib()
    assert lib.lbm_version() >= 10000
    kl = (C.c_int64 * 2)(1, 1)
    ku = (C.c_int64 * 2)(1, 1)
    ol, ou = C.c_int64(), C.c_int64()
    assert lib.lbm_prod_bandwidths(10, 10, 2, kl, ku, C.byref(ol), C.byref(ou)) == 0
    assert (ol.value, ou.value) == (2, 2)                      # README.md:13-26
    assert lib.lbm_prod_bandwidths(2, 2, 2, kl, ku, C.byref(ol), C.byref(ou)) == 0
    assert (ol.value, ou.value) == (1, 1)                      # clipped to size-1
    assert lib.lbm_prod_bandwidths(-1, 2, 2, kl, ku, C.byref(ol), C.byref(ou)) == -1
    assert lib.lbm_prod_bandwidths(2, 2, 0, kl, ku, C.byref(ol), C.byref(ou)) == -3


def test_no_cpu_fallback(L):
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the -m gpu tests")
    with pytest.raises(L.LbmCudaError):
        L.Handle(0)
    A = L.BandedMatrix(torch.zeros(4, 3, dtype=torch.float64), 4, 1, 1)
    with pytest.raises(L.LbmCudaError):
        A @ torch.zeros(4, dtype=torch.float64)


def test_product_never_imports_oracle():
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "lazybandedmatrices.jl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                src = open(os.path.join(dp, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle|liblbm_oracle|oracle/", src, flags=re.M), f
