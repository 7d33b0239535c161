"""CPU: the C-ABI library loads and exports every symbol include/kb200.h declares
(no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from kadal_b200 import _build, _capi


@pytest.fixture(scope="module")
def lib():
    _build.build()   # nvcc cross-compiles without a GPU
    return ctypes.CDLL(_build.LIB)


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "kb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(kb200_[a-z_0-9]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(_capi.SIGNATURES)


def test_library_exports_every_symbol(lib):
    for name in header_symbols():
        assert hasattr(lib, name), name


def test_version_and_no_gpu_is_loud(lib):
    assert lib.kb200_version() >= 100
    lib.kb200_device_count.restype = ctypes.c_int
    if lib.kb200_device_count() == 0:
        with pytest.raises(_capi.Kb200Error):
            _capi.require_gpu()
        import numpy as np
        import kadal_b200
        from oracle import kriging_oracle as ko
        info = ko.make_kriginfo(np.random.default_rng(0).uniform(-1, 1, (8, 2)), np.ones((8, 1)))
        with pytest.raises(_capi.Kb200Error):   # no silent CPU fallback
            kadal_b200.likelihood(np.zeros(2), info, "default", False)
