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


def test_correlation_fast_path_does_not_spill(lib):
    """The Gaussian fast path of corr_tile_kernel runs at an 85-register cap (three CTAs per SM); small source changes
    have made ptxas spill 400-1200 bytes there, which costs 20 % of the assembly and 28 % of the C4 prediction
    (DESIGN.md section 4, r1s).  Resource usage of the built object: stack frame of every fast instantiation <= 64 bytes."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    obj = os.path.join(os.path.dirname(_build.LIB), "kb200_kernels.o")
    if not os.path.exists(cuobjdump) or not os.path.exists(obj):
        pytest.skip("cuobjdump or the object file is not available")
    txt = subprocess.run([cuobjdump, "-res-usage", obj], capture_output=True, text=True).stdout
    # fast instantiations: <1, *> d < 8, <2, predict> 8 <= d < 16, <4, *> all summation schemes in one kernel
    found = re.findall(r"Function (\S*corr_tile_kernelILi[124]ELi[01]E\S*):\s*\n\s*REG:(\d+) STACK:(\d+)", txt)
    assert len(found) == 5, txt[:400]
    for name, reg, stack in found:
        assert int(reg) <= 85 and int(stack) <= 64, (name, reg, stack)
