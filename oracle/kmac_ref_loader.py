"""TEST INFRASTRUCTURE ONLY — ctypes door to oracle/_ref/libkmac_ref.so, the reference's own KMAC C++ compiled
by oracle/kmac_ref/Makefile (the built library travels to the GPU box; /root/reference does not)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libkmac_ref.so")
REF_SRC = "/root/reference/kadal/extern/HDYE_3D_Update"
_lib = None


def build():
    """(Re)build when the reference sources are present (build container); a no-op elsewhere."""
    if os.path.isdir(REF_SRC):
        subprocess.run(["make", "-s", "-C", os.path.join(HERE, "kmac_ref")], check=True)
    return os.path.exists(LIB)


def available():
    return os.path.exists(LIB)


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(LIB)
        _lib.kmac_ref_ehvi3d_multi.restype = ctypes.c_int
        _lib.kmac_ref_ehvi3d_single.restype = ctypes.c_double
    return _lib


def ehvi3d_multi(y_par, ref_point, mean, sdev):
    """The reference's kmac.ehvi3d_sliceupdate_multi.  At most 1000 front points / candidates (static tables
    of ehvi_multi.cc:13-24)."""
    L = _load()
    P = np.ascontiguousarray(np.asarray(y_par, dtype=float).reshape(-1, 3))
    r = np.ascontiguousarray(np.asarray(ref_point, dtype=float).reshape(3))
    MU = np.ascontiguousarray(np.asarray(mean, dtype=float).reshape(-1, 3))
    S = np.ascontiguousarray(np.asarray(sdev, dtype=float).reshape(-1, 3))
    assert len(P) <= 1000 and len(MU) <= 1000
    out = np.empty(len(MU))
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    L.kmac_ref_ehvi3d_multi(p(P), len(P), p(r), p(MU), p(S), len(MU), p(out))
    return out
