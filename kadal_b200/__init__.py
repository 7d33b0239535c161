"""kadal_b200 — B200-native drop-in for KADAL's Kriging hot path.

Public surface (mirrors ``kadal.surrogate_models``):

* :func:`likelihood`, :func:`likelihood_batch`  — ``supports/likelihood_func.py``
* :func:`prediction`, :func:`predict_arrays`   — ``supports/prediction.py``
* :class:`Kriging`, :class:`KPLS`              — ``kriging_model.py``, ``kpls_model.py``
* :func:`install`                              — monkey-patch a live ``kadal`` package

All numerics run in ``_lib/libkb200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/kb200.h``).  There is no CPU fallback.
"""
from .likelihood_func import likelihood, likelihood_batch  # noqa: F401
from .prediction import prediction, predict_arrays  # noqa: F401

__all__ = ["likelihood", "likelihood_batch", "prediction", "predict_arrays"]


def __getattr__(name):
    # heavier host-side mirrors are imported lazily
    if name in ("Kriging", "kriginfocheck"):
        from . import kriging_model
        return getattr(kriging_model, name)
    if name == "KPLS":
        from . import kpls_model
        return kpls_model.KPLS
    if name == "install":
        from .install import install
        return install
    raise AttributeError(name)
