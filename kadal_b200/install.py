"""Monkey-patch a live ``kadal`` package so that its Kriging hot path runs on the GPU.

KADAL binds ``likelihood`` and ``prediction`` by name at import time in three modules
(kriging_model.py:9-10, optim_tools/MOBO.py:11, surrogate_models/supports/krigloocv2.py:3-4);
replacing those module globals is the whole integration.  ``uninstall``
restores the originals.
"""
import importlib

_TARGETS = (
    ("kadal.surrogate_models.kriging_model", ("likelihood", "prediction")),
    ("kadal.surrogate_models.kpls_model", ("likelihood", "prediction")),
    ("kadal.surrogate_models.supports.krigloocv2", ("likelihood", "prediction")),
    ("kadal.optim_tools.MOBO", ("likelihood",)),
    ("kadal.surrogate_models.supports.likelihood_func", ("likelihood",)),
    ("kadal.surrogate_models.supports.prediction", ("prediction",)),
)
_saved = []


def install(package="kadal", strict=False):
    """Patch every module of ``package`` that binds the hot-path callables.  Modules that
    cannot be imported (e.g. MOBO without the ``kmac`` extension) are skipped unless
    ``strict``.  Returns the list of (module, name) pairs that were patched."""
    from .likelihood_func import likelihood
    from .prediction import prediction
    repl = {"likelihood": likelihood, "prediction": prediction}
    done = []
    for modname, names in _TARGETS:
        modname = package + modname[len("kadal"):]
        try:
            mod = importlib.import_module(modname)
        except Exception:
            if strict:
                raise
            continue
        for nm in names:
            if hasattr(mod, nm) and getattr(mod, nm) is not repl[nm]:
                _saved.append((mod, nm, getattr(mod, nm)))
                setattr(mod, nm, repl[nm])
                done.append((modname, nm))
    return done


def uninstall():
    """Undo :func:`install`."""
    while _saved:
        mod, nm, orig = _saved.pop()
        setattr(mod, nm, orig)
