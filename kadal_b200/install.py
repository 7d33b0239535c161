"""Monkey-patch a live ``kadal`` package so that its Kriging hot path runs on the GPU.

KADAL binds ``likelihood`` and ``prediction`` by name at import time in three modules
(kriging_model.py:9-10, optim_tools/MOBO.py:11, surrogate_models/supports/krigloocv2.py:3-4) and the two LOOCV
routines in kriging_model.py:7-8; replacing those module globals is the whole integration.  The EHVI scorers are bound
the same way (optim_tools/acquifunc_opt.py:5 imports ``ehvicalc, ehvicalc_vec, ehvicalc_kmac3d`` from
optim_tools/ehvi/EHVIcomputation.py, which itself needs the compiled extension ``kmac``,
EHVIcomputation.py:3-12): ``install`` puts a ``kmac`` stand-in over ``kb200_ehvi3d`` in place when the
extension has not been built, so that ``kadal.optim_tools`` imports at all, and replaces the scorers.
``uninstall`` restores the originals.
"""
import importlib
import sys
import types

_saved = []
_shimmed = []


def _replacements():
    from .likelihood_func import likelihood
    from .prediction import prediction
    from . import ehvi
    from .loocv import loocv2, loocv_fast
    hot = {"likelihood": likelihood, "prediction": prediction}
    scorers = {"ehvicalc": ehvi.ehvicalc, "ehvicalc_vec": ehvi.ehvicalc_vec, "ehvicalc_kmac3d": ehvi.ehvicalc_kmac3d}
    return (
        # loocvcalc (kriging_model.py:291-294) looks `loocv` / `loocv2` up at call time: both become one closed-form pass
        ("kadal.surrogate_models.kriging_model", dict(hot, loocv=loocv_fast, loocv2=loocv2)),
        ("kadal.surrogate_models.kpls_model", hot),
        ("kadal.surrogate_models.supports.krigloocv2", hot),
        ("kadal.surrogate_models.supports.likelihood_func", {"likelihood": likelihood}),
        ("kadal.surrogate_models.supports.prediction", {"prediction": prediction}),
        ("kadal.optim_tools.ehvi.EHVIcomputation", dict(scorers, exi2d=ehvi.exi2d, kmac=kmac_module())),
        ("kadal.optim_tools.acquifunc_opt", scorers),
        ("kadal.optim_tools.MOBO", {"likelihood": likelihood}),
    )


def kmac_module(name="kmac"):
    """A module object with the two entry points of the reference's pybind11 extension
    (kmac_wrapper.cpp:160-199), served by ``kb200_ehvi3d`` in reference-identical mode."""
    from . import ehvi
    import numpy as np
    mod = types.ModuleType(name)
    mod.__doc__ = "EHVI using KMAC (kadal_b200 stand-in over kb200_ehvi3d)"

    def ehvi3d_sliceupdate_multi(y_par, ref_point, mean_vector, std_dev):
        return ehvi.ehvi3d_sliceupdate_multi(y_par, ref_point, mean_vector, std_dev, mode="reference")

    def ehvi3d_sliceupdate(y_par, ref_point, mean_vector, std_dev):
        mu, s = np.asarray(mean_vector, dtype=float).reshape(1, 3), np.asarray(std_dev, dtype=float).reshape(1, 3)
        return float(ehvi.ehvi3d_sliceupdate_multi(y_par, ref_point, mu, s, mode="reference")[0])

    mod.ehvi3d_sliceupdate_multi = ehvi3d_sliceupdate_multi
    mod.ehvi3d_sliceupdate = ehvi3d_sliceupdate
    return mod


def _ensure_kmac(package):
    """EHVIcomputation.py:3-12 raises ImportError when ``<package>.extern.HDYE_3D_Update.kmac`` was never compiled;
    register the stand-in under that name so the import succeeds."""
    parent = package + ".extern.HDYE_3D_Update"
    full = parent + ".kmac"
    try:
        importlib.import_module(full)
        return False
    except Exception:
        pass
    try:
        pmod = importlib.import_module(parent)
    except Exception:
        return False
    mod = kmac_module(full)
    sys.modules[full] = mod
    setattr(pmod, "kmac", mod)
    _shimmed.append((full, pmod))
    return True


def install(package="kadal", strict=False):
    """Patch every module of ``package`` that binds the hot-path callables.  Modules that cannot be imported are
    skipped unless ``strict``.  Returns the list of (module, name) pairs that were patched."""
    done = []
    if _ensure_kmac(package):
        done.append((package + ".extern.HDYE_3D_Update", "kmac"))
    mods = []
    for modname, repl in _replacements():       # import everything first: a module imported after its source was
        modname = package + modname[len("kadal"):]   # patched would bind the replacement for good
        try:
            mods.append((modname, importlib.import_module(modname), repl))
        except Exception:
            if strict:
                raise
    for modname, mod, repl in mods:
        for nm, new in repl.items():
            if hasattr(mod, nm) and getattr(mod, nm) is not new:
                _saved.append((mod, nm, getattr(mod, nm)))
                setattr(mod, nm, new)
                done.append((modname, nm))
    return done


def uninstall():
    """Undo :func:`install`."""
    while _saved:
        mod, nm, orig = _saved.pop()
        setattr(mod, nm, orig)
    while _shimmed:
        full, pmod = _shimmed.pop()
        sys.modules.pop(full, None)
        if hasattr(pmod, "kmac"):
            delattr(pmod, "kmac")
