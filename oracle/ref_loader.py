"""Import the live KADAL reference (build container only).

TEST INFRASTRUCTURE.  ``/root/reference`` does not exist on the GPU box, so nothing that
runs there (``-m gpu`` tests, ``smoke()``, ``bench.py``) may depend on this module
succeeding; callers must use :func:`available` first.

Three third-party imports of the reference are absent from this image and are replaced
by the no-arithmetic stubs in ``oracle/_stubs`` (``sobolsampling``, ``cma``,
``matplotlib``) – see SURVEY.md Appendix D.
"""
import importlib
import os
import sys

_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_stubs")


def ref_root():
    return os.environ.get("KADAL_REF", "/root/reference")


def available():
    return os.path.isdir(os.path.join(ref_root(), "kadal", "surrogate_models"))


def _ensure_path():
    root = ref_root()
    for p in (root,):
        if p not in sys.path:
            sys.path.insert(0, p)
    # stubs only for modules that are really missing
    for mod in ("sobolsampling", "cma", "matplotlib"):
        try:
            importlib.import_module(mod)
        except Exception:
            if _STUBS not in sys.path:
                sys.path.append(_STUBS)


def load():
    """Returns a namespace with the reference's hot-path callables."""
    if not available():
        raise RuntimeError("KADAL reference not present at %s" % ref_root())
    _ensure_path()
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        lf = importlib.import_module("kadal.surrogate_models.supports.likelihood_func")
        kf = importlib.import_module("kadal.surrogate_models.supports.kernelfunc")
        pr = importlib.import_module("kadal.surrogate_models.supports.prediction")
        km = importlib.import_module("kadal.surrogate_models.kriging_model")
        kp = importlib.import_module("kadal.surrogate_models.kpls_model")
        sp = importlib.import_module("kadal.misc.sampling.samplingplan")

    class NS:
        pass

    ns = NS()
    ns.likelihood = lf.likelihood
    ns.likelihood_module = lf
    ns.calckernel = kf.calckernel
    ns.prediction = pr.prediction
    ns.prediction_module = pr
    ns.Kriging = km.Kriging
    ns.kriging_module = km
    ns.KPLS = kp.KPLS
    ns.standardize = sp.standardize
    return ns
