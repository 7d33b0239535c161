"""install()/uninstall(): the module-global patch points of the reference
(kriging_model.py:9-10, MOBO.py:11, krigloocv2.py:3-4) on a stand-in package tree (the real
reference does not travel to the GPU box; INTEGRATION.md shows the same call on it)."""
import sys
import types


def _fake_kadal(name):
    def mk(mod):
        m = types.ModuleType(mod)
        sys.modules[mod] = m
        return m
    root = mk(name)
    sm = mk(name + ".surrogate_models")
    sup = mk(name + ".surrogate_models.supports")
    km = mk(name + ".surrogate_models.kriging_model")
    lf = mk(name + ".surrogate_models.supports.likelihood_func")
    pr = mk(name + ".surrogate_models.supports.prediction")
    lf.likelihood = lambda *a, **k: "ref-lik"
    pr.prediction = lambda *a, **k: "ref-pred"
    km.likelihood, km.prediction = lf.likelihood, pr.prediction
    km.tune = lambda: km.likelihood()      # looks the global up at call time, like tune_hyperparameters
    root.surrogate_models, sm.supports, sm.kriging_model = sm, sup, km
    return km, lf, pr


def test_install_patches_and_restores():
    from kadal_b200 import install as inst
    import kadal_b200
    km, lf, pr = _fake_kadal("fakekadal")
    assert km.tune() == "ref-lik"
    done = inst.install("fakekadal")
    assert ("fakekadal.surrogate_models.kriging_model", "likelihood") in done
    assert km.likelihood is kadal_b200.likelihood and km.prediction is kadal_b200.prediction
    assert lf.likelihood is kadal_b200.likelihood and pr.prediction is kadal_b200.prediction
    inst.uninstall()
    assert km.tune() == "ref-lik" and pr.prediction() == "ref-pred"
