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
    km.loocv = km.loocv2 = lambda *a, **k: "ref-loocv"
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
    from kadal_b200 import loocv as lo
    assert km.loocv is lo.loocv_fast and km.loocv2 is lo.loocv2          # kriging_model.py:7-8, used by loocvcalc
    inst.uninstall()
    assert km.tune() == "ref-lik" and pr.prediction() == "ref-pred" and km.loocv() == "ref-loocv"


def test_install_supplies_kmac_and_patches_the_ehvi_scorers():
    """EHVIcomputation.py:3-12 cannot be imported without the compiled extension `kmac`; install() registers a
    stand-in under the extension's name and replaces the scorers bound in acquifunc_opt.py:5."""
    from kadal_b200 import install as inst, ehvi
    name = "fakekadal2"
    _fake_kadal(name)
    for sub in (".extern", ".extern.HDYE_3D_Update", ".optim_tools", ".optim_tools.ehvi"):
        sys.modules[name + sub] = types.ModuleType(name + sub)
    ec = types.ModuleType(name + ".optim_tools.ehvi.EHVIcomputation")
    ec.exi2d = ec.ehvicalc = ec.ehvicalc_vec = ec.ehvicalc_kmac3d = lambda *a, **k: "ref"
    ec.kmac = None
    aq = types.ModuleType(name + ".optim_tools.acquifunc_opt")
    aq.ehvicalc, aq.ehvicalc_vec, aq.ehvicalc_kmac3d = ec.ehvicalc, ec.ehvicalc_vec, ec.ehvicalc_kmac3d
    sys.modules[ec.__name__], sys.modules[aq.__name__] = ec, aq
    done = inst.install(name)
    assert (name + ".extern.HDYE_3D_Update", "kmac") in done
    kmac = sys.modules[name + ".extern.HDYE_3D_Update.kmac"]
    assert callable(kmac.ehvi3d_sliceupdate_multi) and callable(kmac.ehvi3d_sliceupdate)
    assert aq.ehvicalc_kmac3d is ehvi.ehvicalc_kmac3d and aq.ehvicalc_vec is ehvi.ehvicalc_vec
    assert ec.exi2d is ehvi.exi2d and hasattr(ec.kmac, "ehvi3d_sliceupdate_multi")
    inst.uninstall()
    assert aq.ehvicalc_kmac3d() == "ref" and name + ".extern.HDYE_3D_Update.kmac" not in sys.modules


def test_install_on_the_live_reference_makes_mobo_importable():
    """In the build container: the real package, whose optim_tools cannot be imported because kmac was never
    compiled, imports after install() and its names point at the CUDA path (no compute here)."""
    import pytest
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("live reference only exists in the build container")
    import warnings
    ref_loader._ensure_path()
    from kadal_b200 import install as inst, ehvi
    import kadal_b200
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            done = inst.install("kadal")
            import kadal.optim_tools.acquifunc_opt as aq
            import kadal.optim_tools.ehvi.EHVIcomputation as ec
            assert aq.ehvicalc_kmac3d is ehvi.ehvicalc_kmac3d and ec.ehvicalc_vec is ehvi.ehvicalc_vec
            assert hasattr(ec.kmac, "ehvi3d_sliceupdate_multi")
            import kadal.surrogate_models.kriging_model as km
            assert km.likelihood is kadal_b200.likelihood
            assert ("kadal.optim_tools.acquifunc_opt", "ehvicalc_kmac3d") in done
        finally:
            inst.uninstall()
    import kadal.surrogate_models.kriging_model as km
    assert km.likelihood is not kadal_b200.likelihood
