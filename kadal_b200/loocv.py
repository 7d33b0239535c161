"""Leave-one-out cross validation (``kadal/surrogate_models/supports/krigloocv.py`` and
``krigloocv2.py``): SURVEY.md section 8 row f2 ("next").  Not built yet -- and because the
product has no CPU compute path, it fails loudly instead of diverting to numpy."""


def loocv(KrigInfo, errtype="rmse", type="standard"):
    raise NotImplementedError(
        "kadal_b200: LOOCV (krigloocv.py / krigloocv2.py) is not on the GPU path yet; "
        "see DESIGN.md 'Out of scope / next'.")
