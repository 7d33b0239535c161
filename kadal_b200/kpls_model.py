"""Drop-in for ``kadal/surrogate_models/kpls_model.py`` (KPLS: Kriging on PLS-weighted distances).

Only the producer of ``KrigInfo['plscoeff']`` lives here: the PLS fit itself is scikit-learn's
``PLSRegression`` on the host, exactly as in the reference (kpls_model.py:84-98) -- it runs once
per model, is O(n d h), and is outside the hot path.  The
likelihood / prediction arithmetic with the projected distances runs in the CUDA library.
"""
import numpy as np

from .kriging_model import Kriging, kriginfocheck


class KPLS(Kriging):
    """KPLS model with the reference's interface (kpls_model.py:11-98).  KrigInfo needs
    ``n_princomp`` (default 1); hyper-parameters are one length scale per PLS component."""

    def __init__(self, KrigInfo, ub=5, lb=-5, standardization=False, standtype="default", normy=True,
                 trainvar=True):
        if "n_princomp" not in KrigInfo:
            KrigInfo["n_princomp"] = 1
        self.n_princomp = KrigInfo["n_princomp"]
        super().__init__(KrigInfo, ub, lb, standardization, standtype, normy, trainvar, inherit=True)
        self.nbhyp = self.n_princomp + 1 if trainvar is True else self.n_princomp
        self.type = "kpls"
        KrigInfo["type"] = self.type
        KrigInfo, scaling = kriginfocheck(KrigInfo, lb, ub, self.nbhyp)
        self.KrigInfo = KrigInfo
        self.scaling = scaling
        self.sigmacmaes = (ub - lb) / 5
        if self.standardization is True:
            self.standardize()

    def standardize(self):
        """Kriging.standardize, then the PLS rotations (kpls_model.py:84-98)."""
        from sklearn.cross_decomposition import PLSRegression
        Kriging.standardize(self)
        K = self.KrigInfo
        X = K["X_norm"] if self.standardization is True else K["X"]
        fit = PLSRegression(self.n_princomp).fit(np.array(X, copy=True), np.array(K["y"], copy=True))
        K["plscoeff"] = fit.x_rotations_
