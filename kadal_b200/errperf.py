"""Error metrics of ``kadal/surrogate_models/supports/errperf.py`` (same names, same
definitions): O(n) scalar post-processing of the leave-one-out predictions on the host."""
import numpy as np


def errperf(T, P, type="rmse"):
    """Metric ``type`` between true values T and predictions P (errperf.py:5-120)."""
    T, P = np.asarray(T, dtype=float), np.asarray(P, dtype=float)
    if T.shape != P.shape:
        raise Exception("The size of T and P is different")
    T, P = T.reshape(-1, 1), P.reshape(-1, 1)
    if not isinstance(type, str):
        raise TypeError("Type should be a string")
    t = type.lower()
    e = T - P
    with np.errstate(divide="ignore", invalid="ignore"):
        re = np.where(T == 0, e, e / np.where(T == 0, 1.0, T))     # errperf.py:57-66: plain error where T == 0
    pe = re * 100
    table = {
        "e": lambda: e, "ae": lambda: np.abs(e), "mae": lambda: np.mean(np.abs(e)),
        "se": lambda: e ** 2, "mse": lambda: np.mean(e ** 2), "rmse": lambda: np.sqrt(np.mean(e ** 2)),
        "re": lambda: re, "are": lambda: np.abs(re), "mare": lambda: np.mean(np.abs(re)),
        "sre": lambda: re ** 2, "msre": lambda: np.mean(re ** 2), "rmsre": lambda: np.sqrt(np.mean(re ** 2)),
        "pe": lambda: pe, "ape": lambda: np.abs(pe), "mape": lambda: np.mean(np.abs(pe)),
        "spe": lambda: pe ** 2, "mspe": lambda: np.mean(pe ** 2), "rmspe": lambda: np.sqrt(np.mean(pe ** 2)),
        "r2": lambda: 1 - np.sum(e ** 2) / np.sum((T - np.mean(T)) ** 2),
    }
    if t not in table:
        raise Exception("Type is invalid")
    return table[t]()
