"""errperf: the metric definitions of kadal/surrogate_models/supports/errperf.py:38-117."""
import numpy as np
import pytest

from kadal_b200.errperf import errperf


def test_metrics_follow_the_reference_definitions():
    T = np.array([[1.0], [2.0], [4.0], [0.0]])
    P = np.array([[1.5], [1.0], [4.0], [0.5]])
    e = T - P
    re = np.array([[-0.5], [0.5], [0.0], [-0.5]])         # plain error where T == 0 (errperf.py:57-64)
    assert np.array_equal(errperf(T, P, "e"), e)
    assert errperf(T, P, "MAE") == np.mean(np.abs(e))
    assert errperf(T, P, "rmse") == np.sqrt(np.mean(e ** 2))
    assert np.array_equal(errperf(T, P, "re"), re)
    assert errperf(T, P, "mape") == np.mean(np.abs(re * 100))
    assert errperf(T, P, "rmspe") == np.sqrt(np.mean((re * 100) ** 2))
    assert errperf(T, P, "r2") == 1 - np.sum(e ** 2) / np.sum((T - T.mean()) ** 2)
    assert errperf(T.ravel(), P.ravel(), "mse") == np.mean(e ** 2)
    with pytest.raises(Exception):
        errperf(T, P[:2], "mse")
    with pytest.raises(Exception):
        errperf(T, P, "nope")
    with pytest.raises(TypeError):
        errperf(T, P, 3)
