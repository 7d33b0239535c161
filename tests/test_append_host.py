"""CPU: the host-side rule that decides when likelihood(x, KrigInfo, 'all') may extend a cached factor by one
row (kadal_b200/model.py:append_source) -- on stand-in models, no device."""
import numpy as np
import pytest

from kadal_b200 import model as M


class _Fake:
    def __init__(self, X, y, x, d=None, kernels=(0,), p=1, h=0, iso=False, device=0, pls=None):
        self.n, self.d = X.shape
        self.p, self.h, self.iso, self.device, self.kernels = p, h, iso, device, tuple(kernels)
        self.handle = object()
        self._keep = (X, y, np.ones((self.n, 1)), pls)
        self.fit_key = None if x is None else (np.asarray(x, dtype=np.float64).tobytes(), False, -6.0)

    def close(self):
        self.handle = None


@pytest.fixture
def cache():
    saved = dict(M._CACHE)
    M._CACHE.clear()
    yield M._CACHE
    M._CACHE.clear()
    M._CACHE.update(saved)


def test_append_source_rules(cache, monkeypatch):
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (11, 3))
    y = rng.standard_normal((11, 1))
    x = np.array([0.1, 0.2, 0.3])
    prev = _Fake(X[:10], y[:10], x)
    nxt = _Fake(X, y, None)
    cache["prev"], cache["next"] = prev, nxt
    assert M.append_source(nxt, x, False, -6.0) is prev
    # other hyper-parameters, trainvar or nugget: no
    assert M.append_source(nxt, x + 1e-12, False, -6.0) is None
    assert M.append_source(nxt, x, True, -6.0) is None
    assert M.append_source(nxt, x, False, -5.0) is None
    # switched off
    monkeypatch.setenv("KB200_NO_APPEND", "1")
    assert M.append_source(nxt, x, False, -6.0) is None
    monkeypatch.delenv("KB200_NO_APPEND")
    # an old row changed (e.g. 'std' standardisation recomputed on the enlarged sample) or y renormalised: no
    X2 = X.copy()
    X2[3, 1] += 1e-9
    assert M.append_source(_Fake(X2, y, None), x, False, -6.0) is None
    y2 = y.copy()
    y2[0] *= 1.0000001
    assert M.append_source(_Fake(X, y2, None), x, False, -6.0) is None
    # two samples more, other kernel list, other device, trend columns: no
    assert M.append_source(_Fake(np.vstack((X, X[:1])), np.vstack((y, y[:1])), None), x, False, -6.0) is None
    assert M.append_source(_Fake(X, y, None, kernels=(3,)), x, False, -6.0) is None
    assert M.append_source(_Fake(X, y, None, device=1), x, False, -6.0) is None
    assert M.append_source(_Fake(X, y, None, p=2), x, False, -6.0) is None
    # a model whose predictor came from KrigInfo['U'] (no device-resident fit) or that was closed: no
    prev.fit_key = None
    assert M.append_source(nxt, x, False, -6.0) is None
    prev.fit_key = (x.tobytes(), False, -6.0)
    prev.close()
    assert M.append_source(nxt, x, False, -6.0) is None
