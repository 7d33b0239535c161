"""CPU: the host-side rule that decides when likelihood(x, KrigInfo, 'all') may extend a cached factor by one
row (kadal_b200/model.py:append_source) -- on stand-in models, no device."""
import numpy as np
import pytest

from kadal_b200 import model as M


class _Fake:
    def __init__(self, X, y, x, d=None, kernels=(0,), p=1, h=0, iso=False, device=0, pls=None):
        self.n, self.d = X.shape
        self.p, self.h, self.iso, self.device, self.kernels = p, h, iso, device, tuple(kernels)
        self.closed = False
        self._keep = (X, y, np.ones((self.n, 1)), pls)
        self._uploaded = tuple(M._content(a) if a is not None else None for a in self._keep)
        self.fit_key = None if x is None else (np.asarray(x, dtype=np.float64).tobytes(), False, -6.0)

    def close(self):
        self.closed = True


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
    # prev's host arrays edited in place after the upload (sum-preserving row swap): the device copy no longer is
    # what the host arrays say, no append from it
    prev2 = _Fake(X[:10].copy(), y[:10].copy(), x)
    cache["prev2"] = prev2
    assert M.append_source(nxt, x, False, -6.0) is prev2
    prev2._keep[0][[0, 1]] = prev2._keep[0][[1, 0]]
    assert M.append_source(nxt, x, False, -6.0) is None


def test_model_cache_key_sees_in_place_edits_that_keep_the_sum():
    a = np.arange(12.0).reshape(4, 3)
    k0 = M._fingerprint(a)
    a[[0, 1]] = a[[1, 0]]
    assert M._fingerprint(a) != k0 and M._fingerprint(a)[0] == k0[0]
    assert M._fingerprint(a[:, ::2])[1] == (4, 2)          # non-contiguous views are hashed by value


def test_pinned_result_arrays_are_recycled(monkeypatch):
    """U / Psi of a mode='all' call live in recycled pinned buffers (kadal_b200/_capi.py:result_empty): a buffer
    returns to the free list when its last numpy view dies, is reused for a slightly larger result, survives
    deepcopy / pickle as an ordinary array, and the budget falls back to pageable memory -- on a stand-in
    allocator, no device."""
    import copy
    import ctypes
    import gc
    import pickle
    from kadal_b200 import _capi

    class FakeLib:
        def __init__(self):
            self.live = {}

        def kb200_host_alloc(self, n):
            b = (ctypes.c_char * n)()
            self.live[ctypes.addressof(b)] = b
            return ctypes.addressof(b)

        def kb200_host_free(self, p):
            del self.live[p.value]

    fl = FakeLib()
    monkeypatch.setattr(_capi, "require_gpu", lambda: fl)
    monkeypatch.setattr(_capi, "load", lambda: fl)
    monkeypatch.setattr(_capi, "_RES_FREE", [])
    monkeypatch.setattr(_capi, "_RES_BYTES", [0])
    monkeypatch.setattr(_capi, "_RES_LIMIT", 4 << 20)
    a = _capi.result_empty((300, 300))
    a[:] = 2.0
    addr = a.__array_interface__["data"][0]
    assert len(fl.live) == 1 and _capi._RES_BYTES[0] >= 720000
    view = a[:5]
    del a
    gc.collect()
    assert _capi._RES_FREE == []                       # a view still holds the buffer
    del view
    gc.collect()
    assert len(_capi._RES_FREE) == 1
    b = _capi.result_empty((301, 301))                 # the next model of a believer sequence
    assert b.__array_interface__["data"][0] == addr and len(fl.live) == 1
    b[:] = 3.0
    assert copy.deepcopy(b).sum() == b.sum() and pickle.loads(pickle.dumps(b)).shape == (301, 301)
    assert _capi.result_empty((4, 4)).base is None      # small results stay ordinary arrays
    big = _capi.result_empty((1000, 1000))              # over the budget: pageable
    assert big.base is None and len(fl.live) == 1
    c = _capi.result_empty((500, 500))                  # fits next to b
    assert len(fl.live) == 2
    del b, c
    gc.collect()
    d = _capi.result_empty((600, 600))                  # needs the room of the two free buffers
    assert d.base is not None and len(fl.live) == 1
