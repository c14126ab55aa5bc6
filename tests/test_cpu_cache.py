"""Engine cache of the functional wrappers (torchpdlp_b200/_cache.py) without a GPU: an entry is reused only for the
very tensors it was built from, unmodified; engines are closed when a source tensor dies or the cache overflows."""
import gc

import torch

from torchpdlp_b200 import _cache


class FakeEngine:
    built = 0

    def __init__(self, K, m_ineq, c, q, l, u, device=None, dtype=None):
        FakeEngine.built += 1
        self.h = object()
        self.q0 = float(q.sum())

    def close(self):
        self.h = None


def lp(seed):
    g = torch.Generator().manual_seed(seed)
    K = torch.rand(6, 9, generator=g, dtype=torch.float64).to_sparse_coo()
    return K, torch.rand(9, dtype=torch.float64), torch.rand(6, dtype=torch.float64), torch.zeros(9, dtype=torch.float64), \
        torch.ones(9, dtype=torch.float64)


def test_hits_only_for_the_same_unmodified_tensors(monkeypatch):
    monkeypatch.setattr(_cache, "Engine", FakeEngine)
    _cache.clear()
    FakeEngine.built = 0
    K, c, q, l, u = lp(0)
    e1 = _cache.engine_for(K, 2, c, q, l, u)
    assert _cache.engine_for(K, 2, c, q, l, u) is e1 and FakeEngine.built == 1
    assert _cache.engine_for(K, 3, c, q, l, u) is not e1                 # another m_ineq is another LP
    # an in-place edit of a source tensor must not return the stale engine
    q.mul_(2.0)
    e2 = _cache.engine_for(K, 2, c, q, l, u)
    assert e2 is not e1 and e2.q0 == float(q.sum())
    # equal contents in different tensors: a different problem as far as the cache can tell
    e3 = _cache.engine_for(K, 2, c.clone(), q, l, u)
    assert e3 is not e2
    # sparse values edited in place
    K._values().mul_(3.0)
    assert _cache.engine_for(K, 2, c, q, l, u) is not e2
    _cache.clear()


def test_engine_is_closed_when_its_tensors_die_or_the_cache_overflows(monkeypatch):
    monkeypatch.setattr(_cache, "Engine", FakeEngine)
    _cache.clear()
    K, c, q, l, u = lp(1)
    e1 = _cache.engine_for(K, 0, c, q, l, u)
    assert e1.h is not None
    del q
    gc.collect()
    assert e1.h is None                                                    # released with its source tensor
    # a same-shaped LP whose tensors may land on the freed addresses is built afresh
    K2, c2, q2, l2, u2 = lp(2)
    e2 = _cache.engine_for(K2, 0, c2, q2, l2, u2)
    assert e2 is not e1 and e2.h is not None
    keep = [lp(s) for s in range(10, 10 + _cache._MAX)]
    engines = [_cache.engine_for(t[0], 0, *t[1:]) for t in keep]
    assert e2.h is None and all(e.h is not None for e in engines)          # oldest evicted, at most _MAX resident
    _cache.clear()
    assert all(e.h is None for e in engines)


def test_engine_for_matrix_reuses_any_engine_of_that_matrix(monkeypatch):
    monkeypatch.setattr(_cache, "Engine", FakeEngine)
    _cache.clear()
    FakeEngine.built = 0
    K, c, q, l, u = lp(3)
    e = _cache.engine_for(K, 1, c, q, l, u)
    assert _cache.engine_for_matrix(K) is e and FakeEngine.built == 1
    K2 = lp(4)[0]
    e2 = _cache.engine_for_matrix(K2)
    assert e2 is not e and _cache.engine_for_matrix(K2) is e2 and FakeEngine.built == 2
    _cache.clear()


def test_numpy_and_scipy_sources_are_fingerprinted(monkeypatch):
    """numpy / scipy inputs have no version counter: an in-place edit is caught by the content fingerprint."""
    import numpy as np
    import scipy.sparse as sp
    monkeypatch.setattr(_cache, "Engine", FakeEngine)
    _cache.clear()
    K = sp.random(8, 12, density=0.4, format="csr", random_state=np.random.default_rng(0))
    c, q, l, u = np.ones(12), np.arange(8.0), np.zeros(12), np.full(12, 5.0)
    e1 = _cache.engine_for(K, 3, c, q, l, u)
    assert _cache.engine_for(K, 3, c, q, l, u) is e1
    q[2] += 1.0
    e2 = _cache.engine_for(K, 3, c, q, l, u)
    assert e2 is not e1
    K.data[0] *= 2.0
    assert _cache.engine_for(K, 3, c, q, l, u) is not e2
    _cache.clear()
