"""Engine lookup for the functional (reference-shaped) entry points.

torchPDLP's step / helper functions take the LP tensors on every call
(PDLP/primal_dual_hybrid_gradient_step.py:3, PDLP/helpers.py:53,98).  The engine
keeps its own copy of the LP resident on the GPU instead, so the functional wrappers
find (or build) the engine for a given (K, c, q, l, u) through this small cache.

An entry is only ever reused for the very tensors it was built from, unmodified: the
cache holds weak references to the source tensors and their version counters, so

* a different tensor that happens to land on a freed tensor's address (the CUDA caching
  allocator recycles blocks between the instances of a ``main.py`` run) is a miss,
* an in-place edit of any source tensor (``q.mul_(2)``) is a miss,
* an engine is closed -- its GPU memory released -- as soon as one of its source
  tensors is garbage collected, and at most ``_MAX`` engines stay resident.
"""
from __future__ import annotations

import weakref
from collections import OrderedDict

import numpy as np
import torch

from .engine import Engine

_MAX = 2
_entries: "OrderedDict[int, _Entry]" = OrderedDict()
_next_id = 0


def _payload(t):
    """The tensor whose version counter moves when the data of ``t`` is edited in place."""
    if isinstance(t, torch.Tensor):
        if t.layout == torch.sparse_coo:
            return t._values()
        if t.layout == torch.sparse_csr:
            return t.values()
    return t


def _version(t):
    """What changes when the data of ``t`` is edited in place: torch's version counter; for numpy arrays and scipy
    matrices (no counter) a content fingerprint -- the sum of all values and of every other value, O(size) per
    lookup (a few milliseconds for the 5M values of config 2, once per solve)."""
    p = _payload(t)
    if isinstance(p, torch.Tensor):
        return int(p._version)
    data = p if isinstance(p, np.ndarray) else getattr(p, "data", None)      # scipy sparse: the values array
    if isinstance(data, np.ndarray) and data.size:
        flat = data.reshape(-1)
        with np.errstate(all="ignore"):
            return (flat.size, float(np.sum(flat, dtype=np.float64)), float(np.sum(flat[::2], dtype=np.float64)))
    return 0


class _Snap:
    """Identity + version of one source object (tensor, scipy matrix, numpy array)."""

    __slots__ = ("ref", "strong", "version")

    def __init__(self, obj, keep_alive=False):
        try:
            self.ref = weakref.ref(obj)
            self.strong = obj if keep_alive else None
        except TypeError:                      # not weak-referenceable (e.g. a tuple): keep it alive instead
            self.ref = None
            self.strong = obj
        self.version = _version(obj)

    def get(self):
        return self.strong if self.ref is None else self.ref()

    def matches(self, obj) -> bool:
        return self.get() is obj and self.version == _version(obj)


class _Entry:
    def __init__(self, engine, K, vectors, m_ineq, keep_alive=False):
        self.engine = engine
        self.K = _Snap(K, keep_alive)
        self.vectors = [_Snap(v, keep_alive) for v in vectors]
        self.m_ineq = int(m_ineq)

    def alive(self) -> bool:
        return bool(self.engine.h) and self.K.get() is not None and all(s.get() is not None for s in self.vectors)

    def matches(self, K, vectors, m_ineq) -> bool:
        return (self.m_ineq == int(m_ineq) and self.K.matches(K)
                and all(s.matches(v) for s, v in zip(self.vectors, vectors)))


def _drop(key) -> None:
    ent = _entries.pop(key, None)
    if ent is not None:
        ent.engine.close()


def _sweep() -> None:
    for key in [k for k, e in _entries.items() if not e.alive()]:
        _drop(key)


def _insert(ent: _Entry) -> None:
    global _next_id
    _sweep()
    key = _next_id
    _next_id += 1
    _entries[key] = ent
    # release the GPU copy as soon as a source tensor dies (not only at the next lookup)
    for snap in [ent.K] + ent.vectors:
        obj = snap.get()
        if snap.ref is not None and snap.strong is None and obj is not None:
            try:
                weakref.finalize(obj, _drop, key)
            except TypeError:
                pass
    while len(_entries) > _MAX:
        _drop(next(iter(_entries)))


def attach(K, engine: Engine, c, q, l, u, m_ineq) -> None:
    """Remember that ``engine`` already holds exactly (K, c, q, l, u)."""
    _insert(_Entry(engine, K, (c, q, l, u), m_ineq))


def _resolve_device(K, device):
    if device is None or (isinstance(device, str) and device in ("cpu", "auto", "gpu")) or \
            (isinstance(device, torch.device) and device.type != "cuda"):
        return K.device if isinstance(K, torch.Tensor) and K.device.type == "cuda" else "cuda"
    return device


def engine_for(K, m_ineq, c, q, l, u, device=None) -> Engine:
    _sweep()
    for key, ent in _entries.items():
        if ent.matches(K, (c, q, l, u), m_ineq):
            _entries.move_to_end(key)
            return ent.engine
    dtype = K.dtype if isinstance(K, torch.Tensor) else None
    eng = Engine(K, m_ineq, c, q, l, u, device=_resolve_device(K, device), dtype=dtype)
    attach(K, eng, c, q, l, u, m_ineq)
    return eng


def engine_for_matrix(K, device=None) -> Engine:
    """An engine whose matrix is ``K`` (any vectors): enough for products and the power iteration
    (PDLP/helpers.py:41-51 takes only K).  Built with zero vectors, which the entry keeps alive."""
    _sweep()
    for key, ent in _entries.items():
        if ent.K.matches(K):
            _entries.move_to_end(key)
            return ent.engine
    dtype = K.dtype if isinstance(K, torch.Tensor) else torch.float64
    zn = torch.zeros(K.shape[1], dtype=dtype)
    zm = torch.zeros(K.shape[0], dtype=dtype)
    eng = Engine(K, 0, zn, zm, zn, zn, device=_resolve_device(K, device), dtype=dtype)
    ent = _Entry(eng, K, (zn, zm, zn, zn), 0)
    for s in ent.vectors:                      # the placeholders live as long as the entry
        s.strong = s.get()
    _insert(ent)
    return eng


def clear() -> None:
    for key in list(_entries):
        _drop(key)
