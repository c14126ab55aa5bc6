"""Engine lookup for the functional (reference-shaped) entry points.

torchPDLP's step / helper functions take the LP tensors on every call
(PDLP/primal_dual_hybrid_gradient_step.py:3, PDLP/helpers.py:53,98).  The engine
keeps the LP resident on the GPU instead, so the functional wrappers find (or
build) the engine for a given (K, c, q, l, u) through this small cache.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from .engine import Engine

_MAX = 4
_engines: "OrderedDict[tuple, Engine]" = OrderedDict()


def _ptr(t) -> int:
    if not isinstance(t, torch.Tensor):
        return id(t)
    if t.layout == torch.sparse_coo:
        return t._values().data_ptr()
    if t.layout == torch.sparse_csr:
        return t.values().data_ptr()
    return t.data_ptr()


def signature(K, c, q, l, u, m_ineq) -> tuple:
    shape = tuple(K.shape)
    return (_ptr(K), shape, _ptr(c), _ptr(q), _ptr(l), _ptr(u), int(m_ineq))


def attach(K, engine: Engine, c, q, l, u, m_ineq) -> None:
    """Remember that ``engine`` already holds exactly (K, c, q, l, u)."""
    sig = signature(K, c, q, l, u, m_ineq)
    _engines[sig] = engine
    _engines.move_to_end(sig)
    while len(_engines) > _MAX:
        _, old = _engines.popitem(last=False)
        old.close()


def engine_for(K, m_ineq, c, q, l, u, device=None) -> Engine:
    sig = signature(K, c, q, l, u, m_ineq)
    eng = _engines.get(sig)
    if eng is not None and eng.h:
        _engines.move_to_end(sig)
        return eng
    if device is None or (isinstance(device, str) and device in ("cpu", "auto", "gpu")) or \
            (isinstance(device, torch.device) and device.type != "cuda"):
        device = K.device if isinstance(K, torch.Tensor) and K.device.type == "cuda" else "cuda"
    dtype = K.dtype if isinstance(K, torch.Tensor) else None
    eng = Engine(K, m_ineq, c, q, l, u, device=device, dtype=dtype)
    attach(K, eng, c, q, l, u, m_ineq)
    return eng


def clear() -> None:
    while _engines:
        _, e = _engines.popitem()
        e.close()
