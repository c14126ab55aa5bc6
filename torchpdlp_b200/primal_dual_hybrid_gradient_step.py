"""Drop-in for torchPDLP's PDLP/primal_dual_hybrid_gradient_step.py: one PDHG
step through the engine (two fused kernels), same arguments and returns."""
from __future__ import annotations

import torch

from . import _cache


def _like(t, ref):
    return t.reshape(ref.shape).to(ref.dtype) if isinstance(ref, torch.Tensor) else t


def fixed_one_step_pdhg(x, y, c, q, K, l, u, m_ineq, eta, omega, theta):
    """step.py:3-40.  Returns (x', y', eta, eta); like the reference, ``y`` is also
    updated in place (step.py:34) when it lives on the engine's device."""
    eng = _cache.engine_for(K, m_ineq, c, q, l, u)
    eng.set_iterate(x, y)
    eng.set_step(float(eta), float(omega), float(theta), adaptive=False)
    eng.iterate(1, use_graph=False)
    xn, yn = eng.get_iterate()
    xn, yn = _like(xn, x), _like(yn, y)
    if isinstance(y, torch.Tensor) and y.device == yn.device:
        y.copy_(yn)
        yn = y
    return xn, yn, eta, eta


def adaptive_one_step_pdhg(x, y, c, q, K, l, u, m_ineq, eta, omega, theta, k, j):
    """step.py:43-115.  Returns (x', y', eta_weight, eta_next, j + 1): a single
    attempt, exactly like the reference whose retry loop returns in its first pass."""
    eng = _cache.engine_for(K, m_ineq, c, q, l, u)
    eng.set_iterate(x, y)
    eng.set_step(float(eta), float(omega), float(theta), adaptive=True)
    eng.set_counters(int(k) - 1, 0)       # the device increments k before applying the rule (:76, step.py:102)
    eng.iterate(1, use_graph=False)
    xn, yn = eng.get_iterate()
    st = eng.state()
    dev = xn.device
    return (_like(xn, x), _like(yn, y), torch.tensor(st.last_weight, dtype=xn.dtype, device=dev),
            torch.tensor(st.eta, dtype=xn.dtype, device=dev), j + 1)
