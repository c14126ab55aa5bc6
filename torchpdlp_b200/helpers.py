"""Drop-in for torchPDLP's PDLP/helpers.py (norm / KKT / residual helpers)."""
from __future__ import annotations

import weakref

import torch

from . import _cache


def project_lambda_box(grad, is_neg_inf, is_pos_inf):
    """helpers.py:3-39.  Element-wise on whatever device ``grad`` lives on; inside
    the solver this projection is fused into the KKT kernel's epilogue."""
    zero = torch.zeros_like(grad)
    return torch.where(is_neg_inf & is_pos_inf, zero,
                       torch.where(is_neg_inf, torch.clamp(grad, max=0.0),
                                   torch.where(is_pos_inf, torch.clamp(grad, min=0.0), grad)))


def _bounds_from_dual(l_dual, u_dual, is_neg_inf, is_pos_inf):
    inf = float("inf")
    l = torch.where(is_neg_inf, torch.full_like(l_dual, -inf), l_dual)
    u = torch.where(is_pos_inf, torch.full_like(u_dual, inf), u_dual)
    return l, u


_bound_cache = []          # [(weakref l_dual, version, weakref u_dual, version, (l, u))]: see _cache for the rationale


def _engine(K, m_ineq, c, q, l_dual, u_dual, is_neg_inf, is_pos_inf):
    lu = None
    for ent in list(_bound_cache):
        a, va, b, vb, val = ent
        if a() is None or b() is None:
            _bound_cache.remove(ent)
        elif a() is l_dual and b() is u_dual and va == l_dual._version and vb == u_dual._version:
            lu = val
    if lu is None:
        lu = _bounds_from_dual(l_dual, u_dual, is_neg_inf, is_pos_inf)
        _bound_cache.append((weakref.ref(l_dual), l_dual._version, weakref.ref(u_dual), u_dual._version, lu))
        del _bound_cache[:-8]
    return _cache.engine_for(K, m_ineq, c, q, lu[0], lu[1])


def spectral_norm_estimate_torch(K, num_iters=10):
    """helpers.py:41-51.  Start vector: torch.randn(n, 1) from the global CPU
    generator in K's dtype."""
    eng = _cache.engine_for_matrix(K)
    b0 = torch.randn(K.shape[1], 1, dtype=K.dtype)
    return torch.tensor(eng.power_iteration(b0, num_iters), dtype=K.dtype, device=eng.device)


def compute_residuals_and_duality_gap(x, y, c, q, K, m_ineq, is_neg_inf, is_pos_inf, l_dual, u_dual):
    """helpers.py:53-96: (primal_residual, dual_residual, duality_gap, prim_obj,
    adjusted_dual) as 1-element tensors."""
    eng = _engine(K, m_ineq, c, q, l_dual, u_dual, is_neg_inf, is_pos_inf)
    r = eng.kkt(x, y, 1.0)
    mk = lambda v: torch.tensor([v], dtype=eng.dtype, device=eng.device)
    return mk(r.pr), mk(r.dr), mk(r.gap), mk(r.prim_obj), mk(r.adj_dual)


def KKT_error(x, y, c, q, K, m_ineq, omega, is_neg_inf, is_pos_inf, l_dual, u_dual, device):
    """helpers.py:98-108."""
    eng = _engine(K, m_ineq, c, q, l_dual, u_dual, is_neg_inf, is_pos_inf)
    r = eng.kkt(x, y, float(omega))
    return torch.tensor([r.kkt], dtype=eng.dtype, device=eng.device)


def check_termination(primal_residual, dual_residual, duality_gap, prim_obj, adjusted_dual, q_norm, c_norm, tol):
    """helpers.py:110-128 (the gap is signed: a negative gap always passes)."""
    cond1 = primal_residual <= tol * (1 + q_norm)
    cond2 = dual_residual <= tol * (1 + c_norm)
    cond3 = duality_gap <= tol * (1 + abs(prim_obj) + abs(adjusted_dual))
    return bool(cond1 and cond2 and cond3)
