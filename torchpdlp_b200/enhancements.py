"""Drop-in for torchPDLP's PDLP/enhancements.py: Ruiz (and Pock-Chambolle)
preconditioning, primal-weight update, infeasibility certificates.

Unlike the reference (dense K only) every function
here works on sparse K: the scaling runs on the engine's CSR(K) / CSR(K^T) copies.
"""
from __future__ import annotations

import time

import torch

from . import _cache
from .engine import Engine


def _precondition(c, K, q, l, u, device, max_iter, eps, pock_chambolle):
    t0 = time.time()
    if isinstance(device, str) and device in ("cpu", "auto", "gpu"):
        device = None
    m_ineq_unknown = 0     # scaling does not depend on the row split
    eng = Engine(K, m_ineq_unknown, c, q, l, u, device=device,
                 dtype=K.dtype if isinstance(K, torch.Tensor) else None)
    eng.ruiz(max_iter, eps)
    if pock_chambolle:
        eng.pock_chambolle()
    D_row, D_col = eng.scaling()
    c_s, q_s, l_s, u_s = eng.problem()
    K_s = eng.matrix()
    col = lambda v: v.reshape(-1, 1)
    torch.cuda.synchronize(eng.device)
    eng.close()
    time_used = time.time() - t0
    return K_s, col(c_s), col(q_s), col(l_s), col(u_s), (col(D_col), col(D_row), K, c, q, l, u), time_used


def ruiz_precondition(c, K, q, l, u, device="cpu", max_iter=20, eps=1e-6):
    """enhancements.py:4-71.  Returns ``(K_s, c_s, q_s, l_s, u_s, (D_col, D_row, K, c,
    q, l, u), time_used)``; K_s is a sparse CSR tensor on the GPU.  Includes the
    reference's convergence quirk (the ROW norms are tested twice, :60-61)."""
    return _precondition(c, K, q, l, u, device, max_iter, eps, pock_chambolle=False)


def preconditioning(c, K, q, l, u, device="cpu", max_iter=20, eps=1e-6, pock_chambolle=True):
    """The symbol PDLP/main.py:8,107 imports but the snapshot does not define.  Ruiz
    then (optionally) Pock-Chambolle alpha=1, returned in the order main.py unpacks:
    ``K, c, q, l, u, dt_precond, time_used``."""
    K_s, c_s, q_s, l_s, u_s, dt, used = _precondition(c, K, q, l, u, device, max_iter, eps, pock_chambolle)
    return K_s, c_s, q_s, l_s, u_s, dt, used


def primal_weight_update(x_prev, x, y_prev, y, omega, smooth_theta):
    """enhancements.py:73-78 (inside the solver this is fused into the restart kernels)."""
    diff_y_norm = torch.linalg.norm((y_prev - y).reshape(-1), 2)
    diff_x_norm = torch.linalg.norm((x_prev - x).reshape(-1), 2)
    omega = torch.as_tensor(omega, dtype=diff_x_norm.dtype, device=diff_x_norm.device)
    if diff_x_norm > 0 and diff_y_norm > 0:
        omega = torch.exp(smooth_theta * torch.log(diff_y_norm / diff_x_norm) + (1 - smooth_theta) * torch.log(omega))
    return omega


def detect_infeasibility(x, y, x_prev, y_prev, lam, lam_prev, c, q, K, l, u, m_ineq, device, tol=1e-2):
    """enhancements.py:80-161 for caller-supplied vectors; the two matrix products
    run on the engine's SpMV kernels.  Inside ``pdlp_algorithm`` the same tests are
    fused into the iteration kernels instead."""
    eng = _cache.engine_for(K, m_ineq, c, q, l, u, device)
    dev, dt = eng.device, eng.dtype
    v = lambda t: torch.as_tensor(t, dtype=dt).reshape(-1).to(dev)
    x, y, x_prev, y_prev, lam, lam_prev, c_, q_, l_, u_ = (v(t) for t in (x, y, x_prev, y_prev, lam, lam_prev, c, q, l, u))
    dx, dy, dlam = x - x_prev, y - y_prev, lam - lam_prev
    Kdx = eng.spmv(dx)
    inf = float("inf")
    if (torch.linalg.norm(Kdx[m_ineq:]) < tol and bool(torch.all(Kdx[:m_ineq] >= -tol))
            and torch.dot(c_, dx) < tol):
        ok = ((~torch.isinf(l_) & ~torch.isinf(u_) & (dx.abs() <= tol))
              | ((u_ == inf) & (c_ >= 0) & (dx >= -tol)) | ((l_ == -inf) & (c_ <= 0) & (dx <= tol)))
        if bool(torch.all(ok)):
            return "DUAL_INFEASIBLE"
    dual_res = eng.spmv(dy, transpose=True) - dlam
    if torch.linalg.norm(dual_res) < tol and bool(torch.all(dy[:m_ineq] >= -tol)):
        combo = float(torch.dot(q_, dy))
        fl = ~torch.isinf(l_) & (l_ != 0)
        fu = ~torch.isinf(u_) & (u_ != 0)
        combo -= float(torch.dot(l_[fl], torch.clamp(dlam[fl], min=0)))
        combo -= float(torch.dot(u_[fu], torch.clamp(-dlam[fu], min=0)))
        if combo > -tol:
            return "PRIMAL_INFEASIBLE"
    return None
