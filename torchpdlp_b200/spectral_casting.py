"""Drop-in for torchPDLP's PDLP/spectral_casting.py ("fishnet" start-point search).

Same call (``spectral_cast(K, c, q, l, u, m_ineq, k, s=2, i=5, device=...)``), same sequence of
random draws (global torch generator on the CPU, in the engine dtype) and the same selection /
breeding rules; the batched PDHG steps run on the engine 8 points at a time, so the matrix is
streamed once per 8 right-hand sides instead of once per point (SpMM).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cache
from ._lib import check

NB = 8      # points per engine pass


def _to_groups(P: torch.Tensor):
    """(len, j) -> list of contiguous (len, 8) blocks, zero-padded."""
    n, j = P.shape
    groups = []
    for g0 in range(0, j, NB):
        blk = torch.zeros(n, NB, dtype=P.dtype, device=P.device)
        w = min(NB, j - g0)
        blk[:, :w] = P[:, g0:g0 + w]
        groups.append(blk)
    return groups


def _from_groups(groups, j):
    return torch.cat(groups, dim=1)[:, :j].contiguous()


def sample_points(K, i, device="cpu", engine=None):
    """spectral_casting.py:32-63: 2^i Gaussian directions scaled to radius sigma_25, centred on the
    positive diagonal.  Returns (points (n, 2^i) on the engine device, r)."""
    eng = engine
    n = eng.n
    b0 = torch.randn(n, 1, dtype=eng.dtype)                         # helpers.py:47 (25 rounds, :47)
    r = eng.power_iteration(b0, 25)
    pts = torch.randn(size=(n, 2 ** i), dtype=eng.dtype).to(eng.device)      # :51
    centre = (r / n ** 0.5) * torch.ones((n, 1), dtype=eng.dtype, device=eng.device)
    pts = pts * r / torch.norm(pts, dim=0, keepdim=True)            # :58
    pts += centre
    return pts, r


def init_PDHG_vars(engine):
    """spectral_casting.py:161-189: eta = 0.9 / sigma_50, omega = ||c|| / ||q||."""
    b0 = torch.randn(engine.n, 1, dtype=engine.dtype)
    sigma = engine.power_iteration(b0, 50)
    st = engine.state()
    eta = (1.0 / sigma) * 0.9
    omega = st.c_norm / st.q_norm if (st.q_norm > 1e-6 and st.c_norm > 1e-6) else 1.0
    return eta, omega


def _batch_step(eng, X8, XB8, Y8, eta, omega):
    check(eng.lib.pdlp_batch_step(eng.h, X8.data_ptr(), XB8.data_ptr(), Y8.data_ptr(), float(eta), float(omega)))


def get_best_pts(eng, pts, pts_y, q, s=2):
    """spectral_casting.py:191-252: keep the 1/s of the points with the smallest SIGNED duality gap."""
    j = pts.shape[1]
    gaps = []
    out = (C.c_double * 24)()
    for X8, Y8 in zip(_to_groups(pts), _to_groups(pts_y)):
        check(eng.lib.pdlp_batch_gap(eng.h, X8.data_ptr(), Y8.data_ptr(), out))
        dual = (q.reshape(1, -1) @ Y8).reshape(-1)                 # q^T y per point
        for v in range(NB):
            prim, lpos, uneg = out[3 * v], out[3 * v + 1], out[3 * v + 2]
            gaps.append(float(dual[v]) + lpos + uneg - prim)
    gaps = torch.tensor(gaps[:j], dtype=torch.float64)
    order = torch.argsort(gaps).to(pts.device)                       # :237
    keep = max(1, j // s)
    return pts[:, order][:, :keep].contiguous(), pts_y[:, order][:, :keep].contiguous()


def fishnet(pts, K, c, q, l, u, m_ineq, s=2, k=32, device="cpu", engine=None):
    """spectral_casting.py:65-159."""
    eng = engine
    groups_x = _to_groups(pts)
    groups_y = []
    for X8 in groups_x:                                              # pts_y = K @ pts, :99
        Y8 = torch.empty(eng.m, NB, dtype=eng.dtype, device=eng.device)
        check(eng.lib.pdlp_batch_spmv(eng.h, X8.data_ptr(), Y8.data_ptr()))
        groups_y.append(Y8)
    pts_y = _from_groups(groups_y, pts.shape[1])
    eta, omega = init_PDHG_vars(eng)
    qd = q.reshape(-1).to(eng.device).to(eng.dtype)
    rnd = 0
    while pts.shape[1] > 1:
        j = pts.shape[1]
        gx, gy = _to_groups(pts), _to_groups(pts_y)
        scratch = torch.empty(eng.n, NB, dtype=eng.dtype, device=eng.device)
        for _ in range(k):                                           # :105-107
            for X8, Y8 in zip(gx, gy):
                _batch_step(eng, X8, scratch, Y8, eta, omega)
        pts, pts_y = _from_groups(gx, j), _from_groups(gy, j)
        old_j = j
        pts, pts_y = get_best_pts(eng, pts, pts_y, qd, s)
        new_j = pts.shape[1]
        if rnd % 2 == 1 and new_j > 1:                               # :117-154
            mid_x, mid_y = pts.mean(dim=1, keepdim=True), pts_y.mean(dim=1, keepdim=True)
            kids_x, kids_y = [], []
            for _ in range(old_j - new_j - 1):
                w = torch.rand(new_j, dtype=eng.dtype)               # CPU generator, like the oracle / a CPU reference run
                w = (w / w.sum()).to(eng.device)
                kids_x.append(pts @ w.view(-1, 1))
                kids_y.append(pts_y @ w.view(-1, 1))
            if kids_x:
                pts = torch.cat([pts] + kids_x + [mid_x], dim=1)
                pts_y = torch.cat([pts_y] + kids_y + [mid_y], dim=1)
        rnd += 1
    torch.cuda.synchronize(eng.device)
    return pts.flatten(), pts_y.flatten()


def spectral_cast(K, c, q, l, u, m_ineq, k, s=2, i=5, device="cpu"):
    """spectral_casting.py:5-29: returns (x0, y0), a start point for ``pdlp_algorithm``."""
    eng = _cache.engine_for(K, m_ineq, c, q, l, u, device)
    pts, radius = sample_points(K, i, device, engine=eng)
    return fishnet(pts, K, c, q, l, u, m_ineq, s, k, device, engine=eng)
