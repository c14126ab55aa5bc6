"""CPU oracle for the restarted-PDHG hot path of torchPDLP.

TEST INFRASTRUCTURE ONLY.  This module is a CPU (torch) restatement of the
reference algorithm; it is imported by ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` and by
nothing else.  The product path (``torchpdlp_b200``) never routes through it and
has no CPU fallback.

Parity pin: the reference ships no tests / golden vectors,
so the oracle is pinned by executing the reference itself in the build container
(``tests/golden/make_golden.py`` imports /root/reference/PDLP, runs it on the
seeded config-1 LP for every option set, and checks this oracle reproduces
k/n/j/status exactly and the objective to 1e-9 before writing the fixtures under
``tests/golden/``).  Pock-Chambolle scaling has no reference implementation in
the snapshot -> "parity unpinned" for that one function.

Every function cites the reference lines it follows (paths relative to
/root/reference).  The restatement is sparse-safe: the reference's Ruiz and
infeasibility code only run on dense K; here every
product goes through :class:`Operator`, which holds K and an explicitly stored
K^T so CSR works too.
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field
from typing import List, Optional

import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore", message=".*Sparse (CSR|invariant).*")

INF = float("inf")


# --------------------------------------------------------------------------
# linear operator: K v and K^T w for dense / COO / CSR / scipy inputs
# --------------------------------------------------------------------------
class Operator:
    """K (m x n) with both products.  ``backend='torch'`` uses ATen (dense mm,
    COO/CSR sparse mm: what the reference executes, step.py:25,33); ``'scipy'``
    uses scipy.sparse CSR matvecs (independent second implementation)."""

    def __init__(self, K, backend: str = "torch"):
        self.backend = backend
        if backend == "scipy":
            import scipy.sparse as sp
            if isinstance(K, torch.Tensor):
                K = _torch_to_scipy(K)
            self.K = sp.csr_matrix(K)
            self.Kt = sp.csr_matrix(self.K.T)
            self.shape = self.K.shape
            self.dtype = torch.float64 if self.K.dtype == np.float64 else torch.float32
            return
        if not isinstance(K, torch.Tensor):  # scipy matrix -> torch CSR
            K = _scipy_to_torch_csr(K)
        self.shape = tuple(K.shape)
        self.dtype = K.dtype
        if K.layout == torch.strided:
            self.K, self.Kt = K, K.T
        elif K.layout == torch.sparse_coo:
            self.K = K.coalesce()
            self.Kt = self.K.t().coalesce()
        elif K.layout == torch.sparse_csr:
            self.K = K
            self.Kt = K.to_sparse_coo().t().coalesce().to_sparse_csr()
        else:
            raise TypeError(f"unsupported layout {K.layout}")

    def mv(self, v: torch.Tensor) -> torch.Tensor:      # K v
        if self.backend == "scipy":
            return torch.from_numpy(self.K @ v.numpy())
        return (self.K @ v.reshape(-1, 1)).reshape(-1)

    def rmv(self, w: torch.Tensor) -> torch.Tensor:     # K^T w
        if self.backend == "scipy":
            return torch.from_numpy(self.Kt @ w.numpy())
        return (self.Kt @ w.reshape(-1, 1)).reshape(-1)

    def mm(self, V: torch.Tensor) -> torch.Tensor:      # K V   (multi right-hand side)
        if self.backend == "scipy":
            return torch.from_numpy(self.K @ V.numpy())
        return self.K @ V

    def rmm(self, W: torch.Tensor) -> torch.Tensor:     # K^T W
        if self.backend == "scipy":
            return torch.from_numpy(self.Kt @ W.numpy())
        return self.Kt @ W

    def bilinear(self, w: torch.Tensor, v: torch.Tensor) -> float:
        """w^T K v evaluated as (w^T K) v, the association the reference uses at
        step.py:96 (matters: the adaptive rule amplifies last-bit differences)."""
        if self.backend == "scipy":
            return float((self.Kt @ w.numpy()) @ v.numpy())
        if self.K.layout == torch.sparse_csr:
            return float(torch.dot(self.rmv(w), v))
        return float(w.reshape(-1, 1).T @ self.K @ v.reshape(-1, 1))

    def triplets(self):
        """(rows, cols, vals) of the nonzeros, for the sparse-safe Ruiz."""
        if self.backend == "scipy":
            coo = self.K.tocoo()
            return (torch.from_numpy(coo.row.astype(np.int64)),
                    torch.from_numpy(coo.col.astype(np.int64)), torch.from_numpy(coo.data.copy()))
        K = self.K
        if K.layout == torch.strided:
            idx = K.nonzero(as_tuple=True)
            return idx[0], idx[1], K[idx]
        coo = K.to_sparse_coo().coalesce() if K.layout != torch.sparse_coo else K
        return coo.indices()[0], coo.indices()[1], coo.values().clone()


def _torch_to_scipy(K: torch.Tensor):
    import scipy.sparse as sp
    if K.layout == torch.strided:
        return sp.csr_matrix(K.numpy())
    coo = K.to_sparse_coo().coalesce()
    return sp.csr_matrix((coo.values().numpy(), (coo.indices()[0].numpy(), coo.indices()[1].numpy())),
                         shape=tuple(K.shape))


def _scipy_to_torch_csr(K):
    K = K.tocsr()
    return torch.sparse_csr_tensor(torch.from_numpy(K.indptr.astype(np.int64)),
                                   torch.from_numpy(K.indices.astype(np.int64)),
                                   torch.from_numpy(K.data.copy()), size=K.shape)


def _vec(v, dtype) -> torch.Tensor:
    return torch.as_tensor(v, dtype=dtype).reshape(-1).clone()


# --------------------------------------------------------------------------
# helpers.py restated
# --------------------------------------------------------------------------
def project_lambda_box(grad, lo_is_inf, up_is_inf):
    """Reduced-cost projection, PDLP/helpers.py:3-39: lambda_i = 0 when both
    bounds are infinite, min(g,0) when only l=-inf, max(g,0) when only u=+inf,
    g when both finite."""
    lam = grad.clone()
    only_lo_inf = lo_is_inf & ~up_is_inf
    only_up_inf = ~lo_is_inf & up_is_inf
    lam[lo_is_inf & up_is_inf] = 0.0
    lam[only_lo_inf] = torch.clamp(grad[only_lo_inf], max=0.0)
    lam[only_up_inf] = torch.clamp(grad[only_up_inf], min=0.0)
    return lam


def spectral_norm_estimate(op: Operator, num_iters: int, b0: Optional[torch.Tensor] = None):
    """Power iteration on K^T K, PDLP/helpers.py:41-51.  The reference draws the
    start vector with torch.randn(n, 1) (default dtype, global RNG, helpers.py:47);
    pass ``b0`` to inject it, else it is drawn here the same way."""
    n = op.shape[1]
    b = torch.randn(n, 1).reshape(-1).to(op.dtype) if b0 is None else b0.reshape(-1).clone()
    for _ in range(num_iters):
        b = op.rmv(op.mv(b))
        b = b / torch.linalg.norm(b)
    return torch.linalg.norm(op.mv(b))


@dataclass
class Residuals:
    pr: float          # ||[r_eq ; min(r_ineq,0)]||_2          helpers.py:88-91
    dr: float          # ||grad - lam||_2                        helpers.py:94
    gap: float         # adjusted_dual - prim_obj (SIGNED)       helpers.py:85
    prim_obj: float    # c^T x                                   helpers.py:76
    adj_dual: float    # q^T y + l_dual^T lam+ + u_dual^T lam-   helpers.py:80-84

    def kkt(self, omega: float) -> float:
        """PDLP/helpers.py:98-108."""
        w2 = omega * omega
        return math.sqrt(w2 * self.pr ** 2 + self.dr ** 2 / w2 + self.gap ** 2)


def compute_residuals_and_duality_gap(x, y, c, q, op, m_ineq, lo_is_inf, up_is_inf, l_dual, u_dual):
    """PDLP/helpers.py:53-96."""
    grad = c - op.rmv(y)
    prim_obj = torch.dot(c, x)
    dual_obj = torch.dot(q, y)
    lam = project_lambda_box(grad, lo_is_inf, up_is_inf)
    adj = dual_obj + torch.dot(l_dual, torch.clamp(lam, min=0.0)) + torch.dot(u_dual, torch.clamp(lam, max=0.0))
    r = op.mv(x) - q
    viol = torch.cat([r[m_ineq:], torch.clamp(r[:m_ineq], max=0.0)])
    return Residuals(float(torch.linalg.norm(viol)), float(torch.linalg.norm(grad - lam)),
                     float(adj - prim_obj), float(prim_obj), float(adj))


def check_termination(res: Residuals, q_norm, c_norm, tol) -> bool:
    """PDLP/helpers.py:110-128 (the gap is signed, so a negative gap passes)."""
    return (res.pr <= tol * (1 + q_norm) and res.dr <= tol * (1 + c_norm)
            and res.gap <= tol * (1 + abs(res.prim_obj) + abs(res.adj_dual)))


# --------------------------------------------------------------------------
# primal_dual_hybrid_gradient_step.py restated
# --------------------------------------------------------------------------
def fixed_one_step_pdhg(x, y, c, q, op, l, u, m_ineq, eta, omega, theta=1.0):
    """PDLP/primal_dual_hybrid_gradient_step.py:22-40.  Returns new (x, y)."""
    grad = c - op.rmv(y)
    x_new = torch.clamp(x - (eta / omega) * grad, min=l, max=u)
    x_bar = x_new + theta * (x_new - x)
    y_new = y + (eta * omega) * (q - op.mv(x_bar))
    if m_ineq > 0:
        y_new[:m_ineq] = torch.clamp(y_new[:m_ineq], min=0.0)
    return x_new, y_new


def adaptive_one_step_pdhg(x, y, c, q, op, l, u, m_ineq, eta, omega, theta, k):
    """PDLP/primal_dual_hybrid_gradient_step.py:64-115.  One attempt only (the
    reference's ``for i in range(200)`` returns in its first pass, :115).
    Returns (x', y', eta_weight, eta_next)."""
    x_new, y_new = fixed_one_step_pdhg(x, y, c, q, op, l, u, m_ineq, eta, omega, theta)
    dx, dy = x_new - x, y_new - y
    den = 2.0 * op.bilinear(dy, dx)                      # step.py:96
    if den != 0.0:
        eta_bar = (omega * float(torch.linalg.norm(dx)) ** 2 + float(torch.linalg.norm(dy)) ** 2 / omega) / abs(den)
        t1 = (1 - (k + 1) ** (-0.3)) * eta_bar
    else:
        eta_bar, t1 = INF, INF
    eta_next = min(t1, (1 + (k + 1) ** (-0.6)) * eta)
    if eta <= eta_bar:                                   # step.py:110-111
        return x_new, y_new, eta, eta_next
    return x_new, y_new, eta_next, eta_next              # step.py:113-115


# --------------------------------------------------------------------------
# enhancements.py restated (sparse-safe)
# --------------------------------------------------------------------------
def primal_weight_update(x_prev, x, y_prev, y, omega, smooth=0.5):
    """PDLP/enhancements.py:73-78."""
    dy = float(torch.linalg.norm(y_prev - y))
    dx = float(torch.linalg.norm(x_prev - x))
    if dx > 0 and dy > 0:
        return math.exp(smooth * math.log(dy / dx) + (1 - smooth) * math.log(omega))
    return omega


@dataclass
class Scaling:
    D_col: torch.Tensor
    D_row: torch.Tensor
    rounds: int
    time_used: float = 0.0


def ruiz_scale_triplets(rows, cols, vals, m, n, max_iter=20, eps=1e-6):
    """Ruiz equilibration on nonzero triplets, following PDLP/enhancements.py:43-62:
    per round rows are divided by sqrt(row inf-norm) then columns by sqrt(col
    inf-norm) of the row-scaled matrix (norms < eps -> 1); the loop stops when
    max|1 - row_norm| < eps (the reference tests the ROW norms twice, :60-61, and
    never the column norms).  Returns (scaled vals, D_row, D_col, rounds)."""
    vals = vals.clone()
    D_row = torch.ones(m, dtype=vals.dtype)
    D_col = torch.ones(n, dtype=vals.dtype)
    rounds = 0
    for _ in range(max_iter):
        rounds += 1
        rn = torch.zeros(m, dtype=vals.dtype).scatter_reduce_(0, rows, vals.abs(), "amax", include_self=True).sqrt()
        rn[rn < eps] = 1.0
        D_row /= rn
        vals /= rn[rows]
        cn = torch.zeros(n, dtype=vals.dtype).scatter_reduce_(0, cols, vals.abs(), "amax", include_self=True).sqrt()
        cn[cn < eps] = 1.0
        D_col /= cn
        vals /= cn[cols]
        if float(torch.max(torch.abs(1 - rn))) < eps:
            break
    return vals, D_row, D_col, rounds


def pock_chambolle_scale_triplets(rows, cols, vals, m, n, alpha=1.0):
    """Pock-Chambolle diagonal scaling (alpha = 1): rows / sqrt(||row||_1),
    cols / sqrt(||col||_1), both norms taken on the INPUT matrix.  Not present in
    the reference snapshot (main.py:8 imports a missing ``preconditioning``);
    specified from the PDLP paper -> parity unpinned."""
    a = vals.abs()
    rn = torch.zeros(m, dtype=vals.dtype).index_add_(0, rows, a ** (2 - alpha)).sqrt()
    cn = torch.zeros(n, dtype=vals.dtype).index_add_(0, cols, a ** alpha).sqrt()
    rn[rn < 1e-6] = 1.0
    cn[cn < 1e-6] = 1.0
    return vals / rn[rows] / cn[cols], 1.0 / rn, 1.0 / cn


def precondition(op: Operator, c, q, l, u, max_iter=20, eps=1e-6, pock_chambolle=False):
    """Sparse-safe ``ruiz_precondition`` (PDLP/enhancements.py:4-71): returns
    (scaled Operator, c_s, q_s, l_s, u_s, Scaling)."""
    t0 = time.time()
    m, n = op.shape
    rows, cols, vals = op.triplets()
    vals, D_row, D_col, rounds = ruiz_scale_triplets(rows, cols, vals, m, n, max_iter, eps)
    if pock_chambolle:
        vals, pr, pc = pock_chambolle_scale_triplets(rows, cols, vals, m, n)
        D_row, D_col = D_row * pr, D_col * pc
    Ks = torch.sparse_coo_tensor(torch.stack([rows, cols]), vals, size=(m, n)).coalesce()
    if op.backend == "torch" and isinstance(op.K, torch.Tensor) and op.K.layout == torch.strided:
        Ks = Ks.to_dense()
    elif op.backend == "torch" and isinstance(op.K, torch.Tensor) and op.K.layout == torch.sparse_csr:
        Ks = Ks.to_sparse_csr()
    sop = Operator(Ks, backend=op.backend)
    return sop, c * D_col, q * D_row, l / D_col, u / D_col, Scaling(D_col, D_row, rounds, time.time() - t0)


def detect_infeasibility(x, y, x_prev, y_prev, lam, lam_prev, c, q, op, l, u, m_ineq, tol):
    """PDLP/enhancements.py:106-161 with the row-sliced products replaced by
    slices of full products and the per-variable Python loop (:126-138) vectorised."""
    dx, dy, dlam = x - x_prev, y - y_prev, lam - lam_prev
    Kdx = op.mv(dx)
    eq_ok = float(torch.linalg.norm(Kdx[m_ineq:])) < tol                # :116
    ineq_ok = bool(torch.all(Kdx[:m_ineq] >= -tol))                     # :119
    obj_ok = float(torch.dot(c, dx)) < tol                              # :122
    if eq_ok and ineq_ok and obj_ok:
        l_inf, u_inf = torch.isinf(l), torch.isinf(u)
        ok = ((~l_inf & ~u_inf & (dx.abs() <= tol))
              | ((u == INF) & (c >= 0) & (dx >= -tol))
              | ((l == -INF) & (c <= 0) & (dx <= tol)))                 # :132-136
        if bool(torch.all(ok)):
            return "DUAL_INFEASIBLE"
    dual_res = op.rmv(dy) - dlam                                        # :145
    if float(torch.linalg.norm(dual_res)) < tol and bool(torch.all(dy[:m_ineq] >= -tol)):
        combo = float(torch.dot(q, dy))                                 # :148
        fin_l = ~torch.isinf(l) & (l != 0)
        fin_u = ~torch.isinf(u) & (u != 0)
        combo -= float(torch.dot(l[fin_l], torch.clamp(dlam[fin_l], min=0.0)))    # :111-112,:154
        combo -= float(torch.dot(u[fin_u], torch.clamp(-dlam[fin_u], min=0.0)))   # :156
        if combo > -tol:
            return "PRIMAL_INFEASIBLE"
    return None


# --------------------------------------------------------------------------
# primal_dual_hybrid_gradient.py restated
# --------------------------------------------------------------------------
@dataclass
class RestartRecord:
    k: int
    t: int
    kind: str          # "Sufficient" | "Necessary" | "Artificial" | "Limit"
    candidate: str     # "Average" | "Current"
    prim_obj: float
    adj_dual: float
    gap: float
    pr: float
    dr: float
    omega: float
    eta: float


@dataclass
class Solution:
    x: torch.Tensor
    y: torch.Tensor
    prim_obj: float
    k: int
    n: int
    j: int
    status: str
    time: float
    sigma_max: float
    trace: List[RestartRecord] = field(default_factory=list)


def pdlp_algorithm(K, m_ineq, c, q, l, u, max_kkt=100_000, tol=1e-4, restart_period=40,
                   primal_update=False, adaptive=False, scaling: Optional[Scaling] = None,
                   unscaled=None, infeasibility_detect=False, infeas_tol=1e-4, time_limit=3600.0,
                   time_used=0.0, x_init=None, y_init=None, b0=None, power_iters=100,
                   backend="torch", verbose=False) -> Solution:
    """Restarted, averaged PDHG: PDLP/primal_dual_hybrid_gradient.py:7-181.

    ``K`` may be an :class:`Operator` or anything one can be built from.  When the
    problem was preconditioned pass ``scaling`` and ``unscaled=(op_u, c_u, q_u,
    l_u, u_u)``: the termination residuals are then evaluated on the unscaled data
    but compared against the scaled ||q||, ||c|| (reference :157-163 vs :19-20)."""
    t_start = time.time()
    op = K if isinstance(K, Operator) else Operator(K, backend=backend)
    dt = op.dtype
    c, q, l, u = (_vec(v, dt) for v in (c, q, l, u))
    m, n = op.shape
    lo_is_inf = torch.isinf(l) & (l < 0)                                # :11-12
    up_is_inf = torch.isinf(u) & (u > 0)
    l_dual = torch.where(lo_is_inf, torch.zeros_like(l), l)            # :14-17
    u_dual = torch.where(up_is_inf, torch.zeros_like(u), u)
    # :19-20 call torch.linalg.norm(v, 2) on (len,1) MATRICES, i.e. the spectral norm
    # through an SVD; equal to the vector 2-norm up to the last bit, which the
    # adaptive rule amplifies -- so the oracle takes the same route.
    q_norm = float(torch.linalg.norm(q.reshape(-1, 1), 2))
    c_norm = float(torch.linalg.norm(c.reshape(-1, 1), 2))
    sigma_max = float(spectral_norm_estimate(op, power_iters, b0))     # :22
    eta = (1.0 / sigma_max) * 0.9      # torch evaluates `0.9 / tensor` as reciprocal(tensor) * 0.9
    omega = c_norm / q_norm if (q_norm > 1e-6 and c_norm > 1e-6) else 1.0   # :23
    theta = 1.0
    beta_suff, beta_nec, beta_art = 0.2, 0.8, 0.36                     # :28
    if x_init is not None and y_init is not None:
        x, y = _vec(x_init, dt), _vec(y_init, dt)
    else:
        x, y = torch.zeros(n, dtype=dt), torch.zeros(m, dtype=dt)
    lam_prev = torch.zeros(n, dtype=dt)                                 # :39-40
    if unscaled is not None:
        op_u, c_u, q_u, l_u, u_u = unscaled
        c_u, q_u, l_u, u_u = (_vec(v, dt) for v in (c_u, q_u, l_u, u_u))
        l_u = torch.where(lo_is_inf, torch.zeros_like(l_u), l_u)       # :159-160
        u_u = torch.where(up_is_inf, torch.zeros_like(u_u), u_u)

    def residuals(xx, yy):
        return compute_residuals_and_duality_gap(xx, yy, c, q, op, m_ineq, lo_is_inf, up_is_inf, l_dual, u_dual)

    n_restart = k = j = 0
    kkt_first = 0.0                                                     # :48
    status = "Unsolved (KKT passes limit exceeded)"
    trace: List[RestartRecord] = []
    prim_obj = float("nan")
    while j < max_kkt:                                                  # :54
        t = 0
        x_sum, y_sum, eta_sum = torch.zeros_like(x), torch.zeros_like(y), 0.0
        x_restart, y_restart = x.clone(), y.clone()
        kind, cand = "Limit", "Current"
        while j < max_kkt:                                              # :67
            if time.time() - t_start + time_used >= time_limit:        # :68-74
                status = "Unsolved (Time limit exceeded)"
                break
            k += 1
            x_prev, y_prev = x, y
            if adaptive:
                x, y, eta_w, eta_next = adaptive_one_step_pdhg(x, y, c, q, op, l, u, m_ineq, eta, omega, theta, k)
            else:
                x, y = fixed_one_step_pdhg(x, y, c, q, op, l, u, m_ineq, eta, omega, theta)
                eta_w = eta_next = eta
            j += 1
            if infeasibility_detect and k > 1:                          # :89-101
                lam = project_lambda_box(c - op.rmv(y), lo_is_inf, up_is_inf)
                verdict = detect_infeasibility(x, y, x_prev, y_prev, lam, lam_prev, c, q, op, l, u, m_ineq, infeas_tol)
                j += 1
                if verdict:
                    return Solution(x, y, float(torch.dot(c, x)), k, n_restart, j, verdict,
                                    time.time() - t_start + time_used, sigma_max, trace)
                lam_prev = lam
            t += 1
            x_sum += eta_w * x                                          # :107-109
            y_sum += eta_w * y
            eta_sum += eta_w
            eta = eta_next
            if t % restart_period == 0:                                 # :115-142
                x_avg, y_avg = x_sum / eta_sum, y_sum / eta_sum
                kkt_cur = residuals(x, y).kkt(omega)
                kkt_avg = residuals(x_avg, y_avg).kkt(omega)
                kkt_min = min(kkt_cur, kkt_avg)
                kkt_prev = residuals(x_prev, y_prev).kkt(omega)
                j += 3
                use_avg = kkt_cur >= kkt_avg
                if kkt_min <= beta_suff * kkt_first:
                    kind = "Sufficient"
                elif kkt_min <= beta_nec * kkt_first and kkt_min > kkt_prev:
                    kind = "Necessary"
                elif t >= beta_art * k:
                    kind = "Artificial"
                else:
                    continue
                cand = "Average" if use_avg else "Current"
                if use_avg:
                    x, y = x_avg, y_avg
                if verbose:
                    print(f"{kind} restart at iteration {t} using the {cand} iterate.")
                break
        if status == "Unsolved (Time limit exceeded)":
            break
        n_restart += 1                                                  # :148
        if primal_update:
            omega = primal_weight_update(x_restart, x, y_restart, y, omega, 0.5)
        kkt_first = residuals(x, y).kkt(omega)                          # :153
        j += 1
        if unscaled is not None:                                        # :157-163
            res = compute_residuals_and_duality_gap(scaling.D_col * x, scaling.D_row * y, c_u, q_u, op_u, m_ineq,
                                                    lo_is_inf, up_is_inf, l_u, u_u)
        else:
            res = residuals(x, y)
        j += 1
        prim_obj = res.prim_obj
        trace.append(RestartRecord(k, t, kind, cand, res.prim_obj, res.adj_dual, res.gap, res.pr, res.dr, omega, eta))
        if verbose:
            print(f"[{k}] Primal Obj: {res.prim_obj:.4f}, Adjusted Dual Obj: {res.adj_dual:.4f}, "
                  f"Gap: {res.gap / (1 + abs(res.prim_obj) + abs(res.adj_dual)):.2e}, "
                  f"Prim Res: {res.pr / (1 + q_norm):.2e}, Dual Res: {res.dr / (1 + c_norm):.2e}")
        if check_termination(res, q_norm, c_norm, tol):                 # :173-177
            status = "Solved"
            break
    return Solution(x, y, prim_obj, k, n_restart, j, status, time.time() - t_start + time_used, sigma_max, trace)


# --------------------------------------------------------------------------
# spectral_casting.py restated ("fishnet" start point)
# --------------------------------------------------------------------------
def _batched_step(op, X, Y, c, q, l, u, eta, omega, m_ineq):
    """PDLP/spectral_casting.py:254-293: one fixed PDHG step on every column."""
    G = c[:, None] - op.rmm(Y)
    Xn = torch.minimum(torch.maximum(X - (eta / omega) * G, l[:, None]), u[:, None])
    Yn = Y + (eta * omega) * (q[:, None] - op.mm(2 * Xn - X))
    if m_ineq > 0:
        Yn[:m_ineq] = torch.clamp(Yn[:m_ineq], min=0.0)
    return Xn, Yn


def _signed_gaps(op, X, Y, c, q, l_dual, u_dual, lo_is_inf, up_is_inf):
    """PDLP/spectral_casting.py:217-234: signed duality gap of every column."""
    G = c[:, None] - op.rmm(Y)
    gaps = []
    for col in range(G.shape[1]):
        lam = project_lambda_box(G[:, col], lo_is_inf, up_is_inf)
        adj = (torch.dot(q, Y[:, col]) + torch.sum(l_dual * torch.clamp(lam, min=0.0))
               + torch.sum(u_dual * torch.clamp(lam, max=0.0)))
        gaps.append(adj - torch.dot(c, X[:, col]))
    return torch.stack(gaps)


def spectral_cast(K, c, q, l, u, m_ineq, k=32, s=2, i=5, backend="torch", rng_log=None):
    """PDLP/spectral_casting.py:5-159.  Draws from the global torch RNG in the same
    order as the reference: randn(n,1) [power iteration, 25 rounds], randn(n,2^i)
    [points], randn(n,1) [power iteration, 50 rounds], then rand(new_j) per bred
    child.  Returns (x0, y0) flattened."""
    op = K if isinstance(K, Operator) else Operator(K, backend=backend)
    dt = op.dtype
    c, q, l, u = (_vec(v, dt) for v in (c, q, l, u))
    m, n = op.shape
    r = spectral_norm_estimate(op, 25)                                  # :47
    npts = 2 ** i
    P = torch.randn(n, npts).to(dt)                                     # :51
    centre = (r / n ** 0.5) * torch.ones(n, 1, dtype=dt)                # :54
    P = P * r / torch.linalg.norm(P, dim=0, keepdim=True) + centre      # :58-60
    Y = op.mm(P)                                                        # :99
    q_norm = float(torch.linalg.norm(q.reshape(-1, 1), 2))              # :178-179 (matrix 2-norm)
    c_norm = float(torch.linalg.norm(c.reshape(-1, 1), 2))
    eta = (1.0 / float(spectral_norm_estimate(op, 50))) * 0.9           # :183 (0.9 / tensor = reciprocal * 0.9)
    omega = c_norm / q_norm if (q_norm > 1e-6 and c_norm > 1e-6) else 1.0
    lo_is_inf = torch.isinf(l) & (l < 0)
    up_is_inf = torch.isinf(u) & (u > 0)
    l_dual = torch.where(lo_is_inf, torch.zeros_like(l), l)
    u_dual = torch.where(up_is_inf, torch.zeros_like(u), u)
    rnd = 0
    while P.shape[1] > 1:                                               # :103
        for _ in range(k):
            P, Y = _batched_step(op, P, Y, c, q, l, u, eta, omega, m_ineq)
        old = P.shape[1]
        gaps = _signed_gaps(op, P, Y, c, q, l_dual, u_dual, lo_is_inf, up_is_inf)
        order = torch.argsort(gaps)                                     # :237
        keep = max(1, old // s)
        P, Y = P[:, order][:, :keep], Y[:, order][:, :keep]
        new = P.shape[1]
        if rnd % 2 == 1 and new > 1:                                    # :117-154
            mid_x, mid_y = P.mean(dim=1, keepdim=True), Y.mean(dim=1, keepdim=True)
            kids_x, kids_y = [], []
            for _ in range(old - new - 1):
                w = torch.rand(new).to(dt)
                w = w / w.sum()
                kids_x.append(P @ w.reshape(-1, 1))
                kids_y.append(Y @ w.reshape(-1, 1))
            if kids_x:
                P = torch.cat([P] + kids_x + [mid_x], dim=1)
                Y = torch.cat([Y] + kids_y + [mid_y], dim=1)
        rnd += 1
    return P.reshape(-1), Y.reshape(-1)
