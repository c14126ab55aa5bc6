"""Seeded synthetic LP families for the configs named in BASELINE.json.

All LPs are in torchPDLP's standard form (reference PDLP/util.py:78-83, 250-261)

    min c^T x   s.t.  G x >= h  (first m_ineq rows of K),  A x = b,  l <= x <= u
    K = [G; A],  q = [h; b]

and are returned as a :class:`SyntheticLP` holding K as a scipy CSR matrix (fp64,
int32 indices, sorted columns) plus dense numpy vectors.  Nothing here imports the
reference; ``feasible_lp`` restates the recipe of the reference's (unseeded,
pulp-dependent) Packages/generate_feasible_lp.py:18-41 with explicit seeds so the
LP is reproducible.
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp


@dataclass
class SyntheticLP:
    name: str
    K: sp.csr_matrix      # (m, n) fp64 CSR, rows [0, m_ineq) are ">=" rows
    c: np.ndarray         # (n,)
    q: np.ndarray         # (m,)
    l: np.ndarray         # (n,) may hold -inf
    u: np.ndarray         # (n,) may hold +inf
    m_ineq: int

    @property
    def m(self) -> int:
        return self.K.shape[0]

    @property
    def n(self) -> int:
        return self.K.shape[1]

    @property
    def nnz(self) -> int:
        return self.K.nnz

    def bytes_per_iteration(self, value_bytes: int = 8) -> int:
        """Algorithmic bytes of one PDHG iteration:
        CSR(K) and CSR(K^T) streamed once (value + int32 index), both row-pointer
        arrays, 9 n-length and 6 m-length vector streams."""
        vb = value_bytes
        return (2 * (vb + 4) * self.nnz + 4 * (self.m + self.n + 2)
                + 9 * vb * self.n + 6 * vb * self.m)

    def torch_tensors(self, dtype=None, device="cpu", layout="csr"):
        """(c, K, q, m_ineq, l, u) like reference util.mps_to_standard_form
        (PDLP/util.py:269): column vectors (n,1)/(m,1); K dense / COO / CSR."""
        import torch
        warnings.filterwarnings("ignore", message=".*Sparse (CSR|invariant).*")
        dtype = dtype or torch.float64
        # copies: the reference mutates its inputs (with precondition=True it overwrites the infinite entries of the
        # caller's l and u, primal_dual_hybrid_gradient.py:157-158) and must not reach this LP's arrays through them
        col = lambda v: torch.tensor(np.asarray(v), dtype=dtype).reshape(-1, 1).to(device)
        K = self.K
        if layout == "dense":
            Kt = torch.as_tensor(K.toarray(), dtype=dtype)
        elif layout == "coo":
            coo = K.tocoo()
            idx = torch.as_tensor(np.vstack([coo.row, coo.col]).astype(np.int64))
            Kt = torch.sparse_coo_tensor(idx, torch.as_tensor(coo.data, dtype=dtype),
                                         size=K.shape).coalesce()
        elif layout == "csr":
            Kt = torch.sparse_csr_tensor(
                torch.as_tensor(K.indptr.astype(np.int64)),
                torch.as_tensor(K.indices.astype(np.int64)),
                torch.as_tensor(K.data, dtype=dtype), size=K.shape)
        else:
            raise ValueError(f"unknown layout {layout!r}")
        return col(self.c), Kt.to(device), col(self.q), self.m_ineq, col(self.l), col(self.u)


def _finish(name, G, A, x_feas, rng, m_ineq, m_eq, n, slack=(0.1, 5.0), box=(1.0, 5.0)):
    """Shared tail of the feasible-by-construction recipe.  Draw order after
    x_feas: slack, l-width, u-width, c (generate_feasible_lp.py:31-41).  The
    reference writes its rows as G x <= h (generate_feasible_lp.py:53) which the
    MPS reader turns into -G x >= -h (PDLP/util.py:228-230), hence K = [-G; A]."""
    h = G @ x_feas + rng.uniform(slack[0], slack[1], size=m_ineq)
    b = A @ x_feas
    l = np.maximum(x_feas - rng.uniform(box[0], box[1], size=n), -1e4)
    u = np.minimum(x_feas + rng.uniform(box[0], box[1], size=n), 1e4)
    c = rng.normal(size=n)
    K = sp.vstack([-G, A], format="csr")
    K.sort_indices()
    K = sp.csr_matrix((K.data.astype(np.float64), K.indices.astype(np.int32),
                       K.indptr.astype(np.int32)), shape=K.shape)
    q = np.concatenate([-h, b])
    return SyntheticLP(name, K, c, q, l, u, m_ineq)


def feasible_lp(n=2000, m_ineq=800, m_eq=200, density=0.01, seed=0, name="cfg1"):
    """Config 1: random feasible LP, scipy.sparse.random matrices with explicit
    seeds (seed+1 for G, seed+2 for A)."""
    rng = np.random.default_rng(seed)
    x_feas = rng.uniform(-10, 10, size=n)
    G = sp.random(m_ineq, n, density=density, format="csr",
                  random_state=np.random.default_rng(seed + 1))
    A = sp.random(m_eq, n, density=density, format="csr",
                  random_state=np.random.default_rng(seed + 2))
    return _finish(name, G, A, x_feas, rng, m_ineq, m_eq, n)


def _stratified_rows(m, n, nnz_per_row, rng):
    """m rows with exactly nnz_per_row entries: one uniformly placed column in
    each of nnz_per_row equal-width strata of [0, n) (distinct, sorted), values
    U[0,1).  Vectorised; O(nnz) memory."""
    k = nnz_per_row
    edges = (np.arange(k + 1, dtype=np.int64) * n) // k
    width = (edges[1:] - edges[:-1]).astype(np.float64)
    cols = edges[:-1][None, :] + np.floor(rng.random((m, k)) * width[None, :]).astype(np.int64)
    vals = rng.random((m, k))
    indptr = np.arange(m + 1, dtype=np.int64) * k
    return sp.csr_matrix((vals.ravel(), cols.ravel().astype(np.int32), indptr), shape=(m, n))


def uniform_rows_lp(m_ineq=80_000, m_eq=20_000, n=200_000, nnz_per_row=50, seed=0, name="cfg2"):
    """Config 2: bounded feasible LP with exactly nnz_per_row entries per row
    (default 100k x 200k, 5M nnz), finite bounds."""
    rng = np.random.default_rng(seed)
    x_feas = rng.uniform(-10, 10, size=n)
    G = _stratified_rows(m_ineq, n, nnz_per_row, np.random.default_rng(seed + 1))
    A = _stratified_rows(m_eq, n, nnz_per_row, np.random.default_rng(seed + 2))
    return _finish(name, G, A, x_feas, rng, m_ineq, m_eq, n)


def infeasible_lp(m_ineq=80_000, m_eq=20_000, n=200_000, nnz_per_row=50, seed=0, delta=1.0,
                  name="cfg3"):
    """Config 3: config-2 LP with a contradictory pair of ">=" rows prepended:
    a^T x >= beta and -a^T x >= -beta + delta (delta > 0), so no x satisfies both
    and y = (1, 1, 0, ...) is a Farkas certificate (K^T y = 0, q^T y = delta > 0)."""
    base = uniform_rows_lp(m_ineq, m_eq, n, nnz_per_row, seed, name)
    rng = np.random.default_rng(seed + 3)
    a = _stratified_rows(1, n, nnz_per_row, rng)
    beta = float(rng.uniform(-1, 1))
    pair = sp.vstack([a, -a], format="csr")
    K = sp.vstack([pair, base.K], format="csr")
    K.sort_indices()
    K = sp.csr_matrix((K.data, K.indices.astype(np.int32), K.indptr.astype(np.int32)), shape=K.shape)
    q = np.concatenate([[beta, -beta + delta], base.q])
    return SyntheticLP(name, K, base.c, q, base.l, base.u, base.m_ineq + 2)


def unbounded_lp(m_ineq=40, m_eq=10, n=100, nnz_per_row=10, seed=0, name="unbounded"):
    """Small dual-infeasible (primal unbounded) LP: config-2 recipe with
    non-negative G, A = 0 columns on a free ray: variable 0 gets u = +inf,
    c_0 = -1 and no constraint entries, so x_0 -> +inf is a primal ray."""
    base = uniform_rows_lp(m_ineq, m_eq, n, nnz_per_row, seed, name)
    K = base.K.tolil()
    K[:, 0] = 0
    K = sp.csr_matrix(K.tocsr())
    K.eliminate_zeros()
    K.sort_indices()
    K = sp.csr_matrix((K.data, K.indices.astype(np.int32), K.indptr.astype(np.int32)), shape=K.shape)
    c = base.c.copy(); c[0] = -1.0
    u = base.u.copy(); u[0] = np.inf
    return SyntheticLP(name, K, c, base.q, base.l, u, base.m_ineq)


def powerlaw_rows_lp(m_ineq=1_600_000, m_eq=400_000, n=4_000_000, mean_nnz=25, alpha=1.7,
                     min_nnz=2, max_nnz=100_000, seed=0, name="cfg4"):
    """Config 4: transportation-style LP with power-law (discrete Pareto) row
    lengths (load-balance stress).  Coefficients are +/-1 flow terms and U[0.5,2)
    capacity terms; feasible by construction like config 1."""
    rng = np.random.default_rng(seed)
    m = m_ineq + m_eq
    x_feas = rng.uniform(-10, 10, size=n)
    mrng = np.random.default_rng(seed + 1)
    raw = min_nnz * (1.0 - mrng.random(m)) ** (-1.0 / alpha)
    raw = np.minimum(raw, max_nnz)
    # rescale the tail (not the minimum) so that the mean hits mean_nnz
    scale = (mean_nnz - min_nnz) / max(raw.mean() - min_nnz, 1e-12)
    lens = np.clip(np.rint(min_nnz + (raw - min_nnz) * scale), min_nnz, min(max_nnz, n)).astype(np.int64)
    # make sure the stress case exists: a handful of rows at the cap
    big = mrng.choice(m, size=max(1, m // 500_000), replace=False)
    lens[big] = min(max_nnz, n)
    indptr = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(lens, out=indptr[1:])
    nnz = int(indptr[-1])
    # stratified columns per row: position j of a row of length L falls in stratum j
    row_of = np.repeat(np.arange(m, dtype=np.int64), lens)
    pos = np.arange(nnz, dtype=np.int64) - indptr[row_of]
    L = lens[row_of]
    lo = (pos * n) // L
    hi = ((pos + 1) * n) // L
    cols = lo + np.floor(mrng.random(nnz) * (hi - lo)).astype(np.int64)
    kind = mrng.random(nnz)
    vals = np.where(kind < 0.4, 1.0, np.where(kind < 0.8, -1.0, 0.5 + 1.5 * mrng.random(nnz)))
    M = sp.csr_matrix((vals, cols.astype(np.int32), indptr), shape=(m, n))
    G, A = M[:m_ineq], M[m_ineq:]
    return _finish(name, G, A, x_feas, rng, m_ineq, m_eq, n)


def mixed_bounds_lp(n=600, m_ineq=220, m_eq=60, density=0.03, seed=5, name="mixed_bounds"):
    """Small feasible, bounded LP whose variables have all four kinds of bounds (box, lower only, upper only,
    free): every 5th variable loses the bound its cost pushes away from (so the LP stays bounded), up to 20
    variables (every 11th with at least 3 constraint entries) lose both and get cost 0.  Exercises the one-sided / free branches of the lambda projection
    and of the dual objective (reference helpers.py:3-39, 70-112)."""
    lp = feasible_lp(n, m_ineq, m_eq, density, seed, name)
    l, u, c = lp.l.copy(), lp.u.copy(), lp.c.copy()
    for j in range(0, n, 5):
        if c[j] > 0:
            u[j] = np.inf
        else:
            l[j] = -np.inf
    # free variables: cost 0 (the objective cannot run away along them), present in at least 3 rows
    col_nnz = np.diff(lp.K.tocsc().indptr)
    for j in [j for j in range(3, n, 11) if col_nnz[j] >= 3][:20]:
        l[j], u[j], c[j] = -np.inf, np.inf, 0.0
    return SyntheticLP(name, lp.K, c, lp.q, l, u, lp.m_ineq)


CONFIGS = {
    "cfg1": lambda **kw: feasible_lp(**kw),
    "cfg2": lambda **kw: uniform_rows_lp(**kw),
    "cfg3": lambda **kw: infeasible_lp(**kw),
    "cfg4": lambda **kw: powerlaw_rows_lp(**kw),
}


def make(name: str, **kw) -> SyntheticLP:
    return CONFIGS[name](**kw)


def isolated_infeasible_lp(m_ineq=80_000, m_eq=20_000, n=200_000, nnz_per_row=50, seed=0, beta=5.0, delta=6.0,
                           name="cfg3_isolated"):
    """Config 3, the variant whose certificate the reference's heuristic reaches at full size: the config-2 LP with
    variable 0 taken out of every row, cost 0, bounds [-10, 10], and a contradictory pair of ">=" rows on it alone,
    x_0 >= beta and -x_0 >= -beta + delta (x_0 <= beta - delta < beta).  The infeasible part is decoupled from the
    rest, so the differences of the iterates (what enhancements.py:106-161 tests) settle as fast as config 2
    converges; y = e_1 + e_2 is a Farkas certificate (K^T y = 0, q^T y = delta > 0) and the violation delta / 2 per
    row is far above the termination tolerance 1e-4 (1 + ||q||)."""
    base = uniform_rows_lp(m_ineq, m_eq, n, nnz_per_row, seed, name)
    K = base.K.tocsc(copy=True)
    K.data[K.indptr[0]:K.indptr[1]] = 0.0
    K = sp.csr_matrix(K.tocsr())
    K.eliminate_zeros()
    pair = sp.csr_matrix((np.array([1.0, -1.0]), (np.array([0, 1]), np.array([0, 0]))), shape=(2, n))
    K = sp.vstack([pair, K], format="csr")
    K.sort_indices()
    K = sp.csr_matrix((K.data, K.indices.astype(np.int32), K.indptr.astype(np.int32)), shape=K.shape)
    q = np.concatenate([[beta, -beta + delta], base.q])
    c, l, u = base.c.copy(), base.l.copy(), base.u.copy()
    c[0], l[0], u[0] = 0.0, -10.0, 10.0
    return SyntheticLP(name, K, c, q, l, u, base.m_ineq + 2)
