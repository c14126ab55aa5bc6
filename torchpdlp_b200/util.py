"""Drop-in for torchPDLP's PDLP/util.py: MPS reader and the Timer helper.

``mps_to_standard_form`` follows the reference's parsing rules (util.py:76-269) but
builds K **sparse** (CSR) instead of going through a dense row dictionary, so
it scales to the sizes the engine is built for, and takes the dtype as an argument
(the reference hard-codes float32, util.py:240-246).  This is host-side I/O: no
kernel is involved and ``device='cpu'`` is fine here.
"""
from __future__ import annotations

import time
from collections import OrderedDict

import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore", message=".*Sparse (CSR|invariant).*")


class Timer:
    """``with Timer("label"): ...`` prints the elapsed wall time (util.py:6-27)."""

    def __init__(self, label="Elapsed time"):
        self.label = label

    def __enter__(self):
        self.start = time.perf_counter()
        return self

    def __exit__(self, *args):
        self.end = time.perf_counter()
        self.elapsed = self.end - self.start
        print(f"{self.label}: {self.elapsed:.6f} seconds")


def sparse_vs_dense(A, device="cpu", kkt_passes=10):
    """util.py:29-74 times dense vs COO mat-vecs and returns the faster tensor.  The
    engine has one data path (sparse, SELL-packed on the GPU), so there is nothing
    to choose: return A as a sparse CSR tensor on ``device``."""
    assert A.dim() == 2, "Input must be a 2D matrix"
    A = A.to(device)
    return A if A.layout == torch.sparse_csr else A.to_sparse_csr()


def mps_to_standard_form(mps_file, device="cpu", support_sparse=False, verbose=False, dtype=torch.float64,
                         reference_bounds=True, dense=False):
    """Parse a free-format MPS file into ``(c, K, q, m_ineq, l, u)`` with
    ``min c^T x  s.t.  K[:m_ineq] x >= q[:m_ineq],  K[m_ineq:] x = q[m_ineq:],  l <= x <= u``
    (util.py:76-269).  Rules kept from the reference: "L" rows are negated into ">="
    rows (:228-230); RANGES rows become two ">=" rows (:195-219); rows keep their
    order of appearance, inequalities first; default bounds [0, +inf) (:235-236);
    bound types LO / UP / FX, and FR mapped to [0, +inf) (:162-164) when
    ``reference_bounds`` is True (False gives the standard FR / MI / PL / BV
    meaning).  K is returned as a sparse CSR tensor (``dense=True`` gives the
    reference's dense layout)."""
    section = None
    row_types: "OrderedDict[str, str]" = OrderedDict()
    obj_row = None
    var_index: "OrderedDict[str, int]" = OrderedDict()
    entries = []                     # (row name, var idx, value)
    rhs, ranges, bounds = {}, {}, {}
    with open(mps_file, "r") as f:
        for raw in f:
            line = raw.strip()
            if not line or raw.startswith("*"):
                continue
            head = line.split()[0]
            if head in ("NAME", "ENDATA"):
                continue
            if line in ("ROWS", "COLUMNS", "RHS", "RANGES", "BOUNDS"):
                section = line
                continue
            tok = line.split()
            if section == "ROWS":
                sense, name = tok[0], tok[1]
                row_types[name] = sense
                if sense == "N":
                    obj_row = name
            elif section == "COLUMNS":
                if len(tok) >= 3 and tok[1] == "'MARKER'":
                    continue
                j = var_index.setdefault(tok[0], len(var_index))
                for i in range(1, len(tok) - 1, 2):
                    entries.append((tok[i], j, float(tok[i + 1])))
            elif section == "RHS":
                for i in range(1, len(tok) - 1, 2):
                    rhs[tok[i]] = float(tok[i + 1])
            elif section == "RANGES":
                for i in range(1, len(tok) - 1, 2):
                    ranges[tok[i]] = float(tok[i + 1])
            elif section == "BOUNDS":
                btype, var = tok[0], tok[2]
                val = float(tok[3]) if len(tok) > 3 else None
                b = bounds.setdefault(var, {})
                if btype == "LO":
                    b["lo"] = val
                elif btype == "UP":
                    b["up"] = val
                elif btype == "FX":
                    b["lo"] = b["up"] = val
                elif btype == "FR":
                    if reference_bounds:
                        b["lo"], b["up"] = 0.0, float("inf")
                    else:
                        b["lo"], b["up"] = -float("inf"), float("inf")
                elif not reference_bounds:
                    if btype == "MI":
                        b["lo"] = -float("inf")
                    elif btype == "PL":
                        b["up"] = float("inf")
                    elif btype == "BV":
                        b["lo"], b["up"] = 0.0, 1.0
    n = len(var_index)
    c = np.zeros(n)
    # map every constraint row to its output rows: list of (out block, out index, sign)
    g_rhs, a_rhs = [], []
    placement = {}
    for name, sense in row_types.items():
        if name == obj_row:
            continue
        r, rng = rhs.get(name, 0.0), ranges.get(name)
        if rng is not None:
            if sense == "G":
                lb, ub = r, r + abs(rng)
            elif sense == "L":
                ub, lb = r, r - abs(rng)
            elif sense == "E":
                lb, ub = (r, r + rng) if rng > 0 else (r + rng, r)
            else:
                raise ValueError(f"Unsupported ranged sense: {sense}")
            placement[name] = [("G", len(g_rhs), 1.0), ("G", len(g_rhs) + 1, -1.0)]
            g_rhs += [lb, -ub]
        elif sense == "E":
            placement[name] = [("A", len(a_rhs), 1.0)]
            a_rhs.append(r)
        elif sense == "G":
            placement[name] = [("G", len(g_rhs), 1.0)]
            g_rhs.append(r)
        elif sense == "L":
            placement[name] = [("G", len(g_rhs), -1.0)]
            g_rhs.append(-r)
    m_ineq, m_eq = len(g_rhs), len(a_rhs)
    coo = {}                                     # (row, col) -> value; a repeated entry overwrites, like util.py:183
    for name, j, v in entries:
        if name == obj_row:
            c[j] = v
            continue
        for block, idx, sign in placement.get(name, ()):
            coo[(idx if block == "G" else m_ineq + idx, j)] = sign * v
    if coo:
        keys = np.array(list(coo.keys()), dtype=np.int64)
        vals = np.array(list(coo.values()), dtype=np.float64)
    else:
        keys, vals = np.zeros((0, 2), dtype=np.int64), np.zeros(0)
    m = m_ineq + m_eq
    import scipy.sparse as sp
    Ks = sp.csr_matrix((vals, (keys[:, 0], keys[:, 1])), shape=(m, n))
    Ks.sort_indices()
    l = np.array([bounds.get(v, {}).get("lo", 0.0) for v in var_index], dtype=np.float64)
    u = np.array([bounds.get(v, {}).get("up", float("inf")) for v in var_index], dtype=np.float64)
    q = np.array(g_rhs + a_rhs, dtype=np.float64)
    col = lambda a: torch.as_tensor(a, dtype=dtype, device=device).reshape(-1, 1)
    K = torch.sparse_csr_tensor(torch.from_numpy(Ks.indptr.astype(np.int64)), torch.from_numpy(Ks.indices.astype(np.int64)),
                                torch.from_numpy(Ks.data).to(dtype), size=(m, n)).to(device)
    if dense:
        K = K.to_dense()
    if verbose:
        print(f"Parsed {mps_file}: {m_ineq} inequality rows, {m_eq} equality rows, {n} variables, {Ks.nnz} nonzeros")
        print("Using Sparse operations" if not dense else "Using Dense operations")
    return col(c), K, col(q), m_ineq, col(l), col(u)


def write_mps(path, c, G, h, A, b, l, u, name="LP"):
    """Free-format MPS of  min c^T x, G x <= h, A x = b, l <= x <= u  (what the reference's generator
    writes through pulp, Packages/generate_feasible_lp.py:43-58).  G, A: scipy sparse or None."""
    import scipy.sparse as sp
    n = len(c)
    G = sp.csc_matrix(G) if G is not None else sp.csc_matrix((0, n))
    A = sp.csc_matrix(A) if A is not None else sp.csc_matrix((0, n))
    with open(path, "w") as f:
        f.write(f"NAME {name}\nROWS\n N OBJ\n")
        for i in range(G.shape[0]):
            f.write(f" L ineq_{i}\n")
        for i in range(A.shape[0]):
            f.write(f" E eq_{i}\n")
        f.write("COLUMNS\n")
        for j in range(n):
            f.write(f" x{j} OBJ {float(c[j])!r}\n")
            for M, tag in ((G, "ineq"), (A, "eq")):
                for p in range(M.indptr[j], M.indptr[j + 1]):
                    f.write(f" x{j} {tag}_{M.indices[p]} {float(M.data[p])!r}\n")
        f.write("RHS\n")
        for i in range(G.shape[0]):
            f.write(f" RHS ineq_{i} {float(h[i])!r}\n")
        for i in range(A.shape[0]):
            f.write(f" RHS eq_{i} {float(b[i])!r}\n")
        f.write("BOUNDS\n")
        for j in range(n):
            lo, up = float(l[j]), float(u[j])
            if np.isfinite(lo):
                f.write(f" LO BND x{j} {lo!r}\n")
            elif not np.isfinite(up):
                f.write(f" FR BND x{j}\n")
            else:
                f.write(f" MI BND x{j}\n")
            if np.isfinite(up):
                f.write(f" UP BND x{j} {up!r}\n")
        f.write("ENDATA\n")
