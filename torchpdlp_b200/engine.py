"""Host-side handle around the C-ABI engine (include/pdlp_b200.h).

PyTorch is used here for device memory, streams and format conversion only; every
PDHG iteration, KKT pass, power-iteration round and scaling pass runs in
libpdlp_b200.so.  There is no CPU path: constructing an :class:`Engine` without a
CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import time
from dataclasses import dataclass
from typing import Optional, Tuple

import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore", message=".*Sparse (CSR|invariant).*")

from . import _lib
from ._lib import Check, Residuals, SolveOptions, SolveResult, State, check

STATUS_STRINGS = {
    0: "Solved",                                   # primal_dual_hybrid_gradient.py:174
    1: "Unsolved (KKT passes limit exceeded)",     # :51
    2: "Unsolved (Time limit exceeded)",           # :71
    3: "PRIMAL_INFEASIBLE",                        # enhancements.py:159
    4: "DUAL_INFEASIBLE",                          # enhancements.py:141
}
RESTART_KINDS = {0: None, 1: "Sufficient", 2: "Necessary", 3: "Artificial", 4: "Limit"}

# how a block of iterations is issued: 0 = one launch per kernel, 1 = CUDA graph of the block's kernels,
# 2 = ONE persistent cooperative kernel per block (grid barriers instead of kernel boundaries)
LOOP_EAGER, LOOP_GRAPH, LOOP_PERSISTENT = 0, 1, 2
# Measured on B200 (round 1): graph + programmatic dependent launch 66 us/iteration on config 2 and 16 us on
# config 1; the persistent kernel 72 us and 16 us (two grid barriers + a redundant grid reduction per iteration
# cost as much as the two dependent launches they replace), so the graph is the default.
DEFAULT_LOOP_MODE = int(os.environ.get("PDLP_LOOP_MODE", LOOP_GRAPH))


def loop_mode(use_graph) -> int:
    """bool / int -> loop mode: True means "the fastest loop" (DEFAULT_LOOP_MODE), False eager."""
    if use_graph is True:
        return DEFAULT_LOOP_MODE
    return int(use_graph)


def _require_cuda(device) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("torchpdlp_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch.device(device if device is not None else "cuda")
    if dev.type != "cuda":
        raise RuntimeError(f"torchpdlp_b200 runs on CUDA devices only, got {dev}")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def to_device_csr(K, device: torch.device, dtype: torch.dtype):
    """(crow int32, col int32, val) of K and of K^T on ``device``, columns sorted
    inside every row.  Accepts a torch tensor (dense / COO / CSR, any device), a
    scipy sparse matrix, or a (crow, col, val, shape) tuple."""
    if isinstance(K, tuple):
        crow, col, val, shape = K
        Kc = torch.sparse_csr_tensor(torch.as_tensor(crow, dtype=torch.int64), torch.as_tensor(col, dtype=torch.int64),
                                     torch.as_tensor(val, dtype=dtype), size=tuple(shape))
    elif isinstance(K, torch.Tensor):
        Kc = K
    else:  # scipy
        Ks = K.tocsr()
        Ks.sort_indices()
        Kc = torch.sparse_csr_tensor(torch.from_numpy(Ks.indptr.astype(np.int64)),
                                     torch.from_numpy(Ks.indices.astype(np.int64)),
                                     torch.from_numpy(np.ascontiguousarray(Ks.data)).to(dtype), size=Ks.shape)
    Kc = Kc.to(device)
    if Kc.dtype != dtype:
        Kc = Kc.to(dtype)

    def csr_of(r, c, v, nrows):
        counts = torch.bincount(r, minlength=nrows)
        crow = torch.zeros(nrows + 1, dtype=torch.int64, device=device)
        torch.cumsum(counts, 0, out=crow[1:])
        return crow.to(torch.int32).contiguous(), c.to(torch.int32).contiguous(), v.contiguous()

    m, n = Kc.shape
    csr = None
    if Kc.layout == torch.sparse_csr:
        # fast path: a canonical CSR (columns strictly increasing inside every row: sorted, no duplicates) is taken
        # as it is -- no round trip through COO and no sort for K itself
        crow, col, vals = Kc.crow_indices(), Kc.col_indices(), Kc.values()
        nnz = int(vals.numel())
        rows = torch.repeat_interleave(torch.arange(m, device=device), crow[1:] - crow[:-1], output_size=nnz)
        canonical = nnz < 2 or bool(torch.all((col[1:] > col[:-1]) | (rows[1:] != rows[:-1])))
        if canonical:
            cols = col
            csr = (crow.to(torch.int32).contiguous(), col.to(torch.int32).contiguous(), vals.contiguous())
    if csr is None:
        if Kc.layout == torch.strided:
            coo = Kc.to_sparse_coo().coalesce()
        elif Kc.layout == torch.sparse_coo:
            coo = Kc.coalesce()
        elif Kc.layout == torch.sparse_csr:
            # (to_sparse_coo() would flag its result as coalesced without looking: build the COO from the indices)
            coo = torch.sparse_coo_tensor(torch.stack([rows, col.to(torch.int64)]), vals, size=(m, n)).coalesce()
        else:
            raise TypeError(f"unsupported K layout {Kc.layout}")
        rows, cols, vals = coo.indices()[0], coo.indices()[1], coo.values()
        csr = csr_of(rows, cols, vals, m)                 # coalesced COO is sorted by (row, col)
    # K^T: entries are in (row, col) order, so a STABLE sort by column alone gives (col, row) order -- 32-bit keys
    order = torch.sort(cols.to(torch.int32), stable=True)[1]
    csc = csr_of(cols[order], rows[order], vals[order], n)
    return (m, n, int(vals.numel())), csr, csc


def host_csr_arrays(K, dtype: torch.dtype):
    """(crow, col, vals, (m, n)) as contiguous CPU tensors when ``K`` is a CSR matrix in HOST memory (torch sparse CSR
    on the CPU, scipy CSR, or a (crow, col, val, shape) tuple of host arrays), else None.  Nothing is converted
    element-wise here except a value dtype that differs from the engine's: the library takes 4- or 8-byte indices."""
    if isinstance(K, tuple):
        crow, col, val, shape = K
        parts = [torch.as_tensor(a) for a in (crow, col, val)]
        if any(t.device.type != "cpu" for t in parts):
            return None
        crow, col, val = parts
    elif isinstance(K, torch.Tensor):
        if K.layout != torch.sparse_csr or K.device.type != "cpu" or K.dim() != 2:
            return None
        crow, col, val, shape = K.crow_indices(), K.col_indices(), K.values(), K.shape
    elif hasattr(K, "indptr") and hasattr(K, "indices") and getattr(K, "format", None) == "csr":   # scipy CSR
        crow, col, val, shape = torch.from_numpy(K.indptr), torch.from_numpy(K.indices), \
            torch.from_numpy(np.ascontiguousarray(K.data)), K.shape
    else:
        return None
    if crow.dtype not in (torch.int32, torch.int64):
        crow = crow.to(torch.int64)
    if col.dtype != crow.dtype:
        col = col.to(crow.dtype)
    if val.dtype != dtype:
        val = val.to(dtype)
    return crow.contiguous(), col.contiguous(), val.contiguous(), (int(shape[0]), int(shape[1]))


@dataclass
class SolveOutput:
    x: torch.Tensor
    y: torch.Tensor
    prim_obj: float
    k: int
    n: int
    j: int
    status: str
    seconds: float
    device_ms: float
    trace: list


class Engine:
    """One LP (c, K, q, m_ineq, l, u) resident on one GPU."""

    def __init__(self, K, m_ineq: int, c, q, l, u, device=None, dtype: Optional[torch.dtype] = None):
        self.device = _require_cuda(device)
        self.lib = _lib.lib()
        if dtype is None:
            dtype = K.dtype if isinstance(K, torch.Tensor) else torch.float64
        if dtype not in (torch.float64, torch.float32):
            raise TypeError("dtype must be float64 or float32")
        self.dtype = dtype
        trace = os.environ.get("PDLP_TRACE_SETUP")
        t_trace = [time.perf_counter()]

        def lap(what):
            if trace:
                t_trace.append(time.perf_counter())
                print(f"[pdlp engine] {what:28s} {1e3 * (t_trace[-1] - t_trace[-2]):8.3f} ms", file=sys.stderr)

        self.h = None
        self._csr = None             # structure of K as torch tensors, fetched lazily for matrix()
        self.ingest = "device"       # "host": the library copied and transposed a host CSR itself (pdlp_create_host)
        with torch.cuda.device(self.device):
            self.m_ineq = int(m_ineq)
            self.stream = torch.cuda.current_stream(self.device)
            dcode = 0 if dtype == torch.float64 else 1
            host = host_csr_arrays(K, dtype) if os.environ.get("PDLP_HOST_INGEST", "1") != "0" else None
            if host is not None:
                crow, col, val, (m, n) = host
                # the vectors stay where they are (host or this device): the library copies with unified addressing
                vec = lambda v: self._place(torch.as_tensor(v, dtype=dtype).reshape(-1).contiguous())
                cc, qq, ll, uu = vec(c), vec(q), vec(l), vec(u)
                if cc.numel() != n or ll.numel() != n or uu.numel() != n or qq.numel() != m or crow.numel() != m + 1 \
                        or col.numel() != val.numel():
                    raise ValueError("vector sizes do not match K")
                h = C.c_void_p()
                lap("host arrays")
                try:
                    check(self.lib.pdlp_create_host(C.byref(h), dcode, m, n, self.m_ineq, int(val.numel()),
                                                    crow.data_ptr(), col.data_ptr(), crow.element_size(), val.data_ptr(),
                                                    cc.data_ptr(), qq.data_ptr(), ll.data_ptr(), uu.data_ptr(),
                                                    self.stream.cuda_stream))
                    self.h, self.ingest = h, "host"
                    self.m, self.n, self.nnz = m, n, int(val.numel())
                except _lib.PdlpError as err:
                    if err.code != _lib.ERR_NOT_CANONICAL:      # unsorted / duplicate entries: coalesce below
                        raise
            if self.h is None:
                (self.m, self.n, self.nnz), csr, csc = to_device_csr(K, self.device, dtype)
                vec = lambda v, ln: torch.as_tensor(v, dtype=dtype).reshape(-1).to(self.device).contiguous()
                c, q, l, u = vec(c, self.n), vec(q, self.m), vec(l, self.n), vec(u, self.n)
                if c.numel() != self.n or l.numel() != self.n or u.numel() != self.n or q.numel() != self.m:
                    raise ValueError("vector sizes do not match K")
                self._csr = csr          # kept for matrix() / debugging (structure only is reused)
                h = C.c_void_p()
                torch.cuda.synchronize(self.device)
                check(self.lib.pdlp_create(C.byref(h), dcode, self.m, self.n, self.m_ineq,
                                           self.nnz, csr[0].data_ptr(), csr[1].data_ptr(), csr[2].data_ptr(),
                                           csc[0].data_ptr(), csc[1].data_ptr(), csc[2].data_ptr(), c.data_ptr(),
                                           q.data_ptr(), l.data_ptr(), u.data_ptr(), self.stream.cuda_stream))
                self.h = h
        lap("create")
        self.bytes_per_iteration = int(self.lib.pdlp_bytes_per_iteration(self.h))

    # -- lifetime ---------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.pdlp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _place(self, t: torch.Tensor) -> torch.Tensor:
        """A vector the library can read from this device: host memory or this GPU."""
        return t if t.device.type == "cpu" or t.device == self.device else t.to(self.device)

    def set_stream(self, stream: "torch.cuda.Stream"):
        check(self.lib.pdlp_set_stream(self.h, stream.cuda_stream))
        self.stream = stream

    def _new(self, length):
        return torch.empty(length, dtype=self.dtype, device=self.device)

    def _dev(self, v, length):
        t = torch.as_tensor(v, dtype=self.dtype).reshape(-1).to(self.device).contiguous()
        if t.numel() != length:
            raise ValueError(f"expected {length} values, got {t.numel()}")
        return t

    # -- iterate / scalars ------------------------------------------------
    def set_iterate(self, x=None, y=None):
        xd = self._dev(x, self.n) if x is not None else None
        yd = self._dev(y, self.m) if y is not None else None
        check(self.lib.pdlp_set_iterate(self.h, xd.data_ptr() if xd is not None else None,
                                        yd.data_ptr() if yd is not None else None))

    def get_iterate(self) -> Tuple[torch.Tensor, torch.Tensor]:
        x, y = self._new(self.n), self._new(self.m)
        check(self.lib.pdlp_get_iterate(self.h, x.data_ptr(), y.data_ptr()))
        return x, y

    def get_average(self) -> Tuple[torch.Tensor, torch.Tensor]:
        x, y = self._new(self.n), self._new(self.m)
        check(self.lib.pdlp_get_average(self.h, x.data_ptr(), y.data_ptr()))
        return x, y

    def set_step(self, eta: float, omega: float, theta: float = 1.0, adaptive: bool = False):
        check(self.lib.pdlp_set_step(self.h, float(eta), float(omega), float(theta), int(bool(adaptive))))

    def set_options(self, tol=1e-4, primal_update=False, infeasibility_detect=False, infeas_tol=1e-4):
        check(self.lib.pdlp_set_options(self.h, float(tol), int(bool(primal_update)),
                                        int(bool(infeasibility_detect)), float(infeas_tol)))

    def set_counters(self, k: int, t: int = 0):
        check(self.lib.pdlp_set_counters(self.h, int(k), int(t)))

    def state(self) -> State:
        s = State()
        check(self.lib.pdlp_get_state(self.h, C.byref(s)))
        return s

    # -- hot loop -----------------------------------------------------------
    def iterate(self, n_iters: int, save_previous: bool = False, use_graph=True):
        check(self.lib.pdlp_iterate(self.h, int(n_iters), int(bool(save_previous)), loop_mode(use_graph)))

    def time_kernels(self, n_iters: int):
        """Average device milliseconds per launch of the (primal, dual) iteration kernels."""
        a, b = C.c_float(), C.c_float()
        check(self.lib.pdlp_time_kernels(self.h, int(n_iters), C.byref(a), C.byref(b)))
        return a.value, b.value

    def time_kernels_in_loop(self, n_iters: int):
        """(primal, dual, whole iteration) milliseconds with the kernels chained as in the block graph: each launch's
        span from leaving griddepcontrol.wait to its finalize (%globaltimer), and the event-timed wall per iteration."""
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        check(self.lib.pdlp_time_kernels_in_loop(self.h, int(n_iters), C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def restart_check(self, force: bool = False) -> Check:
        chk = Check()
        check(self.lib.pdlp_restart_check(self.h, int(bool(force)), C.byref(chk)))
        return chk

    def kkt(self, x, y, omega: float) -> Residuals:
        xd, yd = self._dev(x, self.n), self._dev(y, self.m)
        r = Residuals()
        check(self.lib.pdlp_kkt(self.h, xd.data_ptr(), yd.data_ptr(), float(omega), C.byref(r)))
        return r

    def primal_objective(self) -> float:
        v = C.c_double()
        check(self.lib.pdlp_primal_objective(self.h, C.byref(v)))
        return v.value

    # -- setup-time kernels -------------------------------------------------
    def power_iteration(self, b0, iters: int) -> float:
        b = self._dev(b0, self.n)
        sigma = C.c_double()
        check(self.lib.pdlp_power_iteration(self.h, b.data_ptr(), int(iters), C.byref(sigma)))
        return sigma.value

    def spmv(self, v, transpose: bool = False) -> torch.Tensor:
        vin = self._dev(v, self.m if transpose else self.n)
        out = self._new(self.n if transpose else self.m)
        check(self.lib.pdlp_spmv(self.h, int(bool(transpose)), vin.data_ptr(), out.data_ptr()))
        return out

    def ruiz(self, max_iter: int = 20, eps: float = 1e-6) -> int:
        rounds = C.c_int()
        check(self.lib.pdlp_ruiz(self.h, int(max_iter), float(eps), C.byref(rounds)))
        return rounds.value

    def pock_chambolle(self):
        check(self.lib.pdlp_pock_chambolle(self.h))

    def scaling(self) -> Tuple[torch.Tensor, torch.Tensor]:
        dr, dc = self._new(self.m), self._new(self.n)
        check(self.lib.pdlp_get_scaling(self.h, dr.data_ptr(), dc.data_ptr()))
        return dr, dc

    def clear_scaling(self):
        """Forget a scaling declared with set_scaling: terminate on the problem as held."""
        check(self.lib.pdlp_set_scaling(self.h, None, None))
        torch.cuda.synchronize(self.device)

    def set_scaling(self, D_row, D_col):
        dr, dc = self._dev(D_row, self.m), self._dev(D_col, self.n)
        check(self.lib.pdlp_set_scaling(self.h, dr.data_ptr(), dc.data_ptr()))
        torch.cuda.synchronize(self.device)

    def problem(self):
        c, q, l, u = self._new(self.n), self._new(self.m), self._new(self.n), self._new(self.n)
        check(self.lib.pdlp_get_problem(self.h, c.data_ptr(), q.data_ptr(), l.data_ptr(), u.data_ptr()))
        return c, q, l, u

    def matrix(self) -> torch.Tensor:
        """Current (possibly scaled) K as a torch sparse CSR tensor on the device."""
        vals = self._new(self.nnz)
        check(self.lib.pdlp_get_matrix_values(self.h, vals.data_ptr()))
        if self._csr is None:
            crow = torch.empty(self.m + 1, dtype=torch.int32, device=self.device)
            col = torch.empty(self.nnz, dtype=torch.int32, device=self.device)
            check(self.lib.pdlp_get_matrix_structure(self.h, crow.data_ptr(), col.data_ptr()))
            self._csr = (crow, col)
        return torch.sparse_csr_tensor(self._csr[0].to(torch.int64), self._csr[1].to(torch.int64), vals,
                                       size=(self.m, self.n))

    # -- whole solve ----------------------------------------------------------
    def solve(self, max_kkt=100_000, restart_period=40, time_limit=3600.0, time_used=0.0, use_graph=True,
              max_trace=4096) -> SolveOutput:
        opts = SolveOptions(int(max_kkt), int(restart_period), loop_mode(use_graph), float(time_limit),
                            float(time_used), int(max_trace), 0)
        res = SolveResult()
        trace = (Check * max(1, max_trace))()
        check(self.lib.pdlp_solve(self.h, C.byref(opts), C.byref(res), trace))
        x, y = self.get_iterate()
        return SolveOutput(x, y, res.prim_obj, int(res.k), int(res.n), int(res.j), STATUS_STRINGS[res.status],
                           res.seconds, res.device_ms, [Check.from_buffer_copy(trace[i]) for i in range(res.n_trace)])
