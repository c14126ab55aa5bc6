"""Row-partitioned (multi-GPU) restarted PDHG: SURVEY.md section 8e.

One process per GPU.  Rank g holds a contiguous block of rows K_g of K (and K_g^T),
the matching slices of y, q and the dual sums; x, c, l, u are replicated.  K_g x is
local; K^T y = sum_g K_g^T y_g is the single exchange of an iteration: one
``all_reduce`` of n + 4 doubles (the 4 extra slots carry the partial ||dy||^2 and
dy^T K dx of the previous iteration, so the adaptive step rule needs no second
collective).  With ``exchange="p2p"`` (GPUs of one node) that allreduce is not a host-issued collective at
all: the K_g^T y_g kernel, a segment reduce/push kernel and the step-rule kernel talk through the peers' memory
over NVLink (``pdlp_shard_iterate_p2p``, include/pdlp_b200.h).  The reference has no distributed code; the semantics are those of its
single-device loop (primal_dual_hybrid_gradient.py:54-177), which the sharded run
reproduces up to the summation order of the allreduce.

The collectives are issued by the host through ``torch.distributed`` (NCCL over
NVLink on GPUs, gloo in the CPU tests); the device work between them goes through
the ``pdlp_shard_*`` entry points of the C ABI.  ``prims`` abstracts those entry
points so the orchestration below can be exercised on CPU with a test double
(tests/fake_shard_prims.py) -- the product path always uses :class:`GpuShardPrims`.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import time
from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from ._lib import Check, check


def partition_rows(indptr: np.ndarray, parts: int, row_cost: float = 6.0) -> np.ndarray:
    """Contiguous row blocks balanced by nnz + row_cost * rows.  Returns parts + 1 boundaries."""
    m = len(indptr) - 1
    weight = indptr.astype(np.float64) + row_cost * np.arange(m + 1)
    targets = weight[-1] * np.arange(1, parts) / parts
    cuts = np.searchsorted(weight, targets, side="left")
    return np.concatenate([[0], cuts, [m]]).astype(np.int64)


def ruiz_precondition_sharded(K_local, c, q_local, l, u, group=None, max_iter=20, eps=1e-6, device=None):
    """Ruiz equilibration (PDLP/enhancements.py:4-71) of a ROW-PARTITIONED problem: this rank holds the rows K_local
    (scipy CSR, canonical) with q_local; c, l, u are replicated.  Per round, as in the reference: rows divided by
    sqrt(row inf-norm) -- local --, then columns by sqrt(col inf-norm) of the row-scaled matrix -- the per-rank column
    maxima meet in ONE all_reduce(MAX) of n doubles --, norms below eps count as 1; the loop ends when
    max |1 - row_norm| < eps over ALL ranks (the reference tests the row norms twice and never the column norms,
    :60-61; the rank-local maxima ride in the last slot of the same all_reduce, so the decision is collective and costs
    no collective of its own).  max is associative and commutative, sqrt and division are correctly rounded: every rank
    gets the bits a single process gets on the whole matrix.

    Runs on ``device`` (default: the GPU under NCCL, the CPU under gloo / without a process group).  Returns
    ``(K_s_local, c_s, q_s_local, l_s, u_s, D_row_local, D_col, rounds)`` -- scipy CSR + numpy arrays, ready for
    :func:`build_rank_engine` (``scaling=(D_row_local, D_col)``); the minimizer of the scaled problem times D_col is the
    minimizer of the original one (primal_dual_hybrid_gradient.py:161)."""
    import scipy.sparse as sp
    K_local = sp.csr_matrix(K_local)
    m_loc, n = K_local.shape
    use_dist = dist.is_initialized() and dist.get_world_size(group) > 1
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if (use_dist and dist.get_backend(group) == "nccl") \
            else torch.device("cpu")
    f64 = dict(dtype=torch.float64, device=device)
    vals = torch.as_tensor(np.asarray(K_local.data, dtype=np.float64)).to(device).clone()
    cols = torch.as_tensor(K_local.indices.astype(np.int64)).to(device)
    lengths = torch.as_tensor(np.diff(K_local.indptr).astype(np.int64)).to(device)
    D_row = torch.ones(m_loc, **f64)
    D_col = torch.ones(n, **f64)
    rounds = 0
    for _ in range(int(max_iter)):
        rounds += 1
        rn = torch.segment_reduce(vals.abs(), "max", lengths=lengths, unsafe=True) if vals.numel() else torch.zeros(m_loc, **f64)
        rn = torch.where(lengths > 0, rn, torch.zeros_like(rn)).sqrt()           # an empty row has norm 0 -> 1 below
        rn = torch.where(rn < eps, torch.ones_like(rn), rn)
        D_row /= rn
        vals /= torch.repeat_interleave(rn, lengths)
        buf = torch.zeros(n + 1, **f64)                                          # [column maxima | max |1 - row norm|]
        buf[:n].scatter_reduce_(0, cols, vals.abs(), "amax", include_self=True)
        buf[n] = torch.max(torch.abs(1.0 - rn)) if m_loc else 0.0
        if use_dist:
            dist.all_reduce(buf, op=dist.ReduceOp.MAX, group=group)
        cn = buf[:n].sqrt()
        cn = torch.where(cn < eps, torch.ones_like(cn), cn)
        D_col /= cn
        vals /= cn[cols]
        if float(buf[n]) < eps:
            break
    to_np = lambda t: t.detach().cpu().numpy()
    d_row, d_col = to_np(D_row), to_np(D_col)
    K_s = sp.csr_matrix((to_np(vals), K_local.indices.copy(), K_local.indptr.copy()), shape=K_local.shape)
    as_np = lambda v: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v, dtype=np.float64)).reshape(-1)
    c, q_local, l, u = as_np(c), as_np(q_local), as_np(l), as_np(u)
    return K_s, c * d_col, q_local * d_row, l / d_col, u / d_col, d_row, d_col, rounds


class GpuShardPrims:
    """The C-ABI calls of one rank (a thin veneer over :class:`Engine`)."""

    def __init__(self, engine):
        self.e = engine
        self.lib = engine.lib
        self.h = engine.h
        self.device = engine.device
        self.n, self.m_local = engine.n, engine.m

    def buffer(self, length):
        return torch.zeros(length, dtype=torch.float64, device=self.device)

    def enable(self, q_norm, c_norm):
        check(self.lib.pdlp_shard_enable(self.h, float(q_norm), float(c_norm)))

    def local_q_sumsq(self):
        st = self.e.state()
        return st.q_norm ** 2, st.c_norm

    def spmv(self, v, transpose):
        return self.e.spmv(v, transpose).to(torch.float64)

    def kty(self, buf):
        check(self.lib.pdlp_shard_kty(self.h, buf.data_ptr()))

    def step(self, buf, save_previous):
        check(self.lib.pdlp_shard_step(self.h, buf.data_ptr(), int(bool(save_previous))))

    def rule_pack(self, buf4):
        check(self.lib.pdlp_shard_rule_pack(self.h, buf4.data_ptr()))

    def rule_apply(self, buf4):
        check(self.lib.pdlp_shard_rule_apply(self.h, buf4.data_ptr()))

    def check_begin(self, buf3, have_prev):
        check(self.lib.pdlp_shard_check_begin(self.h, buf3.data_ptr(), int(bool(have_prev))))

    def check_mid(self, buf3, buf9):
        check(self.lib.pdlp_shard_check_mid(self.h, buf3.data_ptr(), buf9.data_ptr()))

    def check_decide(self, buf9, force, buf1):
        check(self.lib.pdlp_shard_check_decide(self.h, buf9.data_ptr(), int(bool(force)), buf1.data_ptr()))

    def check_end(self, buf1):
        chk = Check()
        check(self.lib.pdlp_shard_check_end(self.h, buf1.data_ptr(), C.byref(chk)))
        return chk

    def set_iterate(self, x, y_local):
        self.e.set_iterate(x, y_local)

    def get_iterate(self):
        return self.e.get_iterate()

    def set_step(self, eta, omega, theta, adaptive):
        self.e.set_step(eta, omega, theta, adaptive)

    def set_options(self, tol, primal_update):
        self.e.set_options(tol, primal_update, False, 1e-4)

    def set_counters(self, k, t):
        self.e.set_counters(k, t)

    def state(self):
        return self.e.state()

    def sync(self):
        torch.cuda.synchronize(self.device)

    IPC_HANDLE_BYTES = 64

    def exchange_connect(self, group=None):
        """Allocate this rank's exchange block, all-gather the CUDA IPC handles and map the peers' blocks.  A failure
        on any rank (e.g. CUDA IPC not permitted) is agreed on collectively and raised on EVERY rank, so the
        callers' collectives stay matched."""
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0

        def agree(err):
            if world > 1:
                bad = torch.tensor([0.0 if err is None else 1.0], device=self.device)
                dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=group)
                if float(bad.item()) > 0:
                    raise RuntimeError(f"peer-memory exchange unavailable on rank {rank}: {err}" if err is not None
                                       else "peer-memory exchange unavailable on another rank")
            elif err is not None:
                raise err

        mine = (C.c_ubyte * self.IPC_HANDLE_BYTES)()
        err = None
        try:
            check(self.lib.pdlp_shard_exchange_create(self.h, world, rank, C.cast(mine, C.c_void_p)))
        except Exception as exc:                       # noqa: BLE001 -- reported to every rank below
            err = exc
        agree(err)
        t = torch.tensor(list(mine), dtype=torch.uint8, device=self.device)
        if world > 1:
            every = torch.empty(world * self.IPC_HANDLE_BYTES, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(every, t, group=group)
        else:
            every = t
        raw = bytes(every.cpu().tolist())
        buf = (C.c_ubyte * len(raw)).from_buffer_copy(raw)
        try:
            check(self.lib.pdlp_shard_exchange_connect(self.h, C.cast(buf, C.c_void_p)))
        except Exception as exc:                       # noqa: BLE001
            err = exc
        agree(err)                                     # also the barrier: every rank has mapped every block

    def exchange_disconnect(self):
        check(self.lib.pdlp_shard_exchange_disconnect(self.h))

    def allreduce_p2p(self, t):
        check(self.lib.pdlp_shard_allreduce_p2p(self.h, t.data_ptr(), t.numel()))

    def iterate_p2p(self, n_iters, save_previous):
        check(self.lib.pdlp_shard_iterate_p2p(self.h, int(n_iters), int(bool(save_previous))))

    def capture_block(self, fn):
        """Record fn() -- kernel launches of this engine interleaved with NCCL collectives -- into one CUDA
        graph; returns a callable that replays it on the current stream."""
        torch.cuda.synchronize(self.device)
        home = self.e.stream
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.e.set_stream(torch.cuda.current_stream(self.device))
            fn()
        self.e.set_stream(home)
        return graph.replay


class ShardedSolver:
    """Drives one rank of the row-partitioned solve.  Exposes the three calls the restart loop needs
    (``iterate``, ``restart_check``, ``state``), so ``primal_dual_hybrid_gradient.run_restart_loop`` is shared
    with the single-GPU path."""

    def __init__(self, prims, group=None, exchange="nccl"):
        """exchange: "nccl" = one host-issued allreduce per iteration; "p2p" = the peer-memory kernels (GPUs of one
        node; needs prims.exchange_connect)."""
        if exchange not in ("nccl", "p2p"):
            raise ValueError(f"unknown exchange {exchange!r}")
        self.p = prims
        self.group = group
        self.exchange = exchange
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        n = prims.n
        self.buf = prims.buffer(n + 4)
        self.buf3 = prims.buffer(3 * n)
        self.buf9 = prims.buffer(9)
        self.buf4 = prims.buffer(4)
        self.buf1 = prims.buffer(2)                                 # [||y_restart - y_last||^2 partial, time-limit flag]
        self.timeout_probe = None                                  # set by the restart loop: () -> bool, this rank's clock
        self._agreed_timeout = None
        self.collectives = 0
        self._graphs = {}
        self._have_prev = False
        self._p2p_ready = False
        q2, c_norm = prims.local_q_sumsq()
        t = torch.tensor([q2], dtype=torch.float64, device=self.buf.device)
        self._allreduce(t)
        self.q_norm, self.c_norm = float(t.item()) ** 0.5, c_norm
        prims.enable(self.q_norm, self.c_norm)
        if exchange == "p2p":
            t_x = time.time()
            prims.exchange_connect(group)
            self._p2p_ready = True
            if os.environ.get("PDLP_TRACE_SETUP"):
                print(f"[pdlp host] exchange_connect      {1e3 * (time.time() - t_x):8.2f} ms", file=sys.stderr, flush=True)

    def _allreduce(self, t):
        if self._p2p_ready and t.dtype == torch.float64 and t.is_contiguous():
            self.p.allreduce_p2p(t)                                # also with one rank: same kernels, same bits
        elif self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        self.collectives += 1

    # -- setup: power iteration on K^T K (helpers.py:41-51), start vector replicated ----------------
    def power_iteration(self, b0, iters: int) -> float:
        b = torch.as_tensor(b0, dtype=torch.float64).reshape(-1).to(self.buf.device)
        for _ in range(iters):
            v = self.p.spmv(self.p.spmv(b, False), True)        # K_g^T (K_g b): partial of K^T K b
            self._allreduce(v)
            b = v / torch.linalg.norm(v)
        w = self.p.spmv(b, False)
        s = torch.sum(w * w).reshape(1)
        self._allreduce(s)
        return float(s.item()) ** 0.5

    def initial_step(self, power_iters=100, b0=None):
        if b0 is None:
            b0 = torch.randn(self.p.n, 1, dtype=torch.float64)     # same draw on every rank (same seed)
        sigma = self.power_iteration(b0, power_iters)
        eta = (1.0 / sigma) * 0.9
        omega = self.c_norm / self.q_norm if (self.q_norm > 1e-6 and self.c_norm > 1e-6) else 1.0
        return eta, omega, sigma

    # -- the three calls of the restart loop --------------------------------------------------------
    def _iterate_eager(self, n_iters, save_previous):
        if save_previous:
            self._have_prev = True
        if self.exchange == "p2p":                                 # the exchange happens inside the kernels
            self.p.iterate_p2p(n_iters, save_previous)
            self.collectives += n_iters
            return
        for i in range(n_iters):
            self.p.kty(self.buf)
            self._allreduce(self.buf)                              # THE exchange: n + 4 doubles
            self.p.step(self.buf, save_previous and i == n_iters - 1)

    def iterate(self, n_iters: int, save_previous: bool = False, use_graph: bool = False):
        """n_iters iterations.  With use_graph the whole block -- 5 kernels and one NCCL allreduce per iteration --
        is captured once into a CUDA graph and replayed, so the host issues one launch per restart block."""
        if not (use_graph and n_iters > 1 and hasattr(self.p, "capture_block")):
            return self._iterate_eager(n_iters, save_previous)
        key = (n_iters, bool(save_previous))
        state = self._graphs.get(key)
        if state is None:                                          # first block: eager (sets kernel attributes,
            self._graphs[key] = "warm"                             # opens the NCCL channels), capture on the next one
            return self._iterate_eager(n_iters, save_previous)
        if state == "warm":
            before = self.collectives
            t_cap = time.time()
            state = self.p.capture_block(lambda: self._iterate_eager(n_iters, save_previous))
            if os.environ.get("PDLP_TRACE_SETUP"):
                print(f"[pdlp host] sharded graph capture {1e3 * (time.time() - t_cap):8.2f} ms", file=sys.stderr, flush=True)
            self.collectives = before
            self._graphs[key] = state
        state()
        self.collectives += n_iters
        if save_previous:
            self._have_prev = True

    def _check_enqueue(self, force: bool, have_prev: bool):
        """Everything of the restart test up to its last collective: device work and collectives only, no host
        round trip -- capturable."""
        self.p.rule_pack(self.buf4)                                # step rule of the block's last iteration
        self._allreduce(self.buf4)
        self.p.rule_apply(self.buf4)
        self.p.check_begin(self.buf3, have_prev)
        self._allreduce(self.buf3)                                 # K^T y of current / average / previous iterate
        self.p.check_mid(self.buf3, self.buf9)
        self._allreduce(self.buf9)                                 # row-side KKT sums
        self.p.check_decide(self.buf9, force, self.buf1)           # buf1[0] = this rank's ||y_restart - y_last||^2
        self._allreduce(self.buf1)                                 # ... and buf1[1] = time-limit flags (see below)

    def restart_check(self, force: bool = False):
        """The restart test (primal_dual_hybrid_gradient.py:115-177): four collectives between the device stages.
        From the second use on the whole sequence is replayed as ONE CUDA graph per (force, have_prev) pair, like the
        iteration blocks, so its ~20 launches do not wait for the host.  The last collective also carries every
        rank's "my clock says the time limit is reached" flag (``timeout_probe``, set by the restart loop), so the
        collective time-limit decision of the NEXT block costs neither a collective nor a host round trip."""
        have_prev = bool(self._have_prev and not force)
        self._have_prev = False
        flag = 1.0 if (self.timeout_probe is not None and self.timeout_probe()) else 0.0
        self.buf1[1] = flag                                        # stream-ordered before the graph / the eager launches
        key = ("check", bool(force), have_prev)
        state = self._graphs.get(key)
        if hasattr(self.p, "capture_block") and state is not None:
            if state == "warm":
                before = self.collectives
                state = self.p.capture_block(lambda: self._check_enqueue(force, have_prev))
                self.collectives = before
                self._graphs[key] = state
            state()
            self.collectives += 4
        else:
            if hasattr(self.p, "capture_block"):
                self._graphs[key] = "warm"                         # first use runs eagerly (kernel attributes, channels)
            self._check_enqueue(force, have_prev)
        chk = self.p.check_end(self.buf1)                          # the one host round trip of the test
        self._agreed_timeout = bool(float(self.buf1[1].item()) > 0.5) if self.timeout_probe is not None else None
        return chk

    def state(self):
        return self.p.state()

    def agree_timeout(self, timed_out: bool) -> bool:
        """The wall-clock test of the restart loop (primal_dual_hybrid_gradient.py:68-74) is the only loop decision
        taken from rank-local data: the ranks' clocks differ, and a rank that stopped alone would leave the others
        waiting in the next exchange.  So the limit is reached when ANY rank has reached it (a sum over ranks of
        the 0/1 flags, through the same exchange as every other collective), and all ranks stop in the same
        block."""
        if self._agreed_timeout is not None:                       # agreed on in the last restart test's collective
            agreed, self._agreed_timeout = self._agreed_timeout, None
            return agreed
        if self.world <= 1 and not self._p2p_ready:
            return bool(timed_out)
        t = torch.tensor([1.0 if timed_out else 0.0], dtype=torch.float64, device=self.buf.device)
        self._allreduce(t)
        return bool(float(t.item()) > 0.5)

    def close(self):
        """Drop the captured graphs (they hold NCCL work) and the engine; call before destroy_process_group."""
        self._graphs.clear()
        gpu = hasattr(self.p, "e")
        if gpu:
            if not self.p.e.h:
                return
            torch.cuda.synchronize(self.p.device)
        if self._p2p_ready:
            # every rank unmaps its peers before any rank frees its own exchange block
            self.p.exchange_disconnect()
            self._p2p_ready = False
            if self.world > 1 and dist.is_initialized():
                dist.barrier(group=self.group)
        if gpu:
            self.p.e.close()

    def primal_objective(self):
        raise NotImplementedError("infeasibility detection is not available in the sharded engine")

    # -- whole solve ----------------------------------------------------------------------------------
    def solve(self, max_kkt=100_000, tol=1e-4, restart_period=40, primal_update=False, adaptive=False,
              time_limit=3600.0, verbose=False, power_iters=100, b0=None, x_init=None, y_init_local=None,
              use_graph=True):
        from .primal_dual_hybrid_gradient import run_restart_loop
        t0 = time.time()
        eta, omega, sigma = self.initial_step(power_iters, b0)
        if os.environ.get("PDLP_TRACE_SETUP"):
            self.p.sync()
            print(f"[pdlp host] sharded initial_step {1e3 * (time.time() - t0):8.2f} ms", file=sys.stderr, flush=True)
        self.p.set_iterate(x_init, y_init_local)
        self.p.set_counters(0, 0)
        self.p.set_step(eta, omega, 1.0, adaptive)
        self.p.set_options(tol, primal_update)
        self.p.enable(self.q_norm, self.c_norm)
        rank0 = (not dist.is_initialized()) or dist.get_rank(self.group) == 0
        prim_obj, k, n, j, status, trace = run_restart_loop(self, max_kkt, restart_period, time_limit, 0.0,
                                                            verbose and rank0, False, use_graph=use_graph, t_start=t0)
        x, y_local = self.p.get_iterate()
        if os.environ.get("PDLP_TRACE_SETUP"):
            self.p.sync()
            print(f"[pdlp host] sharded solve total  {1e3 * (time.time() - t0):8.2f} ms (k={k})", file=sys.stderr, flush=True)
        return x, y_local, prim_obj, k, n, j, status, time.time() - t0, sigma


def build_rank_engine(K_local, m_ineq_local: int, c, q_local, l, u, device=None, exchange="nccl", scaling=None):
    """Engine on this rank's rows + the sharded driver.  ``scaling = (D_row_local, D_col)``: the problem passed in is
    the Ruiz-scaled one (:func:`ruiz_precondition_sharded`); the restart tests then terminate on the residuals of the
    ORIGINAL problem, as the reference does with ``precondition=True`` (primal_dual_hybrid_gradient.py:153-177)."""
    from .engine import Engine
    eng = Engine(K_local, m_ineq_local, c, q_local, l, u, device=device, dtype=torch.float64)
    if scaling is not None:
        eng.set_scaling(scaling[0], scaling[1])
    return ShardedSolver(GpuShardPrims(eng), exchange=exchange)
