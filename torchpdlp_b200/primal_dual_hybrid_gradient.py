"""Drop-in for torchPDLP's PDLP/primal_dual_hybrid_gradient.py.

``pdlp_algorithm`` keeps the reference signature and 7-tuple return
(primal_dual_hybrid_gradient.py:7, :181) but runs the restart loop on the B200
engine: two fused kernels per PDHG iteration, scalars resident on the device, one
host read-back per restart check.
"""
from __future__ import annotations

import os
import sys
import threading
import time

import torch

from . import _cache
from .engine import RESTART_KINDS, STATUS_STRINGS, Engine

_BOUND = "Unsolved (KKT passes limit exceeded)"


def initial_step(engine: Engine, power_iters: int = 100, b0=None):
    """eta = 0.9 / sigma_max(K) and omega = ||c|| / ||q|| (reference :19-23).  The
    power iteration starts from torch.randn(n, 1) drawn from the global CPU
    generator in the engine dtype, as helpers.py:47 does on a CPU device."""
    if b0 is None:
        b0 = torch.randn(engine.n, 1, dtype=engine.dtype)
    sigma = engine.power_iteration(b0, power_iters)
    st = engine.state()
    eta = (1.0 / sigma) * 0.9          # torch evaluates `0.9 / tensor` as reciprocal * 0.9
    omega = st.c_norm / st.q_norm if (st.q_norm > 1e-6 and st.c_norm > 1e-6) else 1.0
    return eta, omega, sigma


def _print_restart(chk, verbose):
    if verbose:
        kind = RESTART_KINDS[chk.kind]
        if kind in ("Sufficient", "Necessary", "Artificial"):
            which = "Average iterate." if chk.use_average else "Current iterate."
            print(f"{kind} restart at iteration {chk.t} using the", which)


def _print_point(k, chk, q_norm, c_norm, scaled, verbose):
    if verbose:
        p = chk.point
        pr = p.pr_unscaled if scaled else p.pr
        dr = p.dr_unscaled if scaled else p.dr
        print(f"[{k}] Primal Obj: {p.prim_obj:.4f}, Adjusted Dual Obj: {p.adj_dual:.4f}, "
              f"Gap: {p.gap / (1 + abs(p.prim_obj) + abs(p.adj_dual)):.2e}, Prim Res: {pr / (1 + q_norm):.2e}, "
              f"Dual Res: {dr / (1 + c_norm):.2e}")
        print("")


def run_restart_loop(engine: Engine, max_kkt, restart_period, time_limit, time_used, verbose, scaled, use_graph=True,
                     t_start=None, infeas=False):
    """The outer / inner loops of :54-177 on the host, one engine call per block of
    iterations and one per restart test; counters follow the reference (k per
    iteration, j: +1 per step, +3 per restart test, +2 per restart)."""
    t_start = time.time() if t_start is None else t_start
    st = engine.state()
    q_norm, c_norm = st.q_norm, st.c_norm
    k, n, j = int(st.k), 0, 0
    status = _BOUND
    prim_obj = float("nan")
    trace = []
    if hasattr(engine, "timeout_probe"):       # row-partitioned: the flag rides in the restart test's last collective
        engine.timeout_probe = lambda: time.time() - t_start + time_used >= time_limit
    while j < max_kkt:
        t = 0
        restarted = False
        timed_out = False
        while j < max_kkt:
            timed_out = time.time() - t_start + time_used >= time_limit   # :68-74 (tested per block here)
            if hasattr(engine, "agree_timeout"):                           # row-partitioned: one decision for all ranks
                timed_out = engine.agree_timeout(timed_out)
            if timed_out:
                status = "Unsolved (Time limit exceeded)"
                if verbose:
                    print("Time limit exceeded")
                break
            want = restart_period - (t % restart_period)
            iters, jj, kk = 0, j, k
            while iters < want and jj < max_kkt:                           # guard before every iteration (:67)
                kk += 1
                jj += 1 + (1 if (infeas and kk > 1) else 0)
                iters += 1
            at_check = (t + iters) % restart_period == 0
            engine.iterate(iters, save_previous=at_check, use_graph=use_graph)
            if infeas:                                                     # :89-101
                s2 = engine.state()
                if s2.infeas:
                    while k < s2.infeas_k:
                        k += 1
                        j += 1 + (1 if k > 1 else 0)
                    status = STATUS_STRINGS[3 if s2.infeas == 2 else 4]
                    if verbose:
                        print(f"[PDLP] {status} detected at iteration {k}")
                    return engine.primal_objective(), k, n, j, status, trace
            j, k, t = jj, kk, t + iters
            if at_check:
                chk = engine.restart_check(False)
                j += 3                                                     # :128
                if chk.kind != 0:
                    _print_restart(chk, verbose)
                    restarted = True
                    break
        if timed_out:
            break
        if not restarted:
            chk = engine.restart_check(True)                               # loop left through the guard (:148)
        n += 1
        j += 2                                                             # :154, :165
        prim_obj = chk.point.prim_obj
        trace.append(chk)
        _print_point(k, chk, q_norm, c_norm, scaled, verbose)
        if chk.solved:                                                     # :173-177
            status = "Solved"
            if verbose:
                print(f"Converged at iteration {k} restart loop {n}")
            break
    return prim_obj, k, n, j, status, trace


def pdlp_algorithm(K, m_ineq, c, q, l, u, device, max_kkt=100_000, tol=1e-4, verbose=True, restart_period=40,
                   precondition=False, primal_update=False, adaptive=False, data_precond=None,
                   infeasibility_detect=False, infeas_tol=1e-4, time_limit=3600, time_used=0, x_init=None,
                   y_init=None, power_iters=100, b0=None, use_graph=True, return_info=False):
    """Same call and return as the reference (:7, :181):
    ``(x, prim_obj, k, n, j, status, total_time)`` with ``x`` an (n, 1) tensor on the
    GPU (in the scaled space when preconditioned, like the reference).  Extra keyword
    arguments (power_iters, b0, use_graph, return_info) have reference-compatible
    defaults."""
    t_start = time.time()
    trace_t = [] if os.environ.get("PDLP_TRACE_SETUP") else None

    def lap(what):
        if trace_t is not None:
            torch.cuda.synchronize()
            trace_t.append((what, time.time()))

    lap("entry")
    # the power iteration's start vector is drawn from the global CPU generator (helpers.py:47 on a CPU device); 200k
    # normal deviates take milliseconds, so they are drawn on a helper thread while the library builds the engine
    # (both release the GIL; this is the only draw, so the generator is consumed exactly as in the sequential order)
    draw = None
    if b0 is None:
        n_cols = int(K.shape[1]) if hasattr(K, "shape") else int(K[3][1])
        b0_dtype = K.dtype if isinstance(K, torch.Tensor) and K.dtype in (torch.float32, torch.float64) else torch.float64
        box = {}

        def _draw():
            try:
                box["b0"] = torch.randn(n_cols, 1, dtype=b0_dtype)
            except BaseException as err:      # re-raised on the caller's thread
                box["err"] = err

        draw = threading.Thread(target=_draw, name="pdlp-b0")
        draw.start()
    try:
        engine = _cache.engine_for(K, m_ineq, c, q, l, u, device)
    finally:
        if draw is not None:
            draw.join()
    if draw is not None:
        if "err" in box:
            raise box["err"]
        b0 = box["b0"].to(engine.dtype)
    lap("engine_for")
    scaled = bool(precondition)
    if precondition and data_precond is not None:
        D_col, D_row = data_precond[0], data_precond[1]
        engine.set_scaling(D_row, D_col)
    else:
        engine.clear_scaling()             # a cached engine may carry the scaling of an earlier call
    eta, omega, sigma = initial_step(engine, power_iters, b0)
    lap("initial_step")
    if x_init is not None and y_init is not None:                          # :31-33
        engine.set_iterate(x_init, y_init)
    else:
        engine.set_iterate(None, None)
    engine.set_counters(0, 0)
    engine.set_step(eta, omega, 1.0, adaptive)
    engine.set_options(tol, primal_update, infeasibility_detect, infeas_tol)
    device_ms = None
    if verbose:
        prim_obj, k, n, j, status, trace = run_restart_loop(engine, max_kkt, restart_period, time_limit, time_used,
                                                            verbose, scaled, use_graph, t_start,
                                                            infeas=bool(infeasibility_detect))
    else:
        out = engine.solve(max_kkt, restart_period, time_limit, time_used + (time.time() - t_start), use_graph)
        prim_obj, k, n, j, status, trace = out.prim_obj, out.k, out.n, out.j, out.status, out.trace
        device_ms = out.device_ms
    lap("solve")
    x, y = engine.get_iterate()
    torch.cuda.synchronize(engine.device)
    lap("get_iterate")
    total = time.time() - t_start + time_used
    if trace_t is not None:
        prev = t_start
        for what, t in trace_t:
            print(f"[pdlp host] {what:14s} {1e3 * (t - prev):8.2f} ms", file=sys.stderr)
            prev = t
    if return_info:
        info = {"y": y.reshape(-1, 1), "trace": trace, "sigma_max": sigma, "engine": engine, "device_ms": device_ms,
                "bytes_per_iteration": engine.bytes_per_iteration}
        return x.reshape(-1, 1), prim_obj, k, n, j, status, total, info
    return x.reshape(-1, 1), prim_obj, k, n, j, status, total


def pdlp_solver(mps_file_path, tol=1e-4, restart_period=40, verbose=True, max_iter=1000000, precondition=True,
                adaptive_step=True, primal_update=True):
    """The one-call entry on an MPS file of the reference's older line
    (Packages/PDLP_without_presolve_infeasibility.py:747-789): same arguments, same defaults, same 5-tuple
    ``(minimizer, objective_value, total_iterations, total_restarts, kkt_passes)``.  ``max_iter`` bounds the KKT
    passes j, as there (its loops test ``j < max_iter``).  As in that file the minimizer is returned in the
    ORIGINAL variables (its ``x / D_col`` at :745 undoes the scaling; here ``D_col * x_scaled``,
    PDLP/primal_dual_hybrid_gradient.py:161), an (n, 1) tensor on the GPU.  Preconditioning is Ruiz
    (enhancements.py:4-71).  A file that cannot be read raises (the reference prints and calls exit(1))."""
    from .enhancements import ruiz_precondition
    from .util import mps_to_standard_form

    if not torch.cuda.is_available():
        raise RuntimeError("torchpdlp_b200 runs on a CUDA device (B200) only: no CPU fallback")
    device = torch.device("cuda")
    if verbose:
        print(f"PyTorch is using ROCm/CUDA device: {torch.cuda.get_device_name(0)}")
    c, K, q, m_ineq, l, u = mps_to_standard_form(mps_file_path, device=device, verbose=verbose)
    time_used, data_precond, D_col = 0.0, None, None
    if precondition:
        K, c, q, l, u, data_precond, time_used = ruiz_precondition(c, K, q, l, u, device=device)
        D_col = data_precond[0]
    x, objective_value, k, n, j, status, _ = pdlp_algorithm(
        K, m_ineq, c, q, l, u, device, max_kkt=max_iter, tol=tol, verbose=verbose, restart_period=restart_period,
        precondition=precondition, primal_update=primal_update, adaptive=adaptive_step, data_precond=data_precond,
        time_used=time_used)
    minimizer = x * D_col.to(x.device).reshape(-1, 1) if D_col is not None else x
    if verbose:
        print("Objective Value:", objective_value)
        print("Iterations:", k)
        print("Restarts:", n)
        print("KKT Passes:", j)
        print("\nMinimizer (first 10 variables):")
        print(minimizer[:10].cpu().numpy())
    return minimizer, objective_value, k, n, j
