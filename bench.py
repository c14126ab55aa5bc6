#!/usr/bin/env python
"""bench.py -- PDHG iterations/sec of the B200 engine on BASELINE.json's config 2
(synthetic bounded LP, m=100k, n=200k, 5M nnz, fp64, restarts + adaptive steps).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one restart block of the hot path: 40 PDHG iterations (2 fused kernels
each) followed by the restart test (averages, 3 KKT errors in one pass over K and
one over K^T, criteria, restart).  `value` = iterations/sec with the LP resident in
HBM (CUDA events, max over ranks); `e2e` = iterations/sec of whole solves through
the public API `pdlp_algorithm` starting from HOST tensors (pinned), i.e. including
the H2D copy of the LP, the engine build, the power iteration, every restart test and
the D2H copy of the solution.  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "pdhg_iterations_per_sec"
UNIT = "it/s"
WORKLOAD = "cfg2: synthetic bounded LP m=100000 n=200000 nnz=5000000 (50/row), fp64, restarts + adaptive step + primal weight"
RESTART_PERIOD = 40
# the same `config` on both arms (engine and --impl reference): the reference arm times a bounded sample of this step
CONFIG_N1 = {"workload": WORKLOAD, "step": "40 PDHG iterations + restart test (3 KKT errors)", "parallelism": "single GPU",
             "l2": "working set 140 MB of matrix+vectors per iteration > 126 MB L2; no explicit flush"}
KERNELS_PER_STEP = 2 * RESTART_PERIOD + 6      # 80 iteration kernels + flush, 2 KKT, decide, apply, post


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_lp():
    from torchpdlp_b200 import synthetic
    return synthetic.uniform_rows_lp()


# ----------------------------------------------------------------------------- CPU arms
class OracleStepper:
    """Adaptive PDHG iterations of the oracle port (torch CPU) on the cfg2 LP, state kept between calls."""

    def __init__(self, lp, layout):
        from oracle import pdlp_oracle as orc
        torch.set_default_dtype(torch.float64)
        self.orc = orc
        c, K, q, self.m_ineq, l, u = lp.torch_tensors(layout=layout)
        self.op = orc.Operator(K)
        self.c, self.q, self.l, self.u = (v.reshape(-1) for v in (c, q, l, u))
        self.x, self.y = torch.zeros(lp.n), torch.zeros(lp.m)
        self.eta = 0.9 / 18.2
        self.omega = float(torch.linalg.norm(self.c) / torch.linalg.norm(self.q))
        self.k = 0

    def run(self, iters):
        t0 = time.perf_counter()
        for _ in range(iters):
            self.k += 1
            self.x, self.y, w, self.eta = self.orc.adaptive_one_step_pdhg(
                self.x, self.y, self.c, self.q, self.op, self.l, self.u, self.m_ineq, self.eta, self.omega, 1.0, self.k)
        return time.perf_counter() - t0


def oracle_rate(lp, layout, iters, warm=1):
    st = OracleStepper(lp, layout)
    st.run(warm)
    return iters / st.run(iters)


def run_reference(args):
    """--impl reference: the reference's own CPU path for this workload.  torchPDLP is pure
    Python/torch and /root/reference does not travel to the GPU box, so this times the oracle
    port (oracle/pdlp_oracle.py, validated bit-for-bit against the executed reference on config
    1) with K as a torch sparse COO tensor -- the only sparse layout the reference itself can
    run -- on all host threads.  Each step = a bounded sample of
    adaptive PDHG iterations, sized so the whole run stays within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)       # torchrun exports OMP_NUM_THREADS=1: use the host's cores anyway
    lp = make_lp()
    st = OracleStepper(lp, "coo")
    t_iter = st.run(1)
    t_iter = min(t_iter, st.run(1))
    per_step = int(max(1, min(8, 100.0 / (max(args.steps, 1) * t_iter))))
    for _ in range(min(args.warmup, 2)):
        st.run(1)
    dt = 0.0
    for _ in range(args.steps):
        dt += st.run(per_step)
    value = args.steps * per_step / dt
    csr = oracle_rate(lp, "csr", 20)
    cores = torch.get_num_threads()
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(CONFIG_N1),
        "sample_step": f"{per_step} adaptive PDHG iterations per step (bounded sample of the step in `config`), K as torch sparse "
                       "COO (the reference's sparse layout)",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} x {per_step} adaptive iterations of the oracle port, torch sparse COO, "
                                   f"{cores} threads of {os.cpu_count()} cpus; same port on torch sparse CSR (MKL): {csr:.1f} it/s"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def run_engine(args):
    import torch.distributed as dist
    from torchpdlp_b200.engine import Engine
    from torchpdlp_b200.primal_dual_hybrid_gradient import initial_step, pdlp_algorithm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.set_default_dtype(torch.float64)
    if world > 1:
        return run_sharded(args, world, rank, local, dev)
    lp = make_lp()
    eng = Engine(lp.K, lp.m_ineq, lp.c, lp.q, lp.l, lp.u, device=dev)
    torch.manual_seed(0)
    eta, omega, sigma = initial_step(eng, 100)

    def reset():
        eng.set_iterate(None, None)
        eng.set_counters(0, 0)
        eng.set_step(eta, omega, 1.0, adaptive=True)
        eng.set_options(1e-4, True, False, 1e-4)

    def step():
        eng.iterate(RESTART_PERIOD, save_previous=True, use_graph=True)
        return eng.restart_check(False)

    def barrier():              # single process from here on (N > 1 went to run_sharded above)
        torch.cuda.synchronize(dev)

    reset()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # per-kernel durations, CUDA events on the engine's stream (cannot be placed inside the graph replay)
    iso_prim_ms, iso_dual_ms = eng.time_kernels(200)
    # ... and as they run in the loop: back-to-back launches chained by programmatic dependent launch, every launch
    # stamped with %globaltimer when it leaves griddepcontrol.wait and after its finalize (pdlp_time_kernels_in_loop)
    prim_ms, dual_ms, loop_iter_ms = eng.time_kernels_in_loop(400)
    clocks = sampler.stop()
    iters = RESTART_PERIOD * args.steps
    value = iters / (ms * 1e-3)

    # ---- e2e: whole solves through the public API from pinned host tensors ----
    c, K, q, m_ineq, l, u = lp.torch_tensors(layout="csr")
    host = [t.pin_memory() for t in (K.crow_indices(), K.col_indices(), K.values(), c, q, l, u)]
    h2d = sum(t.numel() * t.element_size() for t in host)
    Kh = torch.sparse_csr_tensor(host[0], host[1], host[2], size=K.shape)
    n_solves = max(1, min(10, args.steps))       # (a shared host adds tens of ms to a solve now and then: 10 solves dilute it)
    from torchpdlp_b200 import _cache
    # one untimed solve first: the first pass through torch's sparse conversions loads their CUDA modules (~0.2 s once)
    _cache.clear()
    torch.manual_seed(0)
    pdlp_algorithm(Kh, m_ineq, host[3], host[4], host[5], host[6], dev, verbose=False, adaptive=True, primal_update=True)
    barrier()
    # (as timeit does: no cyclic garbage collection inside the timed region -- a generation-2 pass over the objects the
    # CPU-side setup left behind costs tens of milliseconds and lands in a random solve)
    import gc
    gc.collect()
    gc.disable()
    t0 = time.perf_counter()
    e2e_iters = 0
    solve_device_ms = []
    solve_wall = []
    for _ in range(n_solves):
        t_solve = time.perf_counter()
        _cache.clear()
        torch.manual_seed(0)
        x, obj, k, n, j, status, secs, info = pdlp_algorithm(Kh, m_ineq, host[3], host[4], host[5], host[6], dev,
                                                             verbose=False, adaptive=True, primal_update=True,
                                                             return_info=True)
        x_host = x.cpu()
        e2e_iters += k
        solve_device_ms.append(info.get("device_ms"))
        solve_wall.append(round(time.perf_counter() - t_solve, 4))
    barrier()
    e2e_s = time.perf_counter() - t0
    gc.enable()
    e2e_value = e2e_iters / e2e_s
    d2h = x_host.numel() * x_host.element_size() + 64

    # ---- config 4 (2M x 4M, ~50M nnz, power-law rows): hot-loop rate on the load-balance stress case ----
    cfg4 = None
    if os.environ.get("PDLP_BENCH_CFG4", "1") != "0" and world == 1:
        eng.close()
        _cache.clear()
        torch.cuda.empty_cache()
        cfg4 = run_cfg4(dev)
    if rank == 0:
        peak, peak_src = measured_peak()
        b_iter = lp.bytes_per_iteration()
        dual_bytes = 12 * lp.nnz + 4 * (lp.m + 1) + 8 * lp.n + 40 * lp.m       # K stream + gather xbar + 5 m-side streams
        primal_bytes = b_iter - dual_bytes                                     # K^T stream + gather y + 8 n-side streams
        # the dominant kernel = the one with the longer launch (K1 on config 2: twice the rows, shorter rows)
        dom = "primal" if prim_ms >= dual_ms else "dual"
        dom_ms, dom_bytes = (prim_ms, primal_bytes) if dom == "primal" else (dual_ms, dual_bytes)
        ach = dom_bytes / (dom_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")) as f:
                traffic = json.load(f).get(f"{dom}_dram_bytes_per_launch")
        except Exception:
            pass
        # bounded CPU sample of the same workload (oracle port, all host threads)
        cores = torch.get_num_threads()
        cpu_coo = oracle_rate(lp, "coo", 8)
        cpu_csr = oracle_rate(lp, "csr", 40)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": dict(CONFIG_N1),
            "sigma_max": sigma, "bytes_per_iteration": b_iter,
            "us_per_iteration": 1e3 * ms / iters,
            "iteration_roofline_frac": b_iter / (ms * 1e-3 / iters) / 1e9 / peak,
            "time_to_tolerance": {"rel_kkt": 1e-4, "iterations": e2e_iters // n_solves, "status": status,
                                  "device_ms": min(v for v in solve_device_ms if v) if any(solve_device_ms) else None,
                                  "wall_s_from_host_tensors": e2e_s / n_solves, "wall_s_each": solve_wall},
            "gpu_launches": KERNELS_PER_STEP * args.steps,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": f"{n_solves} whole solves after one untimed warm-up solve (pdlp_algorithm from pinned host tensors to tolerance 1e-4: "
                            f"{e2e_iters // n_solves} iterations each, status {status}), {e2e_s / n_solves:.3f} s per solve"},
            "roofline": {"bound": "hbm",
                         "kernel": ("stream_kernel<PrimalOp> (K^T y + primal update)" if dom == "primal" else
                                    "stream_kernel<DualOp> (K xbar + dual update + step rule)"),
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes, "launch_ms": dom_ms,
                         "launch_ms_is": "average span of a launch inside the loop (400 iterations chained by programmatic "
                                         "dependent launch as in the block graph): %globaltimer when its first CTA leaves "
                                         "griddepcontrol.wait -> after its last CTA's finalize; isolated launches bracketed "
                                         "by CUDA events (cold start, no overlap with the neighbour) are under `isolated_ms`",
                         "loop_ms_per_iteration": loop_iter_ms,
                         "primal": {"launch_ms": prim_ms, "isolated_ms": iso_prim_ms, "algorithmic_bytes": primal_bytes,
                                    "frac": primal_bytes / (prim_ms * 1e-3) / 1e9 / peak},
                         "dual": {"launch_ms": dual_ms, "isolated_ms": iso_dual_ms, "algorithmic_bytes": dual_bytes,
                                  "frac": dual_bytes / (dual_ms * 1e-3) / 1e9 / peak}},
            "cpu_baseline": {"value": cpu_coo, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"8 adaptive PDHG iterations of the oracle port on the same LP, torch sparse COO "
                                       f"(the reference's sparse layout), {cores} threads; torch sparse CSR (MKL): {cpu_csr:.1f} it/s"},
            "cfg4": cfg4,
        }
        print(json.dumps(out), flush=True)


def run_cfg4(dev):
    """BASELINE.json config 4: transportation-style LP 2M x 4M, ~50M nnz, power-law row lengths (2 ... 100 000):
    microseconds per PDHG iteration of the hot loop (CUDA graph of 80 kernels, adaptive step + primal weight) and the
    fraction of the HBM roofline.  The gathered pair vector (64 MB) competes with the matrix stream for the 126 MB L2."""
    from torchpdlp_b200 import synthetic
    from torchpdlp_b200.engine import Engine
    from torchpdlp_b200.primal_dual_hybrid_gradient import initial_step
    t0 = time.perf_counter()
    lp = synthetic.powerlaw_rows_lp()
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    eng = Engine(lp.K, lp.m_ineq, lp.c, lp.q, lp.l, lp.u, device=dev)
    torch.cuda.synchronize(dev)
    t_build = time.perf_counter() - t0
    torch.manual_seed(0)
    eta, omega, sigma = initial_step(eng, 20)
    eng.set_iterate(None, None); eng.set_counters(0, 0); eng.set_step(eta, omega, 1.0, adaptive=True)
    eng.set_options(1e-4, True, False, 1e-4)
    for _ in range(3):
        eng.iterate(RESTART_PERIOD, save_previous=True, use_graph=True)
        eng.restart_check(False)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    blocks = 5
    e0.record()
    for _ in range(blocks):
        eng.iterate(RESTART_PERIOD, save_previous=True, use_graph=True)
    e1.record()
    torch.cuda.synchronize(dev)
    us = 1e3 * e0.elapsed_time(e1) / (blocks * RESTART_PERIOD)
    prim_ms, dual_ms = eng.time_kernels(20)
    peak, _ = measured_peak()
    b_iter = lp.bytes_per_iteration()
    eng.close()
    return {"workload": f"cfg4: synthetic transportation-style LP m={lp.m} n={lp.n} nnz={lp.nnz}, power-law row lengths, fp64, "
                        "adaptive step + primal weight", "us_per_iteration": us, "iterations_per_sec": 1e6 / us,
            "bytes_per_iteration": b_iter, "iteration_roofline_frac": b_iter / (us * 1e-6) / 1e9 / peak,
            "primal_launch_us": 1e3 * prim_ms, "dual_launch_us": 1e3 * dual_ms, "sigma_max_20_rounds": sigma,
            "setup_s": {"generate": t_gen, "engine_build": t_build}}


def run_sharded(args, world, rank, local, dev):
    """N > 1: the row-partitioned engine (torchpdlp_b200.sharded) on a weak-scaling LP: every rank holds one
    config-2 row block (100k rows, 5M nnz) of a (N*100k) x 200k LP; x replicated; one allreduce of n + 4
    doubles per iteration, done by the iteration kernels through peer memory (PDLP_EXCHANGE=nccl: by NCCL).  value = N x iterations/s (row-block iterations per second), which equals
    the N = 1 metric when N = 1."""
    import torch.distributed as dist
    from torchpdlp_b200 import synthetic
    from torchpdlp_b200.sharded import build_rank_engine
    n, m_blk, nnz_row = 200_000, 100_000, 50
    # rank-local rows: inequality rows first (rank order), generated per block with its own seed
    rng = np.random.default_rng(0)
    x_feas = rng.uniform(-10, 10, size=n)
    l = np.maximum(x_feas - rng.uniform(1, 5, size=n), -1e4)
    u = np.minimum(x_feas + rng.uniform(1, 5, size=n), 1e4)
    c = rng.normal(size=n)
    n_ineq_ranks = max(1, int(round(0.8 * world)))          # 80 % of the rows are ">=" rows, as in config 2
    Kb = synthetic._stratified_rows(m_blk, n, nnz_row, np.random.default_rng(1000 + rank))
    if rank < n_ineq_ranks:
        Kb = -Kb
        q = Kb @ x_feas - np.random.default_rng(2000 + rank).uniform(0.1, 5.0, size=m_blk)     # -G x >= -h
        mi = m_blk
    else:
        q = Kb @ x_feas
        mi = 0
    Kb = Kb.tocsr(); Kb.sort_indices()
    # the exchange: peer-memory kernels over NVLink (default) or one host-issued NCCL allreduce per iteration
    exchange = os.environ.get("PDLP_EXCHANGE", "p2p")
    if exchange == "p2p":
        try:
            solver, ok = build_rank_engine(Kb, mi, c, q, l, u, device=dev, exchange="p2p"), 1.0
        except Exception as exc:                        # e.g. CUDA IPC not permitted in this container
            print(f"[rank {rank}] peer-memory exchange unavailable ({exc}); using NCCL", file=sys.stderr, flush=True)
            solver, ok = None, 0.0
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if float(flag.item()) < 1.0:                    # all ranks take the same path
            if solver is not None:
                solver.close()
            exchange = "nccl"
    if exchange != "p2p":
        solver = build_rank_engine(Kb, mi, c, q, l, u, device=dev, exchange="nccl")
    torch.manual_seed(0)
    eta, omega, sigma = solver.initial_step(100)
    p = solver.p

    def reset():
        p.set_iterate(None, None); p.set_counters(0, 0); p.set_step(eta, omega, 1.0, True); p.set_options(1e-4, True)
        p.enable(solver.q_norm, solver.c_norm)

    def step():
        solver.iterate(RESTART_PERIOD, save_previous=True, use_graph=True)
        return solver.restart_check(False)

    def barrier():
        dist.barrier(); torch.cuda.synchronize(dev)

    def note(msg):
        if os.environ.get("PDLP_BENCH_TRACE"):
            print(f"[rank {rank} t={time.perf_counter():.3f}] {msg}", file=sys.stderr, flush=True)

    reset()
    note("warmup")
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    note("timed region")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    note("timed region done")
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    iters = RESTART_PERIOD * args.steps
    value = world * iters / (ms * 1e-3)
    # e2e: whole sharded solve from host data (engine build on every rank + power iteration + solve)
    barrier()
    t0 = time.perf_counter()
    note("e2e build")
    solver2 = build_rank_engine(Kb, mi, c, q, l, u, device=dev, exchange=exchange)
    torch.manual_seed(0)
    note("e2e solve")
    x, y, obj, k, nres, j, status, secs, sig = solver2.solve(adaptive=True, primal_update=True, max_kkt=20_000)
    x_host = x.cpu()
    note(f"e2e done {status} k={k}")
    barrier()
    e2e_s = time.perf_counter() - t0
    # ---- parity of the row-partitioned engine on this hardware (fails the run when it does not hold) ----
    parity = sharded_parity(world, rank, dev, exchange, x)
    # ---- config 5: the 500M-nnz LP, strong scaling over the same GPUs ----
    cfg5 = None
    if os.environ.get("PDLP_BENCH_CFG5", "1") != "0":
        solver.close()
        solver2.close()
        solver = solver2 = None
        torch.cuda.empty_cache()
        cfg5 = run_cfg5(world, rank, dev, note)
    if rank == 0:
        peak, peak_src = measured_peak()
        b_iter_rank = 24 * Kb.nnz + 4 * (m_blk + n + 2) + 72 * n + 48 * m_blk
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"row-partitioned LP m={world * m_blk} n={n} nnz={world * Kb.nnz}: one config-2 row block "
                                   f"(100k rows, 5M nnz) per GPU, fp64, restarts + adaptive step + primal weight",
                       "step": "40 PDHG iterations + restart test", "parallelism": f"row partition over {world} GPUs, " + (
                           "allreduce of n+4 doubles per iteration inside the iteration kernels (peer memory over NVLink, "
                           "LL protocol)" if exchange == "p2p" else "1 NCCL allreduce of n+4 doubles per iteration"),
                       "exchange": exchange, "value_is": "n_gpus x iterations/s",
                       "note": "weak scaling: the LP grows with N (one row block per GPU), so lines of different N are rates "
                               "on different LPs (and not comparable with the cfg2 CPU arm); with exchange=p2p NCCL is only "
                               "the control plane (handles, barriers), the data plane is peer memory over NVLink; the strong-"
                               "scaling result on the fixed 500M-nnz LP is the cfg5 record",
                       "l2": "per-GPU working set ~140 MB > 126 MB L2; no explicit flush"},
            "us_per_iteration": 1e3 * ms / iters,
            "iteration_roofline_frac": b_iter_rank / (ms * 1e-3 / iters) / 1e9 / peak,
            "gpu_launches": ((3 if exchange == "p2p" else 5) * RESTART_PERIOD + 14) * args.steps, "clocks": clocks,
            "e2e": {"value": world * k / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": int(Kb.nnz * 12 + (m_blk + 1) * 4 + 8 * (3 * n + m_blk)),
                    "d2h_bytes_per_step": x_host.numel() * 8 + 64,
                    "what": f"whole sharded solve from host data: {status} k={k} n={nres}, {e2e_s:.3f} s"},
            "roofline": {"bound": "hbm", "kernel": "per-GPU iteration (" + ("3 kernels, exchange fused" if exchange == "p2p" else "4 kernels + allreduce") + ")", "achieved":
                         b_iter_rank / (ms * 1e-3 / iters) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": b_iter_rank / (ms * 1e-3 / iters) / 1e9 / peak, "traffic": None, "peak_source": peak_src},
            "cpu_baseline": None,
            "parity": parity,
            "cfg5": cfg5,
        }
        print(json.dumps(out), flush=True)
    # captured graphs hold NCCL work: drop them (and the engines) before the communicator goes away
    if solver is not None:
        solver.close()
        solver2.close()
    torch.cuda.synchronize(dev)
    dist.barrier()
    dist.destroy_process_group()
    if not parity["ok"]:
        sys.exit(3)


def sharded_parity(world, rank, dev, exchange, x_replica):
    """Hardware parity of the row-partitioned engine: (1) the replicas of x of the solve just done are bit-identical
    on all ranks; (2) config 2 (the real 100k x 200k LP) cut into `world` row blocks gives the same solve as the
    single-GPU engine on the whole LP: fixed step -> equal k / n / j, objective to 1e-9; adaptive -> equal status and
    n, k within one restart period, objective to 1e-6 (the bands of tests/test_gpu_solve.py)."""
    import torch.distributed as dist
    from torchpdlp_b200 import synthetic
    from torchpdlp_b200.engine import Engine
    from torchpdlp_b200.primal_dual_hybrid_gradient import initial_step
    from torchpdlp_b200.sharded import build_rank_engine, partition_rows
    lo, hi = x_replica.clone(), x_replica.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    identical = bool(torch.equal(lo, hi))
    lp = synthetic.uniform_rows_lp()
    cuts = partition_rows(lp.K.indptr, world)
    r0, r1 = int(cuts[rank]), int(cuts[rank + 1])
    mi = min(max(lp.m_ineq - r0, 0), r1 - r0)
    Kb = lp.K[r0:r1].tocsr()
    Kb.sort_indices()
    runs = []
    ok = identical
    for adaptive, pw in ((False, False), (True, True)):
        solver = build_rank_engine(Kb, mi, lp.c, lp.q[r0:r1], lp.l, lp.u, device=dev, exchange=exchange)
        torch.manual_seed(0)
        xs, ys, obj, k, nres, j, status, secs, sig = solver.solve(adaptive=adaptive, primal_update=pw, max_kkt=20_000)
        a, b = xs.clone(), xs.clone()
        dist.all_reduce(a, op=dist.ReduceOp.MIN)
        dist.all_reduce(b, op=dist.ReduceOp.MAX)
        same_x = bool(torch.equal(a, b))
        solver.close()
        rec = {"adaptive": adaptive, "sharded": {"status": status, "k": int(k), "n": int(nres), "j": int(j), "obj": obj},
               "x_replicas_bit_identical": same_x}
        good = torch.zeros(1, device=dev)
        if rank == 0:                                   # the single-GPU engine on the whole LP
            eng = Engine(lp.K, lp.m_ineq, lp.c, lp.q, lp.l, lp.u, device=dev)
            torch.manual_seed(0)
            eta, omega, sigma = initial_step(eng, 100)
            eng.set_iterate(None, None); eng.set_counters(0, 0); eng.set_step(eta, omega, 1.0, adaptive)
            eng.set_options(1e-4, pw, False, 1e-4)
            ref = eng.solve(20_000, RESTART_PERIOD, 3600.0, 0.0, True)
            eng.close()
            rec["single_gpu"] = {"status": ref.status, "k": int(ref.k), "n": int(ref.n), "j": int(ref.j), "obj": ref.prim_obj}
            rel = abs(obj - ref.prim_obj) / max(1.0, abs(ref.prim_obj))
            if adaptive:
                match = status == ref.status and nres == ref.n and abs(k - ref.k) <= RESTART_PERIOD and rel < 1e-6
            else:
                match = status == ref.status and (k, nres, j) == (ref.k, ref.n, ref.j) and rel < 1e-9
            rec["rel_obj_diff"] = rel
            rec["match"] = bool(match and same_x)
            good[0] = 1.0 if rec["match"] else 0.0
        dist.broadcast(good, src=0)
        ok = ok and bool(good.item() > 0.5)
        runs.append(rec)
    return {"ok": bool(ok), "x_replicas_bit_identical": identical, "cfg2_vs_single_gpu": runs}


def run_cfg5(world, rank, dev, note):
    """BASELINE.json config 5: synthetic LP 20M x 40M, 500M nnz (25 per row), row-partitioned over the GPUs of this
    run with ONE allreduce of the partial K^T y (n + 4 doubles) per iteration -- strong scaling: the LP is the same
    for every N, every rank generates only its own row block.  Reports ms / iteration of the hot loop, the per-GPU
    fraction of the HBM roofline, and the time to the reference's termination test at 1e-4."""
    import torch.distributed as dist
    from torchpdlp_b200 import synthetic
    from torchpdlp_b200.sharded import build_rank_engine
    import scipy.sparse as sp
    m, n, nnz_row = 20_000_000, 40_000_000, 25
    t0 = time.perf_counter()
    rng = np.random.default_rng(0)                       # replicated vectors: the same draw on every rank
    x_feas = rng.uniform(-10, 10, size=n)
    l = np.maximum(x_feas - rng.uniform(1, 5, size=n), -1e4)
    u = np.minimum(x_feas + rng.uniform(1, 5, size=n), 1e4)
    c = rng.normal(size=n)
    m_ineq = int(0.8 * m)
    r0, r1 = rank * m // world, (rank + 1) * m // world
    Kb = synthetic._stratified_rows(r1 - r0, n, nnz_row, np.random.default_rng(1000 + rank))
    mi = min(max(m_ineq - r0, 0), r1 - r0)
    sign = np.ones(r1 - r0)
    sign[:mi] = -1.0                                     # ">=" rows are -G x >= -h
    Kb = sp.csr_matrix(sp.diags(sign) @ Kb)
    Kb.sort_indices()
    q = Kb @ x_feas
    q[:mi] -= np.random.default_rng(2000 + rank).uniform(0.1, 5.0, size=mi)
    t_gen = time.perf_counter() - t0
    note(f"cfg5 generated in {t_gen:.1f}s")
    t0 = time.perf_counter()
    # n = 40M: the LL words of the peer-memory exchange would double the bytes on the wire, NCCL moves plain doubles
    solver = build_rank_engine(Kb, mi, c, q, l, u, device=dev, exchange="nccl")
    torch.cuda.synchronize(dev)
    t_build = time.perf_counter() - t0
    nnz_local = int(Kb.nnz)
    del Kb
    torch.manual_seed(0)
    t0 = time.perf_counter()
    eta, omega, sigma = solver.initial_step(30)
    torch.cuda.synchronize(dev)
    t_power = time.perf_counter() - t0
    p = solver.p
    p.set_iterate(None, None); p.set_counters(0, 0); p.set_step(eta, omega, 1.0, True); p.set_options(1e-4, True)
    p.enable(solver.q_norm, solver.c_norm)
    times = []
    for b in range(4):                                   # block 0 runs eagerly, block 1 captures the graph
        dist.barrier(); torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        solver.iterate(RESTART_PERIOD, save_previous=True, use_graph=True)
        e1.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
        solver.restart_check(False)
    ms_iter = min(times[2:]) / RESTART_PERIOD
    note(f"cfg5 hot loop {ms_iter:.3f} ms/iteration")
    # time to tolerance: a fresh solve (same engine, iterate reset) to the reference's termination test
    dist.barrier(); torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    torch.manual_seed(0)
    x, y, obj, k, nres, j, status, secs, sig = solver.solve(adaptive=True, primal_update=True, max_kkt=6_000, power_iters=30)
    torch.cuda.synchronize(dev)
    dist.barrier()
    t_solve = time.perf_counter() - t0
    a, b = x.clone(), x.clone()
    dist.all_reduce(a, op=dist.ReduceOp.MIN)
    dist.all_reduce(b, op=dist.ReduceOp.MAX)
    same_x = bool(torch.equal(a, b))
    solver.close()
    peak, _ = measured_peak()
    bytes_total = 24 * nnz_row * m + 4 * (m + n + 2) + 72 * n + 48 * m
    bytes_rank = 24 * nnz_local + 4 * ((r1 - r0) + n + 2) + 72 * n + 48 * (r1 - r0)     # x-side work is replicated
    return {"workload": f"cfg5: synthetic LP m={m} n={n} nnz={nnz_row * m}, row-partitioned over {world} GPUs (strong scaling), "
                        "fp64, restarts + adaptive step + primal weight",
            "exchange": "1 NCCL allreduce of n + 4 doubles per iteration (320 MB)", "n_gpus": world,
            "ms_per_iteration": ms_iter, "iterations_per_sec": 1e3 / ms_iter,
            "bytes_per_iteration_total": bytes_total, "aggregate_algorithmic_GBps": bytes_total / (ms_iter * 1e-3) / 1e9,
            "per_gpu_roofline_frac": bytes_rank / (ms_iter * 1e-3) / 1e9 / peak,
            "time_to_tolerance": {"rel_kkt": 1e-4, "status": status, "iterations": int(k), "restarts": int(nres),
                                  "kkt_passes": int(j), "seconds": t_solve, "objective": obj,
                                  "includes": "30 power-iteration rounds + every restart test"},
            "x_replicas_bit_identical": same_x, "sigma_max": sigma,
            "setup_s": {"generate_rank_block": t_gen, "engine_build": t_build, "power_iteration_30": t_power}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
