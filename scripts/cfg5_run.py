"""Config 5 (BASELINE.json): synthetic LP, 25 nnz per row, row-partitioned over N GPUs with one NCCL allreduce
of K^T y per iteration.  torchrun --nproc-per-node N scripts/cfg5_run.py [--m 20000000 --n 40000000 --blocks 5]
Every rank generates only its own row block (seeded per block) and the replicated vectors."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from torchpdlp_b200 import synthetic
from torchpdlp_b200.sharded import build_rank_engine

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=20_000_000)
ap.add_argument("--n", type=int, default=40_000_000)
ap.add_argument("--nnz_row", type=int, default=25)
ap.add_argument("--blocks", type=int, default=5)
ap.add_argument("--power_iters", type=int, default=30)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.set_default_dtype(torch.float64)
t0 = time.time()
n, m = args.n, args.m
rng = np.random.default_rng(0)                       # replicated vectors: same draw on every rank
x_feas = rng.uniform(-10, 10, size=n)
l = np.maximum(x_feas - rng.uniform(1, 5, size=n), -1e4)
u = np.minimum(x_feas + rng.uniform(1, 5, size=n), 1e4)
c = rng.normal(size=n)
m_ineq = int(0.8 * m)
r0, r1 = rank * m // world, (rank + 1) * m // world
Kb = synthetic._stratified_rows(r1 - r0, n, args.nnz_row, np.random.default_rng(1000 + rank))
mi = min(max(m_ineq - r0, 0), r1 - r0)
sign = np.ones(r1 - r0); sign[:mi] = -1.0            # ">=" rows are -G x >= -h
import scipy.sparse as sp
Kb = sp.diags(sign) @ Kb
Kb = sp.csr_matrix(Kb); Kb.sort_indices()
q = Kb @ x_feas
q[:mi] -= np.random.default_rng(2000 + rank).uniform(0.1, 5.0, size=mi)
t_gen = time.time() - t0
t0 = time.time()
solver = build_rank_engine(Kb, mi, c, q, l, u, device=dev)
torch.cuda.synchronize()
t_build = time.time() - t0
if rank == 0:
    print(f"cfg5 world={world}: m={m} n={n} nnz={world * Kb.nnz} (per GPU {Kb.nnz}); generate {t_gen:.1f}s, engine build {t_build:.1f}s, "
          f"GPU memory {torch.cuda.max_memory_allocated() / 1e9:.1f} GB torch + engine", flush=True)
torch.manual_seed(0)
t0 = time.time()
eta, omega, sigma = solver.initial_step(args.power_iters)
torch.cuda.synchronize()
if rank == 0:
    print(f"power iteration ({args.power_iters} rounds): sigma={sigma:.6f} in {time.time() - t0:.2f}s; eta={eta:.3e} omega={omega:.3e}", flush=True)
p = solver.p
p.set_iterate(None, None); p.set_counters(0, 0); p.set_step(eta, omega, 1.0, True); p.set_options(1e-4, True)
p.enable(solver.q_norm, solver.c_norm)
times = []
for b in range(args.blocks):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    solver.iterate(40, save_previous=True, use_graph=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    chk = solver.restart_check(False)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    times.append(float(t.item()))
    if rank == 0:
        pt = chk.point
        print(f"block {b}: {times[-1] / 40:.3f} ms/iteration ({40e3 / times[-1]:.1f} it/s)  restart kind={chk.kind} avg={chk.use_average} "
              f"prim={pt.prim_obj:.6e} gap={pt.gap:.3e} pr={pt.pr:.3e} kkt_cur={chk.kkt_current:.4e}", flush=True)
        assert np.isfinite(pt.prim_obj) and np.isfinite(chk.kkt_current)
if rank == 0:
    best = min(times[1:]) if len(times) > 1 else times[0]
    b_iter = 24 * world * Kb.nnz + 4 * (m + n + 2) + 72 * n + 48 * m
    print(json.dumps({"config": "cfg5", "n_gpus": world, "m": m, "n": n, "nnz": world * Kb.nnz, "ms_per_iteration": best / 40,
                      "it_per_s": 40e3 / best, "bytes_per_iteration_total": b_iter,
                      "aggregate_algorithmic_GBps": b_iter / (best / 40 * 1e-3) / 1e9}), flush=True)
solver.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
