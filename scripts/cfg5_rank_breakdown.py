"""Developer probe: where one rank of config 5 spends an iteration.  Builds the row block rank 0 would hold with
WORLD ranks (no process group: world size 1, NCCL path without a peer) and times the column-side SpMV, the step
(rule + element-wise primal update + K x in column windows) and a full iteration with CUDA events.
usage: python scripts/cfg5_rank_breakdown.py [WORLD=8]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp, torch
from torchpdlp_b200 import synthetic
from torchpdlp_b200.sharded import build_rank_engine

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.set_default_dtype(torch.float64)
m, n, nnz_row = 20_000_000, 40_000_000, 25
rng = np.random.default_rng(0)
x_feas = rng.uniform(-10, 10, size=n)
l = np.maximum(x_feas - rng.uniform(1, 5, size=n), -1e4)
u = np.minimum(x_feas + rng.uniform(1, 5, size=n), 1e4)
c = rng.normal(size=n)
rows = m // world
Kb = sp.csr_matrix(-synthetic._stratified_rows(rows, n, nnz_row, np.random.default_rng(1000)))
Kb.sort_indices()
q = Kb @ x_feas - np.random.default_rng(2000).uniform(0.1, 5.0, size=rows)
t0 = time.time()
solver = build_rank_engine(Kb, rows, c, q, l, u, exchange="nccl")
torch.cuda.synchronize()
print(f"world={world}: block {rows} x {n}, nnz {Kb.nnz}; engine build {time.time() - t0:.1f}s", flush=True)
p = solver.p
p.set_iterate(None, None); p.set_counters(0, 0); p.set_step(0.05, 0.1, 1.0, True); p.set_options(1e-4, True)
p.enable(solver.q_norm, solver.c_norm)
def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t_kty = timed(lambda: p.kty(solver.buf))
t_step = timed(lambda: p.step(solver.buf, False))
t_iter = timed(lambda: (p.kty(solver.buf), p.step(solver.buf, False)))
print(f"K_g^T y_g SpMV {t_kty:.3f} ms | rule + primal update + K_g xbar (windows) {t_step:.3f} ms | iteration without exchange {t_iter:.3f} ms")
pa, du = p.e.time_kernels(5)
print(f"fused single-GPU kernels on this block: primal {pa:.3f} ms, dual (windows) {du:.3f} ms")
