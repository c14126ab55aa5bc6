import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torchpdlp_b200 import synthetic
from torchpdlp_b200.sharded import build_rank_engine
torch.set_default_dtype(torch.float64)
lp = synthetic.make("cfg2")
for exchange in ("nccl", "p2p"):
    solver = build_rank_engine(lp.K, lp.m_ineq, lp.c, lp.q, lp.l, lp.u, exchange=exchange)
    eta, omega, sigma = solver.initial_step(20)
    solver.p.set_iterate(None, None); solver.p.set_counters(0, 0); solver.p.set_step(eta, omega, 1.0, True)
    for _ in range(3):
        solver.iterate(40, True, use_graph=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        solver.iterate(40, True, use_graph=True)
    e1.record(); torch.cuda.synchronize()
    print(f"world=1 {exchange}: {e0.elapsed_time(e1)/400*1e3:.1f} us / iteration")
    solver.close()
