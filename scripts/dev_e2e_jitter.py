"""Developer probe: where do the occasional slow e2e solves of bench.py come from?  Times the teardown of the previous
engine (_cache.clear) and the solve separately, with the first (bench) engine alive as in bench.py."""
import os, sys, time, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torchpdlp_b200 import synthetic, _cache
from torchpdlp_b200.engine import Engine
from torchpdlp_b200.primal_dual_hybrid_gradient import pdlp_algorithm, initial_step
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda")
lp = synthetic.make("cfg2")
eng = Engine(lp.K, lp.m_ineq, lp.c, lp.q, lp.l, lp.u, device=dev)
eta, omega, sigma = initial_step(eng, 100)
eng.set_iterate(None, None); eng.set_counters(0, 0); eng.set_step(eta, omega, 1.0, adaptive=True); eng.set_options(1e-4, True, False, 1e-4)
for _ in range(25):
    eng.iterate(40, save_previous=True, use_graph=True); eng.restart_check(False)
eng.time_kernels(200); eng.time_kernels_in_loop(400)
c, K, q, m_ineq, l, u = lp.torch_tensors(layout="csr")
host = [t.pin_memory() for t in (K.crow_indices(), K.col_indices(), K.values(), c, q, l, u)]
Kh = torch.sparse_csr_tensor(host[0], host[1], host[2], size=K.shape)
gc.collect(); gc.disable()
for rep in range(5):
    t0 = time.perf_counter()
    _cache.clear()
    t1 = time.perf_counter()
    torch.manual_seed(0)
    t2 = time.perf_counter()
    x, obj, k, n, j, status, secs, info = pdlp_algorithm(Kh, m_ineq, host[3], host[4], host[5], host[6], dev, verbose=False,
                                                         adaptive=True, primal_update=True, return_info=True)
    t3 = time.perf_counter()
    xh = x.cpu()
    t4 = time.perf_counter()
    print(f"rep {rep}: clear {1e3*(t1-t0):6.1f}  seed {1e3*(t2-t1):5.1f}  solve {1e3*(t3-t2):6.1f}  x.cpu {1e3*(t4-t3):5.1f} ms", flush=True)
