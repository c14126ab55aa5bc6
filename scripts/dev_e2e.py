"""Developer probe: where does the end-to-end solve time go (host tensors -> solution)?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torchpdlp_b200 import synthetic, _cache
from torchpdlp_b200 import engine as eng_mod
from torchpdlp_b200.primal_dual_hybrid_gradient import pdlp_algorithm, initial_step

torch.set_default_dtype(torch.float64)
lp = synthetic.uniform_rows_lp()
c, K, q, m_ineq, l, u = lp.torch_tensors(layout="csr")
host = [t.pin_memory() for t in (K.crow_indices(), K.col_indices(), K.values(), c, q, l, u)]
Kh = torch.sparse_csr_tensor(host[0], host[1], host[2], size=K.shape)
dev = torch.device("cuda")
torch.zeros(1, device=dev); torch.cuda.synchronize()

def T(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print(f"  {label:38s} {1e3 * (time.perf_counter() - t0):8.2f} ms"); return r

for rep in range(3):
    print(f"--- rep {rep}")
    _cache.clear()
    t_all = time.perf_counter()
    (shape, csr, csc) = T("to_device_csr (H2D + transpose)", lambda: eng_mod.to_device_csr(Kh, dev, torch.float64))
    e = T("Engine() total", lambda: eng_mod.Engine(Kh, m_ineq, host[3], host[4], host[5], host[6], device=dev))
    T("initial_step (power iteration 100)", lambda: initial_step(e, 100))
    def solve():
        e.set_iterate(None, None); e.set_counters(0, 0); e.set_step(0.9 / 18.2221, 0.05, 1.0, True); e.set_options(1e-4, True, False, 1e-4)
        return e.solve()
    out = T("solve()", solve)
    print(f"  k={out.k} device_ms={out.device_ms:.1f}")
    T("get_iterate + D2H", lambda: e.get_iterate()[0].cpu())
    e.close()
    _cache.clear()
    torch.manual_seed(0)
    T("pdlp_algorithm end to end", lambda: pdlp_algorithm(Kh, m_ineq, host[3], host[4], host[5], host[6], dev, verbose=False, adaptive=True, primal_update=True))
