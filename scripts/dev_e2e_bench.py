"""Developer probe: the e2e leg of bench.py (a first engine alive, then whole solves from pinned host tensors)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torchpdlp_b200 import synthetic, _cache
from torchpdlp_b200.engine import Engine
from torchpdlp_b200.primal_dual_hybrid_gradient import pdlp_algorithm, initial_step
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda")
lp = synthetic.make("cfg2")
keep = os.environ.get("KEEP_FIRST", "1") == "1"
eng = Engine(lp.K, lp.m_ineq, lp.c, lp.q, lp.l, lp.u, device=dev)
initial_step(eng, 100)
if not keep:
    eng.close()
c, K, q, m_ineq, l, u = lp.torch_tensors(layout="csr")
host = [t.pin_memory() for t in (K.crow_indices(), K.col_indices(), K.values(), c, q, l, u)]
Kh = torch.sparse_csr_tensor(host[0], host[1], host[2], size=K.shape)
for rep in range(4):
    _cache.clear()
    torch.manual_seed(0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    x, obj, k, n, j, status, secs, info = pdlp_algorithm(Kh, m_ineq, host[3], host[4], host[5], host[6], dev, verbose=False,
                                                         adaptive=True, primal_update=True, return_info=True)
    xh = x.cpu()
    torch.cuda.synchronize()
    print(f"rep {rep}: {1e3 * (time.perf_counter() - t0):.1f} ms total, reported secs {secs * 1e3:.1f} ms, device {info.get('device_ms'):.1f} ms, k={k}", flush=True)
