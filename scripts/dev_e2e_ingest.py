"""Developer probe: end-to-end solve time from pinned host tensors, host ingest (pdlp_create_host) against the torch
conversions + pdlp_create; PDLP_TRACE_SETUP=1 prints the stages."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torchpdlp_b200 import synthetic, _cache
from torchpdlp_b200.primal_dual_hybrid_gradient import pdlp_algorithm

torch.set_default_dtype(torch.float64)
lp = synthetic.uniform_rows_lp()
c, K, q, m_ineq, l, u = lp.torch_tensors(layout="csr")
host = [t.pin_memory() for t in (K.crow_indices(), K.col_indices(), K.values(), c, q, l, u)]
Kh = torch.sparse_csr_tensor(host[0], host[1], host[2], size=K.shape)
dev = torch.device("cuda")
torch.zeros(1, device=dev); torch.cuda.synchronize()
for mode in ("0", "1", "0", "1"):
    os.environ["PDLP_HOST_INGEST"] = mode
    walls = []
    for rep in range(4):
        _cache.clear()
        torch.manual_seed(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x, obj, k, n, j, status, secs, info = pdlp_algorithm(Kh, m_ineq, host[3], host[4], host[5], host[6], dev, verbose=False,
                                                             adaptive=True, primal_update=True, return_info=True)
        xh = x.cpu()
        walls.append(time.perf_counter() - t0)
    print(f"PDLP_HOST_INGEST={mode}: {status} k={k} obj={obj:.6f} sigma={info['sigma_max']:.12f} wall ms " +
          " ".join(f"{1e3 * w:.1f}" for w in walls) + f"  -> {k / min(walls):.0f} it/s best", flush=True)
