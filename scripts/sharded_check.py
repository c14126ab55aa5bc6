"""torchrun --nproc-per-node N scripts/sharded_check.py [cfg]: row-partitioned solve on N GPUs,
compared on rank 0 with the golden numbers (cfg1: executed reference; cfg2: oracle at full size)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from torchpdlp_b200 import synthetic
from torchpdlp_b200.sharded import build_rank_engine, partition_rows

name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.set_default_dtype(torch.float64)
lp = synthetic.make(name)
cuts = partition_rows(lp.K.indptr, world)
r0, r1 = int(cuts[rank]), int(cuts[rank + 1])
mi = min(max(lp.m_ineq - r0, 0), r1 - r0)
solver = build_rank_engine(lp.K[r0:r1], mi, lp.c, lp.q[r0:r1], lp.l, lp.u, device=torch.device("cuda", local))
for adaptive, pw in [(False, False), (True, True)]:
    torch.manual_seed(0)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.time()
    x, y, obj, k, n, j, status, secs, sigma = solver.solve(adaptive=adaptive, primal_update=pw)
    torch.cuda.synchronize()
    if rank == 0:
        print(f"{name} world={world} adaptive={adaptive} pw={pw}: {status} k={k} n={n} j={j} obj={obj:.9f} sigma={sigma:.9f} "
              f"{time.time() - t0:.3f}s ({k / (time.time() - t0):.0f} it/s incl. power iteration)", flush=True)
if rank == 0:
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "cfg1_reference.json" if name == "cfg1" else f"{name}_oracle.json")))
    runs = gold["runs"]
    for r in runs[:4]:
        ref = r.get("reference", r)
        print("golden", r["options"], ref["k"], ref["n"], ref["j"], ref["status"], ref["obj"])
solver.close()
dist.barrier()
dist.destroy_process_group()
