"""torchrun --nproc-per-node N scripts/sharded_check.py [cfg] [nccl|p2p|both]: row-partitioned solve on N GPUs,
compared on rank 0 with the golden numbers (cfg1: executed reference; cfg2: oracle at full size); with both
exchanges also x replicas identical across ranks and a per-iteration timing of the captured restart block."""
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
which = sys.argv[2] if len(sys.argv) > 2 else "nccl"
for exchange in (("nccl", "p2p") if which == "both" else (which,)):
    solver = build_rank_engine(lp.K[r0:r1], mi, lp.c, lp.q[r0:r1], lp.l, lp.u, device=torch.device("cuda", local),
                               exchange=exchange)
    for adaptive, pw in [(False, False), (True, True)]:
        torch.manual_seed(0)
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.time()
        x, y, obj, k, n, j, status, secs, sigma = solver.solve(adaptive=adaptive, primal_update=pw)
        torch.cuda.synchronize()
        lo, hi = x.clone(), x.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"{name} world={world} {exchange} adaptive={adaptive} pw={pw}: {status} k={k} n={n} j={j} obj={obj:.9f} "
                  f"sigma={sigma:.9f} {time.time() - t0:.3f}s ({k / (time.time() - t0):.0f} it/s incl. power iteration) "
                  f"replicas identical: {bool(torch.equal(lo, hi))}", flush=True)
    # timing of the captured restart block
    solver.p.set_iterate(None, None); solver.p.set_counters(0, 0)
    for _ in range(3):
        solver.iterate(40, True, use_graph=True)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        solver.iterate(40, True, use_graph=True)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 400 * 1e3], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{name} world={world} {exchange}: {t.item():.1f} us / iteration (captured block, max over ranks)", flush=True)
    dist.barrier()
    solver.close()
    dist.barrier()
solver = None
if rank == 0:
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "cfg1_reference.json" if name == "cfg1" else f"{name}_oracle.json")))
    runs = gold["runs"]
    for r in runs[:4]:
        ref = r.get("reference", r)
        print("golden", r["options"], ref["k"], ref["n"], ref["j"], ref["status"], ref["obj"])
dist.barrier()
dist.destroy_process_group()
