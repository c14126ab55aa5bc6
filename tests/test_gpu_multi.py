"""-m gpu, needs >= 2 GPUs (skipped otherwise): the row-partitioned engine on real hardware, two ranks over NCCL and
over the peer-memory exchange, against the executed reference's config-1 numbers (tests/golden/cfg1_reference.json)
and for bit-identical replicas of x.  What scripts/sharded_check.py used to print is asserted here."""
import json
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, exchange, adaptive, pw, which, out, ruiz=False):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    torch.set_default_dtype(torch.float64)
    from torchpdlp_b200 import synthetic
    from torchpdlp_b200.sharded import build_rank_engine, partition_rows
    lp = synthetic.feasible_lp() if which == "cfg1" else synthetic.uniform_rows_lp(8000, 2000, 20000, 30, seed=4, name="mid")
    cuts = partition_rows(lp.K.indptr, world)
    r0, r1 = int(cuts[rank]), int(cuts[rank + 1])
    mi = min(max(lp.m_ineq - r0, 0), r1 - r0)
    Kb = lp.K[r0:r1].tocsr()
    Kb.sort_indices()
    if ruiz:
        from torchpdlp_b200.sharded import ruiz_precondition_sharded
        Ks, c, q, l, u, d_row, d_col, _ = ruiz_precondition_sharded(Kb, lp.c, lp.q[r0:r1], lp.l, lp.u)
        solver = build_rank_engine(Ks, mi, c, q, l, u, device=dev, exchange=exchange, scaling=(d_row, d_col))
    else:
        solver = build_rank_engine(Kb, mi, lp.c, lp.q[r0:r1], lp.l, lp.u, device=dev, exchange=exchange)
    torch.manual_seed(0)
    x, y, obj, k, n, j, status, secs, sigma = solver.solve(adaptive=adaptive, primal_update=pw)
    lo, hi = x.clone(), x.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool(torch.equal(lo, hi))
    solver.close()
    if rank == 0:
        out.put(dict(obj=obj, k=int(k), n=int(n), j=int(j), status=status, same_x=same, x=x.cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


def run_world(exchange, adaptive, pw, which="cfg1", world=2, ruiz=False):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, exchange, adaptive, pw, which, out, ruiz)) for r in range(world)]
    for p in procs:
        p.start()
    import queue
    import time
    res, t0 = None, time.time()
    while res is None:                         # fail fast when a rank dies instead of waiting for the queue's timeout
        try:
            res = out.get(timeout=5)
        except queue.Empty:
            dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
            if dead or time.time() - t0 > 600:
                for p in procs:
                    if p.is_alive():
                        p.kill()
                raise AssertionError(f"a rank exited with {dead}" if dead else "timed out waiting for rank 0")
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return res


needs_two = pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")


@needs_two
@pytest.mark.parametrize("exchange", ["nccl", "p2p"])
@pytest.mark.parametrize("adaptive,pw", [(False, False), (True, True)])
def test_two_gpu_solve_matches_executed_reference(exchange, adaptive, pw):
    with open(os.path.join(GOLD, "cfg1_reference.json")) as f:
        runs = json.load(f)["runs"]
    gold = [r for r in runs if r["options"] == dict(ruiz=False, fishnet=False, adaptive=adaptive, primal_weight=pw)][0]["reference"]
    res = run_world(exchange, adaptive, pw)
    assert res["same_x"]                                   # replicas of x bit-identical across the ranks
    assert res["status"] == gold["status"] and res["n"] == gold["n"]
    rel = abs(res["obj"] - gold["obj"]) / max(1.0, abs(gold["obj"]))
    if adaptive:
        assert abs(res["k"] - gold["k"]) <= 40 and rel < 1e-6, (res["k"], gold["k"], rel)
    else:
        assert (res["k"], res["j"]) == (gold["k"], gold["j"]) and rel < 1e-9


@needs_two
def test_two_gpu_solve_in_column_windows(monkeypatch):
    """The config-5 path at small scale: row partition over two ranks AND K_g x in column windows."""
    monkeypatch.setenv("PDLP_WINDOW_MB", "0.008")
    with open(os.path.join(GOLD, "cfg1_reference.json")) as f:
        runs = json.load(f)["runs"]
    gold = [r for r in runs if r["options"] == dict(ruiz=False, fishnet=False, adaptive=False, primal_weight=False)][0]["reference"]
    res = run_world("nccl", False, False)
    assert res["same_x"] and res["status"] == gold["status"]
    assert (res["k"], res["n"], res["j"]) == (gold["k"], gold["n"], gold["j"])
    assert abs(res["obj"] - gold["obj"]) <= 1e-9 * abs(gold["obj"])


@needs_two
@pytest.mark.parametrize("exchange,window_mb", [("nccl", None), ("p2p", None), ("nccl", "0.004")])
@pytest.mark.parametrize("adaptive,pw", [(False, False), (True, True)])
def test_two_gpu_solve_on_dense_xbar(exchange, window_mb, adaptive, pw, monkeypatch):
    """Row partition over two ranks with K_g xbar on the dense xbar vector (DualBarOp; config 5's default): the
    element-wise / fused primal updates write xbar, K_g dx comes from the per-rank recursion on K_g x."""
    monkeypatch.setenv("PDLP_XBAR", "1")
    if window_mb:
        monkeypatch.setenv("PDLP_WINDOW_MB", window_mb)
    with open(os.path.join(GOLD, "cfg1_reference.json")) as f:
        runs = json.load(f)["runs"]
    gold = [r for r in runs if r["options"] == dict(ruiz=False, fishnet=False, adaptive=adaptive, primal_weight=pw)][0]["reference"]
    res = run_world(exchange, adaptive, pw)
    assert res["same_x"] and res["status"] == gold["status"] and res["n"] == gold["n"]
    rel = abs(res["obj"] - gold["obj"]) / max(1.0, abs(gold["obj"]))
    if adaptive:
        assert abs(res["k"] - gold["k"]) <= 40 and rel < 1e-6, (res["k"], gold["k"], rel)
    else:
        assert (res["k"], res["j"]) == (gold["k"], gold["j"]) and rel < 1e-9


@needs_two
@pytest.mark.parametrize("exchange", ["nccl", "p2p"])
def test_two_gpu_solve_with_ruiz(exchange):
    """Ruiz on two ranks (column maxima by all_reduce(MAX) over NCCL, on the GPUs) + the scaled solve that terminates
    on the original problem's residuals, against the executed reference's precondition=True run (fixed step)."""
    with open(os.path.join(GOLD, "cfg1_reference.json")) as f:
        runs = json.load(f)["runs"]
    gold = [r for r in runs if r["options"] == dict(ruiz=True, fishnet=False, adaptive=False, primal_weight=False)][0]["reference"]
    res = run_world(exchange, False, False, ruiz=True)
    assert res["same_x"] and res["status"] == gold["status"]
    assert (res["k"], res["n"], res["j"]) == (gold["k"], gold["n"], gold["j"])
    assert abs(res["obj"] - gold["obj"]) <= 1e-9 * abs(gold["obj"])


@needs_two
def test_two_gpu_exchanges_agree():
    """NCCL allreduce and the fused peer-memory exchange: the same solve (fixed step: bit for bit in k / n / j)."""
    a = run_world("nccl", False, False, "mid")
    b = run_world("p2p", False, False, "mid")
    assert (a["k"], a["n"], a["j"], a["status"]) == (b["k"], b["n"], b["j"], b["status"])
    assert abs(a["obj"] - b["obj"]) <= 1e-9 * max(1.0, abs(a["obj"]))
    np.testing.assert_allclose(a["x"], b["x"], rtol=0, atol=1e-9)
    assert a["same_x"] and b["same_x"]
