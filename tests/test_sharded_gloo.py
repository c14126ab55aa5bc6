"""CPU: the row-partitioned orchestration under torch.distributed/gloo with world_size 2.

The per-rank device work is a torch-CPU test double of the pdlp_shard_* calls
(tests/fake_shard_prims.py); what is tested here is torchpdlp_b200.sharded: the row partition,
the collectives, the deferred step rule and the restart protocol, against the oracle's
single-process solve of the same LP."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, adaptive, primal_update, out, exchange="nccl", time_limit=3600.0, skew=0.0):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_default_dtype(torch.float64)
    from fake_shard_prims import FakeShardPrims
    from torchpdlp_b200 import synthetic
    from torchpdlp_b200.sharded import ShardedSolver, partition_rows
    lp = synthetic.feasible_lp()
    cuts = partition_rows(lp.K.indptr, world)
    r0, r1 = int(cuts[rank]), int(cuts[rank + 1])
    mi_local = min(max(lp.m_ineq - r0, 0), r1 - r0)
    prims = FakeShardPrims(lp.K[r0:r1], mi_local, lp.c, lp.q[r0:r1], lp.l, lp.u)
    solver = ShardedSolver(prims, exchange=exchange)
    torch.manual_seed(0)
    if skew:                             # this rank's clock runs ahead: it alone would see the limit one block early
        import time as _time
        import types
        import torchpdlp_b200.primal_dual_hybrid_gradient as drv
        drv.time = types.SimpleNamespace(time=lambda: _time.time() + (skew if rank == 1 else 0.0))   # the loop's clock only
    x, y_local, obj, k, n, j, status, secs, sigma = solver.solve(adaptive=adaptive, primal_update=primal_update,
                                                                 time_limit=time_limit)
    ys = [torch.zeros(int(cuts[i + 1] - cuts[i])) for i in range(world)] if rank == 0 else None
    dist.gather(y_local, ys, dst=0)
    if rank == 0:
        out.put(dict(obj=obj, k=k, n=n, j=j, status=status, sigma=sigma, x=x.numpy(), y=torch.cat(ys).numpy(),
                     collectives=solver.collectives, cuts=cuts.tolist(),
                     p2p_exchanges=getattr(prims, "p2p_exchanges", None)))
    solver.close()                       # p2p: unmap peers, barrier, then free -- every rank must get here
    if exchange == "p2p":
        assert not prims.connected
    dist.destroy_process_group()


def run_world(world, adaptive, primal_update, exchange="nccl", time_limit=3600.0, skew=0.0):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, adaptive, primal_update, out, exchange, time_limit, skew))
             for r in range(world)]
    for p in procs:
        p.start()
    res = out.get(timeout=600)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


@pytest.mark.parametrize("adaptive,primal_update", [(False, False), (True, True)])
def test_two_rank_solve_matches_oracle(adaptive, primal_update):
    from oracle import pdlp_oracle as orc
    from torchpdlp_b200 import synthetic
    torch.set_default_dtype(torch.float64)
    res = run_world(2, adaptive, primal_update)
    lp = synthetic.feasible_lp()
    c, K, q, m_ineq, l, u = lp.torch_tensors(layout="csr")
    torch.manual_seed(0)
    sol = orc.pdlp_algorithm(K, m_ineq, c, q, l, u, adaptive=adaptive, primal_update=primal_update, backend="scipy")
    assert res["status"] == sol.status == "Solved"
    assert res["n"] == sol.n
    if adaptive:
        assert abs(res["k"] - sol.k) <= 40 and abs(res["obj"] - sol.prim_obj) <= 2e-6 * abs(sol.prim_obj)
    else:
        assert (res["k"], res["j"]) == (sol.k, sol.j)
        assert abs(res["obj"] - sol.prim_obj) <= 1e-9 * abs(sol.prim_obj)
        np.testing.assert_allclose(res["x"], sol.x.numpy(), rtol=0, atol=1e-9)
        np.testing.assert_allclose(res["y"], sol.y.numpy(), rtol=0, atol=1e-9)
    assert abs(res["sigma"] - sol.sigma_max) <= 1e-9 * sol.sigma_max
    # one exchange per iteration + 4 per restart test + setup (norm, 2 per power-iteration round + 1)
    # ... + ONE time-limit agreement (before the first block; afterwards the flag rides in the restart test's last collective)
    assert res["collectives"] == res["k"] + 4 * (res["k"] // 40) + 1 + 1 + 100 + 1


def test_two_rank_peer_exchange_orchestration():
    """exchange="p2p": connect once, every collective of the solve (iterations, restart tests, power iteration)
    goes through the peer-memory entry points, the teardown unmaps before it frees; same numbers as the
    host-allreduce orchestration."""
    a = run_world(2, True, True, "nccl")
    b = run_world(2, True, True, "p2p")
    assert (a["k"], a["n"], a["j"], a["status"]) == (b["k"], b["n"], b["j"], b["status"])
    assert a["obj"] == b["obj"] and np.array_equal(a["x"], b["x"]) and np.array_equal(a["y"], b["y"])
    # everything after the initial ||q|| norm (issued before the exchange is connected) is a peer-memory exchange
    assert b["p2p_exchanges"] == b["collectives"] - 1 == b["k"] + 4 * (b["k"] // 40) + 1 + 100 + 1


def test_time_limit_mid_solve_rides_in_the_restart_test():
    """A limit that runs out DURING the solve on one rank's clock only (rank 1 is 1000 s ahead, limit 1000.2 s): the
    flag travels in the last collective of a restart test and both ranks stop at the same block boundary."""
    res = run_world(2, False, False, time_limit=1000.2, skew=1000.0)
    assert res["status"] == "Unsolved (Time limit exceeded)" and res["k"] > 0 and res["k"] % 40 == 0 and res["k"] < 1840


def test_time_limit_is_a_collective_decision():
    """The ranks' clocks disagree (rank 1 runs 1000 s ahead of a 500 s limit): every rank must leave the loop in the
    same block with the reference's status string, and the teardown collectives still match (no hang: the workers
    exit 0).  Without the agreement rank 1 would stop alone and rank 0 would wait in the next exchange."""
    res = run_world(2, False, False, time_limit=500.0, skew=1000.0)
    assert res["status"] == "Unsolved (Time limit exceeded)" and res["k"] == 0 and res["n"] == 0
    res = run_world(2, False, False, "p2p", time_limit=500.0, skew=1000.0)
    assert res["status"] == "Unsolved (Time limit exceeded)" and res["k"] == 0


def test_partition_rows_balances_nnz():
    from torchpdlp_b200 import synthetic
    from torchpdlp_b200.sharded import partition_rows
    lp = synthetic.powerlaw_rows_lp(1600, 400, 4000, max_nnz=500, seed=1)
    for parts in (2, 4, 8):
        cuts = partition_rows(lp.K.indptr, parts)
        assert cuts[0] == 0 and cuts[-1] == lp.m and np.all(np.diff(cuts) > 0)
        nnz = np.diff(lp.K.indptr[cuts])
        assert nnz.max() <= 1.35 * nnz.mean()


def _ruiz_worker(rank, world, port, which, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_default_dtype(torch.float64)
    from torchpdlp_b200 import synthetic
    from torchpdlp_b200.sharded import partition_rows, ruiz_precondition_sharded
    lp = synthetic.feasible_lp() if which == "cfg1" else synthetic.mixed_bounds_lp()
    cuts = partition_rows(lp.K.indptr, world)
    r0, r1 = int(cuts[rank]), int(cuts[rank + 1])
    Ks, c, q, l, u, d_row, d_col, rounds = ruiz_precondition_sharded(lp.K[r0:r1], lp.c, lp.q[r0:r1], lp.l, lp.u)
    parts = [None] * world if rank == 0 else None
    dist.gather_object(dict(d_row=d_row, q=q, vals=Ks.data, d_col=d_col, c=c, l=l, u=u, rounds=rounds), parts, dst=0)
    if rank == 0:
        out.put(parts)
    dist.destroy_process_group()


@pytest.mark.parametrize("which", ["cfg1", "mixed_bounds"])
def test_two_rank_ruiz_equals_the_single_process_scaling(which):
    """ruiz_precondition_sharded on two row blocks (column maxima and the stopping test meet in one all_reduce(MAX)
    per round) against the oracle's Ruiz on the whole matrix -- which the executed reference pins
    (tests/golden/cfg1_functions.npz): same number of rounds, D_row / D_col / scaled values / c, q, l, u bit for bit
    (max, sqrt and division are exact operations; nothing depends on a summation order); infinite bounds stay infinite."""
    from oracle import pdlp_oracle as orc
    from torchpdlp_b200 import synthetic
    torch.set_default_dtype(torch.float64)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ruiz_worker, args=(r, 2, port, which, out)) for r in range(2)]
    for p in procs:
        p.start()
    parts = out.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lp = synthetic.feasible_lp() if which == "cfg1" else synthetic.mixed_bounds_lp()
    t = lambda v: torch.from_numpy(np.asarray(v, dtype=np.float64))
    op = orc.Operator(lp.K, backend="scipy")
    sop, c_s, q_s, l_s, u_s, sc = orc.precondition(op, t(lp.c), t(lp.q), t(lp.l), t(lp.u))
    assert parts[0]["rounds"] == parts[1]["rounds"] == sc.rounds
    for key in ("d_col", "c", "l", "u"):                       # replicated: both ranks hold the same bits
        assert np.array_equal(parts[0][key], parts[1][key])
    assert np.array_equal(parts[0]["d_col"], sc.D_col.numpy()) and np.array_equal(parts[0]["c"], c_s.numpy())
    assert np.array_equal(parts[0]["l"], l_s.numpy()) and np.array_equal(parts[0]["u"], u_s.numpy())
    assert np.array_equal(np.concatenate([p["d_row"] for p in parts]), sc.D_row.numpy())
    assert np.array_equal(np.concatenate([p["q"] for p in parts]), q_s.numpy())
    rows, cols, vals = sop.triplets()                           # row-major, like the concatenated CSR blocks
    assert np.array_equal(np.concatenate([p["vals"] for p in parts]), vals.numpy())
