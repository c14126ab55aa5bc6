"""-m gpu: whole solves through the reference-shaped API against the golden fixtures
(numbers produced by EXECUTING the reference, tests/golden/make_golden.py) and against
the oracle's full-size runs (tests/golden/make_golden_large.py).

Tolerances.  Termination status, restart count n and restart kinds must match.  With the
FIXED step the trajectory is insensitive to summation order: k and j must be equal and the
objective agree to 1e-9 relative.  With the ADAPTIVE step the rule eta' = f(||dx||, ||dy||,
dy^T K dx) amplifies last-bit differences: the reference's own dense and COO runs differ by
4e-7 relative in the final objective; the band is north_star's 1e-6
relative on the primal AND the (adjusted) dual objective, and one restart period (40) on k.
The dual objective of every restart point is checked against the reference's printed lines
(`[k] Primal Obj: ..., Adjusted Dual Obj: ...`, primal_dual_hybrid_gradient.py:167-170; 4 decimals).
"""
ADAPTIVE_REL = 1e-6
import json
import os

import numpy as np
import pytest
import torch

from torchpdlp_b200 import synthetic

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    return abs(a - b) / max(1.0, abs(b))


def load(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


def solve(lp, ruiz=False, adaptive=False, primal_weight=False, layout="csr", verbose=False, **kw):
    from torchpdlp_b200.enhancements import ruiz_precondition
    from torchpdlp_b200.primal_dual_hybrid_gradient import pdlp_algorithm
    torch.set_default_dtype(torch.float64)
    c, K, q, m_ineq, l, u = lp.torch_tensors(layout=layout, device="cuda")
    torch.manual_seed(0)
    dt, used = None, 0.0
    if ruiz:
        K, c, q, l, u, dt, used = ruiz_precondition(c, K, q, l, u, device="cuda")
    return pdlp_algorithm(K, m_ineq, c, q, l, u, "cuda", verbose=verbose, precondition=ruiz, primal_update=primal_weight,
                          adaptive=adaptive, data_precond=dt, time_used=used, return_info=True, **kw)


CFG1_RUNS = [r for r in load("cfg1_reference.json")["runs"]
             if not r["options"]["fishnet"] and r["options"].get("layout") is None]


@pytest.mark.parametrize("run", CFG1_RUNS, ids=lambda r: "-".join(k for k, v in r["options"].items() if v) or "plain")
def test_cfg1_against_executed_reference(run, cfg1):
    o, ref = run["options"], run["reference"]
    x, obj, k, n, j, status, secs, info = solve(cfg1, o["ruiz"], o["adaptive"], o["primal_weight"])
    assert status == ref["status"]
    assert n == ref["n"]
    kinds = [(c.kind, c.use_average) for c in info["trace"]]
    want = [({"Sufficient": 1, "Necessary": 2, "Artificial": 3}[a], 1 if c == "Average" else 0) for a, _, c in ref["restarts"]]
    assert kinds == want
    if o["adaptive"]:
        assert abs(k - ref["k"]) <= 40 and rel(obj, ref["obj"]) < ADAPTIVE_REL, (k, ref["k"], obj, ref["obj"])
    else:
        assert (k, j) == (ref["k"], ref["j"])
        assert rel(obj, ref["obj"]) < 1e-9, (obj, ref["obj"])
    assert x.shape == (cfg1.n, 1) and x.is_cuda
    # primal and adjusted dual objective of every restart point against the lines the reference printed (:167-170)
    assert len(info["trace"]) == len(ref["lines"])
    for chk, line in zip(info["trace"], ref["lines"]):
        p_ref, d_ref = float(line[1]), float(line[2])
        band = (ADAPTIVE_REL if o["adaptive"] else 1e-9)
        assert abs(chk.point.prim_obj - p_ref) <= 0.51e-4 + band * abs(p_ref), (line, chk.point.prim_obj)
        assert abs(chk.point.adj_dual - d_ref) <= 0.51e-4 + band * abs(d_ref), (line, chk.point.adj_dual)
        if not o["adaptive"]:
            assert int(line[0]) == chk.k


BOUNDS_RUNS = load("bounds_reference.json")["runs"]


@pytest.mark.parametrize("run", BOUNDS_RUNS, ids=lambda r: "-".join(k for k, v in r["options"].items() if v) or "plain")
def test_mixed_infinite_bounds_against_executed_reference(run):
    """Box, lower-only, upper-only and free variables (synthetic.mixed_bounds_lp) against the executed reference:
    fixed step exact in k / n / j and restart kinds, objective 1e-9; adaptive compared by verdict and optimum (the
    reference's own runs part by a factor in k on this LP, tests/golden/make_golden_bounds.py)."""
    lp = synthetic.mixed_bounds_lp()
    o, ref = run["options"], run["reference"]
    x, obj, k, n, j, status, secs, info = solve(lp, o["ruiz"], o["adaptive"], o["primal_weight"])
    assert status == ref["status"] == "Solved"
    if o["adaptive"]:
        assert 0.5 * ref["k"] <= k <= 2 * ref["k"] and rel(obj, ref["obj"]) < 1e-4, (k, ref["k"], obj, ref["obj"])
    else:
        kinds = [(c.kind, c.use_average) for c in info["trace"]]
        want = [({"Sufficient": 1, "Necessary": 2, "Artificial": 3}[a], 1 if c == "Average" else 0)
                for a, _, c in ref["restarts"]]
        assert (k, n, j) == (ref["k"], ref["n"], ref["j"]) and kinds == want
        assert rel(obj, ref["obj"]) < 1e-9, (obj, ref["obj"])


def test_cfg1_layouts_and_verbose_loop_agree(cfg1, capsys):
    """Dense, COO and CSR inputs and the verbose (host-driven) loop give the same solve
    as the C++ driver; the verbose lines have the reference's format (:132-170)."""
    base = solve(cfg1, adaptive=True, primal_weight=True)
    for layout in ("dense", "coo"):
        other = solve(cfg1, adaptive=True, primal_weight=True, layout=layout)
        assert other[2:6] == base[2:6] and other[1] == base[1]
    verbose = solve(cfg1, adaptive=True, primal_weight=True, verbose=True)
    assert verbose[2:6] == base[2:6] and verbose[1] == base[1]
    out = capsys.readouterr().out
    assert "restart at iteration 40 using the" in out and "Primal Obj:" in out and "Converged at iteration" in out


def test_cfg1_kkt_limit_and_warm_start(cfg1):
    x, obj, k, n, j, status, _, info = solve(cfg1, max_kkt=100)
    assert status == "Unsolved (KKT passes limit exceeded)"
    # :67 guard: 40 + 40 iterations with a restart test (+3) and a restart (+2) each = 90,
    # then 10 more iterations reach j = 100; the forced epilogue adds 2
    assert (k, n, j) == (90, 3, 102)
    # warm start from the solution: terminates at the first restart test
    xs, objs, ks, ns, js, st, _, info = solve(cfg1, adaptive=True)
    x2, obj2, k2, n2, j2, st2, _, _ = solve(cfg1, adaptive=True, x_init=xs, y_init=info["y"])
    assert st2 == "Solved" and k2 == 40 and rel(obj2, objs) < 1e-4


def test_time_limit_status(cfg1):
    """The reference tests the clock before every iteration (primal_dual_hybrid_gradient.py:68-74); the engine tests
    it before every block of iterations.  A budget that is already spent returns at once with the reference's
    status string and untouched counters; a budget that runs out mid-solve stops at a block boundary."""
    x, obj, k, n, j, status, secs, info = solve(cfg1, time_limit=0.0)
    assert status == "Unsolved (Time limit exceeded)" and (k, n, j) == (0, 0, 0)
    assert torch.count_nonzero(x).item() == 0
    lp = synthetic.make("cfg2")
    for verbose in (False, True):       # fixed step: 2880 iterations (~0.18 s) to converge, the budget ends earlier
        x, obj, k, n, j, status, secs, info = solve(lp, time_limit=0.1, verbose=verbose, power_iters=5)
        assert status == "Unsolved (Time limit exceeded)"
        assert k < 2880 and k % 40 == 0 and secs >= 0.1
        assert k == 0 or np.isfinite(obj)


@pytest.mark.parametrize("idx", [0, 1])
def test_cfg2_full_size_against_oracle(idx):
    gold = load("cfg2_oracle.json")
    run = gold["runs"][idx]
    lp = synthetic.uniform_rows_lp()
    assert (lp.m, lp.n, lp.nnz) == (gold["lp"]["m"], gold["lp"]["n"], gold["lp"]["nnz"])
    o = run["options"]
    x, obj, k, n, j, status, secs, info = solve(lp, adaptive=o.get("adaptive", False),
                                                primal_weight=o.get("primal_update", False))
    assert status == run["status"] and n == run["n"]
    if o.get("adaptive"):
        assert abs(k - run["k"]) <= 40 and rel(obj, run["obj"]) < ADAPTIVE_REL, (k, run["k"], obj, run["obj"])
    else:
        assert (k, j) == (run["k"], run["j"]) and rel(obj, run["obj"]) < 1e-9
    # dual side: the adjusted dual objective (helpers.py:84) of every restart point against the oracle's trace
    band = ADAPTIVE_REL if o.get("adaptive") else 1e-9
    assert len(info["trace"]) == len(run["trace"])
    for chk, t in zip(info["trace"], run["trace"]):
        assert rel(chk.point.prim_obj, t["prim_obj"]) < band and rel(chk.point.adj_dual, t["adj_dual"]) < band, (t, chk.point.adj_dual)
    # size-independent property: the returned point satisfies the reference's termination test
    xv = x.reshape(-1).cpu().numpy()
    r = lp.K @ xv - lp.q
    viol = np.concatenate([r[lp.m_ineq:], np.minimum(r[:lp.m_ineq], 0.0)])
    assert np.linalg.norm(viol) <= 1e-4 * (1 + np.linalg.norm(lp.q)) * (1 + 1e-9)
    assert np.all(xv >= lp.l - 1e-12) and np.all(xv <= lp.u + 1e-12)
    assert rel(float(lp.c @ xv), obj) < 1e-12


@pytest.mark.parametrize("adaptive,pw", [(False, False), (True, True)])
def test_sharded_peer_exchange_single_rank(adaptive, pw, cfg1):
    """The peer-memory exchange (LL words pushed from the SpMV epilogue, segment reduce / broadcast, deferred rule, all
    in shard_fused_kernel) with world size 1 against the host-allreduce path, which the next test pins to the executed
    reference.  Fixed step: identical k / n / j and objective.  Adaptive: the two paths reduce ||dx||^2 over different
    launch geometries, i.e. in a different (fixed) order, so they agree within the adaptive band."""
    from torchpdlp_b200.sharded import build_rank_engine
    torch.set_default_dtype(torch.float64)
    outs = []
    for exchange in ("nccl", "p2p"):
        solver = build_rank_engine(cfg1.K, cfg1.m_ineq, cfg1.c, cfg1.q, cfg1.l, cfg1.u, exchange=exchange)
        torch.manual_seed(0)
        x, y, obj, k, n, j, status, secs, sigma = solver.solve(adaptive=adaptive, primal_update=pw)
        outs.append((x.cpu(), y.cpu(), obj, k, n, j, status))
        solver.close()
    if adaptive:
        assert outs[0][4] == outs[1][4] and outs[0][6] == outs[1][6]            # restarts, status
        assert abs(outs[0][3] - outs[1][3]) <= 40 and rel(outs[0][2], outs[1][2]) < ADAPTIVE_REL
        return
    assert outs[0][2:] == outs[1][2:]
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_sharded_short_row_transpose_uses_the_run_layout(monkeypatch):
    """A row block of a much larger LP has very short rows in K_g^T; pdlp_shard_kty then multiplies on the run layout
    (rule in engine_create, forced here through PDLP_KT_RUN_MIN_NNZ on an LP with 4 entries per row of K^T).  Same solve
    as without it: fixed step, equal k / n / j, objective to 1e-9."""
    from torchpdlp_b200.sharded import build_rank_engine
    torch.set_default_dtype(torch.float64)
    lp = synthetic.uniform_rows_lp(1600, 400, 10_000, 20, seed=2, name="short_rows")      # K^T: 10 000 rows of ~4 entries
    outs = []
    for knob in (None, "0"):
        if knob is None:
            monkeypatch.delenv("PDLP_KT_RUN_MIN_NNZ", raising=False)
        else:
            monkeypatch.setenv("PDLP_KT_RUN_MIN_NNZ", knob)
        solver = build_rank_engine(lp.K, lp.m_ineq, lp.c, lp.q, lp.l, lp.u, exchange="nccl")
        torch.manual_seed(0)
        x, y, obj, k, n, j, status, secs, sigma = solver.solve(max_kkt=4000)
        outs.append((obj, k, n, j, status))
        solver.close()
    assert outs[0][1:] == outs[1][1:] and rel(outs[0][0], outs[1][0]) < 1e-9, outs


@pytest.mark.parametrize("adaptive,pw", [(False, False), (True, True)])
def test_sharded_entry_points_single_rank(adaptive, pw, cfg1):
    """The pdlp_shard_* kernels (partial SpMV -> [allreduce] -> element-wise epilogues, deferred step
    rule) with world size 1 reproduce the fused single-GPU engine and the executed reference."""
    from torchpdlp_b200.sharded import build_rank_engine
    torch.set_default_dtype(torch.float64)
    gold = [r for r in CFG1_RUNS if r["options"] == dict(ruiz=False, fishnet=False, adaptive=adaptive, primal_weight=pw)][0]["reference"]
    solver = build_rank_engine(cfg1.K, cfg1.m_ineq, cfg1.c, cfg1.q, cfg1.l, cfg1.u)
    torch.manual_seed(0)
    x, y, obj, k, n, j, status, secs, sigma = solver.solve(adaptive=adaptive, primal_update=pw)
    assert status == gold["status"] and n == gold["n"]
    if adaptive:
        assert abs(k - gold["k"]) <= 40 and rel(obj, gold["obj"]) < 2e-6
    else:
        assert (k, j) == (gold["k"], gold["j"]) and rel(obj, gold["obj"]) < 1e-9


@pytest.mark.parametrize("exchange", ["nccl", "p2p"])
@pytest.mark.parametrize("adaptive,pw", [(False, False), (True, True)])
def test_sharded_ruiz_single_rank(adaptive, pw, exchange, cfg1):
    """Ruiz on the row-partitioned engine (sharded.ruiz_precondition_sharded -> build_rank_engine(scaling=...)): the
    scaled problem is solved, the restart tests terminate on the residuals of the original one.  World size 1 against
    the executed reference's precondition=True runs: fixed step k / n / j equal, adaptive within the band; D_col x is
    feasible for the ORIGINAL problem to the tolerance.  (Two ranks: tests/test_gpu_multi.py; the scaling itself on two
    ranks, bit for bit against the single-process one: tests/test_sharded_gloo.py.)"""
    from torchpdlp_b200.sharded import build_rank_engine, ruiz_precondition_sharded
    torch.set_default_dtype(torch.float64)
    gold = [r for r in CFG1_RUNS if r["options"] == dict(ruiz=True, fishnet=False, adaptive=adaptive, primal_weight=pw)][0]["reference"]
    Ks, c, q, l, u, d_row, d_col, rounds = ruiz_precondition_sharded(cfg1.K, cfg1.c, cfg1.q, cfg1.l, cfg1.u)
    solver = build_rank_engine(Ks, cfg1.m_ineq, c, q, l, u, exchange=exchange, scaling=(d_row, d_col))
    torch.manual_seed(0)
    x, y, obj, k, n, j, status, secs, sigma = solver.solve(adaptive=adaptive, primal_update=pw)
    solver.close()
    assert status == gold["status"] and n == gold["n"]
    if adaptive:
        assert abs(k - gold["k"]) <= 40 and rel(obj, gold["obj"]) < 2e-6
    else:
        assert (k, j) == (gold["k"], gold["j"]) and rel(obj, gold["obj"]) < 1e-9
    xv = x.cpu().numpy().reshape(-1) * d_col
    r = cfg1.K @ xv - cfg1.q
    viol = np.concatenate([np.minimum(r[:cfg1.m_ineq], 0.0), r[cfg1.m_ineq:]])
    assert np.linalg.norm(viol) <= 1e-4 * (1 + np.linalg.norm(cfg1.q)) * (1 + 1e-9)
    assert rel(float(cfg1.c @ xv), obj) < 1e-9


def test_infeasibility_verdicts_match_executed_reference():
    """Small LPs from the golden file: the reference's certificate heuristic (enhancements.py:106-161), incl. the
    LPs it mislabels, reproduced verdict for verdict with the tests fused into the iteration kernels."""
    from torchpdlp_b200.primal_dual_hybrid_gradient import pdlp_algorithm
    torch.set_default_dtype(torch.float64)
    cases = {"primal_infeasible": synthetic.infeasible_lp(40, 10, 100, 10, seed=0),
             "dual_infeasible": synthetic.unbounded_lp(seed=0),
             "feasible": synthetic.uniform_rows_lp(40, 10, 100, 10, seed=0, name="small")}
    for rec in load("cfg1_reference.json")["infeasibility"]:
        lp = cases[rec["case"]]
        c, K, q, m_ineq, l, u = lp.torch_tensors(layout="csr", device="cuda")
        torch.manual_seed(0)
        for verbose in (False, True):
            x, obj, k, n, j, status, secs = pdlp_algorithm(K, m_ineq, c, q, l, u, "cuda", verbose=verbose, max_kkt=20000,
                                                           adaptive=rec["adaptive"], infeasibility_detect=True)
            assert status == rec["status"], (rec, status, k, j)
            if rec["adaptive"] and status.startswith("Unsolved"):
                # 20k passes of a non-convergent (infeasible) run with the chaotic adaptive rule: the iterate
                # wanders, only the pass accounting is comparable
                assert abs(k - rec["k"]) <= 40 and rel(obj, rec["obj"]) < 1e-2, (rec, k, obj)
            elif rec["adaptive"]:
                assert abs(k - rec["k"]) <= 40 and rel(obj, rec["obj"]) < 1e-4, (rec, k, obj)
            else:
                assert (k, n, j) == (rec["k"], rec["n"], rec["j"]), (rec, k, n, j)
                assert rel(obj, rec["obj"]) < 1e-8


def test_cfg3_full_size_with_infeasibility_detection():
    """Config 3 (contradictory pair of rows in a 100k x 200k LP) with infeasibility_detect on: the oracle's
    full-size run says what the reference's semantics give (its relative termination test is met before any
    certificate test fires); the engine must agree."""
    gold = load("cfg3_oracle.json")
    lp = synthetic.infeasible_lp()
    for run in gold["runs"]:
        o = run["options"]
        x, obj, k, n, j, status, secs, info = solve(lp, adaptive=o.get("adaptive", False), primal_weight=o.get("primal_update", False),
                                                    infeasibility_detect=True, max_kkt=20_000)
        assert status == run["status"] and n == run["n"]
        if o.get("adaptive"):
            assert abs(k - run["k"]) <= 40 and rel(obj, run["obj"]) < ADAPTIVE_REL, (k, run["k"], obj, run["obj"])
        else:
            assert (k, j) == (run["k"], run["j"]) and rel(obj, run["obj"]) < 1e-9


@pytest.mark.parametrize("case", ["cfg3_infeasible", "cfg3_unbounded"])
def test_cfg3_full_size_certificates_fire(case):
    """Config 3 at full size (100k x 200k) on LPs whose certificate tests FIRE under the reference's semantics
    (enhancements.py:106-161, call site primal_dual_hybrid_gradient.py:89-101): a contradictory pair of rows on an
    isolated variable (violation far above the termination tolerance) and a primal ray.  The oracle's full-size runs
    (tests/golden/make_golden_large.py; 9-33 minutes of CPU each) record what the reference's heuristic returns --
    it tests the dual certificate first and accepts a vanishing dx, so the primal-infeasible LP is labelled
    DUAL_INFEASIBLE, as on the small LP the executed reference pins.  The engine, with the tests fused into the
    iteration kernels, must stop with the same verdict at the same iteration (fixed step: k, n, j exact)."""
    path = os.path.join(GOLD, f"{case}_oracle.json")
    if not os.path.exists(path):
        pytest.skip(f"{case}_oracle.json not generated")
    gold = load(f"{case}_oracle.json")
    lp = (synthetic.isolated_infeasible_lp() if case == "cfg3_infeasible"
          else synthetic.unbounded_lp(80_000, 20_000, 200_000, 50, seed=0, name="cfg3_unbounded"))
    assert (lp.m, lp.n, lp.nnz) == (gold["lp"]["m"], gold["lp"]["n"], gold["lp"]["nnz"])
    run = gold["runs"][0]
    # the isolated pair fires (DUAL_INFEASIBLE at k = 25 881); the primal ray does not within 100 000 KKT passes at this
    # size (it does on the 2000 x 4000 version: PRIMAL_INFEASIBLE at k = 18 428) -- then the 48 180 iterations with the
    # certificate sums evaluated in every one of them must give the same pass accounting and objective
    assert run["status"] == ("DUAL_INFEASIBLE" if case == "cfg3_infeasible" else "Unsolved (KKT passes limit exceeded)")
    x, obj, k, n, j, status, secs, info = solve(lp, infeasibility_detect=True, max_kkt=100_000)
    assert status == run["status"], (status, k, run["k"])
    assert (k, n, j) == (run["k"], run["n"], run["j"]), (k, n, j, run["k"], run["n"], run["j"])
    assert rel(obj, run["obj"]) < 1e-8
    if case == "cfg3_infeasible":      # the verbose (host-driven) loop stops at the same iteration
        x2, obj2, k2, n2, j2, status2, *_ = solve(lp, infeasibility_detect=True, max_kkt=100_000, verbose=True)
        assert (status2, k2, n2, j2) == (status, k, n, j)


@pytest.mark.parametrize("adaptive,pw", [(False, False), (True, True)])
def test_cfg1_solve_in_column_windows(adaptive, pw, cfg1, monkeypatch):
    """Whole solves with K x in 4 column windows (DualWindowOp, forced through PDLP_WINDOW_MB; config 5 uses it
    because its pair vector is beyond L2) against the executed reference: same bands as the single-pass kernels."""
    monkeypatch.setenv("PDLP_WINDOW_MB", "0.008")
    gold = [r for r in CFG1_RUNS if r["options"] == dict(ruiz=False, fishnet=False, adaptive=adaptive, primal_weight=pw)][0]["reference"]
    x, obj, k, n, j, status, secs, info = solve(cfg1, adaptive=adaptive, primal_weight=pw)
    assert status == gold["status"] and n == gold["n"]
    if adaptive:
        assert abs(k - gold["k"]) <= 40 and rel(obj, gold["obj"]) < ADAPTIVE_REL
    else:
        assert (k, j) == (gold["k"], gold["j"]) and rel(obj, gold["obj"]) < 1e-9


@pytest.mark.parametrize("window_mb", [None, "0.004"], ids=["single_pass", "windows"])
@pytest.mark.parametrize("adaptive,pw", [(False, False), (True, True)])
def test_cfg1_solve_on_dense_xbar(adaptive, pw, window_mb, cfg1, monkeypatch):
    """K2 on the dense xbar vector (DualBarOp, forced through PDLP_XBAR=1; configs 4 and 5 use it because their
    (x, dx) pair vector is beyond L2): 8-byte gathers, K dx from the recursion on K x.  Against the executed reference,
    same bands as the pair kernels; K xbar is formed exactly as DualOp forms it, so with the fixed step (which never
    looks at K dx) the single-pass solve equals the pair kernels' bit for bit.  Second case: in 4 column windows."""
    gold = [r for r in CFG1_RUNS if r["options"] == dict(ruiz=False, fishnet=False, adaptive=adaptive, primal_weight=pw)][0]["reference"]
    monkeypatch.setenv("PDLP_XBAR", "0")
    x0, obj0, k0, n0, j0, status0, _, info0 = solve(cfg1, adaptive=adaptive, primal_weight=pw)
    monkeypatch.setenv("PDLP_XBAR", "1")
    if window_mb:
        monkeypatch.setenv("PDLP_WINDOW_MB", window_mb)
    x, obj, k, n, j, status, secs, info = solve(cfg1, adaptive=adaptive, primal_weight=pw)
    assert status == gold["status"] and n == gold["n"]
    if adaptive:
        assert abs(k - gold["k"]) <= 40 and rel(obj, gold["obj"]) < ADAPTIVE_REL
        assert abs(k - k0) <= 40 and rel(obj, obj0) < ADAPTIVE_REL
    else:
        assert (k, j) == (gold["k"], gold["j"]) and rel(obj, gold["obj"]) < 1e-9
        if not window_mb:
            assert torch.equal(x, x0) and torch.equal(info["y"], info0["y"])
    # warm start from the solution (kx is recomputed from the new iterate), Ruiz on the same engine family, and the
    # verbose host-driven loop (restart tests through pdlp_restart_check) take the same path
    x2, obj2, k2, n2, j2, status2, *_ = solve(cfg1, ruiz=True, adaptive=adaptive, primal_weight=pw)
    gold_r = [r for r in CFG1_RUNS if r["options"] == dict(ruiz=True, fishnet=False, adaptive=adaptive, primal_weight=pw)][0]["reference"]
    assert status2 == gold_r["status"] and n2 == gold_r["n"]
    if adaptive:
        assert abs(k2 - gold_r["k"]) <= 40 and rel(obj2, gold_r["obj"]) < ADAPTIVE_REL
    else:
        assert (k2, j2) == (gold_r["k"], gold_r["j"]) and rel(obj2, gold_r["obj"]) < 1e-9
    x3, obj3, k3, n3, j3, status3, *_ = solve(cfg1, adaptive=adaptive, primal_weight=pw, verbose=True)
    assert (status3, n3) == (status, n) and abs(k3 - k) <= (40 if adaptive else 0) and rel(obj3, obj) < (ADAPTIVE_REL if adaptive else 1e-12)
    x4, obj4, k4, n4, j4, status4, *_ = solve(cfg1, adaptive=adaptive, primal_weight=pw, x_init=x, y_init=info["y"])
    assert status4 == "Solved" and k4 == 40 and rel(obj4, obj) < 1e-4          # terminates at the first restart test


def test_cfg2_ruiz_full_size_against_oracle():
    """Ruiz on SPARSE K at full size (the reference cannot run it): oracle restatement,
    validated against the dense reference on config 1, vs the engine."""
    run = load("cfg2_ruiz_oracle.json")["runs"][0]
    lp = synthetic.uniform_rows_lp()
    x, obj, k, n, j, status, secs, info = solve(lp, ruiz=True, adaptive=True, primal_weight=True)
    assert status == run["status"] and n == run["n"]
    assert abs(k - run["k"]) <= 40 and rel(obj, run["obj"]) < 2e-6, (k, run["k"], obj, run["obj"])


FISH_RUNS = [r for r in load("cfg1_reference.json")["runs"] if r["options"]["fishnet"]]


@pytest.mark.parametrize("run", FISH_RUNS, ids=lambda r: "-".join(k for k, v in r["options"].items() if v))
def test_fishnet_against_executed_reference(run, cfg1):
    """spectral_cast (32 points, 8 rounds of 32 batched PDHG steps, selection by signed gap, breeding) followed by
    the warm-started solve, against the reference run with the same seed (CPU generator draws)."""
    from torchpdlp_b200.enhancements import ruiz_precondition
    from torchpdlp_b200.primal_dual_hybrid_gradient import pdlp_algorithm
    from torchpdlp_b200.spectral_casting import spectral_cast
    o, ref = run["options"], run["reference"]
    torch.set_default_dtype(torch.float64)
    c, K, q, m_ineq, l, u = cfg1.torch_tensors(layout="csr", device="cuda")
    torch.manual_seed(0)
    dt, used = None, 0.0
    if o["ruiz"]:
        K, c, q, l, u, dt, used = ruiz_precondition(c, K, q, l, u, device="cuda")
    x0, y0 = spectral_cast(K, c, q, l, u, m_ineq, k=32, device="cuda")
    assert x0.shape == (cfg1.n,) and y0.shape == (cfg1.m,)
    x, obj, k, n, j, status, secs = pdlp_algorithm(K, m_ineq, c, q, l, u, "cuda", verbose=False, precondition=o["ruiz"],
                                                   primal_update=o["primal_weight"], adaptive=o["adaptive"], data_precond=dt,
                                                   time_used=used, x_init=x0, y_init=y0)
    assert status == ref["status"] and n == ref["n"]
    if o["adaptive"]:
        assert abs(k - ref["k"]) <= 40 and rel(obj, ref["obj"]) < 2e-6, (k, ref["k"], obj, ref["obj"])
    else:
        assert (k, j) == (ref["k"], ref["j"]) and rel(obj, ref["obj"]) < 1e-8, (k, j, obj, ref)


def test_console_entry_writes_reference_csv(tmp_path):
    """python -m torchpdlp_b200.main on a folder of MPS files: same CSV schema and statuses as PDLP/main.py:142-172."""
    import shutil
    import pandas as pd
    from scipy.optimize import linprog
    from torchpdlp_b200.main import main
    inst, outdir = tmp_path / "inst", tmp_path / "out"
    inst.mkdir()
    shutil.copy(os.path.join(GOLD, "small_generated.mps"), inst / "a.mps")
    shutil.copy(os.path.join(GOLD, "small_handwritten.mps"), inst / "b.mps")
    (inst / "broken.mps").write_text("ROWS\n Q nonsense extra tokens\n")
    res = main(["--instance_path", str(inst), "--output_path", str(outdir), "--adaptive_stepsize", "--primal_weight_update",
                "--precondition", "--tolerance", "1e-6"])
    df = pd.read_csv(outdir / "solver_results.csv")
    assert list(df.columns) == ["File", "Objective", "Iterations (k)", "Restarts (n)", "KKT Passes (j)", "Time (s)", "Status"]
    assert list(df["File"]) == ["a.mps", "b.mps", "broken.mps"]
    assert df["Status"][0] == "Solved"
    lp = synthetic.feasible_lp(40, 20, 8, 0.2, seed=5)
    G, A = -lp.K[:lp.m_ineq], lp.K[lp.m_ineq:]
    ref = linprog(lp.c, A_ub=G, b_ub=-lp.q[:lp.m_ineq], A_eq=A, b_eq=lp.q[lp.m_ineq:], bounds=list(zip(lp.l, lp.u)), method="highs")
    assert abs(float(df["Objective"][0]) - ref.fun) <= 1e-3 * max(1.0, abs(ref.fun))
    assert str(df["Status"][2]).startswith(("Failed to load", "Solver failed")) or "unpack" in str(df["Status"][2])


def test_cfg4_small_powerlaw_rows_against_oracle():
    """Power-law row lengths (2 .. 20 000 nonzeros per row: sigma-sorted SELL slices + the long-row path), 200k x 400k,
    5M nnz, against the oracle run with the same KKT-pass budget (the LP is far from converged after 3000 passes, so
    the fixed-step run is compared exactly and the adaptive one on its pass accounting)."""
    gold = load("cfg4_small_oracle.json")
    lp = synthetic.powerlaw_rows_lp(160_000, 40_000, 400_000, max_nnz=20_000, name="cfg4_small")
    assert (lp.m, lp.n, lp.nnz) == (gold["lp"]["m"], gold["lp"]["n"], gold["lp"]["nnz"])
    for run in gold["runs"]:
        o = run["options"]
        x, obj, k, n, j, status, secs, info = solve(lp, adaptive=o.get("adaptive", False),
                                                    primal_weight=o.get("primal_update", False), max_kkt=3000)
        assert status == run["status"]
        if o.get("adaptive"):
            assert abs(k - run["k"]) <= 80 and abs(n - run["n"]) <= 1 and rel(obj, run["obj"]) < 5e-2, (k, n, obj, run["k"], run["n"], run["obj"])
        else:
            assert (k, n, j) == (run["k"], run["n"], run["j"]) and rel(obj, run["obj"]) < 1e-8, (k, n, j, obj, run["obj"])


def test_cfg4_full_size_properties():
    """Config 4 at full size (2M x 4M, ~50M nnz, rows up to 100 000 long).  No CPU run is affordable here, so the checks
    are size-independent: pass accounting of the reference (:54-177), x inside its box, the reported objective equals
    c^T x, and the reported primal residual equals ||[ (Kx-q)_eq ; min((Kx-q)_ineq, 0) ]|| recomputed on the host."""
    lp = synthetic.powerlaw_rows_lp()
    assert lp.m == 2_000_000 and lp.n == 4_000_000 and lp.nnz > 45_000_000
    x, obj, k, n, j, status, secs, info = solve(lp, adaptive=True, primal_weight=True, max_kkt=400)
    assert status == "Unsolved (KKT passes limit exceeded)"
    assert j >= 400 and k >= 320 and n >= 1
    xv = x.reshape(-1).cpu().numpy()
    assert np.all(xv >= lp.l - 1e-9) and np.all(xv <= lp.u + 1e-9)
    assert rel(float(lp.c @ xv), obj) < 1e-10
    r = lp.K @ xv - lp.q
    viol = np.concatenate([r[lp.m_ineq:], np.minimum(r[:lp.m_ineq], 0.0)])
    assert rel(float(np.linalg.norm(viol)), info["trace"][-1].point.pr) < 1e-9


def test_pdlp_solver_on_an_mps_file(capsys):
    """The importable one-call entry of the reference's older line (Packages/PDLP_without_presolve_infeasibility.py:
    747-789): same arguments, 5-tuple (minimizer, objective, k, n, j), minimizer in the ORIGINAL variables."""
    import scipy.optimize
    import torchpdlp_b200
    from torchpdlp_b200.util import mps_to_standard_form
    path = os.path.join(GOLD, "small_generated.mps")
    torch.manual_seed(0)
    x, obj, k, n, j = torchpdlp_b200.pdlp_solver(path, tol=1e-6, verbose=False)
    c, K, q, m_ineq, l, u = mps_to_standard_form(path, device="cpu")
    Kd = K.to_dense().numpy()
    cv, qv = c.numpy().ravel(), q.numpy().ravel()
    ref = scipy.optimize.linprog(cv, A_ub=-Kd[:m_ineq], b_ub=-qv[:m_ineq], A_eq=Kd[m_ineq:], b_eq=qv[m_ineq:],
                                 bounds=list(zip(l.numpy().ravel(), u.numpy().ravel())), method="highs")
    assert ref.status == 0
    xv = x.cpu().numpy().ravel()
    assert x.shape == (Kd.shape[1], 1) and x.is_cuda and k > 0 and n > 0 and j > k
    assert abs(obj - ref.fun) <= 1e-4 * (1 + abs(ref.fun))
    assert abs(float(cv @ xv) - obj) <= 1e-9 * (1 + abs(obj))            # x is in the original variables
    r = Kd @ xv - qv
    assert np.linalg.norm(np.concatenate([r[m_ineq:], np.minimum(r[:m_ineq], 0)])) <= 1e-5 * (1 + np.linalg.norm(qv))
    x2, obj2, *_ = torchpdlp_b200.pdlp_solver(path, tol=1e-6, verbose=True, precondition=False, adaptive_step=False,
                                               primal_update=False)
    assert abs(obj2 - ref.fun) <= 1e-4 * (1 + abs(ref.fun))
    assert "Objective Value:" in capsys.readouterr().out


def test_engine_cache_never_serves_a_stale_problem(cfg1):
    """Two different same-shaped LPs back to back, and an in-place edit of q between two solves, must each be solved
    as given (the functional entry points look their engine up in torchpdlp_b200/_cache.py)."""
    from torchpdlp_b200 import _cache
    from torchpdlp_b200.primal_dual_hybrid_gradient import pdlp_algorithm
    torch.set_default_dtype(torch.float64)
    _cache.clear()
    objs = []
    for seed in (0, 1):
        lp = synthetic.feasible_lp(seed=seed)
        c, K, q, m_ineq, l, u = lp.torch_tensors(layout="csr", device="cuda")
        torch.manual_seed(0)
        objs.append(pdlp_algorithm(K, m_ineq, c, q, l, u, "cuda", verbose=False, adaptive=True)[1])
        del c, K, q, l, u                      # the next LP may be allocated on the same addresses
    assert rel(objs[0], objs[1]) > 1e-3        # different problems, different optima
    lp = synthetic.feasible_lp(seed=0)
    c, K, q, m_ineq, l, u = lp.torch_tensors(layout="csr", device="cuda")
    torch.manual_seed(0)
    a = pdlp_algorithm(K, m_ineq, c, q, l, u, "cuda", verbose=False, adaptive=True)[1]
    assert rel(a, objs[0]) < 1e-12             # same LP as the first one: same answer
    q.mul_(1.05)                               # in place: same tensor object, same address
    torch.manual_seed(0)
    b = pdlp_algorithm(K, m_ineq, c, q, l, u, "cuda", verbose=False, adaptive=True)[1]
    assert rel(a, b) > 1e-4
    _cache.clear()


def _solve_record(lp, monkeypatch, pipelined, **kw):
    if pipelined:
        monkeypatch.delenv("PDLP_NO_PIPELINE", raising=False)
    else:
        monkeypatch.setenv("PDLP_NO_PIPELINE", "1")
    x, obj, k, n, j, status, _, info = solve(lp, **kw)
    trace = [(c.kind, c.use_average, c.t, c.solved, c.k, c.omega, c.kkt_current, c.kkt_average, c.point.prim_obj,
              c.point.adj_dual) for c in info["trace"]]
    return x.cpu(), info["y"].cpu(), obj, k, n, j, status, trace


@pytest.mark.parametrize("adaptive,pw", [(False, False), (True, True)])
def test_pipelined_solve_is_bitwise_the_sequential_solve(adaptive, pw, cfg1, monkeypatch):
    """pdlp_solve issues the next block and restart test before it reads the previous test's record (the restart is
    applied on the device; a solved restart point raises `stop` so the work issued ahead runs as no-ops).  Same
    iterate, counters, status and per-restart records as the sequential loop, bit for bit -- to convergence and at
    KKT-pass limits on either side of every place where the budget ends the pipelining."""
    kw = dict(adaptive=adaptive, primal_weight=pw)
    a = _solve_record(cfg1, monkeypatch, True, **kw)
    b = _solve_record(cfg1, monkeypatch, False, **kw)
    assert a[6] == "Solved" and a[2:8] == b[2:8]
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    # the engine is not left stopped: a further block moves the iterate
    from torchpdlp_b200.engine import Engine
    from torchpdlp_b200.primal_dual_hybrid_gradient import initial_step
    monkeypatch.delenv("PDLP_NO_PIPELINE", raising=False)
    eng = Engine(cfg1.K, cfg1.m_ineq, cfg1.c, cfg1.q, cfg1.l, cfg1.u)
    torch.manual_seed(0)
    eta, omega, _ = initial_step(eng, 100)
    eng.set_iterate(None, None)
    eng.set_counters(0, 0)
    eng.set_step(eta, omega, 1.0, adaptive)
    eng.set_options(1e-4, pw, False, 1e-4)
    out = eng.solve()
    assert (out.status, out.k, out.n, out.j) == (a[6], a[3], a[4], a[5]) and torch.equal(out.x.cpu().reshape(-1), a[0].reshape(-1))
    x0, _ = eng.get_iterate()
    eng.iterate(40, save_previous=True)
    x1, _ = eng.get_iterate()
    assert not torch.equal(x0, x1) and eng.state().k == out.k + 40
    eng.close()
    for limit in (1, 39, 40, 44, 45, 46, 84, 85, 86, 88, 89, 90, 91, 95, 130, 131, 135, 136, 137, 200, 333):
        a = _solve_record(cfg1, monkeypatch, True, max_kkt=limit, **kw)
        b = _solve_record(cfg1, monkeypatch, False, max_kkt=limit, **kw)
        assert a[2:8] == b[2:8], (limit, a[2:7], b[2:7])
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]), limit
