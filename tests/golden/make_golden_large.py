"""Full-size golden numbers from the ORACLE (not the reference: the reference
cannot run Ruiz / infeasibility / CSR on sparse K and its COO path needs ~1 s per
iteration at this size).  The oracle was pinned
against the executed reference on config 1 by make_golden.py; this script runs it
at the sizes of configs 2-4 and stores k / n / j / status / objective / trace so
the `-m gpu` tests can compare the CUDA engine at full size without re-running
minutes of CPU work.

Usage:  python tests/golden/make_golden_large.py cfg2 [cfg3 cfg4 ...]
"""
import json
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pdlp_oracle as orc              # noqa: E402
from torchpdlp_b200 import synthetic               # noqa: E402

torch.set_default_dtype(torch.float64)

CASES = {
    "cfg2": (lambda: synthetic.uniform_rows_lp(), [dict(adaptive=True, primal_update=True), dict()]),
    "cfg2_ruiz": (lambda: synthetic.uniform_rows_lp(), [dict(adaptive=True, primal_update=True, ruiz=True)]),
    "cfg3": (lambda: synthetic.infeasible_lp(), [dict(infeasibility_detect=True),
                                                  dict(adaptive=True, primal_update=True, infeasibility_detect=True)]),
    # config 3 at full size with certificates that FIRE under the reference's semantics (enhancements.py:106-161):
    # the decoupled contradictory pair (labelled DUAL_INFEASIBLE by the reference's heuristic, which tests the dual
    # certificate first and accepts a vanishing dx) and a primal ray (labelled PRIMAL_INFEASIBLE the same way)
    "cfg3_infeasible": (lambda: synthetic.isolated_infeasible_lp(), [dict(infeasibility_detect=True, max_kkt=100_000)]),
    "cfg3_unbounded": (lambda: synthetic.unbounded_lp(80_000, 20_000, 200_000, 50, seed=0, name="cfg3_unbounded"),
                       [dict(infeasibility_detect=True, max_kkt=100_000)]),
    "cfg4_small": (lambda: synthetic.powerlaw_rows_lp(160_000, 40_000, 400_000, max_nnz=20_000, name="cfg4_small"),
                   [dict(max_kkt=3000), dict(adaptive=True, primal_update=True, max_kkt=3000)]),
}


def main(names):
    for name in names:
        make, optsets = CASES[name]
        lp = make()
        out = {"lp": {"name": lp.name, "m": lp.m, "n": lp.n, "nnz": lp.nnz, "m_ineq": lp.m_ineq}, "runs": [],
               "torch": torch.__version__}
        for opts in optsets:
            opts = dict(opts)
            ruiz = opts.pop("ruiz", False)
            c, K, q, m_ineq, l, u = lp.torch_tensors(layout="csr")
            c, q, l, u = (v.reshape(-1) for v in (c, q, l, u))
            op = orc.Operator(K)
            torch.manual_seed(0)
            kw = {}
            if ruiz:
                sop, cs, qs, ls, us, sc = orc.precondition(op, c, q, l, u)
                kw = dict(scaling=sc, unscaled=(op, c, q, l, u))
                op, c, q, l, u = sop, cs, qs, ls, us
            t0 = time.time()
            sol = orc.pdlp_algorithm(op, m_ineq, c, q, l, u, max_kkt=opts.pop("max_kkt", 20_000), time_limit=36000,
                                     **opts, **kw)
            dt = time.time() - t0
            print(name, opts, "ruiz" if ruiz else "", sol.k, sol.n, sol.j, sol.status, repr(sol.prim_obj), f"{dt:.1f}s",
                  flush=True)
            out["runs"].append({"options": dict(opts, ruiz=ruiz), "k": sol.k, "n": sol.n, "j": sol.j,
                                "status": sol.status, "obj": sol.prim_obj, "sigma_max": sol.sigma_max,
                                "seconds": dt, "threads": torch.get_num_threads(),
                                "trace": [vars(r) for r in sol.trace]})
            with open(os.path.join(HERE, f"{name}_oracle.json"), "w") as f:
                json.dump(out, f, indent=1)


if __name__ == "__main__":
    main(sys.argv[1:] or ["cfg2"])
