"""Golden numbers for an LP with mixed infinite bounds, by EXECUTING THE REFERENCE (build container only).

torchpdlp_b200.synthetic.mixed_bounds_lp has box, lower-only, upper-only and free variables, so the reference's
one-sided / free branches of project_lambda_box and of the dual objective (helpers.py:3-39, 70-112) and the
infinite-bound handling of Ruiz scaling (enhancements.py:64-67) are pinned too.  For every option set the
unmodified reference pdlp_algorithm and the oracle run on the same data; k / n / j / status / restart kinds must be
identical, then the reference's numbers go to tests/golden/bounds_reference.json.

Usage:  python tests/golden/make_golden_bounds.py
"""
import itertools
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import rel, run_oracle, run_reference      # noqa: E402  (imports the reference from /root/reference)
from torchpdlp_b200 import synthetic                        # noqa: E402

torch.set_default_dtype(torch.float64)


def main():
    lp = synthetic.mixed_bounds_lp()
    kinds = {"free": int(np.sum(np.isinf(lp.l) & np.isinf(lp.u))), "lower_only": int(np.sum(~np.isinf(lp.l) & np.isinf(lp.u))),
             "upper_only": int(np.sum(np.isinf(lp.l) & ~np.isinf(lp.u))), "box": int(np.sum(~np.isinf(lp.l) & ~np.isinf(lp.u)))}
    out = {"lp": {"name": lp.name, "m": lp.m, "n": lp.n, "nnz": lp.nnz, "m_ineq": lp.m_ineq, "bounds": kinds},
           "torch": torch.__version__, "runs": []}
    for ruiz, adaptive, pw in itertools.product([False, True], repeat=3):
        ref = run_reference(lp, ruiz, False, adaptive, pw)
        sol = run_oracle(lp, ruiz, False, adaptive, pw)
        tag = f"ruiz={int(ruiz)} adaptive={int(adaptive)} pw={int(pw)}"
        if adaptive:
            # bitwise equal for the first ~830 iterations, then a last-bit difference is amplified by the adaptive
            # rule (the reference's own dense and COO runs part the same way: k = 1200 vs 1560 with the primal
            # weight): same verdict, same optimum to the solve tolerance, k within the spread such runs show here
            # (with the primal weight on top: 1160 vs 720 iterations after Ruiz -- a factor, not a band)
            assert ref["status"] == sol.status == "Solved" and abs(ref["n"] - sol.n) <= 3, (tag, ref["n"], sol.n)
            assert 0.5 * ref["k"] <= sol.k <= 2 * ref["k"] and rel(sol.prim_obj, ref["obj"]) < 1e-4, (tag, ref["k"], sol.k)
        else:
            assert (ref["k"], ref["n"], ref["j"], ref["status"]) == (sol.k, sol.n, sol.j, sol.status), \
                (tag, ref["k"], ref["n"], ref["j"], ref["status"], sol.k, sol.n, sol.j, sol.status)
            assert rel(sol.prim_obj, ref["obj"]) < 1e-9, (tag, sol.prim_obj, ref["obj"])
            assert [(r.kind, str(r.t), r.candidate) for r in sol.trace] == [tuple(r) for r in ref["restarts"]], tag
        xerr = float(np.max(np.abs(sol.x.numpy() - ref["x"])))
        print(f"{tag}: k={ref['k']} n={ref['n']} j={ref['j']} {ref['status']} obj={ref['obj']:.9f} "
              f"|oracle-ref| obj {rel(sol.prim_obj, ref['obj']):.1e} x {xerr:.1e}", flush=True)
        out["runs"].append({"options": {"ruiz": ruiz, "fishnet": False, "adaptive": adaptive, "primal_weight": pw},
                            "reference": {"k": ref["k"], "n": ref["n"], "j": ref["j"], "status": ref["status"],
                                          "obj": ref["obj"], "restarts": ref["restarts"]},
                            "oracle": {"k": sol.k, "n": sol.n, "j": sol.j, "obj": sol.prim_obj, "sigma_max": sol.sigma_max,
                                       "max_abs_x_err_vs_reference": xerr}})
    with open(os.path.join(HERE, "bounds_reference.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
