"""Generate the golden fixtures under tests/golden/ by EXECUTING THE REFERENCE.

Runs only in the build container (needs /root/reference).  For the seeded
config-1 LP (torchpdlp_b200.synthetic.feasible_lp) and every option set it
  1. runs the unmodified reference ``pdlp_algorithm`` (imported from
     /root/reference/PDLP, torch CPU fp64, dense K, torch.manual_seed(0)),
  2. runs the oracle restatement (oracle/pdlp_oracle.py) on the same data,
  3. asserts k / n / j / status are identical and objectives agree to 1e-9 rel,
  4. writes the reference's numbers + the oracle's per-restart trace to
     tests/golden/cfg1_reference.json.
It also pins the step / KKT / Ruiz / infeasibility / fishnet functions one by one
(tests/golden/cfg1_functions.npz) and records small-LP infeasibility statuses.

Usage:  python tests/golden/make_golden.py
"""
import contextlib
import io
import itertools
import json
import os
import re
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/PDLP")

from oracle import pdlp_oracle as orc                      # noqa: E402
from torchpdlp_b200 import synthetic                       # noqa: E402

import enhancements as ref_enh                             # noqa: E402  (reference)
import helpers as ref_helpers                              # noqa: E402
import primal_dual_hybrid_gradient as ref_pdhg             # noqa: E402
import primal_dual_hybrid_gradient_step as ref_step        # noqa: E402
import spectral_casting as ref_fish                        # noqa: E402

torch.set_default_dtype(torch.float64)


def run_reference(lp, ruiz, fishnet, adaptive, primal_weight, infeas=False, layout="dense", tol=1e-4,
                  max_kkt=100_000):
    c, K, q, m_ineq, l, u = lp.torch_tensors(layout=layout)
    time_used, dt = 0.0, None
    torch.manual_seed(0)
    if ruiz:
        K, c, q, l, u, dt, time_used = ref_enh.ruiz_precondition(c, K, q, l, u, device="cpu")
    x0 = y0 = None
    if fishnet:
        x0, y0 = ref_fish.spectral_cast(K, c, q, l, u, m_ineq, k=32, device="cpu")
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        x, obj, k, n, j, status, _ = ref_pdhg.pdlp_algorithm(
            K, m_ineq, c, q, l, u, "cpu", max_kkt=max_kkt, tol=tol, verbose=True, restart_period=40,
            precondition=ruiz, primal_update=primal_weight, adaptive=adaptive, data_precond=dt,
            infeasibility_detect=infeas, time_limit=36000, time_used=time_used, x_init=x0, y_init=y0)
    restarts = re.findall(r"(\w+) restart at iteration (\d+) using the (\w+) iterate", buf.getvalue())
    lines = re.findall(r"\[(\d+)\] Primal Obj: (\S+), Adjusted Dual Obj: (\S+), Gap: (\S+), Prim Res: (\S+), Dual Res: (\S+)",
                       buf.getvalue())
    return dict(obj=obj, k=k, n=n, j=j, status=status, restarts=restarts, lines=lines,
                x=x.reshape(-1).numpy().copy())


def run_oracle(lp, ruiz, fishnet, adaptive, primal_weight, infeas=False, layout="dense", tol=1e-4,
               max_kkt=100_000, backend="torch"):
    c, K, q, m_ineq, l, u = lp.torch_tensors(layout=layout)
    torch.manual_seed(0)
    op = orc.Operator(K, backend=backend)
    c, q, l, u = (v.reshape(-1) for v in (c, q, l, u))
    scaling = unscaled = None
    if ruiz:
        sop, cs, qs, ls, us, scaling = orc.precondition(op, c, q, l, u)
        unscaled = (op, c, q, l, u)
        op, c, q, l, u = sop, cs, qs, ls, us
    x0 = y0 = None
    if fishnet:
        x0, y0 = orc.spectral_cast(op, c, q, l, u, m_ineq, k=32)
    return orc.pdlp_algorithm(op, m_ineq, c, q, l, u, max_kkt=max_kkt, tol=tol, primal_update=primal_weight,
                              adaptive=adaptive, scaling=scaling, unscaled=unscaled, infeasibility_detect=infeas,
                              time_limit=36000, x_init=x0, y_init=y0)


def rel(a, b):
    return abs(a - b) / max(1.0, abs(b))


def main():
    lp = synthetic.feasible_lp()
    out = {"lp": {"name": "cfg1", "m": lp.m, "n": lp.n, "nnz": lp.nnz, "m_ineq": lp.m_ineq, "seed": 0},
           "torch": torch.__version__, "runs": []}
    torch.manual_seed(0)
    out["b0_head"] = torch.randn(lp.n, 1).reshape(-1)[:4].tolist()
    for ruiz, fishnet, adaptive, pw in itertools.product([False, True], repeat=4):
        ref = run_reference(lp, ruiz, fishnet, adaptive, pw)
        sol = run_oracle(lp, ruiz, fishnet, adaptive, pw)
        tag = f"ruiz={int(ruiz)} fishnet={int(fishnet)} adaptive={int(adaptive)} pw={int(pw)}"
        assert (ref["k"], ref["n"], ref["j"], ref["status"]) == (sol.k, sol.n, sol.j, sol.status), \
            (tag, ref["k"], ref["n"], ref["j"], ref["status"], sol.k, sol.n, sol.j, sol.status)
        # fixed step: bitwise-level agreement.  Adaptive step: the rule amplifies last-bit
        # differences (MKL GEMV summation order depends on buffer alignment), so even the
        # reference is only reproducible to ~1e-6 rel between dense and COO runs of itself.
        assert rel(sol.prim_obj, ref["obj"]) < (2e-6 if adaptive else 1e-9), (tag, sol.prim_obj, ref["obj"])
        assert [(r.kind, str(r.t), r.candidate) for r in sol.trace] == [tuple(r) for r in ref["restarts"]], tag
        xerr = float(np.max(np.abs(sol.x.numpy() - ref["x"])))
        print(f"{tag}: k={ref['k']} n={ref['n']} j={ref['j']} {ref['status']} obj={ref['obj']:.9f} "
              f"|oracle-ref| obj {rel(sol.prim_obj, ref['obj']):.1e} x {xerr:.1e}")
        out["runs"].append({
            "options": {"ruiz": ruiz, "fishnet": fishnet, "adaptive": adaptive, "primal_weight": pw},
            "reference": {"k": ref["k"], "n": ref["n"], "j": ref["j"], "status": ref["status"], "obj": ref["obj"],
                          "restarts": ref["restarts"], "lines": ref["lines"]},
            "oracle": {"obj": sol.prim_obj, "sigma_max": sol.sigma_max, "max_abs_x_err_vs_reference": xerr,
                       "trace": [vars(r) for r in sol.trace]},
        })
    # sparse COO K through the reference's core loop (Ruiz/infeasibility cannot run sparse there)
    for adaptive, pw in [(False, False), (True, True)]:
        ref = run_reference(lp, False, False, adaptive, pw, layout="coo")
        for layout, backend in [("coo", "torch"), ("csr", "torch"), ("csr", "scipy")]:
            sol = run_oracle(lp, False, False, adaptive, pw, layout=layout, backend=backend)
            assert (ref["k"], ref["n"], ref["j"], ref["status"]) == (sol.k, sol.n, sol.j, sol.status)
            assert rel(sol.prim_obj, ref["obj"]) < (2e-6 if adaptive else 1e-9)
        out["runs"].append({"options": {"ruiz": False, "fishnet": False, "adaptive": adaptive, "primal_weight": pw,
                                        "layout": "coo"},
                            "reference": {"k": ref["k"], "n": ref["n"], "j": ref["j"], "status": ref["status"],
                                          "obj": ref["obj"], "restarts": ref["restarts"], "lines": ref["lines"]}})
        print(f"coo adaptive={adaptive} pw={pw}: k={ref['k']} obj={ref['obj']:.9f} (oracle coo/csr/scipy agree)")

    # ---- function-level pins (one call each, dense cfg1) ----
    c, K, q, m_ineq, l, u = lp.torch_tensors(layout="dense")
    g = torch.Generator().manual_seed(123)
    x = (l + (u - l) * torch.rand(lp.n, 1, generator=g))
    y = torch.randn(lp.m, 1, generator=g)
    y[:m_ineq] = y[:m_ineq].abs()
    eta, omega = 0.05, 1.7
    fx, fy, _, _ = ref_step.fixed_one_step_pdhg(x.clone(), y.clone(), c, q, K, l, u, m_ineq, eta, omega, 1.0)
    ax, ay, aw, an, _ = ref_step.adaptive_one_step_pdhg(x.clone(), y.clone(), c, q, K, l, u, m_ineq,
                                                        torch.tensor(eta), torch.tensor(omega), 1.0, 7, 0)
    isn, isp = torch.isinf(l) & (l < 0), torch.isinf(u) & (u > 0)
    res = ref_helpers.compute_residuals_and_duality_gap(x, y, c, q, K, m_ineq, isn, isp, l, u)
    kkt = ref_helpers.KKT_error(x, y, c, q, K, m_ineq, torch.tensor(omega), isn, isp, l, u, "cpu")
    torch.manual_seed(5)
    sig = ref_helpers.spectral_norm_estimate_torch(K, 100)
    Ks, cs, qs, ls, us, dtp, _ = ref_enh.ruiz_precondition(c, K, q, l, u)
    np.savez_compressed(os.path.join(HERE, "cfg1_functions.npz"),
                        x=x.numpy().ravel(), y=y.numpy().ravel(), eta=eta, omega=omega,
                        fixed_x=fx.numpy().ravel(), fixed_y=fy.numpy().ravel(),
                        adaptive_x=ax.numpy().ravel(), adaptive_y=ay.numpy().ravel(),
                        adaptive_w=float(aw), adaptive_next=float(an), adaptive_k=7,
                        residuals=np.array([float(v) for v in res]), kkt=float(kkt), sigma=float(sig),
                        ruiz_D_col=dtp[0].numpy().ravel(), ruiz_D_row=dtp[1].numpy().ravel(),
                        ruiz_c=cs.numpy().ravel(), ruiz_q=qs.numpy().ravel(),
                        ruiz_K_rowsum=Ks.sum(dim=1).numpy().ravel(), ruiz_K_colsum=Ks.sum(dim=0).numpy().ravel())
    # oracle vs those pins
    op = orc.Operator(K)
    v = lambda t: t.reshape(-1)
    ox, oy = orc.fixed_one_step_pdhg(v(x), v(y), v(c), v(q), op, v(l), v(u), m_ineq, eta, omega)
    assert torch.allclose(ox, v(fx), rtol=0, atol=1e-13) and torch.allclose(oy, v(fy), rtol=0, atol=1e-12)
    ox, oy, ow, on = orc.adaptive_one_step_pdhg(v(x), v(y), v(c), v(q), op, v(l), v(u), m_ineq, eta, omega, 1.0, 7)
    assert torch.allclose(ox, v(ax), rtol=0, atol=1e-13) and rel(ow, float(aw)) < 1e-12 and rel(on, float(an)) < 1e-12
    orr = orc.compute_residuals_and_duality_gap(v(x), v(y), v(c), v(q), op, m_ineq, v(isn), v(isp), v(l), v(u))
    for a, b in zip((orr.pr, orr.dr, orr.gap, orr.prim_obj, orr.adj_dual), res):
        assert rel(a, float(b)) < 1e-12
    assert rel(orr.kkt(omega), float(kkt)) < 1e-12
    torch.manual_seed(5)
    assert rel(float(orc.spectral_norm_estimate(op, 100)), float(sig)) < 1e-12
    sop, ocs, oqs, ols, ous, sc = orc.precondition(op, v(c), v(q), v(l), v(u))
    assert torch.allclose(sc.D_col, v(dtp[0]), rtol=1e-13, atol=0) and torch.allclose(sc.D_row, v(dtp[1]), rtol=1e-13, atol=0)
    assert torch.allclose(sop.K, Ks, rtol=1e-13, atol=1e-300)
    print("function-level pins OK")

    # ---- infeasibility statuses on small dense LPs (reference vs oracle) ----
    infeas = []
    cases = {"primal_infeasible": synthetic.infeasible_lp(40, 10, 100, 10, seed=0),
             "dual_infeasible": synthetic.unbounded_lp(seed=0),
             "feasible": synthetic.uniform_rows_lp(40, 10, 100, 10, seed=0, name="small")}
    for name, slp in cases.items():
        for adaptive in (False, True):
            ref = run_reference(slp, False, False, adaptive, False, infeas=True, max_kkt=20000)
            sol = run_oracle(slp, False, False, adaptive, False, infeas=True, max_kkt=20000)
            same = (ref["k"], ref["n"], ref["j"], ref["status"]) == (sol.k, sol.n, sol.j, sol.status)
            print(f"infeas {name} adaptive={adaptive}: ref {ref['status']} k={ref['k']} j={ref['j']} obj={ref['obj']:.6g} same={same}")
            assert same, (sol.k, sol.n, sol.j, sol.status)
            assert rel(sol.prim_obj, ref["obj"]) < 1e-8
            infeas.append({"case": name, "adaptive": adaptive, "k": ref["k"], "n": ref["n"], "j": ref["j"],
                           "status": ref["status"], "obj": ref["obj"]})
    out["infeasibility"] = infeas
    with open(os.path.join(HERE, "cfg1_reference.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.join(HERE, "cfg1_reference.json"))


if __name__ == "__main__":
    main()
