"""Golden vectors for the MPS reader: parse two small MPS files with the REFERENCE's
util.mps_to_standard_form (imported from /root/reference/PDLP; build container only) and store its
dense outputs.  tests/golden/small_generated.mps is written by torchpdlp_b200.util.write_mps from a
seeded feasible LP; tests/golden/small_handwritten.mps exercises G/L/E rows, RANGES on all three senses,
LO/UP/FX/FR bounds, default bounds, a repeated entry and an RHS on the objective row."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/PDLP")
import util as ref_util                                  # noqa: E402  (reference)
from torchpdlp_b200 import synthetic                     # noqa: E402
from torchpdlp_b200.util import write_mps                # noqa: E402

HAND = """NAME hand
ROWS
 N COST
 G g1
 L l1
 E e1
 G rg
 L rl
 E re_pos
 E re_neg
COLUMNS
 x1 COST 1.5 g1 1.0
 x1 l1 2.0 e1 1.0
 x2 COST -2.0 g1 1.0
 x2 rg 3.0 rl 1.0
 x3 COST 0.5 l1 -1.0
 x3 e1 1.0 re_pos 2.0
 x4 re_neg 1.0 rg -1.0
 x4 COST 4.0
 x5 g1 0.25 g1 0.75
 x5 rl 2.0
RHS
 RHS COST -10.0 g1 1.0
 RHS l1 8.0 e1 3.0
 RHS rg 2.0 rl 9.0
 RHS re_pos 4.0 re_neg 5.0
RANGES
 RNG rg 3.0 rl 4.0
 RNG re_pos 2.0 re_neg -1.5
BOUNDS
 LO BND x1 -1.0
 UP BND x1 6.0
 UP BND x2 10.0
 FX BND x3 2.0
 FR BND x4
ENDATA
"""


def main():
    lp = synthetic.feasible_lp(40, 20, 8, 0.2, seed=5, name="small")
    G, A = -lp.K[:lp.m_ineq], lp.K[lp.m_ineq:]
    write_mps(os.path.join(HERE, "small_generated.mps"), lp.c, G, -lp.q[:lp.m_ineq], A, lp.q[lp.m_ineq:], lp.l, lp.u)
    with open(os.path.join(HERE, "small_handwritten.mps"), "w") as f:
        f.write(HAND)
    out = {}
    for tag in ("small_generated", "small_handwritten"):
        c, K, q, m_ineq, l, u = ref_util.mps_to_standard_form(os.path.join(HERE, tag + ".mps"), device="cpu")
        out.update({f"{tag}_c": c.numpy(), f"{tag}_K": K.numpy(), f"{tag}_q": q.numpy(), f"{tag}_l": l.numpy(),
                    f"{tag}_u": u.numpy(), f"{tag}_m_ineq": np.array(m_ineq)})
        print(tag, "K", tuple(K.shape), "m_ineq", m_ineq)
    np.savez_compressed(os.path.join(HERE, "mps_reference.npz"), **out)
    # the generated file round-trips the LP it was written from (fp32 in the reference reader)
    assert np.allclose(out["small_generated_K"], lp.K.toarray(), rtol=1e-6, atol=1e-6)
    assert out["small_generated_m_ineq"] == lp.m_ineq
    print("wrote mps_reference.npz")


if __name__ == "__main__":
    main()
