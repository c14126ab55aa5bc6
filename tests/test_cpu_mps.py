"""CPU: the sparse MPS reader against the outputs of the reference's own reader (tests/golden/mps_reference.npz,
made by tests/golden/make_golden_mps.py), and the console entry's argument surface."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("tag", ["small_generated", "small_handwritten"])
def test_reader_matches_reference_reader(tag):
    from torchpdlp_b200.util import mps_to_standard_form
    z = np.load(os.path.join(GOLD, "mps_reference.npz"))
    c, K, q, m_ineq, l, u = mps_to_standard_form(os.path.join(GOLD, tag + ".mps"), device="cpu", dtype=torch.float64)
    assert K.layout == torch.sparse_csr
    assert m_ineq == int(z[f"{tag}_m_ineq"])
    # the reference reads in float32: compare at float32 resolution
    np.testing.assert_allclose(K.to_dense().numpy(), z[f"{tag}_K"], rtol=1e-6, atol=1e-6)
    for name, t in (("c", c), ("q", q), ("l", l), ("u", u)):
        assert t.shape == z[f"{tag}_{name}"].shape
        np.testing.assert_allclose(t.numpy(), z[f"{tag}_{name}"], rtol=1e-6, atol=1e-6)
    Kd = mps_to_standard_form(os.path.join(GOLD, tag + ".mps"), device="cpu", dense=True, dtype=torch.float32)[1]
    assert Kd.layout == torch.strided and Kd.dtype == torch.float32


def test_standard_bound_types_behind_switch(tmp_path):
    from torchpdlp_b200.util import mps_to_standard_form
    p = tmp_path / "b.mps"
    p.write_text("NAME b\nROWS\n N obj\n G r\nCOLUMNS\n a obj 1 r 1\n b obj 1 r 1\n c obj 1 r 1\n d obj 1 r 1\nRHS\n RHS r 1\n"
                 "BOUNDS\n FR B a\n MI B b\n BV B c\n PL B d\nENDATA\n")
    _, _, _, _, l, u = mps_to_standard_form(str(p))                    # reference rules: FR -> [0, inf), MI/BV/PL ignored
    assert l.reshape(-1).tolist() == [0, 0, 0, 0] and all(np.isinf(u.reshape(-1).numpy()))
    _, _, _, _, l, u = mps_to_standard_form(str(p), reference_bounds=False)
    assert l.reshape(-1).tolist() == [-np.inf, -np.inf, 0, 0] and u.reshape(-1).tolist() == [np.inf, np.inf, 1, np.inf]


def test_main_flags_match_reference():
    from torchpdlp_b200.main import parse_args
    a = parse_args([])
    assert (a.device, a.instance_path, a.tolerance, a.output_path, a.max_kkt, a.time_limit) == \
        ("auto", "feasible", 1e-4, "output", 100_000, 3600)
    assert not any([a.precondition, a.primal_weight_update, a.adaptive_stepsize, a.infeasibility_detect, a.verbose,
                    a.support_sparse, a.fishnet])
    a = parse_args("--precondition --primal_weight_update --adaptive_stepsize --infeasibility_detect --verbose "
                   "--support_sparse --fishnet --max_kkt 5 --time_limit 7 --tolerance 1e-6".split())
    assert all([a.precondition, a.primal_weight_update, a.adaptive_stepsize, a.infeasibility_detect, a.verbose,
                a.support_sparse, a.fishnet]) and (a.max_kkt, a.time_limit, a.tolerance) == (5, 7, 1e-6)
