"""-m gpu: every kernel behind the C ABI against the oracle on the same seeded inputs."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import pdlp_oracle as orc
from torchpdlp_b200 import synthetic

pytestmark = pytest.mark.gpu


def make_engine(lp, dtype=torch.float64):
    from torchpdlp_b200.engine import Engine
    return Engine(lp.K, lp.m_ineq, lp.c, lp.q, lp.l, lp.u, dtype=dtype)


def oracle_parts(lp):
    op = orc.Operator(lp.K, backend="scipy")
    t = lambda v: torch.from_numpy(np.asarray(v, dtype=np.float64))
    return op, t(lp.c), t(lp.q), t(lp.l), t(lp.u)


def ragged_lp(seed=3):
    """Rows of every kind in one matrix: empty rows, 1-entry rows, lane rows up to the 256-entry limit, warp rows
    (257, 2300 and 5000 entries: 32 segments of one row, several pieces) and a 12000-entry CTA row (straight from
    CSR), see stream_spmv.cuh."""
    rng = np.random.default_rng(seed)
    m, n = 700, 20000
    lens = rng.integers(0, 40, size=m)
    lens[::17] = 0
    lens[5] = 5000
    lens[300] = 2300
    lens[301] = 257
    lens[302] = 256
    lens[400] = 12000
    lens[699] = 1
    rows = np.repeat(np.arange(m), lens)
    cols = np.concatenate([np.sort(rng.choice(n, size=L, replace=False)) for L in lens]) if lens.sum() else np.array([])
    vals = rng.normal(size=rows.size)
    K = sp.csr_matrix((vals, (rows, cols)), shape=(m, n))
    K.sort_indices()
    K = sp.csr_matrix((K.data, K.indices.astype(np.int32), K.indptr.astype(np.int32)), shape=K.shape)
    l = rng.uniform(-5, 0, size=n)
    u = l + rng.uniform(0.5, 5, size=n)
    l[::7] = -np.inf
    u[::5] = np.inf
    return synthetic.SyntheticLP("ragged", K, rng.normal(size=n), rng.normal(size=m), l, u, 400)


@pytest.mark.parametrize("which", ["cfg1", "ragged"])
def test_spmv_both_directions(which, cfg1):
    lp = cfg1 if which == "cfg1" else ragged_lp()
    eng = make_engine(lp)
    rng = np.random.default_rng(0)
    v, w = rng.normal(size=lp.n), rng.normal(size=lp.m)
    got = eng.spmv(v).cpu().numpy()
    ref = lp.K @ v
    assert np.max(np.abs(got - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref)))
    got_t = eng.spmv(w, transpose=True).cpu().numpy()
    ref_t = lp.K.T @ w
    assert np.max(np.abs(got_t - ref_t)) <= 1e-12 * max(1.0, np.max(np.abs(ref_t)))
    eng.close()


def test_spmv_fp32_ragged():
    """The fp32 instantiation on every row class (piece sizes differ from fp64: 8 instead of 12 bytes per slot)."""
    lp = ragged_lp()
    eng = make_engine(lp, dtype=torch.float32)
    rng = np.random.default_rng(1)
    v, w = rng.normal(size=lp.n).astype(np.float32), rng.normal(size=lp.m).astype(np.float32)
    K32 = lp.K.astype(np.float32)
    got = eng.spmv(torch.from_numpy(v)).cpu().numpy()
    ref = lp.K @ v.astype(np.float64)
    scale = np.abs(K32) @ np.abs(v)                              # fp32 rounding grows with the row's magnitude sum
    assert np.all(np.abs(got - ref) <= 4e-6 * np.maximum(scale, 1.0))
    got_t = eng.spmv(torch.from_numpy(w), transpose=True).cpu().numpy()
    ref_t = lp.K.T @ w.astype(np.float64)
    scale_t = np.abs(K32).T @ np.abs(w)
    assert np.all(np.abs(got_t - ref_t) <= 4e-6 * np.maximum(scale_t, 1.0))
    eng.close()


@pytest.mark.parametrize("which", ["cfg1", "ragged"])
def test_fixed_and_adaptive_step_match_oracle(which, cfg1):
    lp = cfg1 if which == "cfg1" else ragged_lp()
    op, c, q, l, u = oracle_parts(lp)
    g = torch.Generator().manual_seed(11)
    x = torch.rand(lp.n, generator=g, dtype=torch.float64) * 2 - 1
    y = torch.randn(lp.m, generator=g, dtype=torch.float64)
    y[:lp.m_ineq] = y[:lp.m_ineq].abs()
    eta, omega = 0.07, 1.3
    eng = make_engine(lp)
    # fixed step (step.py:22-40)
    ox, oy = orc.fixed_one_step_pdhg(x, y, c, q, op, l, u, lp.m_ineq, eta, omega)
    eng.set_iterate(x, y)
    eng.set_step(eta, omega, 1.0, adaptive=False)
    eng.iterate(1, use_graph=False)
    gx, gy = (t.cpu() for t in eng.get_iterate())
    assert torch.max(torch.abs(gx - ox)) <= 1e-12 * max(1.0, float(ox.abs().max()))
    assert torch.max(torch.abs(gy - oy)) <= 1e-12 * max(1.0, float(oy.abs().max()))
    st = eng.state()
    assert st.k == 1 and st.t == 1 and st.last_weight == eta and st.eta == eta
    # adaptive step (step.py:64-115), several k so both branches of the rule are hit
    for k, eta_k in [(1, 0.07), (7, 5.0), (400, 1e-3)]:
        ox, oy, ow, on = orc.adaptive_one_step_pdhg(x, y, c, q, op, l, u, lp.m_ineq, eta_k, omega, 1.0, k)
        eng.set_iterate(x, y)
        eng.set_step(eta_k, omega, 1.0, adaptive=True)
        eng.set_counters(k - 1, 0)
        eng.iterate(1, use_graph=False)
        gx, gy = (t.cpu() for t in eng.get_iterate())
        st = eng.state()
        assert torch.max(torch.abs(gx - ox)) <= 1e-12 * max(1.0, float(ox.abs().max()))
        assert torch.max(torch.abs(gy - oy)) <= 1e-11 * max(1.0, float(oy.abs().max()))
        assert abs(st.last_weight - ow) <= 1e-9 * abs(ow), (k, st.last_weight, ow)
        assert abs(st.eta - on) <= 1e-9 * abs(on), (k, st.eta, on)
        assert st.k == k
    eng.close()


@pytest.mark.parametrize("which", ["cfg1", "ragged"])
def test_kkt_matches_oracle(which, cfg1):
    lp = cfg1 if which == "cfg1" else ragged_lp()
    op, c, q, l, u = oracle_parts(lp)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(lp.n, generator=g, dtype=torch.float64)
    y = torch.randn(lp.m, generator=g, dtype=torch.float64)
    lo, up = torch.isinf(l) & (l < 0), torch.isinf(u) & (u > 0)
    ld, ud = torch.where(lo, torch.zeros_like(l), l), torch.where(up, torch.zeros_like(u), u)
    ref = orc.compute_residuals_and_duality_gap(x, y, c, q, op, lp.m_ineq, lo, up, ld, ud)
    eng = make_engine(lp)
    got = eng.kkt(x, y, 0.8)
    for name, a, b in [("pr", got.pr, ref.pr), ("dr", got.dr, ref.dr), ("gap", got.gap, ref.gap),
                       ("prim_obj", got.prim_obj, ref.prim_obj), ("adj_dual", got.adj_dual, ref.adj_dual),
                       ("kkt", got.kkt, ref.kkt(0.8))]:
        assert abs(a - b) <= 1e-11 * max(1.0, abs(b)), (name, a, b)
    eng.close()


def test_power_iteration_matches_oracle(cfg1):
    op, *_ = oracle_parts(cfg1)
    torch.manual_seed(5)
    b0 = torch.randn(cfg1.n, 1, dtype=torch.float64)
    ref = float(orc.spectral_norm_estimate(op, 100, b0=b0))
    eng = make_engine(cfg1)
    got = eng.power_iteration(b0, 100)
    assert abs(got - ref) <= 1e-10 * ref
    npz = np.load(os.path.join(os.path.dirname(__file__), "golden", "cfg1_functions.npz"))
    assert abs(got - float(npz["sigma"])) <= 1e-10 * ref          # the executed reference's value
    eng.close()


def test_ruiz_matches_reference_pins(cfg1):
    npz = np.load(os.path.join(os.path.dirname(__file__), "golden", "cfg1_functions.npz"))
    eng = make_engine(cfg1)
    rounds = eng.ruiz(20, 1e-6)
    D_row, D_col = (t.cpu().numpy() for t in eng.scaling())
    c_s, q_s, l_s, u_s = (t.cpu().numpy() for t in eng.problem())
    assert rounds == 19
    np.testing.assert_allclose(D_col, npz["ruiz_D_col"], rtol=1e-12)
    np.testing.assert_allclose(D_row, npz["ruiz_D_row"], rtol=1e-12)
    np.testing.assert_allclose(c_s, npz["ruiz_c"], rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(q_s, npz["ruiz_q"], rtol=1e-12, atol=1e-300)
    Ks = eng.matrix().to_dense().cpu().numpy()
    np.testing.assert_allclose(Ks.sum(axis=1), npz["ruiz_K_rowsum"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(Ks.sum(axis=0), npz["ruiz_K_colsum"], rtol=1e-10, atol=1e-12)
    # the K^T copy was scaled identically: K^T w through the engine equals K_s^T w
    w = np.random.default_rng(0).normal(size=cfg1.m)
    np.testing.assert_allclose(eng.spmv(w, transpose=True).cpu().numpy(), Ks.T @ w, rtol=1e-10, atol=1e-11)
    eng.close()


@pytest.mark.parametrize("which", ["cfg1", "bounds"])
def test_pock_chambolle_matches_oracle(which, cfg1):
    """pdlp_pock_chambolle (Ruiz, then alpha = 1 row / column L1 scaling: the `preconditioning` PDLP/main.py:8,107
    imports but the snapshot does not define) against the oracle's restatement, orc.precondition(pock_chambolle=True).
    No reference code exists for this function: parity is pinned to the oracle only."""
    lp = cfg1 if which == "cfg1" else synthetic.mixed_bounds_lp()
    op, c, q, l, u = oracle_parts(lp)
    ops, c_s, q_s, l_s, u_s, sc = orc.precondition(op, c, q, l, u, pock_chambolle=True)
    eng = make_engine(lp)
    eng.ruiz(20, 1e-6)
    eng.pock_chambolle()
    D_row, D_col = (t.cpu().numpy() for t in eng.scaling())
    gc, gq, gl, gu = (t.cpu().numpy() for t in eng.problem())
    np.testing.assert_allclose(D_row, sc.D_row.numpy().ravel(), rtol=1e-12)
    np.testing.assert_allclose(D_col, sc.D_col.numpy().ravel(), rtol=1e-12)
    np.testing.assert_allclose(gc, c_s.numpy().ravel(), rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(gq, q_s.numpy().ravel(), rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(gl, l_s.numpy().ravel(), rtol=1e-12, atol=1e-300)      # +-inf stay +-inf
    np.testing.assert_allclose(gu, u_s.numpy().ravel(), rtol=1e-12, atol=1e-300)
    # both packed copies carry the scaled values: K_s v and K_s^T w through the hot-path layouts
    rng = np.random.default_rng(1)
    v, w = rng.normal(size=lp.n), rng.normal(size=lp.m)
    tv, tw = torch.from_numpy(v), torch.from_numpy(w)
    np.testing.assert_allclose(eng.spmv(v).cpu().numpy(), ops.mv(tv).numpy().ravel(), rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(eng.spmv(w, transpose=True).cpu().numpy(), ops.rmv(tw).numpy().ravel(), rtol=1e-10, atol=1e-11)
    # ... including the iteration kernels' own copies (run chunks / panels): one fixed step on the scaled problem
    x0, y0 = torch.from_numpy(rng.uniform(-1, 1, size=lp.n)), torch.from_numpy(np.abs(rng.normal(size=lp.m)))
    ox, oy = orc.fixed_one_step_pdhg(x0.clone(), y0.clone(), c_s.reshape(-1), q_s.reshape(-1), ops, l_s.reshape(-1), u_s.reshape(-1),
                                     lp.m_ineq, 0.05, 0.7, 1.0)
    eng.set_iterate(x0, y0)
    eng.set_step(0.05, 0.7, 1.0, adaptive=False)
    eng.iterate(1, use_graph=False)
    gx, gy = (t.cpu() for t in eng.get_iterate())
    assert float((gx - ox.ravel()).abs().max()) <= 1e-11 * max(1.0, float(ox.abs().max()))
    assert float((gy - oy.ravel()).abs().max()) <= 1e-11 * max(1.0, float(oy.abs().max()))
    eng.close()


def test_step_functions_reference_pins(cfg1, fp64_default):
    """The drop-in step functions against outputs of the EXECUTED reference."""
    from torchpdlp_b200 import primal_dual_hybrid_gradient_step as step
    npz = np.load(os.path.join(os.path.dirname(__file__), "golden", "cfg1_functions.npz"))
    c, K, q, m_ineq, l, u = cfg1.torch_tensors(layout="csr", device="cuda")
    x = torch.from_numpy(npz["x"]).reshape(-1, 1).cuda()
    y = torch.from_numpy(npz["y"]).reshape(-1, 1).cuda()
    eta, omega = float(npz["eta"]), float(npz["omega"])
    fx, fy, e1, e2 = step.fixed_one_step_pdhg(x.clone(), y.clone(), c, q, K, l, u, m_ineq, eta, omega, 1.0)
    np.testing.assert_allclose(fx.cpu().numpy().ravel(), npz["fixed_x"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(fy.cpu().numpy().ravel(), npz["fixed_y"], rtol=0, atol=1e-11)
    ax, ay, aw, an, j = step.adaptive_one_step_pdhg(x.clone(), y.clone(), c, q, K, l, u, m_ineq, eta, omega, 1.0,
                                                    int(npz["adaptive_k"]), 0)
    np.testing.assert_allclose(ax.cpu().numpy().ravel(), npz["adaptive_x"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(ay.cpu().numpy().ravel(), npz["adaptive_y"], rtol=0, atol=1e-11)
    assert abs(float(aw) - float(npz["adaptive_w"])) <= 1e-9 * float(npz["adaptive_w"])
    assert abs(float(an) - float(npz["adaptive_next"])) <= 1e-9 * float(npz["adaptive_next"])
    assert j == 1


@pytest.mark.parametrize("layout", ["sell", "run", "panel", "runpanel"])
@pytest.mark.parametrize("which", ["cfg1", "ragged"])
def test_loop_modes_bitwise_identical(which, layout, cfg1, monkeypatch):
    """Eager launches, the CUDA graph and (SELL layout) the persistent cooperative kernel give bit-identical iterates
    and scalars -- for each layout of the iteration kernels (PDLP_LAYOUT is read when the engine is created; the
    panel kernel is forced on for these small matrices)."""
    monkeypatch.setenv("PDLP_LAYOUT", layout)
    monkeypatch.setenv("PDLP_PANEL_MIN_NNZ", "0")
    lp = cfg1 if which == "cfg1" else ragged_lp()
    outs = []
    for mode in (0, 1, 2):
        eng = make_engine(lp)
        eng.set_iterate(None, None)
        eng.set_step(0.1, 0.5, 1.0, adaptive=True)
        eng.set_options(1e-4, True, False, 1e-4)
        for _ in range(3):
            eng.iterate(40, save_previous=True, use_graph=mode)
            chk = eng.restart_check(False)
        eng.iterate(7, save_previous=False, use_graph=mode)
        x, y = eng.get_iterate()
        st = eng.state()
        outs.append((x.cpu(), y.cpu(), st.eta, st.eta_sum, st.k, st.t, chk.kkt_current, chk.kkt_previous))
        eng.close()
    for o in outs[1:]:
        assert torch.equal(o[0], outs[0][0]) and torch.equal(o[1], outs[0][1]) and o[2:] == outs[0][2:]


@pytest.mark.parametrize("which", ["cfg1", "ragged"])
def test_layouts_agree(which, cfg1, monkeypatch):
    """The three layouts of the iteration kernels (SELL slices, row-major runs, runs + shared-memory panels of K^T)
    sum the same products in different orders: 47 adaptive iterations agree to rounding."""
    lp = cfg1 if which == "cfg1" else ragged_lp()
    outs = []
    monkeypatch.setenv("PDLP_PANEL_MIN_NNZ", "0")
    for layout in ("sell", "run", "panel", "runpanel"):
        monkeypatch.setenv("PDLP_LAYOUT", layout)
        eng = make_engine(lp)
        eng.set_iterate(None, None)
        eng.set_step(0.05, 0.5, 1.0, adaptive=True)
        eng.iterate(40, save_previous=True, use_graph=True)
        eng.iterate(7, save_previous=False, use_graph=False)
        x, y = eng.get_iterate()
        outs.append((x.cpu(), y.cpu(), eng.state().eta))
        eng.close()
    for o in outs[1:]:
        sx, sy = float(outs[0][0].abs().max()) + 1.0, float(outs[0][1].abs().max()) + 1.0
        assert float((o[0] - outs[0][0]).abs().max()) <= 1e-10 * sx
        assert float((o[1] - outs[0][1]).abs().max()) <= 1e-10 * sy
        assert abs(o[2] - outs[0][2]) <= 1e-10 * abs(outs[0][2])


@pytest.mark.parametrize("which,window_mb", [("cfg1", "0.008"), ("ragged", "0.03")])
def test_column_windows_match_the_single_pass(which, window_mb, cfg1, monkeypatch):
    """K x in column windows (DualWindowOp: what config 5 uses because its pair vector is beyond L2), forced here on
    small LPs through PDLP_WINDOW_MB: partial row sums carried through 4 / 11 window launches give the single-pass
    result to rounding, are bitwise reproducible, and drive the same restart decisions."""
    lp = cfg1 if which == "cfg1" else ragged_lp()
    outs = []
    for mb in (None, window_mb, window_mb):
        if mb is None:
            monkeypatch.delenv("PDLP_WINDOW_MB", raising=False)
        else:
            monkeypatch.setenv("PDLP_WINDOW_MB", mb)
        eng = make_engine(lp)
        eng.set_iterate(None, None)
        eng.set_step(0.05, 0.5, 1.0, adaptive=True)
        eng.set_options(1e-4, True, False, 1e-4)
        eng.iterate(40, save_previous=True, use_graph=True)
        chk = eng.restart_check(False)
        eng.iterate(7, save_previous=False, use_graph=False)
        x, y = eng.get_iterate()
        outs.append((x.cpu(), y.cpu(), eng.state().eta, chk.kind, chk.use_average, chk.kkt_current))
        eng.close()
    a, b, c = outs
    assert torch.equal(b[0], c[0]) and torch.equal(b[1], c[1]) and b[2:] == c[2:]          # reproducible
    sx, sy = float(a[0].abs().max()) + 1.0, float(a[1].abs().max()) + 1.0
    assert float((a[0] - b[0]).abs().max()) <= 1e-10 * sx and float((a[1] - b[1]).abs().max()) <= 1e-10 * sy
    assert abs(a[2] - b[2]) <= 1e-10 * abs(a[2]) and a[3:5] == b[3:5] and abs(a[5] - b[5]) <= 1e-9 * abs(a[5])


def test_bitwise_reproducible(cfg1):
    """No atomics on data: two runs of 80 iterations give bit-identical iterates."""
    outs = []
    for _ in range(2):
        eng = make_engine(cfg1)
        eng.set_iterate(None, None)
        eng.set_step(0.1, 0.5, 1.0, adaptive=True)
        eng.iterate(40, save_previous=True, use_graph=True)
        eng.iterate(40, save_previous=True, use_graph=False)
        x, y = eng.get_iterate()
        outs.append((x.cpu(), y.cpu(), eng.state().eta))
        eng.close()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and outs[0][2] == outs[1][2]


def test_fp32_engine_matches_fp32_oracle(cfg1):
    """The reference's MPS path is float32 (util.py:240-246): the fp32 instantiation of every kernel, against the
    oracle run in float32 (tolerances at float32 resolution)."""
    torch.set_default_dtype(torch.float32)
    try:
        lp = cfg1
        f32 = lambda v: torch.from_numpy(np.asarray(v, dtype=np.float32))
        import scipy.sparse as sp
        K32 = sp.csr_matrix((lp.K.data.astype(np.float32), lp.K.indices, lp.K.indptr), shape=lp.K.shape)
        op = orc.Operator(K32, backend="scipy")
        c, q, l, u = f32(lp.c), f32(lp.q), f32(lp.l), f32(lp.u)
        g = torch.Generator().manual_seed(3)
        x = torch.rand(lp.n, generator=g) * 2 - 1
        y = torch.randn(lp.m, generator=g).abs()
        eng = make_engine(lp, dtype=torch.float32)
        ox, oy, ow, on = orc.adaptive_one_step_pdhg(x, y, c, q, op, l, u, lp.m_ineq, 0.05, 1.3, 1.0, 5)
        eng.set_iterate(x, y)
        eng.set_step(0.05, 1.3, 1.0, adaptive=True)
        eng.set_counters(4, 0)
        eng.iterate(1, use_graph=False)
        gx, gy = (t.cpu() for t in eng.get_iterate())
        assert gx.dtype == torch.float32
        assert float((gx - ox).abs().max()) <= 2e-5 * max(1.0, float(ox.abs().max()))
        assert float((gy - oy).abs().max()) <= 2e-5 * max(1.0, float(oy.abs().max()))
        st = eng.state()
        assert abs(st.eta - float(on)) <= 1e-3 * abs(float(on))
        # whole solve in fp32 reaches the same optimum to fp32-level accuracy
        from torchpdlp_b200.primal_dual_hybrid_gradient import pdlp_algorithm
        cc, K, qq, m_ineq, ll, uu = lp.torch_tensors(dtype=torch.float32, layout="csr", device="cuda")
        torch.manual_seed(0)
        xs, obj, k, n, j, status, secs = pdlp_algorithm(K, m_ineq, cc, qq, ll, uu, "cuda", verbose=False, adaptive=True)
        assert status == "Solved" and xs.dtype == torch.float32
        assert abs(obj - (-4273.2619)) <= 1e-3 * 4273.0
        eng.close()
    finally:
        torch.set_default_dtype(torch.float64)


def _same_engine_matrices(a, b, lp):
    """Two engines hold bit-identical K and K^T: structure, values, and both products on random vectors."""
    rng = np.random.default_rng(11)
    v, w = rng.normal(size=lp.n), rng.normal(size=lp.m)
    Ma, Mb = a.matrix().cpu(), b.matrix().cpu()
    assert torch.equal(Ma.crow_indices(), Mb.crow_indices()) and torch.equal(Ma.col_indices(), Mb.col_indices())
    assert torch.equal(Ma.values(), Mb.values())
    assert torch.equal(a.spmv(v).cpu(), b.spmv(v).cpu())
    assert torch.equal(a.spmv(w, transpose=True).cpu(), b.spmv(w, transpose=True).cpu())


@pytest.mark.parametrize("which", ["cfg1", "ragged"])
def test_host_ingest_equals_device_ingest(which, cfg1, monkeypatch):
    """pdlp_create_host (H2D + canonical check + device-side transpose in the library) builds the same engine as
    pdlp_create fed by the torch conversions: scipy int32 / torch int64 host CSR, pinned or pageable, fp64 and fp32."""
    from torchpdlp_b200.engine import Engine
    lp = cfg1 if which == "cfg1" else ragged_lp()
    args = (lp.m_ineq, lp.c, lp.q, lp.l, lp.u)
    host = Engine(lp.K, *args)
    assert host.ingest == "host"
    c, K, q, m_ineq, l, u = lp.torch_tensors(layout="csr", device="cuda")
    dev = Engine(K, m_ineq, c, q, l, u)
    assert dev.ingest == "device"
    _same_engine_matrices(host, dev, lp)
    # the products also agree with scipy (fp64: fixed summation order per row, so only rounding of the order differs)
    v = np.random.default_rng(1).normal(size=lp.n)
    np.testing.assert_allclose(host.spmv(v).cpu().numpy(), lp.K @ v, rtol=1e-12, atol=1e-12)
    # torch CPU CSR with int64 indices in pinned memory, vectors on the host
    ch, Kh, qh, _, lh, uh = lp.torch_tensors(layout="csr")
    Kp = torch.sparse_csr_tensor(Kh.crow_indices().pin_memory(), Kh.col_indices().pin_memory(), Kh.values().pin_memory(),
                                 size=Kh.shape)
    pinned = Engine(Kp, m_ineq, ch, qh, lh, uh, device="cuda")
    assert pinned.ingest == "host"
    _same_engine_matrices(pinned, dev, lp)
    for got, want in zip(pinned.problem(), dev.problem()):
        assert torch.equal(got, want)
    # fp32 engine from the fp64 host matrix
    h32, d32 = Engine(lp.K, *args, dtype=torch.float32), Engine(K.to(torch.float32), m_ineq, c, q, l, u, dtype=torch.float32)
    assert h32.ingest == "host"
    assert torch.equal(h32.spmv(v).cpu(), d32.spmv(v).cpu())
    # switch: PDLP_HOST_INGEST=0 sends host matrices through the torch conversions
    monkeypatch.setenv("PDLP_HOST_INGEST", "0")
    off = Engine(lp.K, *args)
    assert off.ingest == "device"
    _same_engine_matrices(off, dev, lp)
    for e in (host, dev, pinned, h32, d32, off):
        e.close()


def test_host_ingest_rejects_and_falls_back(cfg1):
    """Unsorted columns and duplicate entries are not canonical CSR: the library says so (PDLP_ERR_NOT_CANONICAL) and
    the engine coalesces through the general path; a column outside [0, n) is an error on both."""
    from torchpdlp_b200 import _lib
    from torchpdlp_b200.engine import Engine
    lp = cfg1
    args = (lp.m_ineq, lp.c, lp.q, lp.l, lp.u)
    ref = Engine(lp.K, *args)
    K = lp.K.copy()
    indptr, indices, data = K.indptr.copy(), K.indices.copy(), K.data.copy()
    for r in range(0, lp.m, 3):                     # reverse the entries of every third row
        b, e = indptr[r], indptr[r + 1]
        indices[b:e], data[b:e] = indices[b:e][::-1].copy(), data[b:e][::-1].copy()
    shuffled = (torch.from_numpy(indptr.astype(np.int64)), torch.from_numpy(indices.astype(np.int64)),
                torch.from_numpy(data), K.shape)
    eng = Engine(shuffled, *args)
    assert eng.ingest == "device"
    _same_engine_matrices(eng, ref, lp)
    eng.close()
    # duplicates: every entry split in two halves (first row only gets a real duplicate pair)
    b, e = indptr[0], indptr[1]
    assert e - b >= 1
    dup_ptr = K.indptr.astype(np.int64).copy()
    dup_ptr[1:] += 1
    dup_idx = np.concatenate([K.indices[:b + 1], K.indices[b:]]).astype(np.int64)
    dup_val = np.concatenate([K.data[:b + 1] * 0.25, [K.data[b] * 0.75], K.data[b + 1:]])
    eng = Engine((torch.from_numpy(dup_ptr), torch.from_numpy(dup_idx), torch.from_numpy(dup_val), K.shape), *args)
    assert eng.ingest == "device" and eng.nnz == lp.nnz
    v = np.random.default_rng(2).normal(size=lp.n)
    np.testing.assert_allclose(eng.spmv(v).cpu().numpy(), ref.spmv(v).cpu().numpy(), rtol=1e-13, atol=1e-13)
    eng.close()
    bad = K.indices.astype(np.int64).copy()
    bad[5] = lp.n + 3
    with pytest.raises(_lib.PdlpError, match="outside"):
        Engine((torch.from_numpy(K.indptr.astype(np.int64)), torch.from_numpy(bad), torch.from_numpy(K.data), K.shape), *args)
    ref.close()


def test_in_loop_kernel_timing(cfg1):
    """pdlp_time_kernels_in_loop: every launch leaves its stamps, the two spans are positive and fit inside the
    event-timed iteration, and the call advances the solver like pdlp_iterate (bitwise the same iterate)."""
    a, b = make_engine(cfg1), make_engine(cfg1)
    for e in (a, b):
        e.set_iterate(None, None)
        e.set_step(0.05, 1.0, 1.0, adaptive=True)
    p_ms, d_ms, it_ms = a.time_kernels_in_loop(60)
    b.iterate(60, use_graph=True)
    assert 0 < p_ms < it_ms and 0 < d_ms < it_ms and p_ms + d_ms <= 1.05 * it_ms
    assert a.state().k == b.state().k == 60
    for u, v in zip(a.get_iterate(), b.get_iterate()):
        assert torch.equal(u, v)
    a.set_options(1e-4, False, True, 1e-4)          # infeasibility detection adds a third kernel: not covered
    with pytest.raises(Exception, match="default"):
        a.time_kernels_in_loop(4)
    a.close()
    b.close()


def test_dense_xbar_kernels_track_the_pair_kernels(cfg1, monkeypatch):
    """DualBarOp against DualOp at the kernel level: K2 gathers the dense xbar vector written by K1 and takes K dx from
    the recursion v' = v + (K xbar - v) / (1 + theta) on v = K x.  One iteration: the dual update is bit-identical
    (K xbar is formed by the same operations) and so is dy^T K dx (v = 0).  40 adaptive iterations: the step size -- the
    only consumer of K dx -- and the iterates agree to 1e-10.  (Not longer: unrestarted adaptive iterations amplify
    rounding by ~10^11 in 200 steps -- the oracle's scipy and torch backends, which differ in summation order only, are
    2.5e-5 apart in eta by then -- so long runs are compared through whole solves, tests/test_gpu_solve.py.)  v is
    rebuilt on the device after a restart and after set_iterate: the agreement holds across both."""
    monkeypatch.setenv("PDLP_XBAR", "0")
    a = make_engine(cfg1)
    monkeypatch.setenv("PDLP_XBAR", "1")
    b = make_engine(cfg1)
    rel = lambda u, v: abs(u - v) / max(abs(v), 1e-300)
    # dy^T K dx is a sum with cancellation: compare on the scale of ||dy|| ||dx|| ||K|| (||K|| < 20 here), not its own
    close_inter = lambda sb, sa: abs(sb.inter - sa.inter) <= 1e-8 * (20.0 * (sa.dx2 * sa.dy2) ** 0.5 + abs(sa.inter))

    def same(tol):
        sa, sb = a.state(), b.state()
        assert sa.k == sb.k and rel(sb.eta, sa.eta) < tol and close_inter(sb, sa), (sa.k, sa.eta, sb.eta, sa.inter, sb.inter)
        for u, v in zip(a.get_iterate(), b.get_iterate()):
            assert float((u - v).abs().max()) <= tol * max(1.0, float(u.abs().max()))

    for e in (a, b):
        e.set_iterate(None, None)
        e.set_step(0.05, 1.0, 1.0, adaptive=True)
        e.iterate(1, use_graph=False)
    sa, sb = a.state(), b.state()
    assert torch.equal(a.get_iterate()[1], b.get_iterate()[1]) and torch.equal(a.get_iterate()[0], b.get_iterate()[0])
    assert sa.dy2 == sb.dy2 and sa.dx2 == sb.dx2 and rel(sb.inter, sa.inter) < 1e-12
    for e in (a, b):
        e.iterate(39, save_previous=True, use_graph=True)
    same(1e-10)
    ca, cb = a.restart_check(True), b.restart_check(True)           # forced restart: v = K x recomputed on the device
    assert (ca.kind, ca.use_average) == (cb.kind, cb.use_average) and rel(cb.kkt_current, ca.kkt_current) < 1e-9
    for e in (a, b):
        e.iterate(40, use_graph=True)
    same(1e-8)
    x0, y0 = a.get_iterate()
    for e in (a, b):                                                 # non-zero start through set_iterate
        e.set_iterate(x0, y0)
        e.set_step(0.05, 1.0, 1.0, adaptive=True)
        e.iterate(40, use_graph=True)
    same(1e-10)
    a.close()
    b.close()
