"""Host logic of the C-ABI library without a GPU: the SELL plan builder (row classes, sorting window,
pieces, per-warp ranges) through pdlp_plan_inspect (include/pdlp_b200.h).  No compute calls."""
import ctypes as C

import numpy as np
import pytest

from torchpdlp_b200 import _lib, synthetic


def inspect(lens, value_bytes=8, sms=148):
    lens = np.asarray(lens, dtype=np.int64)
    rowptr = np.zeros(len(lens) + 1, dtype=np.int32)
    np.cumsum(lens, out=rowptr[1:])
    info = _lib.PlanInfo()
    cls = np.full(len(lens), -9, dtype=np.int8)
    rc = _lib.lib().pdlp_plan_inspect(rowptr.ctypes.data, len(lens), value_bytes, sms, C.byref(info),
                                      cls.ctypes.data)
    assert rc == 0, _lib.lib().pdlp_last_error()
    return info, cls


def check_invariants(lens, info, cls):
    lens = np.asarray(lens)
    assert info.rows == len(lens) and info.nnz == lens.sum()
    assert info.covered_once == 1 and info.ranges_ok == 1
    assert info.lane_rows + info.warp_rows + info.cta_rows == len(lens)
    # classes follow the thresholds of stream_spmv.cuh (PDLP_WARP_ROW = 256, PDLP_LONG_ROW = 8192)
    assert np.array_equal(cls == 2, lens > 8192)
    assert np.array_equal(cls == 1, (lens > 256) & (lens <= 8192))
    assert np.array_equal(cls == 0, lens <= 256)
    in_sell = lens[cls != 2].sum()
    assert info.slots >= in_sell                                  # padding only ever adds slots
    assert info.warps == info.grid * 12 and 1 <= info.grid <= 148
    assert info.max_piece_bytes <= 16 + 16 * 128 + 8 * 32 * 12    # header + <=16 slices of row ids + KB steps
    assert info.stages >= 2


def test_uniform_rows_keep_their_order_and_have_no_padding():
    lens = np.full(100_000, 50)
    info, cls = inspect(lens)
    check_invariants(lens, info, cls)
    assert info.sigma == 128                                      # nothing to gain from a wider window
    assert info.slots == lens.sum()                               # 100000 = 3125 full slices, equal lengths
    assert info.max_warp_bytes <= 1.15 * info.mean_warp_bytes


def test_poisson_rows_pick_the_narrow_window():
    """Config 2's K^T: the planner accepts ~10 % padding for the epilogue locality of a 128-row window."""
    lp = synthetic.uniform_rows_lp(8_000, 2_000, 20_000, 50, seed=0)
    lens = np.diff(lp.K.T.tocsr().indptr)
    info, cls = inspect(lens)
    check_invariants(lens, info, cls)
    assert info.sigma == 128
    assert info.slots <= 1.15 * lens.sum()


def test_power_law_rows_get_a_wide_window_and_three_classes():
    lp = synthetic.powerlaw_rows_lp(40_000, 10_000, 100_000, max_nnz=20_000, seed=1, name="pl")
    lens = np.diff(lp.K.indptr)
    info, cls = inspect(lens)
    check_invariants(lens, info, cls)
    assert info.cta_rows >= 1 and info.warp_rows >= 1
    assert info.sigma >= 4096                                     # heavy tail: padding dominates the cost model
    lane_nnz = lens[cls == 0].sum()
    warp_slots = (np.ceil(lens[cls == 1] / 32) * 32).sum()
    assert info.slots - warp_slots <= 1.10 * lane_nnz             # lane rows padded by < 10 %
    # a warp's range is cut where rows start: it may overshoot its share by one chain, at most a row of 8192 entries
    assert info.max_warp_bytes <= 1.25 * info.mean_warp_bytes + 8192 * 12 + 4096


def test_wide_windows_deal_slices_out_longest_first(monkeypatch):
    """Config 4's K at 1/4 scale (power-law rows up to 100 000 entries): contiguous per-warp ranges can only be cut
    where rows start, so a warp overshoots its share by up to one slice of 256 steps (98 KB) -- on the GPU the kernel
    then runs as long as its most loaded warp (config 4: K2 433 us against K1's 297 us for the same bytes).  With a wide
    sorting window (no epilogue locality to keep) the planner deals the slices out longest-first instead: every
    invariant holds and the longest range is within a few percent of the mean; PDLP_NO_LPT=1 restores the ranges."""
    lp = synthetic.powerlaw_rows_lp(400_000, 100_000, 1_000_000, seed=0, name="pl4")
    lens = np.diff(lp.K.indptr)
    monkeypatch.setenv("PDLP_NO_LPT", "1")
    before, cls = inspect(lens)
    check_invariants(lens, before, cls)
    monkeypatch.delenv("PDLP_NO_LPT")
    after, cls = inspect(lens)
    check_invariants(lens, after, cls)
    assert before.sigma >= 1024 and after.sigma == before.sigma
    assert before.max_warp_bytes > 1.2 * before.mean_warp_bytes          # the problem is there ...
    # ... and gone: within a few percent of the mean, or one indivisible unit (a slice of 256 steps x 32 lanes / a warp
    # row of 8192 entries, 98 KB + headers) where the mean itself is below that (this 1/4-scale matrix)
    assert after.max_warp_bytes <= max(1.06 * after.mean_warp_bytes, 256 * 32 * 12 + 4096)
    assert after.max_warp_bytes < 0.8 * before.max_warp_bytes
    assert after.slots <= 1.01 * before.slots                            # same padding (merging of short slices kept)
    # uniform rows keep their contiguous ranges (narrow window: consecutive slices share epilogue cache lines)
    u0, _ = inspect(np.full(100_000, 50))
    monkeypatch.setenv("PDLP_NO_LPT", "1")
    u1, _ = inspect(np.full(100_000, 50))
    assert (u0.pieces, u0.max_warp_bytes, u0.packed_bytes) == (u1.pieces, u1.max_warp_bytes, u1.packed_bytes)


@pytest.mark.parametrize("lens", [[], [0], [0, 0, 0], [1], [257], [8193], [5, 0, 300, 0, 9000, 1, 256, 8192]])
def test_edge_cases(lens):
    info, cls = inspect(lens)
    check_invariants(lens, info, cls)


def test_fp32_plan_and_small_machines():
    rng = np.random.default_rng(0)
    lens = rng.poisson(12, size=5_000)
    for vb, sms in ((4, 148), (8, 4), (8, 1)):
        info, cls = inspect(lens, vb, sms)
        check_invariants(lens, info, cls) if sms == 148 else None
        assert info.covered_once == 1 and info.ranges_ok == 1 and info.grid <= sms


def test_bad_arguments_are_rejected():
    rowptr = np.array([0, 5, 3], dtype=np.int32)                  # not monotone
    info = _lib.PlanInfo()
    lib = _lib.lib()
    assert lib.pdlp_plan_inspect(rowptr.ctypes.data, 2, 8, 148, C.byref(info), None) != 0
    assert b"monotone" in lib.pdlp_last_error()
    assert lib.pdlp_plan_inspect(rowptr.ctypes.data, 2, 3, 148, C.byref(info), None) != 0
