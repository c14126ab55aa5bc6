"""The run layout (torchpdlp_b200/csrc/run_spmv.cuh) without a GPU: pdlp_run_emulate plans and packs a CSR matrix
with the planner and the per-slot packer the device path uses, walks the chunks with a 32-lane emulation of the
kernel's chunk walk (ballots, staging slots, segmented carry scan, epilogue rounds) and returns A x.  Compared
with scipy on matrices that hit every branch: empty rows, rows spanning lanes / chunks / many chunks, more finished
rows per chunk than one staging round holds, rows on the one-row-per-CTA path, tiny grids, and the shared-memory
panel variant with its 16-bit panel-relative columns."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from torchpdlp_b200 import _lib


def emulate(A, x, mode=0, sms=148, long_row=8192):
    A = sp.csr_matrix(A)
    A.sort_indices()
    rowptr = A.indptr.astype(np.int32)
    colidx = A.indices.astype(np.int32)
    vals = A.data.astype(np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty(A.shape[0], dtype=np.float64)
    info = _lib.RunInfo()
    rc = _lib.lib().pdlp_run_emulate(rowptr.ctypes.data, colidx.ctypes.data, vals.ctypes.data, A.shape[0], A.shape[1],
                                     sms, long_row, mode, x.ctypes.data, out.ctypes.data, C.byref(info))
    return rc, out, info


def check(A, mode=0, sms=148, long_row=8192, seed=0):
    A = sp.csr_matrix(A)
    x = np.random.default_rng(seed).standard_normal(A.shape[1])
    rc, out, info = emulate(A, x, mode, sms, long_row)
    assert rc == 0, _lib.lib().pdlp_last_error()
    want = A @ x
    scale = np.abs(A) @ np.abs(x) + 1e-300
    assert not np.isnan(out).any()                              # every row was finished by exactly the walk
    assert np.max(np.abs(out - want) / scale) < 1e-14
    assert info.rows_finished == A.shape[0]
    return info


def random_rows(lens, cols, seed=0):
    rng = np.random.default_rng(seed)
    lens = np.asarray(lens)
    indptr = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=indptr[1:])
    indices = np.concatenate([rng.choice(cols, size=int(l), replace=False) for l in lens]) if indptr[-1] else np.zeros(0, int)
    data = rng.uniform(0.1, 1.0, size=int(indptr[-1]))
    return sp.csr_matrix((data, indices, indptr), shape=(len(lens), cols))


def test_uniform_rows():
    info = check(random_rows(np.full(3000, 50), 5000))
    assert info.stream_entries == 150_000 and info.long_rows == 0
    assert info.chunks * 256 >= info.stream_entries
    assert info.chunks * 256 <= info.stream_entries + info.warps * 256       # padding only behind each warp's last row
    assert info.chunk_bytes == 256 * 12 + 128 + 16             # values, columns, 32 lane words, trailer


def test_ragged_rows_with_empties_and_single_entries():
    rng = np.random.default_rng(3)
    lens = rng.integers(0, 12, size=5000)
    lens[::7] = 0
    lens[5::11] = 1
    lens[:40] = 0                                               # a warp range that starts with empty rows
    lens[-33:] = 0                                              # ... and one that ends with them
    info = check(random_rows(lens, 4000, seed=4))
    assert info.stream_entries == lens.sum() + (lens == 0).sum()


def test_more_finished_rows_per_chunk_than_a_staging_round():
    """Rows of 1 entry: 256 rows end in every chunk, two staging rounds of 128."""
    check(random_rows(np.ones(20_000, dtype=int), 3000, seed=5))
    lens = np.random.default_rng(6).integers(1, 3, size=20_000)
    check(random_rows(lens, 3000, seed=6))


def test_rows_spanning_many_chunks_and_the_cta_path():
    lens = np.array([5, 3000, 0, 700, 9000, 2, 257, 256, 255, 8192, 8193, 1, 0, 40])
    info = check(random_rows(lens, 20_000, seed=7))
    assert info.long_rows == 2                                  # 9000 and 8193 (> 8192)
    assert info.stream_entries == lens[lens <= 8192].sum() + (lens == 0).sum()   # long rows are not in the stream
    info = check(random_rows(lens, 20_000, seed=7), long_row=100)
    assert info.long_rows == 8


def test_power_law_rows_balance():
    rng = np.random.default_rng(8)
    lens = np.minimum((rng.pareto(1.5, size=40_000) * 4 + 2).astype(int), 6000)
    A = random_rows(lens, 30_000, seed=9)
    info = check(A)
    mean = info.chunks / info.warps
    assert info.max_warp_chunks <= mean + 6000 / 256 + 2        # a warp overshoots its share by at most one row


@pytest.mark.parametrize("sms", [1, 2, 148])
def test_tiny_matrices_and_grids(sms):
    check(random_rows([3], 10), sms=sms)
    check(random_rows([0], 10), sms=sms)
    check(random_rows([0, 0, 0], 10), sms=sms)
    check(random_rows(np.full(20, 10), 10), sms=sms)
    info = check(random_rows(np.full(1000, 20), 2000), sms=sms)  # config-1 sized
    assert 1 <= info.grid <= sms


def test_panels_config2_shape_scaled_down():
    """K^T of a uniform-rows LP: n rows of ~25 entries gathering a vector that needs several 12.5k-element panels."""
    m, n = 30_000, 8_000
    K = random_rows(np.full(m, 6), n, seed=10)                  # K: m x n; K^T rows gather from y (m elements)
    Kt = sp.csr_matrix(K.T)
    info = check(Kt, mode=1)
    assert info.panels == 3 and info.panel_cap <= 12800 + 4
    assert Kt.nnz <= info.stream_entries <= 1.25 * Kt.nnz        # padded slots of the per-panel SELL plans
    assert info.grid == 148 and info.warps == 148 * 12
    assert info.max_warp_chunks >= 128                           # the sorting window the cost model chose


def test_panels_one_panel_and_skewed_columns():
    check(sp.csr_matrix(random_rows(np.full(500, 8), 3000, seed=11).T), mode=1)        # y fits one panel
    # columns with very different counts: the nnz-proportional boundaries would exceed the cap -> equal-width panels
    rng = np.random.default_rng(12)
    rows = 4000
    cols = 40_000
    hot = rng.integers(0, 2000, size=(rows, 20))
    cold = rng.integers(2000, cols, size=(rows, 2))
    idx = np.concatenate([hot, cold], axis=1)
    A = sp.csr_matrix((rng.uniform(0.1, 1, size=idx.size), idx.ravel(), np.arange(0, idx.size + 1, 22)), shape=(rows, cols))
    A.sum_duplicates()
    info = check(A, mode=1)
    assert info.panels >= 4
    # rows with a few hundred entries in one panel (warp rows of the SELL plan) and empty rows
    lens = np.random.default_rng(14).integers(0, 30, size=3000)
    lens[7] = 700
    lens[8] = 300
    check(random_rows(lens, 9000, seed=15), mode=1)


def test_panels_refuse_a_vector_that_does_not_fit():
    A = random_rows(np.full(100, 5), 16 * 12800 + 4000, seed=13)
    rc, _, _ = emulate(A, np.zeros(A.shape[1]), mode=1)
    assert rc != 0 and b"fit" in _lib.lib().pdlp_last_error()
