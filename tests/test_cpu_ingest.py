"""LP ingest (torchpdlp_b200.engine.to_device_csr) on the CPU device: every accepted layout of K gives the same
CSR(K) and CSR(K^T) with sorted columns; a canonical CSR input takes the no-sort fast path, a CSR with unsorted
columns or duplicate entries is coalesced like the reference's COO tensors are."""
import numpy as np
import scipy.sparse as sp
import torch

from torchpdlp_b200.engine import to_device_csr


def as_scipy(csr, shape):
    return sp.csr_matrix((csr[2].numpy(), csr[1].numpy(), csr[0].numpy()), shape=shape)


def test_all_layouts_give_the_same_sorted_csr_pair():
    A = sp.random(50, 70, density=0.1, format="csr", random_state=np.random.default_rng(0))
    A.sort_indices()
    csr_t = torch.sparse_csr_tensor(torch.from_numpy(A.indptr.astype(np.int64)), torch.from_numpy(A.indices.astype(np.int64)),
                                    torch.from_numpy(A.data), size=A.shape)
    dense = torch.from_numpy(A.toarray())
    tup = (A.indptr, A.indices, A.data, A.shape)
    for K in (csr_t, dense, dense.to_sparse_coo(), A, tup):
        (m, n, nnz), csr, csc = to_device_csr(K, torch.device("cpu"), torch.float64)
        B, Bt = as_scipy(csr, (m, n)), as_scipy(csc, (n, m))
        assert (m, n, nnz) == (50, 70, A.nnz)
        assert abs(B - A).max() == 0 and abs(Bt - A.T).max() == 0
        for M in (B, Bt):                                      # columns strictly increasing inside every row
            for r in range(M.shape[0]):
                seg = M.indices[M.indptr[r]:M.indptr[r + 1]]
                assert np.all(np.diff(seg) > 0)
        assert csr[0].dtype == torch.int32 and csr[1].dtype == torch.int32 and csr[2].dtype == torch.float64


def test_non_canonical_csr_is_coalesced():
    crow, col, val = torch.tensor([0, 3, 4]), torch.tensor([2, 0, 2, 1]), torch.tensor([1.0, 2.0, 3.0, 4.0])
    (m, n, nnz), csr, csc = to_device_csr(torch.sparse_csr_tensor(crow, col, val, size=(2, 3)), torch.device("cpu"), torch.float64)
    assert nnz == 3 and csr[0].tolist() == [0, 2, 3] and csr[1].tolist() == [0, 2, 1] and csr[2].tolist() == [2.0, 4.0, 4.0]
    assert csc[0].tolist() == [0, 1, 2, 3] and csc[1].tolist() == [0, 1, 0] and csc[2].tolist() == [2.0, 4.0, 4.0]


def test_float32_and_empty_rows():
    A = sp.csr_matrix(np.array([[0, 0, 0], [1.5, 0, 2.5], [0, 0, 0], [0, 3.5, 0]]))
    (m, n, nnz), csr, csc = to_device_csr(A, torch.device("cpu"), torch.float32)
    assert csr[2].dtype == torch.float32 and csr[0].tolist() == [0, 0, 2, 2, 3] and csc[0].tolist() == [0, 1, 2, 3]
