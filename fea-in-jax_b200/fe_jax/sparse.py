"""COO / CSR helpers of the fe_jax call surface (/root/reference/src/fe_jax/sparse_matrix.py).

coo_arrays_sum_duplicates :10-41, coo_arrays_sum_duplicates_jit :44-92, coo_sum_duplicates :95-110,
coo_to_csr :113-153, apply_dirichlet_bcs_lhs :157-184, apply_dirichlet_bcs_rhs :187-202.
Matrices that live on a NodePattern (everything the assembly produces) never materialise their
int64 index arrays unless asked; generic COO input goes through the segmented-sum kernels.
"""
from __future__ import annotations

import torch

from . import _lib
from .containers import COO, CSR, NodePattern
from .mesh import as_device


def _lexsort_rc(row, col):
    """permutation sorting by (row, col), stable -- jnp.lexsort((col, row)); index plumbing via torch.sort"""
    p1 = torch.sort(col, stable=True).indices
    p2 = torch.sort(row[p1], stable=True).indices
    return p1[p2]


def coo_arrays_sum_duplicates_jit(A: COO, result_length: int):
    """(data, row, col) of the row-then-column sorted matrix with duplicates summed; like the reference
    it expects `A` already sorted (fea.py:651-654 sorts first).  sparse_matrix.py:44-92."""
    lib = _lib.load()
    row, col, data = A.row.contiguous(), A.col.contiguous(), A.data.contiguous()
    n = int(data.shape[0])
    dev = data.device
    head = torch.empty(n, dtype=torch.int32, device=dev)
    seg = torch.empty(n, dtype=torch.int64, device=dev)
    _lib.call("feb200_coo_mark_heads", n, _lib.ptr(row), _lib.ptr(col), _lib.ptr(head))
    ws = _lib.workspace(lib.feb200_scan_workspace_bytes(n), dev)
    _lib.call("feb200_coo_scan_heads", n, _lib.ptr(head), _lib.ptr(seg), _lib.ptr(ws), ws.numel())
    if result_length is None or result_length <= 0:
        result_length = int(seg[-1].item()) + 1 if n > 0 else 0
    out_d = torch.zeros(result_length, dtype=torch.float64, device=dev)
    out_r = torch.zeros(result_length, dtype=torch.int64, device=dev)
    out_c = torch.zeros(result_length, dtype=torch.int64, device=dev)
    _lib.call("feb200_coo_segment_sum", n, _lib.ptr(row), _lib.ptr(col), _lib.ptr(data), _lib.ptr(head), _lib.ptr(seg),
              result_length, _lib.ptr(out_d), _lib.ptr(out_r), _lib.ptr(out_c))
    return out_d, out_r, out_c


def coo_arrays_sum_duplicates(A: COO):
    """Dynamic-length variant (sparse_matrix.py:10-41): sorts, then sums duplicates."""
    perm = _lexsort_rc(A.row, A.col)
    B = COO((A.data[perm].contiguous(), A.row[perm].contiguous(), A.col[perm].contiguous()), A.shape, rows_sorted=True)
    return coo_arrays_sum_duplicates_jit(B, 0)


def coo_sum_duplicates(A: COO, result_length: int) -> COO:
    """sparse_matrix.py:95-110."""
    if A._pattern is not None:           # already sorted-unique by construction
        return A
    d, r, c = coo_arrays_sum_duplicates_jit(A, result_length)
    return COO((d, r, c), shape=A.shape, rows_sorted=True)


def coo_to_csr(A: COO) -> CSR:
    """indptr = [0, cumsum(bincount(rows))], indices = cols (sparse_matrix.py:113-153; the stray
    debug print of indptr at :151 is not reproduced)."""
    if A._pattern is not None:
        return CSR((A.data,), shape=A.shape, pattern=A._pattern)
    data, row, col = A.data, A.row, A.col
    if not A._rows_sorted:
        perm = _lexsort_rc(row, col)
        data, row, col = data[perm].contiguous(), row[perm].contiguous(), col[perm].contiguous()
    n_rows = A.shape[0]
    indptr = torch.empty(n_rows + 1, dtype=torch.int64, device=data.device)
    _lib.call("feb200_coo_rows_to_indptr", int(row.shape[0]), _lib.ptr(row.contiguous()), n_rows, _lib.ptr(indptr))
    return CSR((data, col, indptr), shape=A.shape)


def _mask_from_dofs(n, dirichlet_dofs, device):
    m = torch.zeros(n, dtype=torch.float64, device=device)
    m[as_device(dirichlet_dofs, torch.int64)] = 1.0
    return m


def apply_dirichlet_bcs_lhs(A: COO, dirichlet_dofs) -> COO:
    """Same sparsity, constrained rows / columns zeroed, ones on their diagonals (sparse_matrix.py:157-184)."""
    mask = _mask_from_dofs(A.shape[0], dirichlet_dofs, A.data.device)
    if A._pattern is not None:
        pat = A._pattern
        vals = A.data.clone()
        _lib.call("feb200_dirichlet_lhs", pat.V, pat.U, _lib.ptr(pat.nindptr), _lib.ptr(pat.nindices), _lib.ptr(mask),
                  _lib.ptr(vals))
        return COO((vals,), shape=A.shape, rows_sorted=A._rows_sorted, cols_sorted=A._cols_sorted, pattern=pat)
    out = torch.empty_like(A.data)
    _lib.call("feb200_dirichlet_lhs_coo", int(A.data.shape[0]), _lib.ptr(A.row.contiguous()), _lib.ptr(A.col.contiguous()),
              _lib.ptr(mask), _lib.ptr(A.data.contiguous()), _lib.ptr(out))
    return COO((out, A.row, A.col), shape=A.shape, rows_sorted=A._rows_sorted, cols_sorted=A._cols_sorted)


def coo_matvec(A: COO, v):
    if A._pattern is not None:
        return pattern_matvec(A._pattern, A.data, v)
    return csr_matvec(coo_to_csr(A), v)


def pattern_matvec(pat: NodePattern, vals, v):
    x = as_device(v, torch.float64)
    y = torch.empty_like(x)
    _lib.call("feb200_spmv", pat.V, pat.U, _lib.ptr(pat.nindptr), _lib.ptr(pat.nindices), _lib.ptr(vals), _lib.ptr(x),
              _lib.ptr(y), None, None, 0)
    return y


def csr_matvec(A: CSR, v):
    if A._pattern is not None:
        return pattern_matvec(A._pattern, A.data, v)
    x = as_device(v, torch.float64)
    y = torch.empty(A.shape[0], dtype=torch.float64, device=x.device)
    _lib.call("feb200_spmv_csr", A.shape[0], _lib.ptr(A.indptr.contiguous()), _lib.ptr(A.indices.contiguous()),
              _lib.ptr(A.data.contiguous()), _lib.ptr(x), _lib.ptr(y))
    return y


def apply_dirichlet_bcs_rhs(A: COO, b, dirichlet_dofs, dirichlet_values):
    """b - A @ lift, then b[dofs] = values (sparse_matrix.py:187-202)."""
    b = as_device(b, torch.float64)
    dofs = as_device(dirichlet_dofs, torch.int64)
    vals = as_device(dirichlet_values, torch.float64)
    tmp = torch.zeros_like(b)
    tmp[dofs] = vals
    out = b - coo_matvec(A, tmp)
    out[dofs] = vals
    return out
