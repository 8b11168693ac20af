"""Light stand-ins for the `jax.experimental.sparse` containers on the fe_jax call surface.

JAX is not available in this image, so the drop-in functions return these objects, which expose the
same attribute names (`COO.data/.row/.col/.shape/._rows_sorted`, `CSR.data/.indices/.indptr/.shape`,
`BCSR.data/.indices/.indptr/.shape` with the leading batch dimension 1) holding torch CUDA tensors.
Where JAX exists the FFI shim returns the real containers.
"""
from __future__ import annotations

import torch


class BCSR:
    """Assembly map (1, V, E*N): what mesh_to_sparse_assembly_map returns
    (/root/reference/src/fe_jax/setup.py:250-285)."""

    def __init__(self, args, shape, indices_sorted=True, unique_indices=True):
        self.data, self.indices, self.indptr = args
        self.shape = tuple(shape)
        self.indices_sorted, self.unique_indices = indices_sorted, unique_indices
        # device-side extras used by the kernels (not part of the reference surface)
        self._cells32 = None   # (E, N) int32
        self._N = None

    @property
    def nse(self):
        return self.indices.shape[-1]


class NodePattern:
    """Node-level sparsity shared by every Jacobian of one problem (see include/feb200.h):
    nindptr (V+1) int64, nindices int32, diag_rank (V) int32, max_deg."""

    def __init__(self, V, U, nindptr, nindices, diag_rank, max_deg):
        self.V, self.U = V, U
        self.nindptr, self.nindices, self.diag_rank, self.max_deg = nindptr, nindices, diag_rank, max_deg
        self.nnz_nodes = int(nindices.shape[0])
        self.nnz = U * U * self.nnz_nodes
        self._row = self._col = self._indptr = None

    def _expand(self):
        from . import _lib
        if self._row is None:
            dev = self.nindices.device
            self._row = torch.empty(self.nnz, dtype=torch.int64, device=dev)
            self._col = torch.empty(self.nnz, dtype=torch.int64, device=dev)
            self._indptr = torch.empty(self.U * self.V + 1, dtype=torch.int64, device=dev)
            _lib.call("feb200_pattern_expand", self.V, self.U, _lib.ptr(self.nindptr), _lib.ptr(self.nindices),
                      _lib.ptr(self._row), _lib.ptr(self._col), _lib.ptr(self._indptr))
        return self._row, self._col, self._indptr


class COO:
    """Sorted-unique COO Jacobian (what calculate_jacobian_wo_dirichlet returns, fea.py:613-660).
    `row` / `col` (int64) are materialised on first access when the matrix lives on a NodePattern."""

    def __init__(self, args, shape, rows_sorted=False, cols_sorted=False, pattern: NodePattern | None = None):
        self.data = args[0]
        self._row_arr, self._col_arr = (args[1], args[2]) if len(args) == 3 else (None, None)
        self.shape = tuple(shape)
        self._rows_sorted, self._cols_sorted = rows_sorted, cols_sorted
        self._pattern = pattern

    @property
    def row(self):
        if self._row_arr is None:
            self._row_arr = self._pattern._expand()[0]
        return self._row_arr

    @property
    def col(self):
        if self._col_arr is None:
            self._col_arr = self._pattern._expand()[1]
        return self._col_arr

    @property
    def nse(self):
        return self.data.shape[0]

    def todense(self):
        out = torch.zeros(self.shape, dtype=self.data.dtype, device=self.data.device)
        out.index_put_((self.row, self.col), self.data, accumulate=True)
        return out

    def _sort_indices(self):
        from .sparse import _lexsort_rc
        perm = _lexsort_rc(self.row, self.col)
        return COO((self.data[perm], self.row[perm], self.col[perm]), self.shape, rows_sorted=True)

    def __matmul__(self, v):
        from .sparse import coo_matvec
        return coo_matvec(self, v)


class CSR:
    """What coo_to_csr returns (sparse_matrix.py:113-153)."""

    def __init__(self, args, shape, pattern: NodePattern | None = None):
        self.data = args[0]
        self._indices, self._indptr = (args[1], args[2]) if len(args) == 3 else (None, None)
        self.shape = tuple(shape)
        self._pattern = pattern

    @property
    def indices(self):
        if self._indices is None:
            self._indices = self._pattern._expand()[1]
        return self._indices

    @property
    def indptr(self):
        if self._indptr is None:
            self._indptr = self._pattern._expand()[2]
        return self._indptr

    def __matmul__(self, v):
        from .sparse import csr_matvec
        return csr_matvec(self, v)
