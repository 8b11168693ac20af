"""Grids, assembly-map builders and the gather / scatter transforms.

Call surface of /root/reference/src/fe_jax/setup.py (same names, argument meaning, dtypes):
uniform_quad_grid :17-55, uniform_tri_grid :58-102, mesh_to_jax :151-186, get_n_cells_per_vert :189-209,
build_row_ind :212-224, build_col_ind_and_data :227-246, mesh_to_sparse_assembly_map :250-285 and the
transform_* family :288-387.  The builders and transforms run on the GPU through libfeb200; grid
generation is host NumPy (setup cost, not on the hot path).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .containers import BCSR


def _dev():
    _lib.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def as_device(a, dtype=None):
    """numpy / torch (any device) -> contiguous CUDA tensor"""
    if isinstance(a, torch.Tensor):
        t = a
    else:
        arr = np.asarray(a)
        if arr.dtype == np.uint64:
            arr = arr.astype(np.int64)
        t = torch.from_numpy(np.ascontiguousarray(arr))
    if t.dtype == torch.uint64:
        t = t.view(torch.int64)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.to(_dev()).contiguous()


# ------------------------------------------------------------------------------------------ grids
def uniform_quad_grid(n_rows: int, n_cols: int, bbox):
    """Vertices (n_rows*n_cols, 3) and int64 quads (basix tensor order); node id = i*n_cols + j,
    x follows i and y follows j (setup.py:41-53)."""
    bbox = np.asarray(bbox, dtype=np.float64)
    x = np.arange(n_rows) / float(n_rows - 1) * (bbox[0, 1] - bbox[0, 0]) + bbox[0, 0]
    y = np.arange(n_cols) / float(n_cols - 1) * (bbox[1, 1] - bbox[1, 0]) + bbox[1, 0]
    V = np.zeros((n_rows * n_cols, 3), dtype=np.float64)
    V[:, 0] = np.repeat(x, n_cols)
    V[:, 1] = np.tile(y, n_rows)
    base = (np.arange(n_rows - 1)[:, None] * n_cols + np.arange(n_cols - 1)[None, :]).ravel()
    F = np.stack([base, base + n_cols, base + 1, base + n_cols + 1], axis=1).astype(np.int64)
    return V, F


def uniform_tri_grid(n_rows: int, n_cols: int):
    """Two triangles per cell on the unit square (setup.py:58-102)."""
    V, Fq = uniform_quad_grid(n_rows, n_cols, [[0.0, 1.0], [0.0, 1.0]])
    F = np.empty((2 * Fq.shape[0], 3), dtype=np.int64)
    F[0::2] = Fq[:, [0, 1, 2]]
    F[1::2] = Fq[:, [1, 3, 2]]
    return V, F


def uniform_hex_grid(nx: int, ny: int, nz: int, bbox=((0.0, 1.0), (0.0, 1.0), (0.0, 1.0))):
    """3-D extension of uniform_quad_grid's convention (nx, ny, nz = ELEMENTS per axis): node id =
    (i*(ny+1) + j)*(nz+1) + k with x following i; hex8 connectivity in basix tensor order."""
    bbox = np.asarray(bbox, dtype=np.float64)
    axes = [np.arange(n + 1) / float(n) * (bbox[a, 1] - bbox[a, 0]) + bbox[a, 0] for a, n in enumerate((nx, ny, nz))]
    X, Y, Z = np.meshgrid(*axes, indexing="ij")
    V = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    sy, sx = nz + 1, (ny + 1) * (nz + 1)
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    base = (i * sx + j * sy + k).ravel()
    offs = np.array([(c & 1) * sx + ((c >> 1) & 1) * sy + ((c >> 2) & 1) for c in range(8)])
    return V, (base[:, None] + offs[None, :]).astype(np.int64)


# ------------------------------------------------------------------------------------------ maps
def _cells_device(cells):
    c = as_device(cells, torch.int64)
    assert c.ndim == 2, "cells must be (E, N)"
    return c


def build_row_ind(n_vertices: int, cells):
    """uint64 (V+1,) exclusive scan of the vertex valence (setup.py:212-224)."""
    c = _cells_device(cells)
    E, N = c.shape
    lib = _lib.load()
    row = torch.empty(n_vertices + 1, dtype=torch.int64, device=c.device)
    ws = _lib.workspace(lib.feb200_build_row_ind_workspace_bytes(n_vertices), c.device)
    _lib.call("feb200_build_row_ind", _lib.ptr(c), E, N, n_vertices, _lib.ptr(row), _lib.ptr(ws), ws.numel())
    return row.view(torch.uint64)


def build_col_ind_and_data(csr_row_ind, cells):
    """(uint64 (E*N,), float64 (E*N,)): per vertex its slots e*N+n ascending, data = 1 (setup.py:227-246)."""
    c = _cells_device(cells)
    E, N = c.shape
    row = as_device(csr_row_ind, torch.int64)
    V = row.shape[0] - 1
    lib = _lib.load()
    col = torch.empty(E * N, dtype=torch.int64, device=c.device)
    data = torch.empty(E * N, dtype=torch.float64, device=c.device)
    ws = _lib.workspace(lib.feb200_build_col_ind_workspace_bytes(V), c.device)
    _lib.call("feb200_build_col_ind", _lib.ptr(c), E, N, V, _lib.ptr(row), _lib.ptr(col), _lib.ptr(data),
              _lib.ptr(ws), ws.numel())
    return col.view(torch.uint64), data


def mesh_to_sparse_assembly_map(n_vertices: int, cells) -> BCSR:
    """BCSR (1, V, E*N) map between assembled vectors and the element-node layout (setup.py:250-285)."""
    c = _cells_device(cells)
    row = build_row_ind(n_vertices, c)
    col, data = build_col_ind_and_data(row, c)
    m = BCSR((data.reshape(1, -1), col.reshape(1, -1), row.reshape(1, -1)),
             shape=(1, n_vertices, c.shape[0] * c.shape[1]), indices_sorted=True, unique_indices=True)
    m._cells32 = c.to(torch.int32).contiguous()
    m._N = c.shape[1]
    return m


def mesh_to_jax(vertices, cells):
    """Element-major coordinates (E, N, 2) -- the reference keeps two columns (setup.py:156-160);
    all D columns of `vertices` are kept when it has D == 3 *and* a non-trivial third column is
    requested through `mesh_to_element_coords`."""
    return mesh_to_element_coords(np.asarray(vertices)[:, 0:2] if not isinstance(vertices, torch.Tensor)
                                  else vertices[:, 0:2], cells)


def mesh_to_element_coords(vertices_vd, cells):
    """x_end[e, n, :] = vertices[cells[e, n], :] for any D (gather kernel)."""
    v = as_device(vertices_vd, torch.float64)
    c = _cells_device(cells).to(torch.int32).contiguous()
    E, N = c.shape
    D = v.shape[1]
    out = torch.empty((E, N, D), dtype=torch.float64, device=v.device)
    _lib.call("feb200_gather", _lib.ptr(c), E, N, D, _lib.ptr(v), _lib.ptr(out))
    return out


def get_n_cells_per_vert(vertices, cells):
    """Number of cells touching each vertex (setup.py:189-209), uint64."""
    n = as_device(vertices).shape[0] if not isinstance(vertices, int) else vertices
    row = build_row_ind(n, cells).view(torch.int64)
    return (row[1:] - row[:-1]).view(torch.uint64)


# ------------------------------------------------------------------------------------------ transforms
def _map_parts(assembly_map: BCSR):
    assert assembly_map.shape[0] == 1
    row = assembly_map.indptr.reshape(-1).view(torch.int64)
    col = assembly_map.indices.reshape(-1).view(torch.int64)
    return row, col


def transform_global_unraveled_to_element_node(assembly_map: BCSR, v_g, E: int):
    """(F,) -> (E, N, U) gather (setup.py:304-322)."""
    v = as_device(v_g, torch.float64)
    V = assembly_map.shape[1]
    U = v.shape[0] // V
    N = assembly_map.shape[2] // E
    if assembly_map._cells32 is None:
        raise _lib.Feb200Error("assembly map was not built by mesh_to_sparse_assembly_map")
    out = torch.empty((E, N, U), dtype=torch.float64, device=v.device)
    _lib.call("feb200_gather", _lib.ptr(assembly_map._cells32), E, N, U, _lib.ptr(v), _lib.ptr(out))
    return out


def transform_global_to_element_node(assembly_map: BCSR, v_g, E: int, V: int, U: int):
    """(V, U) -> (E, N, U) (setup.py:288-301)."""
    return transform_global_unraveled_to_element_node(assembly_map, as_device(v_g, torch.float64).reshape(-1), E)


def _scatter(assembly_map: BCSR, v_en, out=None, accumulate=False, mask=None, values=None, u=None):
    v = as_device(v_en, torch.float64)
    U = v.shape[-1]
    V = assembly_map.shape[1]
    row, col = _map_parts(assembly_map)
    if out is None:
        out = torch.empty(V * U, dtype=torch.float64, device=v.device)
    _lib.call("feb200_scatter_sum", _lib.ptr(row), _lib.ptr(col), V, U, _lib.ptr(v), _lib.ptr(out),
              1 if accumulate else 0, _lib.ptr(mask), _lib.ptr(values), _lib.ptr(u))
    return out


def transform_element_node_to_global_unraveled_sum(assembly_map: BCSR, v_en):
    """(E, N, U) -> (F,) row-ordered segmented sum (setup.py:344-356)."""
    return _scatter(assembly_map, v_en)


def transform_element_node_to_global_sum(assembly_map: BCSR, v_en):
    """(E, N, U) -> (V, U) (setup.py:358-370)."""
    U = v_en.shape[-1]
    return _scatter(assembly_map, v_en).reshape(-1, U)


def _valence(assembly_map: BCSR):
    row, _ = _map_parts(assembly_map)
    return (row[1:] - row[:-1]).to(torch.float64)


def transform_element_node_to_global_unraveled_nosum(assembly_map: BCSR, v_en):
    """Scatter-mean, unravelled (setup.py:325-341)."""
    U = v_en.shape[-1]
    s = _scatter(assembly_map, v_en).reshape(-1, U)
    return (s / _valence(assembly_map)[:, None]).reshape(-1)


def transform_element_node_to_global_nosum(assembly_map: BCSR, v_en):
    """Scatter-mean (V, U) (setup.py:372-387)."""
    U = v_en.shape[-1]
    s = _scatter(assembly_map, v_en).reshape(-1, U)
    return s / _valence(assembly_map)[:, None]
