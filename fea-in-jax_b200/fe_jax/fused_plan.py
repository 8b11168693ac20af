"""Tile plan for the fused hex8 assembly kernel (csrc/fused.cu, feb200_assemble_fused_hex8).

Built once per mesh on the device with torch index operations (setup plumbing, not on the hot path):
vertices are sorted along a Morton curve of their coordinates and cut into tiles of at most `tn_cap`
vertices that share a Morton cell; each tile lists the elements touching it and one packed item per
(vertex, incident slot).  Works for any hex8 connectivity whose vertex valence is <= 8; returns None when
the mesh is not eligible (the caller then uses the two-pass path).
"""
from __future__ import annotations

import torch

from . import _lib

TN_CAP = 32
MAXVAL = 8
SMEM_LIMIT = 227 * 1024


def _spread3(v):
    """insert two zero bits between the low 21 bits of v (int64)"""
    v = v & 0x1FFFFF
    v = (v | (v << 32)) & 0x1F00000000FFFF
    v = (v | (v << 16)) & 0x1F0000FF0000FF
    v = (v | (v << 8)) & 0x100F00F00F00F00F
    v = (v | (v << 4)) & 0x10C30C30C30C30C3
    v = (v | (v << 2)) & 0x1249249249249249
    return v


def morton_codes(vertices: torch.Tensor) -> torch.Tensor:
    """Morton code of every vertex on a grid whose cell is the mean vertex spacing."""
    V, D = vertices.shape
    lo = vertices.min(dim=0).values
    hi = vertices.max(dim=0).values
    ext = torch.clamp(hi - lo, min=1e-300)
    vol = float(torch.prod(ext).item())
    cell = (vol / max(V, 1)) ** (1.0 / D)
    # vertices per axis ~ ext / cell; spacing = ext / (count - 1), exact for uniform grids (jitter < half a cell)
    counts = torch.clamp(torch.round(ext / cell), min=2.0)
    spacing = ext / (counts - 1.0)
    q = torch.floor((vertices - lo) / spacing + 0.5).to(torch.int64).clamp_(0, (1 << 21) - 1)
    code = torch.zeros(V, dtype=torch.int64, device=vertices.device)
    for d in range(D):
        code |= _spread3(q[:, d]) << d
    return code


class FusedPlan:
    def __init__(self, struct, keep, n_tiles, tn_cap, te_cap, stats):
        self.struct, self._keep = struct, keep
        self.n_tiles, self.tn_cap, self.te_cap, self.stats = n_tiles, tn_cap, te_cap, stats


def build_fused_plan(vertices, conn32, row_ind, col_ind, nindptr=None):
    """Largest tile size whose tiles fit in shared memory: 32 (4x4x2 Morton cells), else 16, else 8 vertices."""
    import os
    first = int(os.environ.get("FEB200_FUSED_TN", "32"))
    for tn_cap, shift in ((32, 5), (16, 4), (8, 3)):
        if tn_cap > first:
            continue
        plan = _build(vertices, conn32, row_ind, col_ind, tn_cap, shift, nindptr)
        if plan is not None:
            return plan
    return None


def _build(vertices: torch.Tensor, conn32: torch.Tensor, row_ind: torch.Tensor, col_ind: torch.Tensor, tn_cap: int,
           shift: int, nindptr=None):
    """vertices (V,3) f64, conn32 (E,8) i32, row_ind (V+1) / col_ind (E*8) int64 views of the assembly map."""
    lib = _lib.load()
    dev = vertices.device
    V = int(vertices.shape[0])
    E, N = int(conn32.shape[0]), int(conn32.shape[1])
    if N != 8 or vertices.shape[1] != 3 or V == 0 or E == 0:
        return None
    val = row_ind[1:] - row_ind[:-1]
    if int(val.max().item()) > MAXVAL:
        return None
    # ---- tiles: Morton order, cut where the cell prefix (`shift` low bits dropped: 4x4x2 cells for 5) changes or at tn_cap
    code = morton_codes(vertices)
    order = torch.argsort(code, stable=True)
    prefix = (code[order] >> shift)
    new_cell = torch.ones(V, dtype=torch.bool, device=dev)
    new_cell[1:] = prefix[1:] != prefix[:-1]
    cell_id = torch.cumsum(new_cell.to(torch.int64), 0) - 1
    cell_start = torch.nonzero(new_cell).reshape(-1)
    pos_in_cell = torch.arange(V, device=dev) - cell_start[cell_id]
    new_tile = new_cell | (pos_in_cell % tn_cap == 0)
    tile_of_sorted = torch.cumsum(new_tile.to(torch.int64), 0) - 1
    n_tiles = int(tile_of_sorted[-1].item()) + 1
    tile_start = torch.nonzero(new_tile).reshape(-1)
    tile_node_ptr = torch.cat([tile_start, torch.tensor([V], device=dev)]).to(torch.int32)
    tile_of_node = torch.empty(V, dtype=torch.int64, device=dev)
    tile_of_node[order] = tile_of_sorted
    a_loc_of_node = torch.empty(V, dtype=torch.int64, device=dev)
    a_loc_of_node[order] = torch.arange(V, device=dev) - tile_start[tile_of_sorted]
    # ---- (vertex, slot) entries in assembly-map order
    n_ent = int(col_ind.shape[0])
    node_of_entry = torch.repeat_interleave(torch.arange(V, device=dev), val)
    k_of_entry = torch.arange(n_ent, device=dev) - row_ind[node_of_entry]
    elem = torch.div(col_ind, N, rounding_mode="floor")
    n_a = col_ind - elem * N
    tile = tile_of_node[node_of_entry]
    key = tile * E + elem
    ukey, inverse = torch.unique(key, sorted=True, return_inverse=True)
    tile_of_u = torch.div(ukey, E, rounding_mode="floor")
    tile_elems = (ukey - tile_of_u * E).to(torch.int32)
    tile_elem_ptr = torch.searchsorted(tile_of_u, torch.arange(n_tiles + 1, device=dev)).to(torch.int32)
    te_cap = int((tile_elem_ptr[1:] - tile_elem_ptr[:-1]).max().item())
    le = inverse - tile_elem_ptr[tile].to(torch.int64)
    # both kernels that consume the plan must fit: fused.cu (per-point material) and fused_mma.cu (the default)
    if te_cap > 96 or max(int(lib.feb200_fused_smem_bytes(tn_cap, te_cap)),
                          int(lib.feb200_fused_mma_smem_bytes(tn_cap, te_cap, _lib.LAYOUT_EM))) > SMEM_LIMIT:
        return None
    packed = (le | (n_a << 12) | (a_loc_of_node[node_of_entry] << 16) | (k_of_entry << 24))
    perm = torch.argsort(inverse * N + n_a, stable=True)           # tile-major, then element, then local node
    tile_items = packed[perm].to(torch.int32)                       # bit pattern of the uint32 the kernel unpacks
    tile_item_ptr = torch.searchsorted(tile[perm].contiguous(), torch.arange(n_tiles + 1, device=dev)).to(torch.int32)
    tile_nodes = order.to(torch.int32)
    keep = [t.contiguous() for t in (tile_node_ptr, tile_nodes, tile_elem_ptr, tile_elems, tile_item_ptr, tile_items)]
    if not vertices.is_cuda:        # host-side inspection only (tests): arrays without the kernel-side struct
        return FusedPlan(None, keep, n_tiles, tn_cap, te_cap, dict(n_tiles=n_tiles, te_cap=te_cap))
    deg = nindptr[1:] - nindptr[:-1]
    if int(deg.max().item()) > 32:
        return None
    keep.append(nindptr[:-1][order].contiguous())                                     # tile_node_p0
    keep.append((deg[order] | (val[order] << 8)).to(torch.int32).contiguous())        # tile_node_meta
    struct = _lib.FusedPlan(n_tiles=n_tiles, tn_cap=tn_cap, te_cap=te_cap, tile_node_ptr=_lib.ptr(keep[0]),
                            tile_nodes=_lib.ptr(keep[1]), tile_elem_ptr=_lib.ptr(keep[2]), tile_elems=_lib.ptr(keep[3]),
                            tile_item_ptr=_lib.ptr(keep[4]), tile_items=_lib.ptr(keep[5]),
                            tile_node_p0=_lib.ptr(keep[6]), tile_node_meta=_lib.ptr(keep[7]))
    stats = dict(n_tiles=n_tiles, te_cap=te_cap, mean_nodes_per_tile=V / n_tiles,
                 element_visits_per_element=float(tile_elems.shape[0]) / E,
                 smem_bytes=int(lib.feb200_fused_smem_bytes(tn_cap, te_cap)),
                 smem_bytes_mma=int(lib.feb200_fused_mma_smem_bytes(tn_cap, te_cap, _lib.LAYOUT_EM)),
                 smem_bytes_mma_per_point=int(lib.feb200_fused_mma_smem_bytes(tn_cap, te_cap, _lib.LAYOUT_EQM)))
    return FusedPlan(struct, keep, n_tiles, tn_cap, te_cap, stats)
