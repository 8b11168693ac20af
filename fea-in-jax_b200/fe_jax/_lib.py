"""ctypes binding of libfeb200.so (the C ABI declared in include/feb200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, an
exception is raised.  PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_int32, c_int64, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FEB200_LIB", os.path.join(_HERE, "..", "lib", "libfeb200.so"))

FEB200_MAX_BATCHES = 16
KIND_LINEAR, KIND_HYPER = 0, 1
MAT_ISOTROPIC, MAT_ORTHOTROPIC, MAT_SVK, MAT_MOONEY_RIVLIN = 0, 1, 2, 3
LAYOUT_M, LAYOUT_EM, LAYOUT_EQM = 0, 1, 2
HINT_UNIFORM_POISSON = 1
CG_INIT_SPMV, CG_INIT_VECTORS, CG_INIT_FINALIZE, CG_SPMV, CG_PAP, CG_UPDATE_XR, CG_UPDATE_P = range(7)


class Feb200Error(RuntimeError):
    pass


class Batch(Structure):
    _fields_ = [("E", c_int64), ("N", c_int32), ("Q", c_int32), ("D", c_int32), ("kind", c_int32),
                ("material", c_int32), ("layout", c_int32), ("M", c_int32), ("reserved", c_int32),
                ("conn", c_void_p), ("dphi", c_void_p), ("W", c_void_p), ("mp", c_void_p)]


class BatchMap(Structure):
    _fields_ = [("E", c_int64), ("N", c_int32), ("reserved", c_int32),
                ("conn", c_void_p), ("row_ind", c_void_p), ("col_ind", c_void_p)]


class FusedPlan(Structure):
    _fields_ = [("n_tiles", c_int32), ("tn_cap", c_int32), ("te_cap", c_int32), ("reserved", c_int32),
                ("tile_node_ptr", c_void_p), ("tile_nodes", c_void_p), ("tile_elem_ptr", c_void_p),
                ("tile_elems", c_void_p), ("tile_item_ptr", c_void_p), ("tile_items", c_void_p),
                ("tile_node_p0", c_void_p), ("tile_node_meta", c_void_p)]


class SweepPlan(Structure):
    _fields_ = [("n_segments", c_int32), ("max_deg", c_int32), ("reserved0", c_int32), ("reserved1", c_int32),
                ("seg_visit_ptr", c_void_p), ("seg_chunk_ptr", c_void_p), ("visits", c_void_p), ("chunk_rec", c_void_p)]


MAX_NEIGHBOURS = 8


class Halo(Structure):
    _fields_ = [("n_neighbours", c_int32), ("U", c_int32), ("peer", c_int32 * MAX_NEIGHBOURS), ("reserved", c_int32 * 2),
                ("count", c_int64 * MAX_NEIGHBOURS), ("idx", c_void_p * MAX_NEIGHBOURS), ("send", c_void_p * MAX_NEIGHBOURS),
                ("recv", c_void_p * MAX_NEIGHBOURS)]


class CGState(Structure):
    _fields_ = [("V", c_int64), ("U", c_int32), ("maxiter", c_int32),
                ("nindptr", c_void_p), ("nindices", c_void_p), ("values", c_void_p), ("inv_diag", c_void_p),
                ("weights", c_void_p), ("b", c_void_p), ("x", c_void_p), ("r", c_void_p), ("p", c_void_p), ("Ap", c_void_p),
                ("hist", c_void_p), ("scalars", c_void_p), ("counters", c_void_p), ("ws", c_void_p),
                ("ws_bytes", c_size_t)]


_P = c_void_p
_I64, _I32, _F64, _SZ = c_int64, c_int32, c_double, c_size_t

# name -> (restype, argtypes); must list every symbol include/feb200.h declares
SIGNATURES = {
    "feb200_last_error": (c_char_p, []),
    "feb200_version": (c_int32, []),
    "feb200_build_row_ind_workspace_bytes": (_SZ, [_I64]),
    "feb200_build_row_ind": (c_int32, [_P, _P, _I64, _I32, _I64, _P, _P, _SZ]),
    "feb200_build_col_ind_workspace_bytes": (_SZ, [_I64]),
    "feb200_build_col_ind": (c_int32, [_P, _P, _I64, _I32, _I64, _P, _P, _P, _P, _SZ]),
    "feb200_gather": (c_int32, [_P, _P, _I64, _I32, _I32, _P, _P]),
    "feb200_scatter_sum": (c_int32, [_P, _P, _P, _I64, _I32, _P, _P, _I32, _P, _P, _P]),
    "feb200_dirichlet_blend": (c_int32, [_P, _I64, _P, _P, _P, _P]),
    "feb200_dirichlet_residual": (c_int32, [_P, _I64, _P, _P, _P, _P]),
    "feb200_element_residual": (c_int32, [_P, POINTER(Batch), _P, _P, _P, _P, _P]),
    "feb200_element_tangent": (c_int32, [_P, POINTER(Batch), _P, _P, _P, _P, _P]),
    "feb200_element_tangent_diag": (c_int32, [_P, POINTER(Batch), _P, _P]),
    "feb200_pattern_count": (c_int32, [_P, POINTER(BatchMap), _I32, _I64, _P, _P]),
    "feb200_scan_workspace_bytes": (_SZ, [_I64]),
    "feb200_pattern_scan": (c_int32, [_P, _P, _I64, _P, _P, _SZ]),
    "feb200_pattern_fill": (c_int32, [_P, POINTER(BatchMap), _I32, _I64, _P, _P, _P]),
    "feb200_slot_rank": (c_int32, [_P, _P, _I64, _I32, _P, _P, _P]),
    "feb200_pattern_expand": (c_int32, [_P, _I64, _I32, _P, _P, _P, _P, _P]),
    "feb200_assemble_csr": (c_int32, [_P, POINTER(BatchMap), _I64, _I32, _P, _P, _P, _I32, _P, _I32, _P, _P]),
    "feb200_assemble_csr_flags": (c_int32, [_P, POINTER(BatchMap), _I64, _I32, _P, _P, _P, _I32, _P, _I32, _P, _P]),
    "feb200_fused_smem_bytes": (_SZ, [_I32, _I32]),
    "feb200_assemble_fused_hex8": (c_int32, [_P, POINTER(Batch), POINTER(FusedPlan), _P, _P, _P, _P, _P, _P, _P]),
    "feb200_fused_mma_smem_bytes": (_SZ, [_I32, _I32, _I32]),
    "feb200_assemble_fused_hex8_mma": (c_int32, [_P, POINTER(Batch), POINTER(FusedPlan), _P, _P, _P, _P, _I32, _P]),
    "feb200_fused_sweep_warps": (c_int32, [_I32]),
    "feb200_sweep_block_flags": (c_int32, [_P, _I64, _P, _P, _P, _P]),
    "feb200_assemble_fused_hex8_sweep": (c_int32, [_P, POINTER(Batch), POINTER(SweepPlan), _P, _P, _P, _I32, _P, _I32]),
    "feb200_dirichlet_lhs": (c_int32, [_P, _I64, _I32, _P, _P, _P, _P]),
    "feb200_dirichlet_lhs_coo": (c_int32, [_P, _I64, _P, _P, _P, _P, _P]),
    "feb200_csr_diagonal": (c_int32, [_P, _I64, _I32, _P, _P, _P, _P]),
    "feb200_coo_mark_heads": (c_int32, [_P, _I64, _P, _P, _P]),
    "feb200_coo_scan_heads": (c_int32, [_P, _I64, _P, _P, _P, _SZ]),
    "feb200_coo_segment_sum": (c_int32, [_P, _I64, _P, _P, _P, _P, _P, _I64, _P, _P, _P]),
    "feb200_coo_rows_to_indptr": (c_int32, [_P, _I64, _P, _I64, _P]),
    "feb200_reduce_workspace_bytes": (_SZ, []),
    "feb200_spmv": (c_int32, [_P, _I64, _I32, _P, _P, _P, _P, _P, _P, _P, _SZ]),
    "feb200_spmv_csr": (c_int32, [_P, _I64, _P, _P, _P, _P, _P]),
    "feb200_dot": (c_int32, [_P, _I64, _P, _P, _P, _P, _SZ]),
    "feb200_cg_init": (c_int32, [_P, POINTER(CGState), _F64, _F64]),
    "feb200_cg_iterate": (c_int32, [_P, POINTER(CGState), _I32]),
    "feb200_cg_stage": (c_int32, [_P, POINTER(CGState), _I32, _F64, _F64]),
    "feb200_nccl_available": (c_int32, []),
    "feb200_nccl_unique_id": (c_int32, [_P]),
    "feb200_nccl_comm_init": (c_int32, [_I32, _I32, _P, POINTER(c_void_p)]),
    "feb200_nccl_comm_destroy": (c_int32, [_P]),
    "feb200_halo_pack": (c_int32, [_P, POINTER(Halo), _P]),
    "feb200_halo_unpack_add": (c_int32, [_P, POINTER(Halo), _P]),
    "feb200_halo_sum": (c_int32, [_P, _P, POINTER(Halo), _P]),
    "feb200_dot_allreduce": (c_int32, [_P, _P, _P, _I32]),
    "feb200_cg_init_dist": (c_int32, [_P, _P, POINTER(CGState), POINTER(Halo), _F64, _F64]),
    "feb200_cg_iterate_dist": (c_int32, [_P, _P, POINTER(CGState), POINTER(Halo), _I32]),
    "feb200_bcg_init": (c_int32, [_P, _I64, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _F64, _F64, _I32]),
    "feb200_bcg_step": (c_int32, [_P, _I64, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _I32]),
    "feb200_mg_prolong_add": (c_int32, [_P, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "feb200_mg_restrict": (c_int32, [_P, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "feb200_axpby": (c_int32, [_P, _I64, _F64, _P, _F64, _P, _P]),
    "feb200_scatter_set": (c_int32, [_P, _I64, _P, _P, _F64, _P, _P]),
}

_lib = None
launches = 0   # number of C-ABI calls that enqueue kernels (bench.py reports it)
launches_by_name = {}   # the same, per entry point (tests check which kernel ran)


def load():
    """Load libfeb200.so; raises Feb200Error when it has not been built (no CPU fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.abspath(LIB_PATH)
    if not os.path.exists(path):
        raise Feb200Error(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(the fe_jax hot path has no CPU fallback)")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def require_cuda():
    if not torch.cuda.is_available():
        raise Feb200Error("fe_jax (feb200) needs a CUDA device: the hot path has no CPU fallback")


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    """device pointer of a tensor (None -> NULL)"""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "feb200 needs contiguous CUDA tensors"
    return t.data_ptr()


def call(name, *args):
    """Call an int-returning entry point on the current stream and raise on a non-zero status."""
    global launches
    lib = load()
    rc = getattr(lib, name)(stream(), *args)
    launches += 1
    launches_by_name[name] = launches_by_name.get(name, 0) + 1
    if rc != 0:
        msg = lib.feb200_last_error()
        raise Feb200Error(f"{name} failed ({rc}): {msg.decode() if msg else ''}")


def workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 8), dtype=torch.uint8, device=device)
