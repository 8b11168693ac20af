"""Batch data model and the residual / Jacobian assembly entry points.

Call surface of /root/reference/src/fe_jax/fea.py:
  ElementBatch :18-73, MaterialPropertyArrayType / QuadratureArrayType :77-85,
  ElementBatchCollection :88-345, batch_to_collection :348-487, _calculate_jacobian_unique_nnz :490-516,
  calculate_jacobian_wo_dirichlet :613-660, calculate_jacobian_diag_wo_dirichlet :759-797,
  calculate_residual_wo_dirichlet :850-919, calculate_residual_w_dirichlet :922-965.
Device work goes through libfeb200 (include/feb200.h); arrays are torch CUDA float64 tensors.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from enum import Enum
from typing import Any, Callable

import numpy as np
import torch

from . import _lib
from .containers import COO, NodePattern
from .mesh import as_device, mesh_to_element_coords
from .physics import resolve_model, resolve_residual
from .tables import FiniteElementType, eval_basis_and_derivatives, get_quadrature


@dataclass
class ElementBatch:
    """fea.py:18-73 (same fields, same shape checks)."""
    fe_type: FiniteElementType
    n_dofs_per_basis: int
    connectivity_en: Any
    constitutive_model: Callable
    material_params: Any
    internal_state: Any = None

    def __post_init__(self):
        Q = get_quadrature(fe_type=self.fe_type)[0].shape[0]
        E = self.connectivity_en.shape[0]
        mp = self.material_params
        if len(mp.shape) == 2:
            assert mp.shape[0] == E, (f"`material_params` had dimension 2, which means the shape should be (E, M). However, "
                                      f"`connectivity_en.shape[0]` ({E}) did not match `material_params.shape[0]` ({mp.shape[0]})")
        elif len(mp.shape) == 3:
            assert mp.shape[0] == E, (f"`material_params` had dimension 3, which means the shape should be (E, Q, M). However, "
                                      f"`connectivity_en.shape[0]` ({E}) did not match `material_params.shape[0]` ({mp.shape[0]})")
            assert mp.shape[1] == Q, (f"`material_params` had dimension 3, which means the shape should be (E, Q, M). However, "
                                      f"`fe_type` results in Q = {Q}, which did not match `material_params.shape[1]` ({mp.shape[1]})")
        st = self.internal_state
        if st is not None:
            if len(st.shape) in (2, 3):
                assert st.shape[0] == E, f"`internal_state.shape[0]` ({st.shape[0]}) did not match the number of elements ({E})"
            if len(st.shape) == 3:
                assert st.shape[1] == Q, f"`internal_state.shape[1]` ({st.shape[1]}) did not match Q = {Q}"


class MaterialPropertyArrayType(Enum):
    EQM = 3
    EM = 2
    M = 1


class QuadratureArrayType(Enum):
    EQ = 2
    Q = 1


_LAYOUT = {MaterialPropertyArrayType.M: _lib.LAYOUT_M, MaterialPropertyArrayType.EM: _lib.LAYOUT_EM,
           MaterialPropertyArrayType.EQM: _lib.LAYOUT_EQM}


def _full_weights():
    return os.environ.get("FEB200_FULL_WEIGHTS", "0") == "1"


@dataclass
class ElementBatchCollection:
    """fea.py:88-345: static shape tuples + flattened arrays with offset tables + getters."""
    D: int
    B: int
    E: tuple
    N: tuple
    U: tuple
    Q: tuple
    P: tuple
    M: tuple
    I: tuple
    connectivity: torch.Tensor
    material_params: torch.Tensor
    internal_state: torch.Tensor
    xi: torch.Tensor
    weights: torch.Tensor
    phi: torch.Tensor
    dphi_dxi: torch.Tensor
    constitutive_models: tuple
    EN_offsets: tuple
    material_params_types: list
    material_params_offsets: tuple
    material_params_sizes: tuple
    internal_state_offsets: tuple
    internal_state_sizes: tuple
    quadrature_types: tuple
    xi_offsets: tuple
    weights_offsets: tuple
    phi_offsets: tuple
    dphi_dxi_offsets: tuple
    # device-side extras (not on the reference surface)
    _vertices: torch.Tensor = None
    _conn32: list = field(default_factory=list)
    _W_applied: list = field(default_factory=list)
    _x: torch.Tensor = None
    _plans: dict = field(default_factory=dict)
    _scratch: dict = field(default_factory=dict)

    @property
    def x(self):
        """Unravelled element-major coordinates, shape (sum(E*N*D),) -- built on first access."""
        if self._x is None:
            self._x = torch.cat([mesh_to_element_coords(self._vertices, self._conn32[i]).reshape(-1)
                                 for i in range(self.B)])
        return self._x

    def get_connectivity(self, i):
        o = self.EN_offsets[i]
        return self.connectivity[o:o + self.E[i] * self.N[i]].reshape(self.E[i], self.N[i])

    def get_material_params(self, i):
        o = self.material_params_offsets[i]
        t = self.material_params_types[i]
        if t == MaterialPropertyArrayType.EQM:
            return self.material_params[o:o + self.E[i] * self.Q[i] * self.M[i]].reshape(self.E[i], self.Q[i], self.M[i])
        if t == MaterialPropertyArrayType.EM:
            return self.material_params[o:o + self.E[i] * self.M[i]].reshape(self.E[i], self.M[i])
        return self.material_params[o:o + self.M[i]].reshape(self.M[i])

    def get_internal_state(self, i):
        o = self.internal_state_offsets[i]
        return self.internal_state[o:o + self.E[i] * self.Q[i] * self.I[i]].reshape(self.E[i], self.Q[i], self.I[i])

    def get_x(self, i):
        o = self.D * self.EN_offsets[i]
        return self.x[o:o + self.E[i] * self.N[i] * self.D].reshape(self.E[i], self.N[i], self.D)

    def get_weights(self, i):
        """Length-1 slice for the (only used) QuadratureArrayType.Q case, exactly as fea.py:285-297."""
        o = self.weights_offsets[i]
        n = self.E[i] if self.quadrature_types[i] == QuadratureArrayType.EQ else 1
        return self.weights[o:o + n]

    def get_dphi_dxi(self, i):
        o = self.dphi_dxi_offsets[i]
        return self.dphi_dxi[o:o + self.Q[i] * self.N[i] * self.P[i]].reshape(self.Q[i], self.N[i], self.P[i])

    def get_dof_map(self, i):
        """dof[e, n*U + j] = U*conn[e, n] + j, int64 (fea.py:325-345)."""
        c = self.get_connectivity(i)
        U = self.U[i]
        return (U * c[:, :, None] + torch.arange(U, device=c.device)[None, None, :]).reshape(self.E[i], self.N[i] * U)

    # ---- kernel-side description of batch i
    def _batch_struct(self, i, kind):
        model = resolve_model(self.constitutive_models[i])
        if model.kind != kind:
            raise _lib.Feb200Error(f"constitutive model {model.__name__} cannot be used with this element residual")
        if self.U[i] != self.D:
            raise _lib.Feb200Error(f"batch {i}: n_dofs_per_basis={self.U[i]} but the element kernels write D={self.D} components "
                                   "per node (the reference assumes U == D too, fea.py:550)")
        mp = self.get_material_params(i)
        need = model.n_params(self.D)
        if self.M[i] < need:
            raise _lib.Feb200Error(f"{model.__name__} needs {need} material parameters in {self.D}-D, got {self.M[i]}")
        dphi = self.get_dphi_dxi(i)
        b = _lib.Batch(E=self.E[i], N=self.N[i], Q=self.Q[i], D=self.D, kind=kind, material=model.material,
                       layout=_LAYOUT[self.material_params_types[i]], M=self.M[i],
                       conn=_lib.ptr(self._conn32[i]), dphi=_lib.ptr(dphi), W=_lib.ptr(self._W_applied[i]),
                       mp=_lib.ptr(mp))
        b._keep = (mp, dphi)
        return b

    def _buffer(self, key, numel):
        t = self._scratch.get(key)
        if t is None or t.numel() < numel:
            t = torch.empty(int(numel), dtype=torch.float64, device=self._vertices.device)
            self._scratch[key] = t
        return t


def _cumsum0(sizes):
    out = [0]
    for s in sizes:
        out.append(out[-1] + int(s))
    return tuple(out)


def batch_to_collection(vertices_vd, element_batches) -> ElementBatchCollection:
    """fea.py:348-487."""
    vertices = as_device(vertices_vd, torch.float64)
    dev = vertices.device
    E = tuple(int(b.connectivity_en.shape[0]) for b in element_batches)
    N = tuple(int(b.connectivity_en.shape[1]) for b in element_batches)
    U = tuple(int(b.n_dofs_per_basis) for b in element_batches)
    quad = [get_quadrature(fe_type=b.fe_type) for b in element_batches]
    tabs = [eval_basis_and_derivatives(fe_type=b.fe_type, xi_qp=q[0]) for b, q in zip(element_batches, quad)]
    Q = tuple(int(q[0].shape[0]) for q in quad)
    P = tuple(int(q[0].shape[-1]) for q in quad)
    M = tuple(int(b.material_params.shape[-1]) for b in element_batches)
    I = tuple(int(b.internal_state.shape[-1]) if b.internal_state is not None else 0 for b in element_batches)
    D = int(vertices.shape[1])
    for n, t in zip(N, tabs):
        assert n == t[1].shape[1], f"Number of nodes per element {n} must match the number of basis functions {t[1].shape[1]}."

    conn64 = [as_device(b.connectivity_en, torch.int64) for b in element_batches]
    mps = [as_device(b.material_params, torch.float64) for b in element_batches]
    states = [as_device(b.internal_state, torch.float64).reshape(-1) if b.internal_state is not None
              else torch.zeros(E[i] * Q[i] * I[i], dtype=torch.float64, device=dev)
              for i, b in enumerate(element_batches)]
    f64 = lambda arrs: torch.cat([torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).reshape(-1) for a in arrs]).to(dev)
    state_sizes = tuple(int(b.internal_state.size if isinstance(b.internal_state, np.ndarray) else b.internal_state.numel())
                        if b.internal_state is not None else 0 for b in element_batches)
    ebc = ElementBatchCollection(
        D=D, B=len(element_batches), E=E, N=N, U=U, Q=Q, P=P, M=M, I=I,
        connectivity=torch.cat([c.reshape(-1) for c in conn64]),
        material_params=torch.cat([m.reshape(-1) for m in mps]),
        internal_state=torch.cat(states) if states else torch.zeros(0, dtype=torch.float64, device=dev),
        xi=f64([q[0] for q in quad]), weights=f64([q[1] for q in quad]),
        phi=f64([t[0] for t in tabs]), dphi_dxi=f64([t[1] for t in tabs]),
        constitutive_models=tuple(b.constitutive_model for b in element_batches),
        EN_offsets=_cumsum0([e * n for e, n in zip(E, N)]),
        material_params_types=[MaterialPropertyArrayType(len(b.material_params.shape)) for b in element_batches],
        material_params_offsets=_cumsum0([m.numel() for m in mps]),
        material_params_sizes=tuple(int(m.numel()) for m in mps),
        internal_state_offsets=_cumsum0(state_sizes),
        internal_state_sizes=state_sizes,
        quadrature_types=tuple(QuadratureArrayType.Q for _ in element_batches),
        xi_offsets=_cumsum0([q[0].size for q in quad]),
        weights_offsets=_cumsum0([q[1].size for q in quad]),
        phi_offsets=_cumsum0([t[0].size for t in tabs]),
        dphi_dxi_offsets=_cumsum0([t[1].size for t in tabs]),
    )
    ebc._vertices = vertices
    ebc._conn32 = [c.to(torch.int32).contiguous() for c in conn64]
    # weights as the kernels apply them: the reference hands W[0:1] to the element function and einsum
    # broadcasts it (fea.py:285-297, linear_elasticity.py:203).  FEB200_FULL_WEIGHTS=1 uses the full rule.
    for q in quad:
        W = np.asarray(q[1], dtype=np.float64)
        Wa = W if _full_weights() else np.full_like(W, W[0])
        ebc._W_applied.append(torch.from_numpy(Wa).to(dev))
    return ebc


# --------------------------------------------------------------------------------------------- sparsity plan
class _Plan:
    """Everything derived once from (ebc, assembly maps): node pattern + per-batch slot-rank maps."""

    def __init__(self, ebc: ElementBatchCollection, assembly_map_b):
        lib = _lib.load()
        dev = ebc._vertices.device
        V = int(ebc._vertices.shape[0])
        U = ebc.U[0]
        self.maps = (_lib.BatchMap * ebc.B)()
        self._keep = []
        for i, am in enumerate(assembly_map_b):
            row = am.indptr.reshape(-1).view(torch.int64)
            col = am.indices.reshape(-1).view(torch.int64)
            self._keep.append((row, col))
            self.maps[i] = _lib.BatchMap(E=ebc.E[i], N=ebc.N[i], conn=_lib.ptr(ebc._conn32[i]), row_ind=_lib.ptr(row),
                                         col_ind=_lib.ptr(col))
        deg = torch.empty(V, dtype=torch.int32, device=dev)
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.call("feb200_pattern_count", self.maps, ebc.B, V, _lib.ptr(deg), _lib.ptr(err))
        nindptr = torch.empty(V + 1, dtype=torch.int64, device=dev)
        ws = _lib.workspace(lib.feb200_scan_workspace_bytes(V), dev)
        _lib.call("feb200_pattern_scan", _lib.ptr(deg), V, _lib.ptr(nindptr), _lib.ptr(ws), ws.numel())
        if int(err.item()) != 0:
            raise _lib.Feb200Error("a vertex has more neighbours than FEB200_MAX_NODE_DEGREE")
        nnz_nodes = int(nindptr[-1].item())
        max_deg = int(deg.max().item()) if V > 0 else 1
        nindices = torch.empty(nnz_nodes, dtype=torch.int32, device=dev)
        diag_rank = torch.empty(V, dtype=torch.int32, device=dev)
        _lib.call("feb200_pattern_fill", self.maps, ebc.B, V, _lib.ptr(nindptr), _lib.ptr(nindices), _lib.ptr(diag_rank))
        self.pattern = NodePattern(V, U, nindptr, nindices, diag_rank, max(max_deg, 1))
        self.slot_rank = []
        for i in range(ebc.B):
            sr = torch.empty(ebc.E[i] * ebc.N[i] * ebc.N[i], dtype=torch.int16, device=dev)
            _lib.call("feb200_slot_rank", _lib.ptr(ebc._conn32[i]), ebc.E[i], ebc.N[i], _lib.ptr(nindptr),
                      _lib.ptr(nindices), _lib.ptr(sr))
            self.slot_rank.append(sr)
        self._fused = {}
        self._sweep = {}

    def sweep_plan(self, ebc, i):
        """Streaming plan of feb200_assemble_fused_hex8_sweep for batch i (None when the mesh is not eligible)."""
        if i not in self._sweep:
            plan = None
            if (os.environ.get("FEB200_FUSED", "1") != "0" and ebc.N[i] == 8 and ebc.Q[i] == 8
                    and ebc.D == 3 and self.pattern.max_deg <= 32):
                from .sweep_plan import build_sweep_plan
                row, col = self._keep[i]
                plan = build_sweep_plan(ebc._vertices, ebc._conn32[i], row, col, self.pattern.nindptr)
            if plan is not None:
                plan.rank8 = self.slot_rank[i].to(torch.uint8).contiguous()      # ranks < 32: one byte per node pair for the kernel
            self._sweep[i] = plan
        return self._sweep[i]

    def fused_plan(self, ebc, i):
        """Tile plan of the fused hex8 kernel for batch i (None when the batch / mesh is not eligible)."""
        if i not in self._fused:
            plan = None
            if (os.environ.get("FEB200_FUSED", "1") != "0" and ebc.B == 1 and ebc.N[i] == 8 and ebc.Q[i] == 8
                    and ebc.D == 3 and self.pattern.max_deg <= 32):
                from .fused_plan import build_fused_plan
                row, col = self._keep[i]
                plan = build_fused_plan(ebc._vertices, ebc._conn32[i], row, col, self.pattern.nindptr)
            self._fused[i] = plan
        return self._fused[i]


def _plan(ebc, assembly_map_b) -> _Plan:
    key = tuple(id(m) for m in assembly_map_b)
    p = ebc._plans.get(key)
    if p is None:
        assert (np.array(ebc.U) == ebc.U[0]).all(), "The number of DoFs per a point (U) must be the same across all batches."
        p = _Plan(ebc, assembly_map_b)
        p._maps_ref = list(assembly_map_b)      # keeps the ids in `key` alive
        ebc._plans[key] = p
    return p


def _default_maps(ebc):
    """assembly maps built from the collection itself (used when the caller has none at hand)"""
    maps = ebc._scratch.get("_default_maps")
    if maps is None:
        from .mesh import mesh_to_sparse_assembly_map
        V = int(ebc._vertices.shape[0])
        maps = [mesh_to_sparse_assembly_map(V, ebc.get_connectivity(i)) for i in range(ebc.B)]
        ebc._scratch["_default_maps"] = maps
    return maps


def _calculate_jacobian_unique_nnz(n_vertices: int, ebc: ElementBatchCollection):
    """Number of structural non-zeros of the assembled Jacobian (fea.py:490-516)."""
    return _plan(ebc, _default_maps(ebc)).pattern.nnz


# --------------------------------------------------------------------------------------------- residual
def _as_u(ebc, u_f):
    u = as_device(u_f, torch.float64).reshape(-1)
    F = int(ebc._vertices.shape[0]) * ebc.U[0]
    assert u.shape[0] == F, f"solution vector has {u.shape[0]} entries, expected {F}"
    return u


def _residual(element_residual_func, ebc, assembly_map_b, u, mask, values, fuse_dirichlet=True):
    res = resolve_residual(element_residual_func)
    F = u.shape[0]
    R = torch.empty(F, dtype=torch.float64, device=u.device)
    states = []
    for i in range(ebc.B):
        b = ebc._batch_struct(i, res.kind)
        R_e = ebc._buffer("R_e", ebc.E[i] * ebc.N[i] * ebc.U[i])
        _lib.call("feb200_element_residual", b, _lib.ptr(ebc._vertices), _lib.ptr(u), _lib.ptr(mask), _lib.ptr(values),
                  _lib.ptr(R_e))
        am = assembly_map_b[i]
        row = am.indptr.reshape(-1).view(torch.int64)
        col = am.indices.reshape(-1).view(torch.int64)
        last = (i == ebc.B - 1) and mask is not None and fuse_dirichlet
        _lib.call("feb200_scatter_sum", _lib.ptr(row), _lib.ptr(col), am.shape[1], ebc.U[i], _lib.ptr(R_e), _lib.ptr(R),
                  1 if i > 0 else 0, _lib.ptr(mask) if last else None, _lib.ptr(values) if last else None,
                  _lib.ptr(u) if last else None)
        states.append(torch.zeros((ebc.E[i], ebc.Q[i], 0), dtype=torch.float64, device=u.device))
    return R, states


def calculate_residual_wo_dirichlet(element_residual_func, ebc: ElementBatchCollection, assembly_map_b, u_f):
    """R_f (F,), [state_b] without any Dirichlet modification (fea.py:850-919)."""
    return _residual(element_residual_func, ebc, assembly_map_b, _as_u(ebc, u_f), None, None)


def calculate_residual_w_dirichlet(element_residual_func, ebc: ElementBatchCollection, assembly_map_b, u_f,
                                   dirichlet_values_g, dirichlet_mask_g):
    """u_w = (1-m) u + m g ; R = (1-m) R(u_w) + m (u - g)   (fea.py:922-965).  The blend is applied while
    the element kernel gathers u, the second formula inside the last scatter."""
    u = _as_u(ebc, u_f)
    g = as_device(dirichlet_values_g, torch.float64)
    m = as_device(dirichlet_mask_g, torch.float64)
    return _residual(element_residual_func, ebc, assembly_map_b, u, m, g)


# --------------------------------------------------------------------------------------------- Jacobian
def _assemble_values(element_residual_func, ebc, assembly_map_b, u, mask=None, values=None, out=None, eliminate=False):
    """CSR values (reference order) of J(u) [or J(u_w) when a mask is given] on the node pattern.  With
    `eliminate` the Dirichlet elimination (apply_dirichlet_bcs_lhs) is fused into the last batch's store."""
    res = resolve_residual(element_residual_func)
    plan = _plan(ebc, assembly_map_b)
    pat = plan.pattern
    vals = out if out is not None else torch.empty(pat.nnz, dtype=torch.float64, device=u.device)
    # Several batches (fea.py:885-889 sums them): the values start from zero, every batch accumulates -- the streaming
    # fused kernel for its hex8 batches, the two-pass path for the rest -- and the Dirichlet elimination runs once on the
    # finished sums.  One batch: the kernels overwrite and fuse the elimination into their store.
    multi = ebc.B > 1
    if multi:
        vals.zero_()
    for i in range(ebc.B):
        b = ebc._batch_struct(i, res.kind)
        # fused paths: linear isotropic hex8 (any material layout; the tangent does not depend on u).
        # Streaming kernel when the batch has one Poisson ratio, else (single batch) the tile kernels.
        if (b.kind == _lib.KIND_LINEAR and b.material == _lib.MAT_ISOTROPIC and _use_sweep_kernel(ebc, i, b)
                and plan.sweep_plan(ebc, i) is not None):
            flags = _block_flags(ebc, pat, mask) if (eliminate and mask is not None and not multi) else None
            sp = plan.sweep_plan(ebc, i)
            _lib.call("feb200_assemble_fused_hex8_sweep", b, sp.struct, _lib.ptr(ebc._vertices), _lib.ptr(sp.rank8),
                      _lib.ptr(flags), _material_hints(ebc, i, b), _lib.ptr(vals), 1 if multi else 0)
            continue
        if (not multi and b.kind == _lib.KIND_LINEAR and b.material == _lib.MAT_ISOTROPIC
                and plan.fused_plan(ebc, i) is not None):
            fp = plan.fused_plan(ebc, i)
            flags = _vertex_flags(ebc, mask) if (eliminate and mask is not None) else None
            if _use_tensor_core_kernel(ebc, i, b, fp):
                _lib.call("feb200_assemble_fused_hex8_mma", b, fp.struct, _lib.ptr(ebc._vertices),
                          _lib.ptr(plan.slot_rank[i]), _lib.ptr(pat.nindices), _lib.ptr(flags),
                          _material_hints(ebc, i, b), _lib.ptr(vals))
                continue
            _lib.call("feb200_assemble_fused_hex8", b, fp.struct, _lib.ptr(ebc._vertices), _lib.ptr(plan._keep[i][0]),
                      _lib.ptr(plan.slot_rank[i]), _lib.ptr(pat.nindptr), _lib.ptr(pat.nindices),
                      _lib.ptr(flags), _lib.ptr(vals))
            continue
        K_e = ebc._buffer("K_e", ebc.E[i] * ebc.N[i] * ebc.N[i] * ebc.U[i] * ebc.U[i])
        _lib.call("feb200_element_tangent", b, _lib.ptr(ebc._vertices), _lib.ptr(u), _lib.ptr(mask), _lib.ptr(values),
                  _lib.ptr(K_e))
        bm = plan.maps[i]
        elim = eliminate and mask is not None and not multi
        # fused elimination from the per-vertex flag bytes (one byte per CSR block instead of U mask doubles)
        _lib.call("feb200_assemble_csr_flags", bm, pat.V, pat.U, _lib.ptr(K_e), _lib.ptr(plan.slot_rank[i]),
                  _lib.ptr(pat.nindptr), pat.max_deg, _lib.ptr(vals), 1 if multi else 0,
                  _lib.ptr(pat.nindices) if elim else None, _lib.ptr(_vertex_flags(ebc, mask)) if elim else None)
    if multi and eliminate and mask is not None:
        _lib.call("feb200_dirichlet_lhs", pat.V, pat.U, _lib.ptr(pat.nindptr), _lib.ptr(pat.nindices), _lib.ptr(mask), _lib.ptr(vals))
    return vals, pat


def _poisson_facts(ebc, i):
    """(constant within every element, identical everywhere) for the Poisson ratio of batch i; one reduction per
    material array, cached on the collection and redone when the array is replaced or written to."""
    cache = ebc.__dict__.setdefault("_nu_facts", {})
    mp = ebc.get_material_params(i)
    key = (i, mp.data_ptr(), mp._version)
    if key not in cache:
        cache.clear()
        nu = mp[..., 1]
        if nu.dim() == 0:
            cache[key] = (True, True)
        else:
            per_el = True if nu.dim() == 1 else bool((nu.max(dim=1).values == nu.min(dim=1).values).all().item())
            cache[key] = (per_el, bool((nu.max() == nu.min()).item()))
    return cache[key]


def _material_hints(ebc, i, b) -> int:
    if b.layout == _lib.LAYOUT_M:
        return 0
    return _lib.HINT_UNIFORM_POISSON if _poisson_facts(ebc, i)[1] else 0


def _use_sweep_kernel(ebc, i, b) -> bool:
    """feb200_assemble_fused_hex8_sweep (the default) needs one Poisson ratio for the batch: layout (M,), or a material
    array whose Poisson ratio is the same everywhere (checked once per array, cached).  FEB200_FUSED_IMPL=mma / v4 select
    the tile kernels instead."""
    if os.environ.get("FEB200_FUSED_IMPL", "sweep") != "sweep":
        return False
    if os.environ.get("FEB200_FUSED", "1") == "0" or not (ebc.N[i] == 8 and ebc.Q[i] == 8 and ebc.D == 3):
        return False
    return b.layout == _lib.LAYOUT_M or _poisson_facts(ebc, i)[1]


def _use_tensor_core_kernel(ebc, i, b, fp) -> bool:
    """feb200_assemble_fused_hex8_mma handles materials constant per batch or per element, and per-point fields whose
    Poisson ratio is constant within every element (one check per material array, cached on the collection);
    anything else, or FEB200_FUSED_IMPL=v4, goes to the scalar fused kernel."""
    if os.environ.get("FEB200_FUSED_IMPL", "sweep") == "v4":
        return False
    if b.layout != _lib.LAYOUT_EQM:
        return True
    if fp.stats.get("smem_bytes_mma_per_point", 1 << 30) > 227 * 1024:
        return False
    return _poisson_facts(ebc, i)[0]


def _vertex_flags(ebc, mask):
    """uint8 per vertex, bit j set when DoF U*v+j is constrained (input of the fused Dirichlet elimination).
    Cached on the mask tensor object itself (invalidated by in-place modification through `_version`)."""
    hit = getattr(mask, "_feb200_vflags", None)
    if hit is not None and hit[0] == mask._version:
        return hit[1]
    U = ebc.U[0]
    bits = torch.tensor([1 << j for j in range(U)], dtype=torch.int32, device=mask.device)
    flags = ((mask.view(-1, U) != 0).to(torch.int32) * bits).sum(dim=1).to(torch.uint8).contiguous()
    try:
        mask._feb200_vflags = (mask._version, flags)
    except AttributeError:
        pass
    return flags


def _block_flags(ebc, pat, mask):
    """(V, 32) uint8 Dirichlet flags per CSR block row (input of the streaming kernel's fused elimination), derived from
    the per-vertex flags and the node pattern; cached on the mask tensor like `_vertex_flags`."""
    hit = getattr(mask, "_feb200_bflags", None)
    if hit is not None and hit[0] == mask._version and hit[1] is pat:
        return hit[2]
    vf = _vertex_flags(ebc, mask)
    out = torch.empty((pat.V, 32), dtype=torch.uint8, device=mask.device)
    _lib.call("feb200_sweep_block_flags", pat.V, _lib.ptr(pat.nindptr), _lib.ptr(pat.nindices), _lib.ptr(vf), _lib.ptr(out))
    try:
        mask._feb200_bflags = (mask._version, pat, out)
    except AttributeError:
        pass
    return out


def calculate_jacobian_wo_dirichlet(element_residual_func, ebc: ElementBatchCollection, assembly_map_b, u_f,
                                    precomputed_jacobian_nnz: int):
    """Sorted-unique COO Jacobian (rows_sorted) with `precomputed_jacobian_nnz` entries (fea.py:613-660)."""
    u = _as_u(ebc, u_f)
    vals, pat = _assemble_values(element_residual_func, ebc, assembly_map_b, u)
    if precomputed_jacobian_nnz is not None and int(precomputed_jacobian_nnz) != pat.nnz:
        raise _lib.Feb200Error(f"precomputed_jacobian_nnz={precomputed_jacobian_nnz} but the pattern has {pat.nnz} entries")
    F = u.shape[0]
    return COO((vals,), shape=(F, F), rows_sorted=True, pattern=pat)


def calculate_jacobian_w_dirichlet(element_residual_func, ebc: ElementBatchCollection, assembly_map_b, u_f,
                                   dirichlet_values_g, dirichlet_mask_g, precomputed_jacobian_nnz: int = None, out=None):
    """J_w = (I-M) J(u_w) (I-M) + M with u_w = (1-m) u + m g, as a sorted-unique COO on the node pattern: the Jacobian
    builder that goes with calculate_residual_w_dirichlet.  It is the operator the reference's CG applies as a JVP of
    R_w (sparse_linear_solve.py:251-258) and equals apply_dirichlet_bcs_lhs(calculate_jacobian_wo_dirichlet(u_w))
    (sparse_matrix.py:157-184); here the elimination is fused into the assembly store.  `out` (nnz doubles) is reused
    for the values when given."""
    u = _as_u(ebc, u_f)
    g = as_device(dirichlet_values_g, torch.float64)
    m = as_device(dirichlet_mask_g, torch.float64)
    vals, pat = _assemble_values(element_residual_func, ebc, assembly_map_b, u, m, g, out=out, eliminate=True)
    if precomputed_jacobian_nnz is not None and int(precomputed_jacobian_nnz) != pat.nnz:
        raise _lib.Feb200Error(f"precomputed_jacobian_nnz={precomputed_jacobian_nnz} but the pattern has {pat.nnz} entries")
    F = u.shape[0]
    return COO((vals,), shape=(F, F), rows_sorted=True, pattern=pat)


def calculate_jacobian_diag_wo_dirichlet(element_residual_func, ebc: ElementBatchCollection, assembly_map_b, u_f):
    """diag(J) (F,) (fea.py:759-797).  Linear isotropic batches: element diagonals from a diagonal-only kernel, scatter-summed
    in ascending slot order (no matrix is assembled); anything else: read off the deterministically assembled CSR."""
    u = _as_u(ebc, u_f)
    res = resolve_residual(element_residual_func)
    structs = [ebc._batch_struct(i, res.kind) for i in range(ebc.B)]
    if all(b.kind == _lib.KIND_LINEAR and b.material == _lib.MAT_ISOTROPIC for b in structs):
        diag = torch.empty(u.shape[0], dtype=torch.float64, device=u.device)
        for i, b in enumerate(structs):
            d_e = ebc._buffer("R_e", ebc.E[i] * ebc.N[i] * ebc.U[i])
            _lib.call("feb200_element_tangent_diag", b, _lib.ptr(ebc._vertices), _lib.ptr(d_e))
            am = assembly_map_b[i]
            row = am.indptr.reshape(-1).view(torch.int64)
            col = am.indices.reshape(-1).view(torch.int64)
            _lib.call("feb200_scatter_sum", _lib.ptr(row), _lib.ptr(col), am.shape[1], ebc.U[i], _lib.ptr(d_e), _lib.ptr(diag),
                      1 if i > 0 else 0, None, None, None)
        return diag
    vals, pat = _assemble_values(element_residual_func, ebc, assembly_map_b, u)
    return csr_diagonal(pat, vals)


def csr_diagonal(pat: NodePattern, vals):
    diag = torch.empty(pat.U * pat.V, dtype=torch.float64, device=vals.device)
    _lib.call("feb200_csr_diagonal", pat.V, pat.U, _lib.ptr(pat.nindptr), _lib.ptr(pat.diag_rank), _lib.ptr(vals),
              _lib.ptr(diag))
    return diag
