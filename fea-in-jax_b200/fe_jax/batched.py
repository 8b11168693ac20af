"""Many small boundary-value problems on one mesh.

The reference runs `vmap` over thousands of microscale RVE solves that share a mesh and differ in their
material fields and Dirichlet values (/root/reference/tests/test_microscale_bvp.py,
tests/test_multiscale_prnn.py).  First cut here: the copies are laid side by side as ONE block-diagonal
problem (vertices and connectivity replicated with an offset, material parameters concatenated along the
element axis), so the whole batch goes through the same assembly kernels and one CSR.  The linear solves run as a
batched CG (csrc/krylov.cu k_bcg_*: one SpMV over all copies + one launch with a CTA per copy per iteration), so every
copy has its own alpha / beta, stopping test and iteration count, as under the reference's vmap; the Newton loop is
shared (it runs until every copy meets the nonlinear tolerances; converged copies receive zero updates).
`solve_bvp_batched.last_iterations` holds the per-copy CG iteration counts of each Newton iteration.
"""
from __future__ import annotations

import numpy as np
import torch

from .assembly import ElementBatch
from .newton import solve_bvp
from .solver import SolverOptions


def _np(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


def solve_bvp_batched(vertices_vd, element_batches, element_residual_func, dirichlet_bcs, dirichlet_values_b,
                      material_params_b, u_0_bg=None, solver_options: SolverOptions = SolverOptions()):
    """Solve `n_copies` BVPs that share `vertices_vd`, the connectivity of `element_batches` and `dirichlet_bcs`.

    material_params_b : list (one per element batch) of arrays with a leading copy axis, (C, E, M) or (C, E, Q, M)
    dirichlet_values_b: (C, n_bc)
    Returns (u (C, V*D), residual (C, V*D)).
    """
    V, D = int(vertices_vd.shape[0]), int(vertices_vd.shape[1])
    vals = _np(dirichlet_values_b)
    C = vals.shape[0]
    verts = np.tile(_np(vertices_vd), (C, 1))
    bcs = _np(dirichlet_bcs).astype(np.int64)
    offs = (np.arange(C, dtype=np.int64) * V)
    bcs_all = np.concatenate([np.stack([bcs[:, 0] + o, bcs[:, 1]], axis=1) for o in offs])
    batches = []
    for b, mp in zip(element_batches, material_params_b):
        mp = _np(mp)
        assert mp.shape[0] == C and mp.ndim in (3, 4), "material_params_b entries must be (C, E, M) or (C, E, Q, M)"
        conn = _np(b.connectivity_en).astype(np.int64)
        conn_all = np.concatenate([conn + o for o in offs])
        batches.append(ElementBatch(fe_type=b.fe_type, n_dofs_per_basis=b.n_dofs_per_basis, connectivity_en=conn_all,
                                    constitutive_model=b.constitutive_model,
                                    material_params=mp.reshape((C * mp.shape[1],) + mp.shape[2:])))
    u0 = None if u_0_bg is None else _np(u_0_bg).reshape(-1)
    from . import solver as _solver
    _solver._BATCH["copies"], _solver._BATCH["iterations"] = C, []
    try:
        u, res, _ = solve_bvp(verts, batches, element_residual_func, u0, bcs_all, vals.reshape(-1),
                              solver_options=solver_options)
    finally:
        solve_bvp_batched.last_iterations = [k.cpu().numpy() for k in _solver._BATCH["iterations"]]
        _solver._BATCH["copies"], _solver._BATCH["iterations"] = None, []
    return u.reshape(C, V * D), res.reshape(C, V * D)


solve_bvp_batched.last_iterations = []
