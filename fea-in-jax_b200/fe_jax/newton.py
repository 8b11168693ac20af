"""Newton loop and `solve_bvp` (/root/reference/src/fe_jax/fea.py:968-1309)."""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _lib
from .assembly import (ElementBatchCollection, _plan, batch_to_collection, calculate_jacobian_diag_wo_dirichlet,
                       calculate_jacobian_w_dirichlet, calculate_residual_w_dirichlet)
from .containers import COO
from .mesh import as_device, mesh_to_sparse_assembly_map
from .solver import Jacobian, JacobianDiagonl, Residual, SolverOptions, init_solver_info, linear_solve


def _quiet():
    return os.environ.get("FEB200_QUIET", "0") == "1"


def solve_nonlinear_step(element_residual_func, ebc: ElementBatchCollection, assembly_map_b, jacobian_nnz: int,
                         u_0_g, dirichlet_values_g, dirichlet_mask_g, dirichlet_dofs, dirichlet_values,
                         solver_options: SolverOptions):
    """Newton's method on R_w(u) = 0 (fea.py:968-1129); returns (u, [state_b], R, relative_error, info)."""
    if not _quiet():
        print(f"Global dimensionality : {ebc.D}")
        print(f"# of batches : {ebc.B}")
        for i in range(ebc.B):
            print(f"For batch {i}:\n\t", f"Number of elements : {ebc.E[i]}\n\t",
                  f"Number of nodes / element : {ebc.N[i]}\n\t", f"Number of quadrature points : {ebc.Q[i]}\n\t",
                  f"Parametric dimensionality: {ebc.P[i]}\n\t",
                  f"Number of material parameters per quad point: {ebc.M[i]}")
    values_g = as_device(dirichlet_values_g, torch.float64)
    mask_g = as_device(dirichlet_mask_g, torch.float64)

    def residual_isv_func_w_dirichlet(u_f):
        return calculate_residual_w_dirichlet(element_residual_func=element_residual_func, ebc=ebc,
                                              assembly_map_b=assembly_map_b, u_f=u_f, dirichlet_values_g=values_g,
                                              dirichlet_mask_g=mask_g)

    def residual_func_w_dirichlet(u_f):
        return residual_isv_func_w_dirichlet(u_f)[0]

    def jacobian_func_w_dirichlet(u_f):
        # J_w = (I-M) J(u_w) (I-M) + M, assembled at u_w = (1-m) u + m g with the elimination fused into the
        # assembly store: this is the operator the reference's CG branch applies (JVP of R_w at x_0,
        # sparse_linear_solve.py:251-258); identical to apply_dirichlet_bcs_lhs(J(u_f)) for linear problems
        return calculate_jacobian_w_dirichlet(element_residual_func=element_residual_func, ebc=ebc,
                                              assembly_map_b=assembly_map_b, u_f=u_f, dirichlet_values_g=values_g,
                                              dirichlet_mask_g=mask_g, precomputed_jacobian_nnz=jacobian_nnz)

    def jacobian_diag_func_wo_dirichlet(u_f):
        return calculate_jacobian_diag_wo_dirichlet(element_residual_func=element_residual_func, ebc=ebc,
                                                    assembly_map_b=assembly_map_b, u_f=u_f)

    u_f = as_device(u_0_g, torch.float64).clone()
    R_f, new_internal_state_beqi = residual_isv_func_w_dirichlet(u_f)
    initial_R_f_norm = float(torch.linalg.norm(R_f).item())
    info = init_solver_info(solver_options, device=u_f.device)
    nl_iteration = 0
    while True:
        absolute_error = float(torch.linalg.norm(R_f).item())
        relative_error = absolute_error / initial_R_f_norm if initial_R_f_norm != 0.0 else float("nan")
        if not _quiet():
            print(f"End of iteration {nl_iteration - 1} rel error {relative_error}, abs error {absolute_error}")
        if not (nl_iteration < solver_options.nonlinear_max_iter
                and relative_error > solver_options.nonlinear_relative_tol
                and absolute_error > solver_options.nonlinear_absolute_tol):
            break
        delta_u, info = linear_solve(
            residual=Residual(function=residual_func_w_dirichlet, dirichlet_bcs_builtin=True),
            jacobian=Jacobian(function=jacobian_func_w_dirichlet, dirichlet_bcs_builtin=True),
            jacobian_diagonal=JacobianDiagonl(function=jacobian_diag_func_wo_dirichlet, dirichlet_bcs_builtin=False),
            dirichlet_dofs=dirichlet_dofs, dirichlet_values=dirichlet_values, solver_options=solver_options,
            solver_info_0=info, check_consistency=False, x_0=u_f)
        u_f = u_f + delta_u
        R_f = residual_isv_func_w_dirichlet(u_f)[0]
        nl_iteration += 1
        info = info.increment_nl_iteration()
    absolute_error = float(torch.linalg.norm(R_f).item())
    relative_error = absolute_error / initial_R_f_norm if initial_R_f_norm != 0.0 else float("nan")
    return u_f, new_internal_state_beqi, R_f, relative_error, info


def solve_bvp(vertices_vd, element_batches, element_residual_func, u_0_g, dirichlet_bcs, dirichlet_values,
              solver_options: SolverOptions = SolverOptions(), plot_convergence: bool = False,
              profile_memory: bool = False):
    """Static BVP: returns (u (V*D,), residual (V*D,), element_batches) (fea.py:1132-1309)."""
    vertices_np = vertices_vd if isinstance(vertices_vd, torch.Tensor) else np.asarray(vertices_vd)
    V, D = int(vertices_np.shape[0]), int(vertices_np.shape[1])
    dirichlet_bcs = np.asarray(dirichlet_bcs.cpu() if isinstance(dirichlet_bcs, torch.Tensor) else dirichlet_bcs)
    dirichlet_values_np = np.asarray(dirichlet_values.cpu() if isinstance(dirichlet_values, torch.Tensor) else dirichlet_values,
                                     dtype=np.float64)
    assert D <= 3
    assert dirichlet_bcs.shape[0] <= D * V
    assert dirichlet_bcs.shape[1] == 2
    assert dirichlet_values_np.shape[0] == dirichlet_bcs.shape[0]
    for b in element_batches:
        assert b.connectivity_en.shape[1] <= V

    assembly_map_b = [mesh_to_sparse_assembly_map(n_vertices=V, cells=b.connectivity_en) for b in element_batches]
    ebc = batch_to_collection(vertices_vd=vertices_vd, element_batches=element_batches)
    assert (np.array(ebc.U) == ebc.U[0]).all(), "The number of DoFs per a point (U) must be the same across all batches."
    F = V * ebc.U[0]
    dev = ebc._vertices.device
    if u_0_g is None:
        u_0_g = torch.zeros(F, dtype=torch.float64, device=dev)
    else:
        u_0_g = as_device(u_0_g, torch.float64)
        assert tuple(u_0_g.shape) == (F,)

    jacobian_nnz = _plan(ebc, assembly_map_b).pattern.nnz

    # dof = D * vertex + component (fea.py:1219-1225)
    dofs_np = (D * dirichlet_bcs[:, 0].astype(np.int64) + dirichlet_bcs[:, 1].astype(np.int64))
    dirichlet_dofs = torch.from_numpy(dofs_np).to(dev)
    values = torch.from_numpy(dirichlet_values_np).to(dev)
    dirichlet_values_g = torch.zeros(F, dtype=torch.float64, device=dev)
    dirichlet_values_g[dirichlet_dofs] = values
    dirichlet_mask_g = torch.zeros(F, dtype=torch.float64, device=dev)
    dirichlet_mask_g[dirichlet_dofs] = 1.0

    u, internal_state_beqi, residual, relative_error, info = solve_nonlinear_step(
        element_residual_func=element_residual_func, ebc=ebc, assembly_map_b=assembly_map_b, jacobian_nnz=jacobian_nnz,
        u_0_g=u_0_g, dirichlet_values_g=dirichlet_values_g, dirichlet_mask_g=dirichlet_mask_g,
        dirichlet_dofs=dirichlet_dofs, dirichlet_values=values, solver_options=solver_options)

    if info.cumulative_linear_iterations > 0 and not _quiet():
        print(f"Cumulative # of linear solver iterations: {info.cumulative_linear_iterations}")
    solve_bvp.last_info = info
    return (u, residual, element_batches)
