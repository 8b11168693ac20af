"""fe_jax call surface on the B200-native hot path (libfeb200).

Drop-in for the data-parallel path of KeithBallard/fea-in-jax: batched element residual / tangent
evaluation, deterministic sparse assembly, CSR SpMV + CG inside the Newton loop.  Same function names,
argument meaning and return structure as /root/reference/src/fe_jax (see the per-module docstrings);
arrays are torch CUDA float64 tensors because JAX is not available in this image (INTEGRATION.md shows
the jax.ffi registration).  FP64 is canonical (reference: src/fe_jax/__init__.py:3).
"""
from .tables import (CellType, ElementFamily, LagrangeVariant, QuadratureType, PolysetType, FiniteElementType,
                     eval_basis_and_derivatives, get_quadrature)
from .containers import BCSR, COO, CSR, NodePattern
from .mesh import (uniform_quad_grid, uniform_tri_grid, uniform_hex_grid, mesh_to_jax, mesh_to_element_coords,
                   get_n_cells_per_vert, build_row_ind, build_col_ind_and_data, mesh_to_sparse_assembly_map,
                   transform_global_to_element_node, transform_global_unraveled_to_element_node,
                   transform_element_node_to_global_unraveled_nosum, transform_element_node_to_global_unraveled_sum,
                   transform_element_node_to_global_sum, transform_element_node_to_global_nosum)
from .physics import (elastic_isotropic, elastic_orthotropic, st_venant_kirchhoff, mooney_rivlin,
                      linear_elasticity_residual, hyperelasticity_residual)
from .assembly import (ElementBatch, ElementBatchCollection, MaterialPropertyArrayType, QuadratureArrayType,
                       batch_to_collection, _calculate_jacobian_unique_nnz, calculate_jacobian_wo_dirichlet,
                       calculate_jacobian_w_dirichlet, calculate_jacobian_diag_wo_dirichlet,
                       calculate_residual_wo_dirichlet, calculate_residual_w_dirichlet)
from .sparse import (coo_arrays_sum_duplicates, coo_arrays_sum_duplicates_jit, coo_sum_duplicates, coo_to_csr,
                     apply_dirichlet_bcs_lhs, apply_dirichlet_bcs_rhs)
from .krylov import cg, JacobiPreconditioner
from .krylov import cg as cg_w_info
from .solver import (PreconditionerType, LinearSolverType, SolverOptions, SolverResultInfo, init_solver_info,
                     Residual, Jacobian, JacobianDiagonl, linear_solve, set_multigrid_factory)
from .newton import solve_nonlinear_step, solve_bvp
from . import partition
from .vtk_io import read_vtk, write_vtk
from .batched import solve_bvp_batched
from .multigrid import GeometricMultigrid
