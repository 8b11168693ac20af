"""Solver options, result records and `linear_solve` (CG branches).

Call surface of /root/reference/src/fe_jax/sparse_linear_solve.py: PreconditionerType / LinearSolverType
:75-99, SolverOptions :102-135, SolverResultInfo :138-172, init_solver_info, Residual / Jacobian /
JacobianDiagonl :175-199, linear_solve :202-529.  In scope: CG_JAX_SCIPY(_W_INFO) with NONE / JACOBI.
The reference applies the operator as a JVP of the residual (:251-258); here the assembled,
Dirichlet-eliminated CSR (same operator (I-M) J (I-M) + M) is used with SpMV, so the `jacobian`
argument is required.  Every other solver type is third-party dispatch and raises.
"""
from __future__ import annotations

from dataclasses import dataclass, replace
from enum import Enum, auto
from typing import Callable, Optional

import torch

from . import _lib
from .krylov import JacobiPreconditioner, cg_on_pattern, cg_on_pattern_batched

# set by batched.solve_bvp_batched: the system is `copies` independent problems side by side; CG then runs per copy
_BATCH = {"copies": None, "iterations": []}
from .mesh import as_device
from .sparse import apply_dirichlet_bcs_lhs


class PreconditionerType(Enum):
    NONE = auto()
    JACOBI = auto()
    ILU_CUPY = auto()
    GEOMETRIC_MULTIGRID = auto()   # not in the reference (it reaches for AMGX, sparse_linear_solve.py:704-729): fe_jax.multigrid


class LinearSolverType(Enum):
    DENSE_INVERSE_JNP = auto()
    CG_JAX_SCIPY = auto()
    CG_JAX_SCIPY_W_INFO = auto()
    GMRES_JAX_SCIPY = auto()
    BICGSTAB_JAX_SCIPY = auto()
    DENSE_INVERSE_JAXOPT = auto()
    LU_JAXOPT = auto()
    CHOLESKY_JAXOPT = auto()
    CG_JAXOPT = auto()
    GMRES_JAXOPT = auto()
    BICGSTAB_JAXOPT = auto()
    SPSOLVE_CUPY = auto()
    LU_CUPY = auto()
    AMGX = auto()
    SPSOLVE_PYPARDISO = auto()


_IN_SCOPE = (LinearSolverType.CG_JAX_SCIPY, LinearSolverType.CG_JAX_SCIPY_W_INFO)


@dataclass(eq=True, frozen=True)
class SolverOptions:
    linear_precond_type: PreconditionerType = PreconditionerType.NONE
    linear_solve_type: LinearSolverType = LinearSolverType.CG_JAX_SCIPY_W_INFO
    linear_max_iter: int = 1000
    linear_relative_tol: float = 1e-14
    linear_absolute_tol: float = 1e-10
    nonlinear_max_iter: int = 10
    nonlinear_relative_tol: float = 1e-10
    nonlinear_absolute_tol: float = 1e-8

    def __post_init__(self):
        assert self.linear_precond_type != PreconditionerType.ILU_CUPY, \
            f"The selected preconditioner ({self.linear_precond_type.name}) requires cupy, which is not available."
        assert self.linear_solve_type in _IN_SCOPE, \
            (f"The selected solver ({self.linear_solve_type.name}) is third-party dispatch outside the feb200 hot path; "
             "use CG_JAX_SCIPY or CG_JAX_SCIPY_W_INFO.")


@dataclass
class SolverResultInfo:
    nonlinear_iterations: int
    cumulative_linear_iterations: int
    linear_iterations_per_nonlinear_iteration: torch.Tensor
    # length nonlinear_iterations + cumulative_linear_iterations: each nonlinear iteration's history
    # starts with the residual norm before its linear solve
    cumulative_residual_norm_history: torch.Tensor

    def increment_nl_iteration(self):
        return replace(self, nonlinear_iterations=self.nonlinear_iterations + 1)


def init_solver_info(opts: SolverOptions, device=None):
    if device is None:
        _lib.require_cuda()
        device = torch.device("cuda", torch.cuda.current_device())
    return SolverResultInfo(
        nonlinear_iterations=0, cumulative_linear_iterations=0,
        linear_iterations_per_nonlinear_iteration=torch.zeros(opts.nonlinear_max_iter, dtype=torch.float64, device=device),
        cumulative_residual_norm_history=torch.zeros(opts.linear_max_iter * opts.nonlinear_max_iter + 1,
                                                     dtype=torch.float64, device=device))


@dataclass
class Residual:
    function: Callable
    dirichlet_bcs_builtin: bool


@dataclass
class Jacobian:
    function: Callable
    dirichlet_bcs_builtin: bool


@dataclass
class JacobianDiagonl:
    function: Callable
    dirichlet_bcs_builtin: bool


def linear_solve(residual: Residual, jacobian: Optional[Jacobian], jacobian_diagonal: Optional[JacobianDiagonl],
                 dirichlet_dofs, dirichlet_values, solver_options: SolverOptions, solver_info_0: SolverResultInfo,
                 check_consistency: bool, x_0, *args, **kwargs):
    """Solve J(x_0) dx = -R(x_0) (sparse_linear_solve.py:202-529, CG branches); returns (delta_x, info)."""
    if not residual.dirichlet_bcs_builtin:
        raise Exception("TODO (straightforward) implementation needed")
    if jacobian is None:
        raise _lib.Feb200Error("linear_solve on the feb200 path needs the `jacobian` argument: the CG operator is the "
                               "assembled CSR (SpMV), not a JVP of the residual")
    x_0 = as_device(x_0, torch.float64)
    dofs = as_device(dirichlet_dofs, torch.int64)
    values = as_device(dirichlet_values, torch.float64)

    R_0 = residual.function(x_0, *args, **kwargs)
    J = jacobian.function(x_0, *args, **kwargs)
    if not jacobian.dirichlet_bcs_builtin:
        J = apply_dirichlet_bcs_lhs(J, dofs)
    if J._pattern is None:
        raise _lib.Feb200Error("linear_solve expects a Jacobian assembled by calculate_jacobian_wo_dirichlet")
    pat = J._pattern

    if check_consistency:
        from .sparse import coo_matvec
        v = torch.rand(x_0.shape, dtype=torch.float64, device=x_0.device)
        dense = J.todense()
        print("Jacobian-vector product via SpMV matches product via the dense Jacobian:",
              bool(torch.isclose(coo_matvec(J, v), dense @ v).all()))
        if jacobian_diagonal is not None:
            d = jacobian_diagonal.function(x_0, *args, **kwargs)
            if not jacobian_diagonal.dirichlet_bcs_builtin:
                d = d.clone()
                d[dofs] = 1.0
            print("Jacobian diagonal function matches the diagonal of the assembled Jacobian:",
                  bool(torch.isclose(torch.diag(dense), d).all()))

    # delta_x seeded at the constrained DoFs (:287-292)
    delta_x = torch.zeros_like(R_0)
    delta_x[dofs] = values - x_0[dofs]

    inv_diag = None
    if solver_options.linear_precond_type == PreconditionerType.JACOBI:
        assert jacobian_diagonal is not None, \
            f"{solver_options.linear_precond_type} requires the `jacobian_diagonal` argument to be provided."
        from .assembly import csr_diagonal
        inv_diag = JacobiPreconditioner(csr_diagonal(pat, J.data)).inv_diag     # diag of the eliminated matrix
    elif solver_options.linear_precond_type == PreconditionerType.GEOMETRIC_MULTIGRID:
        if _MULTIGRID["factory"] is None:
            raise _lib.Feb200Error("PreconditionerType.GEOMETRIC_MULTIGRID needs set_multigrid_factory(lambda J: "
                                   "GeometricMultigrid(...)) first (the grid shape is not part of the reference call surface)")
        from .krylov import cg as _cg
        M = _MULTIGRID["factory"](J)
        b = (-R_0).contiguous()
        delta_x, cg_info = _cg(J, b, delta_x, tol=solver_options.linear_relative_tol, atol=solver_options.linear_absolute_tol,
                               maxiter=solver_options.linear_max_iter, M=M)
        k, hist = int(cg_info["iterations"]), cg_info["residual_norm_history"]
        return _finish(delta_x, k, hist, dofs, values, x_0, solver_options, solver_info_0)
    elif solver_options.linear_precond_type != PreconditionerType.NONE:
        raise Exception(f"Preconditioner type {solver_options.linear_precond_type} is not implemented")

    b = (-R_0).contiguous()
    maxiter = solver_options.linear_max_iter if solver_options.linear_solve_type == LinearSolverType.CG_JAX_SCIPY_W_INFO \
        else 10 * b.shape[0]
    if _BATCH["copies"]:
        delta_x, ks, hist_c = cg_on_pattern_batched(pat, J.data, b, delta_x, solver_options.linear_relative_tol,
                                                    solver_options.linear_absolute_tol, maxiter, _BATCH["copies"], inv_diag)
        _BATCH["iterations"].append(ks)
        k = int(ks.max().item())
        hist = torch.sqrt((hist_c ** 2).sum(dim=0))          # batch residual norm per iteration (finished copies keep their last value 0)
    else:
        delta_x, k, hist = cg_on_pattern(pat, J.data, b, delta_x, solver_options.linear_relative_tol,
                                         solver_options.linear_absolute_tol, maxiter, inv_diag)
    return _finish(delta_x, k, hist, dofs, values, x_0, solver_options, solver_info_0)


_MULTIGRID = {"factory": None}


def set_multigrid_factory(factory):
    """factory(J: COO on a NodePattern) -> preconditioner callable, used by PreconditionerType.GEOMETRIC_MULTIGRID
    (None removes it)."""
    _MULTIGRID["factory"] = factory


def _finish(delta_x, k, hist, dofs, values, x_0, solver_options, solver_info_0):
    info = solver_info_0
    if solver_options.linear_solve_type == LinearSolverType.CG_JAX_SCIPY_W_INFO:
        per_nl = solver_info_0.linear_iterations_per_nonlinear_iteration.clone()
        if solver_info_0.nonlinear_iterations < per_nl.shape[0]:
            per_nl[solver_info_0.nonlinear_iterations] = k
        cum = solver_info_0.cumulative_residual_norm_history.clone()
        start = solver_info_0.cumulative_linear_iterations + solver_info_0.nonlinear_iterations
        n = max(0, min(hist.shape[0], cum.shape[0] - start))
        cum[start:start + n] = hist[:n]
        info = SolverResultInfo(nonlinear_iterations=solver_info_0.nonlinear_iterations,
                                cumulative_linear_iterations=solver_info_0.cumulative_linear_iterations + k,
                                linear_iterations_per_nonlinear_iteration=per_nl,
                                cumulative_residual_norm_history=cum)
    # the iterative solve does not land exactly on the constraints: overwrite them (:527)
    delta_x[dofs] = values - x_0[dofs]
    return delta_x, info
