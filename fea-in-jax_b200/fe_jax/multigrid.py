"""Geometric multigrid preconditioner for the CG of the Newton loop on structured hex8 grids.

The reference offers the Jacobi preconditioner only (`lambda x: x / diag_J(x_0)`,
/root/reference/src/fe_jax/sparse_linear_solve.py:301-306) and hands the system to AMGX when a multilevel method is
wanted (:704-729).  At 24 M DoFs Jacobi-CG needs O(10^3) iterations, so the Newton-CG step is all SpMV; this module
cuts the iteration count instead:

* levels: the (nx, ny, nz)-element grid is coarsened by two per axis while all three counts stay even; a coarse
  level keeps every other vertex of the finer one (coordinates injected, so a jittered grid stays consistent), its
  Young's modulus is the mean over the child elements, its Dirichlet mask the injected one;
* coarse operators are RE-DISCRETISED: J_w of the coarse mesh through the same public builder the fine level uses
  (calculate_jacobian_w_dirichlet -> the fused assembly kernel), no sparse triple product;
* transfers: trilinear interpolation P and its transpose (csrc/multigrid.cu, feb200_mg_prolong_add / feb200_mg_restrict),
  constrained DoFs excluded on both sides;
* smoother: Chebyshev polynomial in D^-1 A of degree `degree` on [lmax / ratio, 1.1 lmax] (lmax from a few power
  iterations at setup) -- only SpMV (feb200_spmv) and vector updates, and a fixed symmetric linear operator, so the V-cycle
  is a valid preconditioner for the plain CG recurrence of solve_cg.py:20-63;
* coarsest level: the same Chebyshev iteration with a wide interval and a higher degree.

Use: `M = GeometricMultigrid(...); x, info = fe_jax.cg(J, b, x0, tol=..., M=M)` (the general-M branch of fe_jax.cg), or
`solve_linear_system_mg(...)` below.  Applies to linear-elastic and hyperelastic tangents alike (the coarse operators are
those of the linear isotropic law with the averaged moduli: a preconditioner, not the operator).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .assembly import ElementBatch, batch_to_collection, calculate_jacobian_w_dirichlet, csr_diagonal
from .mesh import as_device, mesh_to_sparse_assembly_map, uniform_hex_grid
from .physics import elastic_isotropic, linear_elasticity_residual
from .sparse import pattern_matvec
from .tables import CellType, ElementFamily, FiniteElementType, LagrangeVariant, QuadratureType


class _Level:
    pass


def _cheb(level, b, x, degree, lo, hi):
    """`degree` Chebyshev steps on A x = b, preconditioned by D^-1, eigenvalue interval [lo, hi] of D^-1 A.
    x is None for a zero initial guess (saves the first residual evaluation)."""
    theta, delta = 0.5 * (hi + lo), 0.5 * (hi - lo)
    sigma = theta / delta
    rho = 1.0 / sigma
    if x is None:
        r = b.clone()
        x = torch.zeros_like(b)
    else:
        r = b - pattern_matvec(level.pat, level.vals, x)
    d = (level.inv_diag * r) * (1.0 / theta)
    for k in range(degree):
        x += d
        if k == degree - 1:
            break
        r -= pattern_matvec(level.pat, level.vals, d)
        rho_new = 1.0 / (2.0 * sigma - rho)
        d = (rho_new * rho) * d + (2.0 * rho_new / delta) * (level.inv_diag * r)
        rho = rho_new
    return x


class GeometricMultigrid:
    """V-cycle preconditioner z = M r for the eliminated Jacobian J_w of a structured hex8 grid.

    vertices (V, 3), grid = (nx, ny, nz) elements with vertex id (i*(ny+1)+j)*(nz+1)+k, young (E,) or (E, Q) or scalar
    Young's modulus of the fine elements (element id (i*ny+j)*nz+k), nu the Poisson ratio, dirichlet_mask_g (3V,) 0/1;
    fine_pattern / fine_values: the assembled fine operator (NodePattern, CSR values of J_w)."""

    def __init__(self, vertices, grid, young, nu, dirichlet_mask_g, fine_pattern, fine_values, degree=2, ratio=6.0,
                 coarse_degree=24, coarse_ratio=400.0, min_coarse=4):
        dev = fine_values.device
        self.degree, self.ratio, self.coarse_degree, self.coarse_ratio = degree, ratio, coarse_degree, coarse_ratio
        nx, ny, nz = (int(g) for g in grid)
        Vx = np.asarray(vertices.cpu() if isinstance(vertices, torch.Tensor) else vertices, dtype=np.float64)
        assert Vx.shape[0] == (nx + 1) * (ny + 1) * (nz + 1), "vertices do not match the grid"
        mask = as_device(dirichlet_mask_g, torch.float64).contiguous()
        Ee = np.asarray(young.cpu() if isinstance(young, torch.Tensor) else young, dtype=np.float64)
        Ee = np.full(nx * ny * nz, float(Ee)) if Ee.ndim == 0 else (Ee.reshape(nx * ny * nz, -1).mean(axis=1))
        fe = FiniteElementType(CellType.hexahedron, ElementFamily.P, 1, LagrangeVariant.equispaced, QuadratureType.default, 2)
        self.levels = []
        lv = _Level()
        lv.grid, lv.pat, lv.vals, lv.mask = (nx, ny, nz), fine_pattern, fine_values, mask
        self.levels.append(lv)
        Vg = Vx.reshape(nx + 1, ny + 1, nz + 1, 3)
        Mg = mask.view(nx + 1, ny + 1, nz + 1, 3)
        Eg = Ee.reshape(nx, ny, nz)
        while all(g % 2 == 0 and g // 2 >= min_coarse for g in (nx, ny, nz)):
            nx, ny, nz = nx // 2, ny // 2, nz // 2
            Vg = Vg[::2, ::2, ::2]
            Mg = Mg[::2, ::2, ::2]
            Eg = Eg.reshape(nx, 2, ny, 2, nz, 2).mean(axis=(1, 3, 5))
            _, cells = uniform_hex_grid(nx, ny, nz)
            mp = np.stack([Eg.reshape(-1), np.full(nx * ny * nz, float(nu))], axis=1)
            batch = ElementBatch(fe_type=fe, n_dofs_per_basis=3, connectivity_en=cells, constitutive_model=elastic_isotropic,
                                 material_params=mp)
            Vc = np.ascontiguousarray(Vg.reshape(-1, 3))
            ebc = batch_to_collection(Vc, [batch])
            maps = [mesh_to_sparse_assembly_map(Vc.shape[0], cells)]
            mc = Mg.reshape(-1).contiguous()
            zero = torch.zeros_like(mc)
            J = calculate_jacobian_w_dirichlet(linear_elasticity_residual, ebc, maps, zero, zero, mc)
            lv = _Level()
            lv.grid, lv.pat, lv.vals, lv.mask = (nx, ny, nz), J._pattern, J.data, mc
            lv._keep = (ebc, maps)
            self.levels.append(lv)
        for lv in self.levels:
            lv.inv_diag = (1.0 / csr_diagonal(lv.pat, lv.vals)).contiguous()
            lv.lmax = self._power_iteration(lv)
        self.free = (1.0 - mask).contiguous()          # 1 on free DoFs: the cycle works on them only (J_w is the identity on the rest)
        self.dev = dev

    @staticmethod
    def _power_iteration(lv, iters=12):
        g = torch.Generator(device=lv.vals.device).manual_seed(7)
        v = torch.rand(lv.inv_diag.shape[0], dtype=torch.float64, device=lv.vals.device, generator=g)
        lam = 1.0
        for _ in range(iters):
            v = v / torch.linalg.norm(v)
            w = lv.inv_diag * pattern_matvec(lv.pat, lv.vals, v)
            lam = float(torch.linalg.norm(w).item())
            v = w
        return lam

    def _transfer(self, name, lc, src, mask, dst):
        nxc, nyc, nzc = lc.grid
        _lib.call(name, nxc, nyc, nzc, 3, _lib.ptr(src), _lib.ptr(mask), _lib.ptr(dst))

    def _vcycle(self, l, b):
        lv = self.levels[l]
        hi = 1.1 * lv.lmax
        if l == len(self.levels) - 1:
            return _cheb(lv, b, None, self.coarse_degree, hi / self.coarse_ratio, hi)
        x = _cheb(lv, b, None, self.degree, hi / self.ratio, hi)
        r = b - pattern_matvec(lv.pat, lv.vals, x)      # (zero on constrained DoFs: b and x are, and J_w decouples them)
        lc = self.levels[l + 1]
        rc = torch.empty(lc.inv_diag.shape[0], dtype=torch.float64, device=b.device)
        self._transfer("feb200_mg_restrict", lc, r, lc.mask, rc)
        ec = self._vcycle(l + 1, rc)
        self._transfer("feb200_mg_prolong_add", lc, ec, lv.mask, x)
        return _cheb(lv, b, x, self.degree, hi / self.ratio, hi)

    def __call__(self, r):
        """z = M r (one V-cycle with a zero initial guess); constrained DoFs pass through the unit diagonal of J_w."""
        r = as_device(r, torch.float64).contiguous()
        z = self._vcycle(0, r * self.free)
        return z + r * self.levels[0].mask

    @property
    def shape_summary(self):
        return [lv.grid for lv in self.levels]
