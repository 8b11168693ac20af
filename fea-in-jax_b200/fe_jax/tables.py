"""Basis / quadrature tables (host side, tiny) -- the role basix plays for the reference.

Mirrors /root/reference/src/fe_jax/basis_quadrature.py:17-87: `FiniteElementType`,
`eval_basis_and_derivatives` -> (phi (Q,N), dphi_dxi (Q,N,P)) and `get_quadrature` -> (xi (Q,P), W (Q,)).
basix is not available in this image, so the enums below are stand-ins carrying basix' names and the
tables are generated here in basix' reference-cell ordering (vertices, edges, faces, interior;
tensor-product cells x-fastest).  Only the P1 triangle is pinned by a reference test
(tests/test_basix_interface.py:59-81); see DESIGN.md.
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum

import numpy as np


class CellType(Enum):
    interval = 1
    triangle = 2
    tetrahedron = 3
    quadrilateral = 4
    hexahedron = 5


class ElementFamily(Enum):
    P = 1


class LagrangeVariant(Enum):
    equispaced = 1
    gll_warped = 2
    unset = 0


class QuadratureType(Enum):
    default = 0
    gauss_jacobi = 1
    gll = 2
    xiao_gimbutas = 3


class PolysetType(Enum):
    standard = 0


@dataclass
class FiniteElementType:
    """basis_quadrature.py:17-28"""
    cell_type: CellType
    family: ElementFamily
    basis_degree: int
    lagrange_variant: LagrangeVariant
    quadrature_type: QuadratureType
    quadrature_degree: int


_SIMPLEX_EDGES = {CellType.triangle: ((1, 2), (0, 2), (0, 1)),
                  CellType.tetrahedron: ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1))}
_DIM = {CellType.triangle: 2, CellType.quadrilateral: 2, CellType.tetrahedron: 3, CellType.hexahedron: 3}


def _tensor_node_indices(cell: CellType, degree: int):
    """(i, j[, k]) lattice index of every node of a tensor-product cell, basix order."""
    P = _DIM[cell]
    corner = [tuple(((c >> p) & 1) * degree for p in range(P)) for c in range(2 ** P)]
    if degree == 1:
        return corner
    if degree != 2:
        raise NotImplementedError("tensor-product Lagrange tables exist for degree 1 and 2")
    if P == 2:
        edges = ((0, 1), (0, 2), (1, 3), (2, 3))
        faces = ((0, 1, 2, 3),)
        volume = []
    else:
        edges = ((0, 1), (0, 2), (0, 4), (1, 3), (1, 5), (2, 3), (2, 6), (3, 7), (4, 5), (4, 6), (5, 7), (6, 7))
        faces = ((0, 1, 2, 3), (0, 1, 4, 5), (0, 2, 4, 6), (1, 3, 5, 7), (2, 3, 6, 7), (4, 5, 6, 7))
        volume = [(1, 1, 1)]
    mid = lambda ids: tuple(sum(corner[v][p] for v in ids) // len(ids) for p in range(P))
    return corner + [mid(e) for e in edges] + [mid(f) for f in faces] + volume


def _lagrange_1d_table(degree: int, t: np.ndarray):
    """values / derivatives of the equispaced 1-D Lagrange basis of `degree` on [0,1]"""
    if degree == 1:
        return np.stack([1.0 - t, t], axis=1), np.stack([-np.ones_like(t), np.ones_like(t)], axis=1)
    val = np.stack([(2 * t - 1) * (t - 1), 4 * t * (1 - t), t * (2 * t - 1)], axis=1)
    der = np.stack([4 * t - 3, 4 - 8 * t, 4 * t - 1], axis=1)
    return val, der


def eval_basis_and_derivatives(fe_type: FiniteElementType, xi_qp: np.ndarray):
    """phi (Q,N), dphi_dxi (Q,N,P).  The reference slices r[1:3] (2-D only,
    basis_quadrature.py:60); here all P derivative components are returned."""
    xi = np.asarray(xi_qp, dtype=np.float64)
    Q, P = xi.shape
    cell, deg = fe_type.cell_type, fe_type.basis_degree
    if cell in (CellType.quadrilateral, CellType.hexahedron):
        idx = _tensor_node_indices(cell, deg)
        tabs = [_lagrange_1d_table(deg, xi[:, p]) for p in range(P)]
        phi = np.empty((Q, len(idx)))
        dphi = np.empty((Q, len(idx), P))
        for a, ijk in enumerate(idx):
            vals = [tabs[p][0][:, ijk[p]] for p in range(P)]
            ders = [tabs[p][1][:, ijk[p]] for p in range(P)]
            phi[:, a] = np.prod(vals, axis=0)
            for r in range(P):
                dphi[:, a, r] = np.prod([ders[p] if p == r else vals[p] for p in range(P)], axis=0)
        return phi, dphi
    if cell not in _SIMPLEX_EDGES:
        raise NotImplementedError(f"cell type {cell}")
    lam = np.concatenate([1.0 - xi.sum(axis=1, keepdims=True), xi], axis=1)
    dlam = np.vstack([-np.ones((1, P)), np.eye(P)])
    if deg == 1:
        return lam, np.tile(dlam[None], (Q, 1, 1))
    if deg != 2:
        raise NotImplementedError("simplex Lagrange tables exist for degree 1 and 2")
    edges = _SIMPLEX_EDGES[cell]
    phi = np.concatenate([lam * (2 * lam - 1)] + [4 * lam[:, [a]] * lam[:, [b]] for a, b in edges], axis=1)
    dv = (4 * lam - 1)[:, :, None] * dlam[None]
    de = [4 * (lam[:, [a], None] * dlam[None, [b]] + lam[:, [b], None] * dlam[None, [a]]) for a, b in edges]
    return phi, np.concatenate([dv] + de, axis=1)


def _gauss_unit(m: int):
    x, w = np.polynomial.legendre.leggauss(m)
    return 0.5 * (x + 1.0), 0.5 * w


def _collapsed(dim: int, m: int):
    from scipy.special import roots_jacobi

    rules = [roots_jacobi(m, float(alpha), 0.0) for alpha in range(dim)]        # alpha = 0, 1, (2)
    unit = [(0.5 * (x + 1.0), w / 2.0 ** (alpha + 1)) for alpha, (x, w) in enumerate(rules)]
    pts, wts = [], []
    if dim == 2:
        (x0, w0), (x1, w1) = unit
        for i in range(m):
            for j in range(m):
                pts.append((x1[i], x0[j] * (1 - x1[i])))
                wts.append(w1[i] * w0[j])
    else:
        (x0, w0), (x1, w1), (x2, w2) = unit
        for i in range(m):
            for j in range(m):
                for k in range(m):
                    pts.append((x2[i], x1[j] * (1 - x2[i]), x0[k] * (1 - x2[i]) * (1 - x1[j])))
                    wts.append(w2[i] * w1[j] * w0[k])
    return np.array(pts), np.array(wts)


def get_quadrature(fe_type: FiniteElementType):
    """xi (Q,P), W (Q,).  basis_quadrature.py:63-87.  Tensor cells: Gauss-Legendre with
    m = (degree + 2) // 2 points per axis on [0,1], last axis fastest.  Simplices: the
    low-degree symmetric rules (degree 3 triangle: 6 points, sum W = 1/2), else collapsed
    Gauss-Jacobi."""
    cell, deg = fe_type.cell_type, fe_type.quadrature_degree
    if cell in (CellType.quadrilateral, CellType.hexahedron):
        P = _DIM[cell]
        x, w = _gauss_unit((deg + 2) // 2)
        grids = np.meshgrid(*([x] * P), indexing="ij")
        wgrid = np.meshgrid(*([w] * P), indexing="ij")
        return np.stack([g.ravel() for g in grids], axis=1), np.prod([g.ravel() for g in wgrid], axis=0)
    if cell == CellType.triangle:
        if deg <= 1:
            return np.array([[1 / 3, 1 / 3]]), np.array([0.5])
        if deg == 2:
            return np.array([[1 / 6, 1 / 6], [1 / 6, 2 / 3], [2 / 3, 1 / 6]]), np.full(3, 1 / 6)
        if deg == 3:
            a, b, c = 0.659027622374092, 0.231933368553031, 0.109039009072877
            return np.array([[a, b], [a, c], [b, a], [b, c], [c, a], [c, b]]), np.full(6, 1 / 12)
        return _collapsed(2, (deg + 2) // 2)
    if cell == CellType.tetrahedron:
        if deg <= 1:
            return np.array([[0.25, 0.25, 0.25]]), np.array([1 / 6])
        if deg == 2:
            a, b = 0.5854101966249685, 0.1381966011250105
            return np.array([[b, b, b], [a, b, b], [b, a, b], [b, b, a]]), np.full(4, 1 / 24)
        return _collapsed(3, (deg + 2) // 2)
    raise NotImplementedError(f"cell type {cell}")
