"""
CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

A NumPy restatement of the fea-in-jax hot path (element residual / tangent ->
sparse assembly -> CG inside Newton).  Nothing in the product package imports
this file: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may call it, and only as the checker.

Every function cites the reference lines it restates (paths relative to
/root/reference).  The restatement is written independently (vectorised over
the element axis with NumPy) -- it is not a copy of the JAX sources.

Pinning status (see DESIGN.md "Oracle"):
  * pinned against the reference's own known answers and against outputs of the
    reference's *source text* executed under a NumPy stand-in for `jax.numpy`
    (tests/golden/make_golden.py -> tests/golden/*.npz): index-map builders,
    grids, Voigt maps, all four constitutive laws (2-D and 3-D), both element
    residuals, the CG recurrence, COO duplicate summation, Dirichlet
    elimination, the two-triangle BVP of tests/test_simple_fea_solve.py and the
    textbook C / J / det J / B values of tests/fea_traditional.py.
  * PARITY UNPINNED (no reference test, and basix is not installable here):
    the basix tables for anything other than the P1 triangle (hex8 / quad4 /
    tet / quadratic elements; node and quadrature-point ordering follow basix'
    documented tensor-product / sub-entity conventions).
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# Voigt helpers                                         src/fe_jax/utils.py:202-253
# --------------------------------------------------------------------------------------

_V3_I = np.array([0, 1, 2, 1, 0, 0])
_V3_J = np.array([0, 1, 2, 2, 2, 1])
_V2_I = np.array([0, 1, 0])
_V2_J = np.array([0, 1, 1])


def rank2_tensor_to_voigt(t):
    """(..., D, D) -> (..., S); shear entries doubled.  utils.py:202-225."""
    D = t.shape[-1]
    if D == 1:
        return t
    if D == 2:
        v = t[..., _V2_I, _V2_J].copy()
        v[..., 2] *= 2.0
        return v
    if D == 3:
        v = t[..., _V3_I, _V3_J].copy()
        v[..., 3:] *= 2.0
        return v
    raise RuntimeError("tensor must be 1D, 2D or 3D")


def rank2_voigt_to_tensor(v):
    """(..., S) -> (..., D, D).  utils.py:228-253 with the `shape[2] == 6`
    test read as `shape[-1] == 6`."""
    S = v.shape[-1]
    if S == 1:
        return v
    if S == 3:
        return v[..., [0, 2, 2, 1]].reshape(*v.shape[:-1], 2, 2)
    if S == 6:
        return v[..., [0, 5, 4, 5, 1, 3, 4, 3, 2]].reshape(*v.shape[:-1], 3, 3)
    raise RuntimeError("invalid Voigt size")


# --------------------------------------------------------------------------------------
# Basis / quadrature tables (stand-in for basix)     src/fe_jax/basis_quadrature.py:31-87
# --------------------------------------------------------------------------------------

def _gauss01(m):
    x, w = np.polynomial.legendre.leggauss(m)
    return 0.5 * (x + 1.0), 0.5 * w


def _simplex_collapsed(dim, m):
    """Collapsed Gauss-Jacobi rule on the unit simplex (UNPINNED ordering)."""
    from scipy.special import roots_jacobi

    x0, w0 = roots_jacobi(m, 0.0, 0.0)
    x1, w1 = roots_jacobi(m, 1.0, 0.0)
    if dim == 2:
        pts, wts = [], []
        for i in range(m):
            for j in range(m):
                a = 0.5 * (1 + x1[i])
                b = 0.5 * (1 + x0[j])
                pts.append([a, b * (1 - a)])
                wts.append(w1[i] * w0[j] * 0.125)
        return np.array(pts), np.array(wts)
    x2, w2 = roots_jacobi(m, 2.0, 0.0)
    pts, wts = [], []
    for i in range(m):
        for j in range(m):
            for k in range(m):
                a = 0.5 * (1 + x2[i])
                b = 0.5 * (1 + x1[j])
                c = 0.5 * (1 + x0[k])
                pts.append([a, b * (1 - a), c * (1 - a) * (1 - b)])
                wts.append(w2[i] * w1[j] * w0[k] * 0.125 * 0.25 * 0.5)
    return np.array(pts), np.array(wts)


def get_quadrature(cell: str, degree: int):
    """(xi (Q,P), W (Q,)) for `cell` in {triangle, quadrilateral, tetrahedron,
    hexahedron}.  basis_quadrature.py:63-87 (basix `make_quadrature`, default
    rule).  Triangle degree 3 = 6 points, sum W = 0.5
    (tests/test_basix_interface.py:59-67)."""
    if cell == "triangle":
        if degree <= 1:
            return np.array([[1 / 3, 1 / 3]]), np.array([0.5])
        if degree == 2:
            return (np.array([[1 / 6, 1 / 6], [1 / 6, 2 / 3], [2 / 3, 1 / 6]]),
                    np.full(3, 1 / 6))
        if degree == 3:
            a, b, c = 0.659027622374092, 0.231933368553031, 0.109039009072877
            return (np.array([[a, b], [a, c], [b, a], [b, c], [c, a], [c, b]]),
                    np.full(6, 1 / 12))
        return _simplex_collapsed(2, (degree + 2) // 2)
    if cell == "tetrahedron":
        if degree <= 1:
            return np.array([[0.25, 0.25, 0.25]]), np.array([1 / 6])
        if degree == 2:
            a, b = 0.5854101966249685, 0.1381966011250105
            return (np.array([[b, b, b], [a, b, b], [b, a, b], [b, b, a]]),
                    np.full(4, 1 / 24))
        return _simplex_collapsed(3, (degree + 2) // 2)
    m = (degree + 2) // 2
    x, w = _gauss01(m)
    if cell == "quadrilateral":
        pts = np.array([[x[i], x[j]] for i in range(m) for j in range(m)])
        wts = np.array([w[i] * w[j] for i in range(m) for j in range(m)])
        return pts, wts
    if cell == "hexahedron":
        pts = np.array([[x[i], x[j], x[k]] for i in range(m) for j in range(m) for k in range(m)])
        wts = np.array([w[i] * w[j] * w[k] for i in range(m) for j in range(m) for k in range(m)])
        return pts, wts
    raise ValueError(cell)


def _lagrange_1d(nodes, x):
    """values (len(x), n) and derivatives of the 1-D Lagrange basis on `nodes`."""
    n = len(nodes)
    x = np.asarray(x, dtype=np.float64)
    val = np.ones((x.shape[0], n))
    der = np.zeros((x.shape[0], n))
    for a in range(n):
        for b in range(n):
            if b != a:
                val[:, a] *= (x - nodes[b]) / (nodes[a] - nodes[b])
        for b in range(n):
            if b == a:
                continue
            term = np.full(x.shape[0], 1.0 / (nodes[a] - nodes[b]))
            for c in range(n):
                if c != a and c != b:
                    term *= (x - nodes[c]) / (nodes[a] - nodes[c])
            der[:, a] += term
    return val, der


# basix sub-entity tables (vertex lists of edges / faces) used to place the
# quadratic nodes.  UNPINNED (no reference test covers them).
_TRI_EDGES = [(1, 2), (0, 2), (0, 1)]
_TET_EDGES = [(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)]
_QUAD_EDGES = [(0, 1), (0, 2), (1, 3), (2, 3)]
_HEX_EDGES = [(0, 1), (0, 2), (0, 4), (1, 3), (1, 5), (2, 3), (2, 6), (3, 7), (4, 5), (4, 6), (5, 7), (6, 7)]
_HEX_FACES = [(0, 1, 2, 3), (0, 1, 4, 5), (0, 2, 4, 6), (1, 3, 5, 7), (2, 3, 6, 7), (4, 5, 6, 7)]


def reference_nodes(cell: str, degree: int):
    """Reference-cell node coordinates in basix ordering (vertices, then edge,
    face, interior nodes)."""
    if cell == "triangle":
        v = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])
        edges, faces = _TRI_EDGES, []
    elif cell == "tetrahedron":
        v = np.array([[0.0, 0, 0], [1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0]])
        edges, faces = _TET_EDGES, []
    elif cell == "quadrilateral":
        v = np.array([[0.0, 0], [1.0, 0], [0, 1.0], [1.0, 1.0]])
        edges, faces = _QUAD_EDGES, [(0, 1, 2, 3)]
    elif cell == "hexahedron":
        v = np.array([[(c >> 0) & 1, (c >> 1) & 1, (c >> 2) & 1] for c in range(8)], dtype=np.float64)
        edges, faces = _HEX_EDGES, _HEX_FACES
    else:
        raise ValueError(cell)
    if degree == 1:
        return v
    if degree != 2:
        raise NotImplementedError("only degree 1 and 2 tables are restated")
    pts = [v]
    pts.append(np.array([0.5 * (v[a] + v[b]) for a, b in edges]))
    if cell in ("quadrilateral", "hexahedron"):
        if faces:
            pts.append(np.array([v[list(f)].mean(axis=0) for f in faces]))
        if cell == "hexahedron":
            pts.append(np.array([[0.5, 0.5, 0.5]]))
    return np.vstack(pts)


def eval_basis_and_derivatives(cell: str, degree: int, xi_qp):
    """phi (Q,N), dphi_dxi (Q,N,P) with P = cell dimension (the reference keeps
    only r[1:3], i.e. two derivative components -- basis_quadrature.py:60; here
    generalised to r[1:1+P])."""
    xi = np.asarray(xi_qp, dtype=np.float64)
    Q, P = xi.shape
    nodes = reference_nodes(cell, degree)
    N = nodes.shape[0]
    if cell in ("quadrilateral", "hexahedron"):
        n1 = np.array([0.0, 1.0]) if degree == 1 else np.array([0.0, 0.5, 1.0])
        vals, ders = zip(*[_lagrange_1d(n1, xi[:, p]) for p in range(P)])
        phi = np.ones((Q, N))
        dphi = np.ones((Q, N, P))
        for a in range(N):
            idx = [int(round(nodes[a, p] * degree)) for p in range(P)]
            for p in range(P):
                phi[:, a] *= vals[p][:, idx[p]]
                for r in range(P):
                    dphi[:, a, r] *= ders[p][:, idx[p]] if r == p else vals[p][:, idx[p]]
        return phi, dphi
    # simplices: barycentric formulas
    lam = np.concatenate([1.0 - xi.sum(axis=1, keepdims=True), xi], axis=1)      # (Q, P+1)
    dlam = np.concatenate([-np.ones((1, P)), np.eye(P)], axis=0)                  # (P+1, P)
    if degree == 1:
        return lam.copy(), np.broadcast_to(dlam, (Q, P + 1, P)).copy()
    edges = _TRI_EDGES if cell == "triangle" else _TET_EDGES
    phi = np.zeros((Q, N))
    dphi = np.zeros((Q, N, P))
    for a in range(P + 1):
        phi[:, a] = lam[:, a] * (2 * lam[:, a] - 1)
        dphi[:, a, :] = (4 * lam[:, a] - 1)[:, None] * dlam[a][None, :]
    for k, (a, b) in enumerate(edges):
        phi[:, P + 1 + k] = 4 * lam[:, a] * lam[:, b]
        dphi[:, P + 1 + k, :] = 4 * (lam[:, a][:, None] * dlam[b][None, :] + lam[:, b][:, None] * dlam[a][None, :])
    return phi, dphi


# --------------------------------------------------------------------------------------
# Grids                                                src/fe_jax/setup.py:17-102
# --------------------------------------------------------------------------------------

def uniform_quad_grid(n_rows, n_cols, bbox):
    """setup.py:17-55: node id = i*n_cols + j, x follows i, y follows j; cell =
    [(i,j),(i+1,j),(i,j+1),(i+1,j+1)] (basix tensor order). 3-column vertices."""
    bbox = np.asarray(bbox, dtype=np.float64)
    i = np.arange(n_rows)
    j = np.arange(n_cols)
    x = i / float(n_rows - 1) * (bbox[0, 1] - bbox[0, 0]) + bbox[0, 0]
    y = j / float(n_cols - 1) * (bbox[1, 1] - bbox[1, 0]) + bbox[1, 0]
    V = np.zeros((n_rows * n_cols, 3))
    V[:, 0] = np.repeat(x, n_cols)
    V[:, 1] = np.tile(y, n_rows)
    ii, jj = np.meshgrid(np.arange(n_rows - 1), np.arange(n_cols - 1), indexing="ij")
    ii, jj = ii.ravel(), jj.ravel()
    F = np.stack([ii * n_cols + jj, (ii + 1) * n_cols + jj, ii * n_cols + jj + 1,
                  (ii + 1) * n_cols + jj + 1], axis=1).astype(np.int64)
    return V, F


def uniform_tri_grid(n_rows, n_cols):
    """setup.py:58-102: unit square, two triangles per cell."""
    V, _ = uniform_quad_grid(n_rows, n_cols, [[0.0, 1.0], [0.0, 1.0]])
    ii, jj = np.meshgrid(np.arange(n_rows - 1), np.arange(n_cols - 1), indexing="ij")
    ii, jj = ii.ravel(), jj.ravel()
    t0 = np.stack([ii * n_cols + jj, (ii + 1) * n_cols + jj, ii * n_cols + jj + 1], axis=1)
    t1 = np.stack([(ii + 1) * n_cols + jj, (ii + 1) * n_cols + jj + 1, ii * n_cols + jj + 1], axis=1)
    F = np.empty((2 * t0.shape[0], 3), dtype=np.int64)
    F[0::2] = t0
    F[1::2] = t1
    return V, F


def uniform_hex_grid(nx, ny, nz, bbox=((0.0, 1.0), (0.0, 1.0), (0.0, 1.0))):
    """3-D extension of uniform_quad_grid's convention: nx,ny,nz
    = number of ELEMENTS per axis; node id = (i*(ny+1)+j)*(nz+1)+k; hex
    connectivity in basix tensor order (x fastest)."""
    bbox = np.asarray(bbox, dtype=np.float64)
    gx = np.arange(nx + 1) / float(nx) * (bbox[0, 1] - bbox[0, 0]) + bbox[0, 0]
    gy = np.arange(ny + 1) / float(ny) * (bbox[1, 1] - bbox[1, 0]) + bbox[1, 0]
    gz = np.arange(nz + 1) / float(nz) * (bbox[2, 1] - bbox[2, 0]) + bbox[2, 0]
    X, Y, Z = np.meshgrid(gx, gy, gz, indexing="ij")
    V = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    ii, jj, kk = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    ii, jj, kk = ii.ravel(), jj.ravel(), kk.ravel()
    cols = []
    for c in range(8):
        di, dj, dk = c & 1, (c >> 1) & 1, (c >> 2) & 1
        cols.append(((ii + di) * (ny + 1) + (jj + dj)) * (nz + 1) + (kk + dk))
    return V, np.stack(cols, axis=1).astype(np.int64)


# --------------------------------------------------------------------------------------
# Assembly-map builders and gather / scatter      src/fe_jax/setup.py:212-246,304-356
# --------------------------------------------------------------------------------------

def build_row_ind(n_vertices, cells):
    """uint64 (V+1,) exclusive scan of vertex valence.  setup.py:212-224."""
    cnt = np.bincount(np.asarray(cells).astype(np.int64).ravel(), minlength=n_vertices)
    out = np.zeros(n_vertices + 1, dtype=np.uint64)
    out[1:] = np.cumsum(cnt).astype(np.uint64)
    return out


def build_col_ind_and_data(csr_row_ind, cells):
    """For every vertex the slots e*N+n that touch it, ascending; data = 1.0.
    setup.py:227-246 (a stable counting sort of slot ids by vertex)."""
    flat = np.asarray(cells).astype(np.int64).ravel()
    col = np.argsort(flat, kind="stable").astype(np.uint64)
    assert col.shape[0] == int(csr_row_ind[-1])
    return col, np.ones(col.shape[0], dtype=np.float64)


def gather_global_to_element_node(cells, v_g, U):
    """u_enu[e,n,:] = v[conn[e,n],:].  setup.py:304-322."""
    return np.asarray(v_g).reshape(-1, U)[np.asarray(cells).astype(np.int64)]


def scatter_element_node_to_global_sum(row_ind, col_ind, v_enu):
    """Row-ordered segmented sum (ascending slot e*N+n).  setup.py:344-356."""
    E, N, U = v_enu.shape
    flat = v_enu.reshape(E * N, U)[col_ind.astype(np.int64)]
    V = row_ind.shape[0] - 1
    out = np.zeros((V, U))
    starts = row_ind[:-1].astype(np.int64)
    nonempty = row_ind[1:] > row_ind[:-1]
    if flat.shape[0]:
        out[nonempty] = np.add.reduceat(flat, starts[nonempty], axis=0)
    return out.reshape(V * U)


def get_dof_map(cells, U):
    """dof[e, n*U+j] = U*conn[e,n]+j (int64).  fea.py:325-345."""
    c = np.asarray(cells).astype(np.int64)
    return (U * c[:, :, None] + np.arange(U)[None, None, :]).reshape(c.shape[0], -1)


# --------------------------------------------------------------------------------------
# Constitutive laws        src/fe_jax/linear_elasticity.py:9-142, hyperelasticity.py:9-91
# --------------------------------------------------------------------------------------

def _iso_C(E, nu, D):
    G = 0.5 * E / (1.0 + nu)
    z = np.zeros_like(nu)
    if D == 2:
        comp = np.array([[1.0 / E, -nu / E, z], [-nu / E, 1.0 / E, z], [z, z, 1.0 / G]])
    elif D == 3:
        comp = np.array([
            [1.0 / E, -nu / E, -nu / E, z, z, z],
            [-nu / E, 1.0 / E, -nu / E, z, z, z],
            [-nu / E, -nu / E, 1.0 / E, z, z, z],
            [z, z, z, 1.0 / G, z, z],
            [z, z, z, z, 1.0 / G, z],
            [z, z, z, z, z, 1.0 / G]])
    else:
        raise RuntimeError("1-D not restated")
    comp = np.moveaxis(comp, (0, 1), (-2, -1))
    return np.linalg.inv(comp)


def elastic_isotropic(eps, mp):
    """sigma = inv(compliance) : eps, plane stress in 2-D.  eps (...,D,D),
    mp (...,2) = (E, nu).  linear_elasticity.py:9-59."""
    C = _iso_C(mp[..., 0], mp[..., 1], eps.shape[-1])
    return rank2_voigt_to_tensor(np.einsum("...si,...i->...s", C, rank2_tensor_to_voigt(eps)))


def elastic_orthotropic(eps, mp):
    """linear_elasticity.py:62-142.  2-D params (Exx,Eyy,nuxy,Gxy); 3-D params
    (Exx,Eyy,Ezz,nuxy,nuyz,nuxz,Gxy,Gyz,Gxz)."""
    D = eps.shape[-1]
    z = np.zeros_like(mp[..., 0])
    if D == 2:
        Exx, Eyy, nxy, Gxy = (mp[..., k] for k in range(4))
        comp = np.array([[1 / Exx, -nxy / Exx, z], [-nxy / Exx, 1 / Eyy, z], [z, z, 1 / Gxy]])
    else:
        Exx, Eyy, Ezz, nxy, nyz, nxz, Gxy, Gyz, Gxz = (mp[..., k] for k in range(9))
        comp = np.array([
            [1 / Exx, -nxy / Exx, -nxz / Exx, z, z, z],
            [-nxy / Exx, 1 / Eyy, -nyz / Eyy, z, z, z],
            [-nxz / Exx, -nyz / Eyy, 1 / Ezz, z, z, z],
            [z, z, z, 1 / Gyz, z, z],
            [z, z, z, z, 1 / Gxz, z],
            [z, z, z, z, z, 1 / Gxy]])
    C = np.linalg.inv(np.moveaxis(comp, (0, 1), (-2, -1)))
    return rank2_voigt_to_tensor(np.einsum("...si,...i->...s", C, rank2_tensor_to_voigt(eps)))


def st_venant_kirchhoff(F, mp):
    """S = C : 1/2 (F^T F - I), used as is (no F.S).  hyperelasticity.py:9-53."""
    D = F.shape[-1]
    C = _iso_C(mp[..., 0], mp[..., 1], D)
    egl = 0.5 * (np.einsum("...di,...dj->...ij", F, F) - np.eye(D))
    return rank2_voigt_to_tensor(np.einsum("...si,...i->...s", C, rank2_tensor_to_voigt(egl)))


def mooney_rivlin(F, mp):
    """hyperelasticity.py:56-91, literally (S(I) = 4*C2*I)."""
    C1, C2, D1 = mp[..., 0], mp[..., 1], mp[..., 2]
    J = np.linalg.det(F)
    a = J ** (-1.0 / 3.0)
    Fb = a[..., None, None] * F
    Bb = np.einsum("...dj,...dk->...jk", Fb, Fb)
    I1 = np.einsum("...dd->...", Bb)
    I2 = 0.5 * (I1 ** 2 - np.einsum("...ij,...ji->...", Bb, Bb))
    BF = np.einsum("...jd,...dk->...jk", Bb, Fb)
    Finv = np.linalg.inv(Fb)
    c0 = 2.0 / D1 * a ** (-2) * (J - 1.0) - 2.0 / 3.0 * a * (C1 * I1 + 2.0 * C2 * I2)
    c1 = 2.0 * a * (C1 + I1 * C2)
    c2 = 2.0 * a * C2
    return c0[..., None, None] * Finv + c1[..., None, None] * Fb + c2[..., None, None] * BF


LINEAR_MODELS = {"elastic_isotropic": elastic_isotropic, "elastic_orthotropic": elastic_orthotropic}
HYPER_MODELS = {"st_venant_kirchhoff": st_venant_kirchhoff, "mooney_rivlin": mooney_rivlin}


# --------------------------------------------------------------------------------------
# Element residuals (batched over e)   linear_elasticity.py:145-206, hyperelasticity.py:93-139
# --------------------------------------------------------------------------------------

def _geometry(x_end, dphi_qnp):
    """J[e,q,p,d] = sum_n x[n,d] dphi[q,n,p]; g = dphi/dx; detJ (signed).
    linear_elasticity.py:180-184."""
    J = np.einsum("end,qnp->eqpd", x_end, dphi_qnp)
    G = np.swapaxes(np.linalg.inv(J), -1, -2)                 # G[p,d] = inv(J)[d,p]
    detJ = np.linalg.det(J)
    g = np.einsum("eqpd,qnp->eqnd", G, dphi_qnp)
    return g, detJ


def expand_material(mp, E, Q):
    """Material layouts (fea.py:34-38): (M,), (E,M) or (E,Q,M) -> (E,Q,M)."""
    mp = np.asarray(mp)
    if mp.ndim == 1:
        return np.broadcast_to(mp, (E, Q, mp.shape[0]))
    if mp.ndim == 2:
        return np.broadcast_to(mp[:, None, :], (E, Q, mp.shape[1]))
    return mp


def weights_as_reference(W_q, reference_quirk=True):
    """fea.py:285-297 hands the kernel W[0:1]; einsum('q,q->q') broadcasts it
    (SURVEY.md 0.5).  reference_quirk=False uses the full rule."""
    W_q = np.asarray(W_q, dtype=np.float64)
    return np.full_like(W_q, W_q[0]) if reference_quirk else W_q


def element_residuals(kind, model, u_end, x_end, dphi_qnp, W_q, mp_eqm):
    """R_end for all elements.  kind in {'linear', 'hyper'}; W_q is the full (Q,)
    vector actually applied (call weights_as_reference first)."""
    g, detJ = _geometry(x_end, dphi_qnp)
    wdet = detJ * W_q[None, :]
    if kind == "linear":
        H = np.einsum("eqnd,eni->eqid", g, u_end)
        eps = 0.5 * (H + np.swapaxes(H, -1, -2))
        S = LINEAR_MODELS[model](eps, mp_eqm)
    else:
        Fd = np.einsum("eqnd,eni->eqid", g, u_end + x_end)
        S = HYPER_MODELS[model](Fd, mp_eqm)
    t = np.einsum("eqni,eqid->eqnd", g, S)
    return np.einsum("eqnd,eq->end", t, wdet)


def element_tangents(kind, model, u_end, x_end, dphi_qnp, W_q, mp_eqm):
    """K_e = d vec(R_e) / d vec(u_e), (E, N*U, N*U), what vmap(jacfwd) yields
    (fea.py:519-572).  Linear kinds: exact columns R(e_j) (R is linear in u);
    hyper kinds: complex-step derivative (exact to rounding)."""
    E, N, D = u_end.shape
    NU = N * D
    K = np.zeros((E, NU, NU))
    if kind == "linear":
        zero = np.zeros_like(u_end)
        for j in range(NU):
            du = zero.copy()
            du.reshape(E, NU)[:, j] = 1.0
            K[:, :, j] = element_residuals(kind, model, du, x_end, dphi_qnp, W_q, mp_eqm).reshape(E, NU)
        return K
    h = 1e-30
    for j in range(NU):
        uc = u_end.astype(np.complex128).copy()
        uc.reshape(E, NU)[:, j] += 1j * h
        Rc = element_residuals(kind, model, uc, x_end.astype(np.complex128), dphi_qnp, W_q,
                               mp_eqm.astype(np.complex128))
        K[:, :, j] = Rc.reshape(E, NU).imag / h
    return K


# --------------------------------------------------------------------------------------
# A problem description mirroring ElementBatch / ElementBatchCollection (fea.py:18-487)
# --------------------------------------------------------------------------------------

class Batch:
    def __init__(self, cells, kind, model, mp, dphi_qnp, W_q, reference_weights=True):
        self.cells = np.asarray(cells).astype(np.int64)
        self.kind, self.model = kind, model
        self.dphi = np.asarray(dphi_qnp, dtype=np.float64)
        self.W = weights_as_reference(W_q, reference_weights)
        self.mp = expand_material(mp, self.cells.shape[0], self.dphi.shape[0]).astype(np.float64)


class Problem:
    """vertices (V,D) + batches; provides the global operations of fea.py."""

    def __init__(self, vertices_vd, batches, U=None):
        self.x = np.asarray(vertices_vd, dtype=np.float64)
        self.V, self.D = self.x.shape
        self.U = self.D if U is None else U
        self.F = self.V * self.U
        self.batches = batches
        self.maps = []
        for b in batches:
            ri = build_row_ind(self.V, b.cells)
            ci, _ = build_col_ind_and_data(ri, b.cells)
            self.maps.append((ri, ci))

    # fea.py:850-919
    def residual_wo_dirichlet(self, u_f):
        R = np.zeros(self.F)
        for b, (ri, ci) in zip(self.batches, self.maps):
            u_e = gather_global_to_element_node(b.cells, u_f, self.U)
            R_e = element_residuals(b.kind, b.model, u_e, self.x[b.cells], b.dphi, b.W, b.mp)
            R = R + scatter_element_node_to_global_sum(ri, ci, R_e)
        return R

    # fea.py:922-965
    def residual_w_dirichlet(self, u_f, values_g, mask_g):
        u_w = (1.0 - mask_g) * u_f + mask_g * values_g
        R = self.residual_wo_dirichlet(u_w)
        return (1.0 - mask_g) * R + mask_g * (u_f - values_g)

    # fea.py:575-610 (rows[e,i,j] = dof[e,i], cols[e,i,j] = dof[e,j])
    def jacobian_coo_terms(self, u_f):
        vals, rows, cols = [], [], []
        for b in self.batches:
            u_e = gather_global_to_element_node(b.cells, u_f, self.U)
            K = element_tangents(b.kind, b.model, u_e, self.x[b.cells], b.dphi, b.W, b.mp)
            dof = get_dof_map(b.cells, self.U)
            NU = dof.shape[1]
            rows.append(np.repeat(dof[:, :, None], NU, axis=2).ravel())
            cols.append(np.repeat(dof[:, None, :], NU, axis=1).ravel())
            vals.append(K.ravel())
        return np.concatenate(vals), np.concatenate(rows), np.concatenate(cols)

    # fea.py:490-516
    def jacobian_unique_nnz(self):
        rows, cols = [], []
        for b in self.batches:
            dof = get_dof_map(b.cells, self.U)
            NU = dof.shape[1]
            rows.append(np.repeat(dof[:, :, None], NU, axis=2).ravel())
            cols.append(np.repeat(dof[:, None, :], NU, axis=1).ravel())
        r, c = np.concatenate(rows), np.concatenate(cols)
        perm = np.lexsort((c, r))
        r, c = r[perm], c[perm]
        uniq = np.append(True, (np.diff(r) != 0) | (np.diff(c) != 0))
        return int(uniq.sum())

    # fea.py:613-660
    def jacobian_wo_dirichlet(self, u_f):
        data, row, col = self.jacobian_coo_terms(u_f)
        return coo_sum_duplicates(data, row, col)

    # fea.py:759-797
    def jacobian_diag_wo_dirichlet(self, u_f):
        diag = np.zeros(self.F)
        for b in self.batches:
            u_e = gather_global_to_element_node(b.cells, u_f, self.U)
            K = element_tangents(b.kind, b.model, u_e, self.x[b.cells], b.dphi, b.W, b.mp)
            dof = get_dof_map(b.cells, self.U)
            np.add.at(diag, dof.ravel(), np.einsum("eii->ei", K).ravel())
        return diag


# --------------------------------------------------------------------------------------
# Sparse helpers                                      src/fe_jax/sparse_matrix.py
# --------------------------------------------------------------------------------------

def coo_sum_duplicates(data, row, col):
    """Row-then-column sorted unique (row, col) with summed data; duplicates are
    added in stable order (ascending original position), which is what lexsort
    + bincount produce on CPU.  sparse_matrix.py:44-110."""
    perm = np.lexsort((col, row))
    r, c, d = row[perm], col[perm], data[perm]
    uniq = np.append(True, (np.diff(r) != 0) | (np.diff(c) != 0))
    inv = np.cumsum(uniq) - 1
    out = np.bincount(inv, weights=d, minlength=int(uniq.sum()))
    return out, r[uniq].astype(np.int64), c[uniq].astype(np.int64)


def coo_to_csr(data, row, col, n_rows):
    """indptr = [0, cumsum(bincount(rows))], indices = cols.  sparse_matrix.py:113-153."""
    indptr = np.concatenate([[0], np.cumsum(np.bincount(row, minlength=n_rows))]).astype(np.int64)
    return data, col.astype(np.int64), indptr


def apply_dirichlet_bcs_lhs(data, row, col, dirichlet_dofs):
    """Zero constrained rows / columns, one on their diagonals.  sparse_matrix.py:157-184."""
    rm = np.isin(row, dirichlet_dofs)
    cm = np.isin(col, dirichlet_dofs)
    out = np.where(~(rm | cm), data, 0.0)
    return np.where((row == col) & rm, 1.0, out)


def apply_dirichlet_bcs_rhs(matvec, b, dirichlet_dofs, dirichlet_values):
    """sparse_matrix.py:187-202."""
    tmp = np.zeros_like(b)
    tmp[dirichlet_dofs] = dirichlet_values
    out = b - matvec(tmp)
    out[dirichlet_dofs] = dirichlet_values
    return out


def csr_matvec(data, indices, indptr, x):
    prod = data * x[indices]
    out = np.zeros(indptr.shape[0] - 1)
    nz = indptr[1:] > indptr[:-1]
    if prod.shape[0]:
        out[nz] = np.add.reduceat(prod, indptr[:-1][nz])
    return out


# --------------------------------------------------------------------------------------
# CG                                                   src/fe_jax/solve_cg.py:20-63
# --------------------------------------------------------------------------------------

def cg(A, b, x0=None, *, tol=1e-5, atol=0.0, maxiter=None, M=None):
    """Returns (x, {'iterations', 'residual_norm_history' (maxiter,)}).  The write
    hist[k+1] at k+1 == maxiter is dropped (JAX out-of-bounds update)."""
    matvec = A if callable(A) else (lambda v: A @ v)
    prec = (lambda v: v) if M is None else (M if callable(M) else (lambda v: M @ v))
    b = np.asarray(b, dtype=np.float64)
    x = np.zeros_like(b) if x0 is None else np.array(x0, dtype=np.float64)
    if maxiter is None:
        maxiter = 10 * b.size
    atol2 = max(tol * tol * float(b @ b), atol * atol)
    r = b - matvec(x)
    z = prec(r)
    p = z.copy()
    gamma = float(r @ z)
    hist = np.zeros(maxiter)
    if maxiter > 0:
        hist[0] = np.linalg.norm(r)
    k = 0
    while True:
        rs = gamma if M is None else float(r @ r)
        if not (rs > atol2 and k < maxiter):
            break
        Ap = matvec(p)
        alpha = gamma / float(p @ Ap)
        x = x + alpha * p
        r = r - alpha * Ap
        z = prec(r)
        gamma_new = float(r @ z)
        beta = gamma_new / gamma
        p = z + beta * p
        if k + 1 < maxiter:
            hist[k + 1] = np.linalg.norm(r)
        gamma = gamma_new
        k += 1
    return x, {"iterations": k, "residual_norm_history": hist}


# --------------------------------------------------------------------------------------
# linear_solve (CG branches) and the Newton loop
#   sparse_linear_solve.py:202-529, fea.py:968-1129, 1132-1309
# --------------------------------------------------------------------------------------

DEFAULT_OPTIONS = dict(precond="NONE", linear_max_iter=1000, linear_relative_tol=1e-14,
                       linear_absolute_tol=1e-10, nonlinear_max_iter=10,
                       nonlinear_relative_tol=1e-10, nonlinear_absolute_tol=1e-8)


def dirichlet_arrays(F, D, dirichlet_bcs, dirichlet_values):
    """fea.py:1219-1225 (dof = D*vertex + component)."""
    bcs = np.asarray(dirichlet_bcs).astype(np.int64)
    dofs = D * bcs[:, 0] + bcs[:, 1]
    values_g = np.zeros(F)
    values_g[dofs] = dirichlet_values
    mask_g = np.zeros(F)
    mask_g[dofs] = 1.0
    return dofs, values_g, mask_g


def jvp_operator(prob, x0, values_g, mask_g):
    """z -> d/dx R_w(x0) . z = (I-M) J(u_w) (I-M) z + M z, with the exact element
    tangents (sparse_linear_solve.py:251-258 builds it with jax.jvp)."""
    u_w = (1.0 - mask_g) * x0 + mask_g * values_g
    data, row, col = prob.jacobian_wo_dirichlet(u_w)
    d, idx, ptr = coo_to_csr(data, row, col, prob.F)

    def op(z):
        return (1.0 - mask_g) * csr_matvec(d, idx, ptr, (1.0 - mask_g) * z) + mask_g * z
    return op


def linear_solve(prob, x0, dofs, values, values_g, mask_g, opts, info):
    """CG_JAX_SCIPY_W_INFO with NONE / JACOBI.  sparse_linear_solve.py:287-306,
    349-376, 527."""
    R0 = prob.residual_w_dirichlet(x0, values_g, mask_g)
    dx = np.zeros_like(R0)
    dx[dofs] = values - x0[dofs]
    M = None
    if opts["precond"] == "JACOBI":
        diag = prob.jacobian_diag_wo_dirichlet(x0)
        diag[dofs] = 1.0
        M = lambda v: v / diag
    A = jvp_operator(prob, x0, values_g, mask_g)
    dx, cg_info = cg(A, -R0, x0=dx, M=M, tol=opts["linear_relative_tol"],
                     atol=opts["linear_absolute_tol"], maxiter=opts["linear_max_iter"])
    k = cg_info["iterations"]
    start = info["cumulative_linear_iterations"] + info["nonlinear_iterations"]
    hist = info["cumulative_residual_norm_history"]
    n = min(opts["linear_max_iter"], hist.shape[0] - start)
    hist[start:start + n] = cg_info["residual_norm_history"][:n]
    info["linear_iterations_per_nonlinear_iteration"][info["nonlinear_iterations"]] = k
    info["cumulative_linear_iterations"] += k
    dx[dofs] = values - x0[dofs]
    return dx, info


def solve_nonlinear_step(prob, u0, dofs, values, values_g, mask_g, options=None):
    """fea.py:968-1129.  Returns (u, R, relative_error, info)."""
    opts = dict(DEFAULT_OPTIONS)
    opts.update(options or {})
    info = dict(nonlinear_iterations=0, cumulative_linear_iterations=0,
                linear_iterations_per_nonlinear_iteration=np.zeros(opts["nonlinear_max_iter"]),
                cumulative_residual_norm_history=np.zeros(
                    opts["linear_max_iter"] * opts["nonlinear_max_iter"] + 1))
    u = np.array(u0, dtype=np.float64)
    R = prob.residual_w_dirichlet(u, values_g, mask_g)
    r0 = np.linalg.norm(R)
    it = 0
    while True:
        a = np.linalg.norm(R)
        with np.errstate(invalid="ignore", divide="ignore"):
            rel = a / r0
        if not (it < opts["nonlinear_max_iter"] and rel > opts["nonlinear_relative_tol"]
                and a > opts["nonlinear_absolute_tol"]):
            break
        du, info = linear_solve(prob, u, dofs, values, values_g, mask_g, opts, info)
        u = u + du
        R = prob.residual_w_dirichlet(u, values_g, mask_g)
        info["nonlinear_iterations"] += 1
        it += 1
    a = np.linalg.norm(R)
    with np.errstate(invalid="ignore", divide="ignore"):
        rel = a / r0
    return u, R, rel, info


def solve_bvp(vertices_vd, batches, u0, dirichlet_bcs, dirichlet_values, options=None):
    """fea.py:1132-1309.  Returns (u, R, info)."""
    prob = Problem(vertices_vd, batches)
    if u0 is None:
        u0 = np.zeros(prob.F)
    dofs, values_g, mask_g = dirichlet_arrays(prob.F, prob.D, dirichlet_bcs, dirichlet_values)
    u, R, rel, info = solve_nonlinear_step(prob, u0, dofs, np.asarray(dirichlet_values, dtype=np.float64),
                                           values_g, mask_g, options)
    return u, R, info


# --------------------------------------------------------------------------------------
# Synthetic workload of BASELINE.json / SURVEY.md 8(d)
# --------------------------------------------------------------------------------------

def synthetic_hex_problem(n, material_layout="M", perturb=0.2, seed=0):
    """uniform_hex_grid(n,n,n) on the unit cube, interior nodes jittered by
    perturb*h*U(-1,1) (rng seed), u = 1e-3*U(0,1) (seed+1), E = 10e9+90e9*U(0,1)
    (seed+2), nu = 0.25; Dirichlet: face x=0 fixed, face x=1 u_x = 0.01."""
    Vx, cells = uniform_hex_grid(n, n, n)
    h = 1.0 / n
    rng = np.random.default_rng(seed)
    jitter = perturb * h * rng.uniform(-1.0, 1.0, size=Vx.shape)
    interior = np.all((Vx > 1e-12) & (Vx < 1.0 - 1e-12), axis=1)
    Vx = Vx + jitter * interior[:, None]
    u = 1e-3 * np.random.default_rng(seed + 1).uniform(0.0, 1.0, size=Vx.size)
    E = cells.shape[0]
    rm = np.random.default_rng(seed + 2)
    if material_layout == "EQM":
        mp = np.empty((E, 8, 2))
        mp[..., 0] = 10e9 + 90e9 * rm.uniform(0.0, 1.0, size=(E, 8))
        mp[..., 1] = 0.25
    elif material_layout == "EM":
        mp = np.empty((E, 2))
        mp[:, 0] = 10e9 + 90e9 * rm.uniform(0.0, 1.0, size=E)
        mp[:, 1] = 0.25
    else:
        mp = np.array([70e9, 0.25])
    left = np.nonzero(Vx[:, 0] < 1e-12)[0]
    right = np.nonzero(Vx[:, 0] > 1.0 - 1e-12)[0]
    bcs = np.concatenate([
        np.stack([np.repeat(left, 3), np.tile(np.arange(3), left.shape[0])], axis=1),
        np.stack([right, np.zeros_like(right)], axis=1)]).astype(np.int64)
    vals = np.concatenate([np.zeros(3 * left.shape[0]), np.full(right.shape[0], 0.01)])
    return Vx, cells, u, mp, bcs, vals
