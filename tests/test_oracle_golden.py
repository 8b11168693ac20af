"""The CPU oracle against (a) outputs of the reference's own source text
(tests/golden/reference_outputs.npz, made by tests/golden/make_golden.py) and
(b) the known answers the reference's tests hold."""
import numpy as np
import pytest

from oracle import fe_oracle as orc

RTOL = 1e-12


def close(a, b, rtol=RTOL):
    a, b = np.asarray(a), np.asarray(b)
    scale = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() <= rtol * scale


def test_grids_bit_exact(golden):
    V, F = orc.uniform_quad_grid(5, 7, [[0.0, 2.0], [-1.0, 1.0]])
    assert np.array_equal(F, golden["quad_grid_F"]) and np.array_equal(V, golden["quad_grid_V"])
    V, F = orc.uniform_tri_grid(7, 10)
    assert np.array_equal(F, golden["tri_grid_F"]) and np.array_equal(V, golden["tri_grid_V"])


@pytest.mark.parametrize("tag", ["tri", "quad", "hex"])
def test_index_maps_bit_exact(golden, tag):
    cells = golden[f"maps_{tag}_cells"]
    ri = orc.build_row_ind(int(cells.max()) + 1, cells)
    ci, data = orc.build_col_ind_and_data(ri, cells)
    assert ri.dtype == np.uint64 and ci.dtype == np.uint64 and data.dtype == np.float64
    assert np.array_equal(ri, golden[f"maps_{tag}_row_ind"])
    assert np.array_equal(ci, golden[f"maps_{tag}_col_ind"])
    assert np.all(data == 1.0)


def test_mesh_to_element_coords(golden):
    V, F = orc.uniform_tri_grid(7, 10)
    assert np.array_equal(V[F][:, :, :2], golden["mesh_to_jax_tri"])


def test_voigt(golden):
    assert np.array_equal(orc.rank2_tensor_to_voigt(golden["voigt_t2"]), golden["voigt_v2"])
    assert np.array_equal(orc.rank2_tensor_to_voigt(golden["voigt_t3"]), golden["voigt_v3"])
    assert np.array_equal(orc.rank2_voigt_to_tensor(golden["voigt_v2"]), golden["voigt_back2"])
    assert np.array_equal(orc.rank2_voigt_to_tensor(golden["voigt_v3"]), golden["voigt_back3"])


@pytest.mark.parametrize("D", [2, 3])
def test_constitutive_laws(golden, D):
    g = golden
    assert close(orc.elastic_isotropic(g[f"mat{D}_eps"], g[f"mat{D}_iso_mp"]), g[f"mat{D}_iso_sig"])
    assert close(orc.elastic_orthotropic(g[f"mat{D}_eps"], g[f"mat{D}_orth_mp"]), g[f"mat{D}_orth_sig"])
    assert close(orc.st_venant_kirchhoff(g[f"mat{D}_F"], g[f"mat{D}_iso_mp"]), g[f"mat{D}_svk_S"])
    assert close(orc.mooney_rivlin(g[f"mat{D}_F"], g[f"mat{D}_mr_mp"]), g[f"mat{D}_mr_S"])


@pytest.mark.parametrize("tag", ["tri3", "quad4", "hex8", "tet4"])
def test_element_residuals(golden, tag):
    g = golden
    x, u, dphi, W = g[f"res_{tag}_x"], g[f"res_{tag}_u"], g[f"res_{tag}_dphi"], g[f"res_{tag}_W"]
    Wr = orc.weights_as_reference(W)
    R = orc.element_residuals("linear", "elastic_isotropic", u[None], x[None], dphi, Wr, g[f"res_{tag}_iso_mp"][None])
    assert close(R[0], g[f"res_{tag}_lin_iso_R"])
    uh = g[f"res_{tag}_uh"]
    R = orc.element_residuals("hyper", "st_venant_kirchhoff", uh[None], x[None], dphi, Wr, g[f"res_{tag}_iso_mp"][None])
    assert close(R[0], g[f"res_{tag}_hyp_svk_R"])
    R = orc.element_residuals("hyper", "mooney_rivlin", uh[None], x[None], dphi, Wr, g[f"res_{tag}_mr_mp"][None])
    assert close(R[0], g[f"res_{tag}_hyp_mr_R"])


def test_coo_sum_duplicates_and_dirichlet(golden):
    g = golden
    d, r, c = orc.coo_sum_duplicates(g["coo_in_data"], g["coo_in_row"], g["coo_in_col"])
    for tag in ("out", "jit"):
        assert np.array_equal(r, g[f"coo_{tag}_row"]) and np.array_equal(c, g[f"coo_{tag}_col"])
        assert close(d, g[f"coo_{tag}_data"], 1e-15)
    out = orc.apply_dirichlet_bcs_lhs(d, r, c, g["dirichlet_dofs"])
    assert np.array_equal(out, g["dirichlet_lhs_data"])


@pytest.mark.parametrize("tag,maxiter", [("cg3", 30), ("cgspd", 12)])
@pytest.mark.parametrize("prec", ["none", "jacobi"])
def test_cg_recurrence(golden, tag, maxiter, prec):
    A, b = golden[f"{tag}_A"], golden[f"{tag}_b"]
    M = None if prec == "none" else (lambda v: v / np.diag(A))
    x, info = orc.cg(A, b, x0=np.zeros_like(b), tol=1e-12, atol=0.0, maxiter=maxiter, M=M)
    assert info["iterations"] == int(golden[f"{tag}_{prec}_iters"])
    assert close(x, golden[f"{tag}_{prec}_x"], 1e-13)
    assert close(info["residual_norm_history"], golden[f"{tag}_{prec}_hist"], 1e-10)


def test_cg_known_answer():
    # tests/test_solve_cg.py:46-59 : x = [3, 4, 1]
    A = np.array([[2.0, -1.0, 0.0], [-1.0, 2.0, 0.0], [0.0, 0.0, 1.0]])
    b = np.array([2.0, 5.0, 1.0])
    for M in (None, np.diag(1.0 / np.diag(A)), np.linalg.inv(A)):
        x, info = orc.cg(A, b, tol=1e-12, atol=0.0, M=M)
        assert np.allclose(x, [3.0, 4.0, 1.0], rtol=0, atol=1e-12)


def test_basis_tables_triangle_pinned():
    # tests/test_basix_interface.py:59-81
    xi, W = orc.get_quadrature("triangle", 3)
    assert xi.shape == (6, 2) and W.shape == (6,) and np.isclose(W.sum(), 0.5)
    phi, dphi = orc.eval_basis_and_derivatives("triangle", 1, xi)
    ref = np.stack([1 - xi[:, 0] - xi[:, 1], xi[:, 0], xi[:, 1]], axis=1)
    assert np.allclose(phi, ref)
    assert np.allclose(dphi, np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])[None])


@pytest.mark.parametrize("cell,deg,qdeg", [("triangle", 1, 3), ("triangle", 2, 4), ("quadrilateral", 1, 2),
                                           ("quadrilateral", 2, 4), ("tetrahedron", 1, 2), ("tetrahedron", 2, 4),
                                           ("hexahedron", 1, 2), ("hexahedron", 2, 4)])
def test_basis_tables_consistency(cell, deg, qdeg):
    xi, W = orc.get_quadrature(cell, qdeg)
    vol = {"triangle": 0.5, "quadrilateral": 1.0, "tetrahedron": 1 / 6, "hexahedron": 1.0}[cell]
    assert np.isclose(W.sum(), vol)
    nodes = orc.reference_nodes(cell, deg)
    phi, dphi = orc.eval_basis_and_derivatives(cell, deg, nodes)
    assert np.allclose(phi, np.eye(nodes.shape[0]), atol=1e-13)          # Kronecker property
    phi, dphi = orc.eval_basis_and_derivatives(cell, deg, xi)
    assert np.allclose(phi.sum(axis=1), 1.0) and np.allclose(dphi.sum(axis=1), 0.0, atol=1e-12)
    # linear field reproduced: sum_n x_n dphi_n/dxi = I
    assert np.allclose(np.einsum("nd,qnp->qpd", nodes, dphi), np.eye(nodes.shape[1])[None], atol=1e-12)
    # hex8 degree-2 rule: 8 points, weights 1/8
    if cell == "hexahedron" and qdeg == 2:
        assert W.shape == (8,) and np.allclose(W, 0.125)
        assert np.allclose(xi[1] - xi[0], [0, 0, xi[1, 2] - xi[0, 2]])    # z fastest


def _two_triangle_problem(layout="EQM"):
    # tests/test_simple_fea_solve.py:25-75
    pts = np.array([[3.0, 0.0], [3.0, 2.0], [0.0, 2.0], [0.0, 0.0]])
    cells = np.array([[0, 1, 3], [2, 3, 1]])
    xi, W = orc.get_quadrature("triangle", 3)
    _, dphi = orc.eval_basis_and_derivatives("triangle", 1, xi)
    mp = {"EQM": np.tile([30e6, 0.25], (1, 6, 1)), "EM": np.array([[30e6, 0.25]]), "M": np.array([30e6, 0.25])}[layout]
    batches = [orc.Batch(cells[:1], "linear", "elastic_isotropic", mp, dphi, W),
               orc.Batch(cells[1:], "linear", "elastic_isotropic", mp, dphi, W)]
    bcs = np.array([[0, 0], [0, 1], [1, 0], [2, 0], [3, 0], [3, 1]])
    vals = np.array([1.0, 0.0, 1.0, 0.0, 0.0, 0.0])
    return pts, batches, bcs, vals


@pytest.mark.parametrize("layout", ["EQM", "EM", "M"])
def test_two_triangle_bvp_known_answer(layout):
    pts, batches, bcs, vals = _two_triangle_problem(layout)
    u, R, info = orc.solve_bvp(pts, batches, None, bcs, vals)
    # tests/test_simple_fea_solve.py:109-113
    assert abs(u[0] - 1.0) < 1e-6 and abs(u[2] - 1.0) < 1e-6
    assert abs(-(u[3] / 2.0) / (u[0] / 3.0) - 0.25) < 1e-6
    assert abs(-(u[5] / 2.0) / (u[0] / 3.0) - 0.25) < 1e-6
    assert u[1] == u[7] == 0.0 and u[4] == u[6] == 0.0
    assert np.allclose(u, [1, 0, 1, -1 / 6, 0, -1 / 6, 0, 0], atol=1e-9)
    assert info["nonlinear_iterations"] >= 1


def test_textbook_values_fea_traditional():
    # tests/fea_traditional.py:312-315 (C), :320-330 (J, det J = 6), :339-359 (B)
    pts, batches, _, _ = _two_triangle_problem()
    C = orc._iso_C(np.array(30e6), np.array(0.25), 2)
    assert np.allclose(C, [[3.2e7, 0.8e7, 0.0], [0.8e7, 3.2e7, 0.0], [0.0, 0.0, 1.2e7]])
    for b, Bref in zip(batches, (
            [[2 / 6, 0, 0, 0, -2 / 6, 0], [0, -3 / 6, 0, 3 / 6, 0, 0], [-3 / 6, 2 / 6, 3 / 6, 0, 0, -2 / 6]],
            [[-2 / 6, 0, 0, 0, 2 / 6, 0], [0, 3 / 6, 0, -3 / 6, 0, 0], [3 / 6, -2 / 6, -3 / 6, 0, 0, 2 / 6]])):
        g, detJ = orc._geometry(pts[b.cells], b.dphi)
        assert np.allclose(detJ, 6.0)
        gn = g[0, 0]
        B = np.zeros((3, 6))
        B[0, 0::2], B[1, 1::2] = gn[:, 0], gn[:, 1]
        B[2, 0::2], B[2, 1::2] = gn[:, 1], gn[:, 0]
        assert np.allclose(B, Bref)


def test_global_K_times_u_equals_residual():
    # tests/fea_traditional.py:415-418
    pts, batches, _, _ = _two_triangle_problem()
    prob = orc.Problem(pts, batches)
    u = np.random.default_rng(0).random(prob.F)
    d, r, c = prob.jacobian_wo_dirichlet(np.zeros(prob.F))
    K = np.zeros((prob.F, prob.F))
    K[r, c] = d
    assert np.allclose(K @ u, prob.residual_wo_dirichlet(u), rtol=1e-12, atol=1e-3)
    assert np.allclose(K, K.T)
    assert prob.jacobian_unique_nnz() == d.shape[0]
    assert np.allclose(prob.jacobian_diag_wo_dirichlet(u), np.diag(K))


def test_scatter_gather_roundtrip():
    # tests/test_maps.py:27-34 : scatter-mean(gather(x)) == x
    V, F = orc.uniform_tri_grid(6, 9)
    ri = orc.build_row_ind(V.shape[0], F)
    ci, _ = orc.build_col_ind_and_data(ri, F)
    x_e = orc.gather_global_to_element_node(F, V[:, :2].ravel(), 2)
    s = orc.scatter_element_node_to_global_sum(ri, ci, x_e).reshape(-1, 2)
    val = (ri[1:] - ri[:-1]).astype(np.float64)
    assert np.allclose(s / val[:, None], V[:, :2])


def test_duplicate_sum_doubles_matrix():
    # tests/test_sparse_linalg.py:7-23
    rng = np.random.default_rng(100)
    n = 60
    mask = rng.random((n, n)) < 0.2
    r, c = np.nonzero(mask)
    d = rng.normal(size=r.shape[0])
    d2, r2, c2 = orc.coo_sum_duplicates(np.hstack([d, d]), np.hstack([r, r]), np.hstack([c, c]))
    assert np.array_equal(r2, r) and np.array_equal(c2, c) and np.allclose(d2, 2 * d)


def test_hex_pattern_nnz_formula():
    # nnz = 9 (3n+1)^3 on a uniform n^3 hex grid
    for n in (2, 3, 4):
        V, cells = orc.uniform_hex_grid(n, n, n)
        xi, W = orc.get_quadrature("hexahedron", 2)
        _, dphi = orc.eval_basis_and_derivatives("hexahedron", 1, xi)
        prob = orc.Problem(V, [orc.Batch(cells, "linear", "elastic_isotropic", np.array([1.0, 0.3]), dphi, W)])
        assert prob.jacobian_unique_nnz() == 9 * (3 * n + 1) ** 3


def test_hyper_tangent_complex_step_vs_fd():
    V, cells = orc.uniform_hex_grid(1, 1, 1)
    xi, W = orc.get_quadrature("hexahedron", 2)
    _, dphi = orc.eval_basis_and_derivatives("hexahedron", 1, xi)
    rng = np.random.default_rng(3)
    x = V[cells] + 0.05 * rng.normal(size=(1, 8, 3))
    u = 0.02 * rng.normal(size=(1, 8, 3))
    for model, mp in (("st_venant_kirchhoff", [70e9, 0.25]), ("mooney_rivlin", [0.5e6, 0.1e6, 2e-7])):
        mpe = orc.expand_material(np.array(mp), 1, 8)
        K = orc.element_tangents("hyper", model, u, x, dphi, W, mpe)[0]
        h = 1e-6
        Kfd = np.zeros_like(K)
        for j in range(24):
            du = np.zeros(24)
            du[j] = h
            Rp = orc.element_residuals("hyper", model, u + du.reshape(1, 8, 3), x, dphi, W, mpe)
            Rm = orc.element_residuals("hyper", model, u - du.reshape(1, 8, 3), x, dphi, W, mpe)
            Kfd[:, j] = (Rp - Rm).ravel() / (2 * h)
        assert np.abs(K - Kfd).max() <= 1e-6 * np.abs(K).max()
