"""The reference's own unstructured test meshes (tests/meshes/*.vtk of the reference, converted to
tests/golden/reference_meshes.npz by tests/golden/make_meshes.py) through the legacy-VTK reader and the
CUDA path: the scatter-mean / gather identity of tests/test_maps.py, the Dirichlet check of
tests/test_fea_solve.py:90 and assembly parity on the quadratic-triangle RVE mesh of tests/test_microscale_bvp.py."""
import os

import numpy as np
import pytest
import torch

from helpers import build_pair, material_params, rel_err
from oracle import fe_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("FEB200_QUIET", "1")


@pytest.fixture(scope="module")
def meshes():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_meshes.npz"))


def test_vtk_roundtrip_and_fixture(tmp_path, meshes):
    import __graft_entry__  # noqa: F401  (puts the package on sys.path)
    from fe_jax.vtk_io import read_vtk, write_vtk
    pts = meshes["polygon_mesh_0p5__points"]
    blocks = [("line", meshes["polygon_mesh_0p5__line"]), ("triangle", meshes["polygon_mesh_0p5__triangle"]),
              ("vertex", meshes["polygon_mesh_0p5__vertex"])]
    for binary in (True, False):
        f = tmp_path / f"m{int(binary)}.vtk"
        write_vtk(str(f), pts, blocks, binary=binary)
        p2, b2 = read_vtk(str(f))
        assert np.array_equal(p2, pts)
        assert [n for n, _ in b2] == ["line", "triangle", "vertex"]
        assert all(np.array_equal(x[1], y[1]) for x, y in zip(blocks, b2))
    src = "/root/reference/tests/meshes/polygon_mesh_0.5.vtk"
    if os.path.exists(src):                      # build container only
        p3, b3 = read_vtk(src)
        assert np.array_equal(p3, pts) and np.array_equal(dict(b3)["triangle"], meshes["polygon_mesh_0p5__triangle"])


@pytest.fixture(scope="module")
def fj():
    import __graft_entry__ as graft
    graft.build()
    assert torch.cuda.is_available()
    import fe_jax
    return fe_jax


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["polygon_mesh_0p5", "polygon_mesh_0p1", "polygon_mesh_0p05", "rectangle_with_hole"])
def test_maps_on_reference_meshes(fj, meshes, name):
    # tests/test_maps.py:27-34 of the reference
    points = meshes[name + "__points"]
    cells = meshes[name + "__triangle"].astype(np.uint64)
    A = fj.mesh_to_sparse_assembly_map(points.shape[0], cells)
    ri = orc.build_row_ind(points.shape[0], cells)
    ci, _ = orc.build_col_ind_and_data(ri, cells)
    assert np.array_equal(A.indptr.view(torch.int64).cpu().numpy()[0].astype(np.uint64), ri)
    assert np.array_equal(A.indices.view(torch.int64).cpu().numpy()[0].astype(np.uint64), ci)
    x_n = fj.mesh_to_jax(points, cells)
    back = fj.transform_element_node_to_global_nosum(A, x_n).cpu().numpy()
    used = np.unique(cells.astype(np.int64))
    assert np.allclose(back[used], points[used, :2], atol=1e-14)


@pytest.mark.gpu
def test_bvp_on_reference_polygon_mesh(fj, meshes):
    # problem of tests/test_fea_solve.py: random Dirichlet values on the boundary, per-point random modulus
    points = meshes["polygon_mesh_0p1__points"][:, :2]
    cells = meshes["polygon_mesh_0p1__triangle"]
    boundary = np.unique(meshes["polygon_mesh_0p1__line"])
    rng = np.random.default_rng(0)
    bcs = np.stack([np.repeat(boundary, 2), np.tile([0, 1], boundary.shape[0])], axis=1)
    vals = 0.001 * rng.random(bcs.shape[0])
    mp = material_params("elastic_isotropic", 2, (cells.shape[0], 6), seed=1)
    ebc, maps, token, prob, fb = build_pair(fj, points, [dict(cell="triangle", degree=1, qdeg=3, cells=cells,
                                                             model="elastic_isotropic", mp=mp)])
    opts = fj.SolverOptions(linear_precond_type=fj.PreconditionerType.JACOBI, linear_max_iter=3000)
    u, res, _ = fj.solve_bvp(points, fb, fj.linear_elasticity_residual, None, bcs, vals, solver_options=opts)
    dofs = 2 * bcs[:, 0] + bcs[:, 1]
    assert np.array_equal(u.cpu().numpy()[dofs], vals)                       # tests/test_fea_solve.py:90
    u_o, _, _ = orc.solve_bvp(points, prob.batches, None, bcs, vals, options={"precond": "JACOBI", "linear_max_iter": 3000})
    assert rel_err(u, u_o) <= 1e-8


@pytest.mark.gpu
def test_assembly_on_reference_rve_mesh_tri6(fj, meshes):
    # microscale_2D_r0.vtk: 11158 quadratic triangles (gmsh node order -> basix order), orthotropic material
    points = meshes["microscale_2D_r0__points"][:, :2]
    cells = meshes["microscale_2D_r0__triangle6"][:, [0, 1, 2, 4, 5, 3]]
    mp = material_params("elastic_orthotropic", 2, (cells.shape[0],), seed=2)
    ebc, maps, token, prob, _ = build_pair(fj, points, [dict(cell="triangle", degree=2, qdeg=4, cells=cells,
                                                            model="elastic_orthotropic", mp=mp)])
    u = 1e-4 * np.random.default_rng(3).normal(size=points.size)
    R, _ = fj.calculate_residual_wo_dirichlet(token, ebc, maps, u)
    assert rel_err(R, prob.residual_wo_dirichlet(u)) <= 1e-11
    nnz = fj._calculate_jacobian_unique_nnz(prob.V, ebc)
    assert nnz == prob.jacobian_unique_nnz()
    J = fj.calculate_jacobian_wo_dirichlet(token, ebc, maps, u, nnz)
    d_o, r_o, c_o = prob.jacobian_wo_dirichlet(u)
    assert np.array_equal(J.row.cpu().numpy(), r_o) and np.array_equal(J.col.cpu().numpy(), c_o)
    assert rel_err(J.data, d_o) <= 1e-11


@pytest.mark.gpu
def test_batched_rve_solves_match_individual_solves(fj):
    """Config 5 (vmap of many small BVPs): 12 copies of a quad4 RVE with per-copy modulus fields and boundary
    values, solved as one block-diagonal problem, against the same problems solved one by one."""
    V3, cells = fj.uniform_quad_grid(9, 9, np.array([[0.0, 1.0], [0.0, 1.0]]))
    pts = V3[:, :2]
    C, E = 12, cells.shape[0]
    rng = np.random.default_rng(3)
    mp = np.empty((C, E, 4, 2))
    mp[..., 0] = 10e9 + 90e9 * rng.uniform(size=(C, E, 4))
    mp[..., 1] = 0.25
    edge = np.nonzero((pts[:, 0] < 1e-12) | (pts[:, 0] > 1 - 1e-12))[0]
    bcs = np.stack([np.repeat(edge, 2), np.tile([0, 1], edge.shape[0])], axis=1)
    vals = 1e-3 * rng.uniform(size=(C, bcs.shape[0]))
    fe = fj.FiniteElementType(fj.CellType.quadrilateral, fj.ElementFamily.P, 1, fj.LagrangeVariant.equispaced,
                              fj.QuadratureType.default, 2)
    batch = fj.ElementBatch(fe_type=fe, n_dofs_per_basis=2, connectivity_en=cells, constitutive_model=fj.elastic_isotropic,
                            material_params=mp[0])
    opts = fj.SolverOptions(linear_precond_type=fj.PreconditionerType.JACOBI, linear_max_iter=4000)
    vals[3] *= 1e-3          # copies of different scale: every copy must meet ITS OWN relative tolerance
    u_b, res_b = fj.solve_bvp_batched(pts, [batch], fj.linear_elasticity_residual, bcs, vals, [mp], solver_options=opts)
    assert u_b.shape == (C, pts.size)
    k_b = fj.solve_bvp_batched.last_iterations[0]          # per-copy CG iterations of the first Newton iteration
    assert k_b.shape == (C,) and k_b.min() > 0
    for c in (0, 3, 5, 11):
        one = fj.ElementBatch(fe_type=fe, n_dofs_per_basis=2, connectivity_en=cells, constitutive_model=fj.elastic_isotropic,
                              material_params=mp[c])
        u1, _, _ = fj.solve_bvp(pts, [one], fj.linear_elasticity_residual, None, bcs, vals[c], solver_options=opts)
        assert rel_err(u_b[c], u1.cpu().numpy()) <= 1e-7
        k1 = int(fj.solve_bvp.last_info.linear_iterations_per_nonlinear_iteration[0])
        assert abs(int(k_b[c]) - k1) <= 1, (c, int(k_b[c]), k1)     # same iteration as the stand-alone solve (vmap semantics)


@pytest.mark.gpu
def test_batched_rve_solves_match_the_oracle_per_copy(fj):
    """Config 5 against the oracle (not against the GPU itself): 64 copies of a quad4 RVE with per-copy modulus fields and
    boundary values through solve_bvp_batched; six of them are solved one by one by orc.solve_bvp (the reference's Newton
    + CG, fea.py:1132-1309 / solve_cg.py:20-63) and must agree to the solver tolerance, iteration counts included."""
    V3, cells = fj.uniform_quad_grid(7, 7, np.array([[0.0, 1.0], [0.0, 1.0]]))
    pts = V3[:, :2]
    C, E = 64, cells.shape[0]
    rng = np.random.default_rng(5)
    mp = np.empty((C, E, 4, 2))
    mp[..., 0] = 10e9 + 90e9 * rng.uniform(size=(C, E, 4))
    mp[..., 1] = 0.25
    edge = np.nonzero((pts[:, 0] < 1e-12) | (pts[:, 0] > 1 - 1e-12))[0]
    bcs = np.stack([np.repeat(edge, 2), np.tile([0, 1], edge.shape[0])], axis=1)
    vals = 1e-3 * rng.uniform(size=(C, bcs.shape[0]))
    fe = fj.FiniteElementType(fj.CellType.quadrilateral, fj.ElementFamily.P, 1, fj.LagrangeVariant.equispaced,
                              fj.QuadratureType.default, 2)
    batch = fj.ElementBatch(fe_type=fe, n_dofs_per_basis=2, connectivity_en=cells, constitutive_model=fj.elastic_isotropic,
                            material_params=mp[0])
    opts = fj.SolverOptions(linear_precond_type=fj.PreconditionerType.JACOBI, linear_max_iter=4000)
    u_b, res_b = fj.solve_bvp_batched(pts, [batch], fj.linear_elasticity_residual, bcs, vals, [mp], solver_options=opts)
    k_b = fj.solve_bvp_batched.last_iterations[0]
    xi, W = orc.get_quadrature("quadrilateral", 2)
    _, dphi = orc.eval_basis_and_derivatives("quadrilateral", 1, xi)
    for c in (0, 7, 21, 40, 55, 63):
        ob = [orc.Batch(cells, "linear", "elastic_isotropic", mp[c], dphi, W)]
        u_o, R_o, info = orc.solve_bvp(pts, ob, None, bcs, vals[c], options={"precond": "JACOBI", "linear_max_iter": 4000})
        assert rel_err(u_b[c], u_o) <= 1e-7, c
        k_o = int(info["linear_iterations_per_nonlinear_iteration"][0])
        assert abs(int(k_b[c]) - k_o) <= 2, (c, int(k_b[c]), k_o)
