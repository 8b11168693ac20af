"""Timing of the element kernels on quadratic elements (SURVEY config 4): python profiles/profile_quadratic.py"""
import os, sys, time
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path[:0] = [ROOT, os.path.join(ROOT, "fea-in-jax_b200"), os.path.join(ROOT, "tests")]
os.environ.setdefault("FEB200_QUIET", "1")
import numpy as np, torch
import fe_jax as fj
from fe_jax import _lib
from helpers import lattice_mesh, material_params
def run(cell, degree, qdeg, n, model, token):
    pts, cells = lattice_mesh(cell, degree, n)
    E = cells.shape[0]
    D = pts.shape[1]
    fe = fj.FiniteElementType(getattr(fj.CellType, cell), fj.ElementFamily.P, degree, fj.LagrangeVariant.equispaced, fj.QuadratureType.default, qdeg)
    mp = material_params(model.__name__, D, (), seed=1)
    batch = fj.ElementBatch(fe_type=fe, n_dofs_per_basis=D, connectivity_en=cells, constitutive_model=model, material_params=mp)
    ebc = fj.batch_to_collection(pts, [batch]); maps = [fj.mesh_to_sparse_assembly_map(pts.shape[0], cells)]
    u = torch.tensor(1e-5 * np.random.default_rng(0).uniform(size=pts.size), device="cuda")
    nnz = fj._calculate_jacobian_unique_nnz(pts.shape[0], ebc)
    for f, name in ((lambda: fj.calculate_residual_wo_dirichlet(token, ebc, maps, u), "residual"), (lambda: fj.calculate_jacobian_wo_dirichlet(token, ebc, maps, u, nnz), "jacobian")):
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): f()
        e1.record(); torch.cuda.synchronize()
        print(f"{cell} deg{degree} {model.__name__} E={E} {name}: {e0.elapsed_time(e1)/5:.3f} ms")
run("tetrahedron", 2, 4, 28, fj.elastic_isotropic, fj.linear_elasticity_residual)
run("tetrahedron", 2, 4, 28, fj.st_venant_kirchhoff, fj.hyperelasticity_residual)
run("quadrilateral", 2, 4, 300, fj.elastic_isotropic, fj.linear_elasticity_residual)
run("quadrilateral", 2, 4, 300, fj.st_venant_kirchhoff, fj.hyperelasticity_residual)
run("hexahedron", 2, 4, 20, fj.elastic_isotropic, fj.linear_elasticity_residual)
run("triangle", 2, 4, 400, fj.elastic_isotropic, fj.linear_elasticity_residual)
run("triangle", 2, 4, 400, fj.st_venant_kirchhoff, fj.hyperelasticity_residual)
