"""BASELINE configs[3] and [4] on one B200: quadratic elements with 27-point quadrature (hex27 32^3, tet10 from the 6-tet split
of 32^3) and thousands of small RVE solves (4096 copies of a 32 x 32 quad4 RVE with per-copy modulus fields).
python profiles/profile_config45.py [quadratic|batched|all]"""
import os, sys, time, json
ROOT = os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fea-in-jax_b200"), os.path.join(ROOT, "tests")]
os.environ.setdefault("FEB200_QUIET", "1")
import numpy as np, torch
import fe_jax as fj

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def quadratic(cell, n):
    from helpers import lattice_mesh, material_params
    pts, cells = lattice_mesh(cell, 2, n)
    E = cells.shape[0]
    fe = fj.FiniteElementType(getattr(fj.CellType, cell), fj.ElementFamily.P, 2, fj.LagrangeVariant.equispaced, fj.QuadratureType.default, 4)
    rng = np.random.default_rng(2)
    out = {"config": 4, "cell": cell, "degree": 2, "elements": E, "dofs": int(pts.size),
           "full_weights": os.environ.get("FEB200_FULL_WEIGHTS", "0") == "1"}
    for model, token in ((fj.elastic_isotropic, fj.linear_elasticity_residual), (fj.st_venant_kirchhoff, fj.hyperelasticity_residual)):
        mp = np.array([70e9, 0.25])
        batch = fj.ElementBatch(fe_type=fe, n_dofs_per_basis=3, connectivity_en=cells, constitutive_model=model, material_params=mp)
        ebc = fj.batch_to_collection(pts, [batch]); maps = [fj.mesh_to_sparse_assembly_map(pts.shape[0], cells)]
        out["quadrature_points"] = int(ebc.Q[0])
        u = torch.tensor(1e-5 * rng.uniform(size=pts.size), device="cuda")
        nnz = fj._calculate_jacobian_unique_nnz(pts.shape[0], ebc)
        r = timed(lambda: fj.calculate_residual_wo_dirichlet(token, ebc, maps, u))
        j = timed(lambda: fj.calculate_jacobian_wo_dirichlet(token, ebc, maps, u, nnz))
        out[model.__name__] = {"residual_ms": r, "jacobian_ms": j, "assembly_elements_per_s": E / ((r + j) * 1e-3), "nnz": int(nnz)}
        del ebc, maps
        torch.cuda.empty_cache()
    print(json.dumps(out), flush=True)

def batched(C=4096, m=32):
    V3, cells = fj.uniform_quad_grid(m + 1, m + 1, np.array([[0.0, 1.0], [0.0, 1.0]]))
    pts = V3[:, :2]
    E = cells.shape[0]
    rng = np.random.default_rng(3)
    mp = np.empty((C, E, 4, 2)); mp[..., 0] = 10e9 + 90e9 * rng.uniform(size=(C, E, 4)); mp[..., 1] = 0.25
    edge = np.nonzero((pts[:, 0] < 1e-12) | (pts[:, 0] > 1 - 1e-12))[0]
    bcs = np.stack([np.repeat(edge, 2), np.tile([0, 1], edge.shape[0])], axis=1)
    vals = 1e-3 * rng.uniform(size=(C, bcs.shape[0]))
    fe = fj.FiniteElementType(fj.CellType.quadrilateral, fj.ElementFamily.P, 1, fj.LagrangeVariant.equispaced, fj.QuadratureType.default, 2)
    batch = fj.ElementBatch(fe_type=fe, n_dofs_per_basis=2, connectivity_en=cells, constitutive_model=fj.elastic_isotropic, material_params=mp[0])
    opts = fj.SolverOptions(linear_precond_type=fj.PreconditionerType.JACOBI, linear_max_iter=4000, linear_relative_tol=1e-10)
    best = None
    for rep in range(2):                                  # first call builds maps / pattern / plans; the second is the steady state
        torch.cuda.synchronize(); t0 = time.perf_counter()
        u, res = fj.solve_bvp_batched(pts, [batch], fj.linear_elasticity_residual, bcs, vals, [mp], solver_options=opts)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    k = fj.solve_bvp_batched.last_iterations[0]
    print(json.dumps({"config": 5, "copies": C, "rve": f"{m}x{m} quad4", "dofs_per_copy": int(pts.size), "total_dofs": int(C * pts.size),
                      "wall_s_first_call": dt if rep == 0 else None, "wall_s": best, "solves_per_s": C / best,
                      "cg_iterations_min_mean_max": [int(k.min()), float(k.mean()), int(k.max())],
                      "residual_max": float(res.abs().max().item()),
                      "note": "solve_bvp_batched: host setup (block-diagonal mesh, maps, pattern) included in wall_s"}), flush=True)

if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("quadratic", "all"):
        quadratic("hexahedron", 32)
        quadratic("tetrahedron", 32)
    if what in ("batched", "all"):
        batched()
