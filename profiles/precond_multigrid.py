"""Jacobi-CG against multigrid-preconditioned CG on the bench workload (hex8 n^3, per-point moduli, Dirichlet faces):
iterations to a relative residual of 1e-8 and time to solution.  python profiles/precond_multigrid.py [n ...]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fea-in-jax_b200")]
os.environ.setdefault("FEB200_QUIET", "1")
import numpy as np, torch
import fe_jax as fj
from bench import make_problem

def run(n, tol=1e-8, degree=2, ratio=6.0):
    Vx, cells, mp = make_problem(n, n, "EQM")
    V = Vx.shape[0]; F = 3 * V
    fe = fj.FiniteElementType(fj.CellType.hexahedron, fj.ElementFamily.P, 1, fj.LagrangeVariant.equispaced, fj.QuadratureType.default, 2)
    batch = fj.ElementBatch(fe_type=fe, n_dofs_per_basis=3, connectivity_en=cells, constitutive_model=fj.elastic_isotropic, material_params=mp)
    ebc = fj.batch_to_collection(Vx, [batch]); maps = [fj.mesh_to_sparse_assembly_map(V, cells)]
    left, right = np.nonzero(Vx[:, 0] < 1e-12)[0], np.nonzero(Vx[:, 0] > 1 - 1e-12)[0]
    dofs = np.concatenate([3 * left, 3 * left + 1, 3 * left + 2, 3 * right])
    mask = torch.zeros(F, dtype=torch.float64, device="cuda"); mask[torch.from_numpy(dofs).cuda()] = 1.0
    vals = torch.zeros_like(mask); vals[torch.from_numpy(3 * right).cuda()] = 0.01
    u0 = torch.zeros_like(mask)
    tok = fj.linear_elasticity_residual
    R, _ = fj.calculate_residual_w_dirichlet(tok, ebc, maps, u0, vals, mask)
    J = fj.calculate_jacobian_w_dirichlet(tok, ebc, maps, u0, vals, mask)
    b = (-R).contiguous()
    x0 = torch.zeros_like(b); x0[mask != 0] = vals[mask != 0]
    from fe_jax.assembly import csr_diagonal
    out = {"n": n, "dofs": F}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    xj, ij = fj.cg(J, b, x0, tol=tol, maxiter=20000, M=fj.JacobiPreconditioner(csr_diagonal(J._pattern, J.data)))
    torch.cuda.synchronize(); out["jacobi"] = {"iterations": int(ij["iterations"]), "solve_s": time.perf_counter() - t0}
    t0 = time.perf_counter()
    mg = fj.GeometricMultigrid(Vx, (n, n, n), mp[..., 0], 0.25, mask, J._pattern, J.data, degree=degree, ratio=ratio)
    torch.cuda.synchronize(); setup = time.perf_counter() - t0
    t0 = time.perf_counter()
    xm, im = fj.cg(J, b, x0, tol=tol, maxiter=500, M=mg)
    torch.cuda.synchronize()
    out["multigrid"] = {"iterations": int(im["iterations"]), "solve_s": time.perf_counter() - t0, "setup_s": setup,
                        "levels": mg.shape_summary, "lmax": [round(l.lmax, 3) for l in mg.levels], "degree": degree, "ratio": ratio}
    out["solutions_rel_diff"] = float(((xj - xm).abs().max() / xj.abs().max()).item())
    return out

if __name__ == "__main__":
    ns = [int(a) for a in sys.argv[1:]] or [64]
    deg = int(os.environ.get("MG_DEGREE", "2")); ratio = float(os.environ.get("MG_RATIO", "6"))
    for n in ns:
        print(json.dumps(run(n, degree=deg, ratio=ratio)), flush=True)
        torch.cuda.empty_cache()
