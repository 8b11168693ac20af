import os, sys
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path[:0] = [ROOT, os.path.join(ROOT, "fea-in-jax_b200")]
os.environ.setdefault("FEB200_QUIET", "1")
import numpy as np, torch
import fe_jax as fj
from fe_jax.assembly import _assemble_values, _plan, csr_diagonal
from fe_jax.sparse import pattern_matvec
from bench import make_problem
n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
for layout in ("EQM", "M"):
    Vx, cells, mp = make_problem(n, n, layout)
    fe = fj.FiniteElementType(fj.CellType.hexahedron, fj.ElementFamily.P, 1, fj.LagrangeVariant.equispaced, fj.QuadratureType.default, 2)
    batch = fj.ElementBatch(fe_type=fe, n_dofs_per_basis=3, connectivity_en=cells, constitutive_model=fj.elastic_isotropic, material_params=mp)
    ebc = fj.batch_to_collection(Vx, [batch]); maps = [fj.mesh_to_sparse_assembly_map(Vx.shape[0], cells)]
    V = Vx.shape[0]; F = 3 * V
    dev = "cuda"
    mask = torch.zeros(F, dtype=torch.float64, device=dev); values = torch.zeros_like(mask)
    xs = torch.from_numpy(Vx[:, 0]).to(dev)
    left = (xs < 1e-12).nonzero().reshape(-1); right = (xs > 1 - 1e-12).nonzero().reshape(-1)
    for c in range(3): mask[3 * left + c] = 1.0
    mask[3 * right] = 1.0; values[3 * right] = 0.01
    u = torch.zeros(F, dtype=torch.float64, device=dev)
    vals, pat = _assemble_values(fj.linear_elasticity_residual, ebc, maps, u, mask, values, eliminate=True)
    R, _ = fj.calculate_residual_w_dirichlet(fj.linear_elasticity_residual, ebc, maps, u, values, mask)
    b = -R
    A = lambda v: pattern_matvec(pat, vals, v)
    diag = csr_diagonal(pat, vals)
    # diagonal 3x3 blocks from the node-pattern layout
    nindptr = pat.nindptr; deg = (nindptr[1:] - nindptr[:-1]); dr = pat.diag_rank.to(torch.int64)
    base = 9 * nindptr[:-1]
    blk = torch.empty(V, 3, 3, dtype=torch.float64, device=dev)
    for i in range(3):
        for j in range(3):
            blk[:, i, j] = vals[base + i * 3 * deg + 3 * dr + j]
    binv = torch.linalg.inv(blk)
    def cg(M, tol=1e-8, maxiter=20000):
        x = torch.zeros_like(b); r = b.clone(); z = M(r); p = z.clone(); g = torch.dot(r, z); bb = torch.dot(b, b)
        for k in range(maxiter):
            Ap = A(p); a = g / torch.dot(p, Ap); x += a * p; r -= a * Ap
            if torch.dot(r, r) <= tol * tol * bb: return k + 1
            z = M(r); gn = torch.dot(r, z); p = z + (gn / g) * p; g = gn
        return maxiter
    print(layout, n, "none", cg(lambda r: r), "jacobi", cg(lambda r: r / diag), "block-jacobi", cg(lambda r: torch.bmm(binv, r.view(V, 3, 1)).reshape(-1)))
