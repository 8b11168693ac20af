"""A/B of the CG operator on the bench workload: assembled CSR SpMV (feb200_spmv) against a matrix-free application of the
same operator.  For linear elasticity J p = R(p) (the residual is linear in u), so the matrix-free apply IS the residual
path: feb200_element_residual + feb200_scatter_sum with p in place of u -- what the reference's CG does through a JVP of
the whole assembly in every iteration (sparse_linear_solve.py:251-258).  python profiles/matrix_free_ab.py [n]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fea-in-jax_b200")]
os.environ.setdefault("FEB200_QUIET", "1")
import numpy as np, torch
import fe_jax as fj
from fe_jax import _lib
from fe_jax.sparse import pattern_matvec
from bench import make_problem

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
Vx, cells, mp = make_problem(n, n, "EQM")
V = Vx.shape[0]; F = 3 * V
fe = fj.FiniteElementType(fj.CellType.hexahedron, fj.ElementFamily.P, 1, fj.LagrangeVariant.equispaced, fj.QuadratureType.default, 2)
batch = fj.ElementBatch(fe_type=fe, n_dofs_per_basis=3, connectivity_en=cells, constitutive_model=fj.elastic_isotropic, material_params=mp)
ebc = fj.batch_to_collection(Vx, [batch]); maps = [fj.mesh_to_sparse_assembly_map(V, cells)]
tok = fj.linear_elasticity_residual
p = torch.rand(F, dtype=torch.float64, device="cuda")
nnz = fj._calculate_jacobian_unique_nnz(V, ebc)
J = fj.calculate_jacobian_wo_dirichlet(tok, ebc, maps, p, nnz)

def timed(fn, reps=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

y_spmv = pattern_matvec(J._pattern, J.data, p)
y_mf = fj.calculate_residual_wo_dirichlet(tok, ebc, maps, p)[0]
out = {"n": n, "dofs": F, "nnz": nnz,
       "spmv_ms": timed(lambda: pattern_matvec(J._pattern, J.data, p)),
       "matrix_free_ms": timed(lambda: fj.calculate_residual_wo_dirichlet(tok, ebc, maps, p)),
       "max_rel_diff": float(((y_spmv - y_mf).abs().max() / y_spmv.abs().max()).item()),
       "bytes_spmv": 8 * nnz + 4 * (nnz // 9) + 8 * (V + 1) + 16 * F,
       "bytes_matrix_free": 4 * 8 * cells.shape[0] + 8 * 3 * V + 16 * F + 8 * 16 * cells.shape[0] + 2 * 192 * cells.shape[0]}
print(json.dumps(out))
