"""SpMV timing driver: python profiles/profile_spmv.py [n]  (hex8 n^3 pattern, random values; FEB200_SPMV_BULK=0 selects the plain kernel)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fea-in-jax_b200")]
os.environ.setdefault("FEB200_QUIET", "1")
import torch
import fe_jax as fj
from fe_jax import _lib
from fe_jax.assembly import _plan
from bench import make_problem

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
Vx, cells, mp = make_problem(n, n, "M")
fe = fj.FiniteElementType(fj.CellType.hexahedron, fj.ElementFamily.P, 1, fj.LagrangeVariant.equispaced, fj.QuadratureType.default, 2)
batch = fj.ElementBatch(fe_type=fe, n_dofs_per_basis=3, connectivity_en=cells, constitutive_model=fj.elastic_isotropic, material_params=mp)
ebc = fj.batch_to_collection(Vx, [batch]); maps = [fj.mesh_to_sparse_assembly_map(Vx.shape[0], cells)]
pat = _plan(ebc, maps).pattern
F = 3 * pat.V
vals = torch.rand(pat.nnz, dtype=torch.float64, device="cuda")
x = torch.rand(F, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
spmv = lambda: _lib.call("feb200_spmv", pat.V, pat.U, _lib.ptr(pat.nindptr), _lib.ptr(pat.nindices), _lib.ptr(vals), _lib.ptr(x), _lib.ptr(y), None, None, 0)
bytes_ = 8 * pat.nnz + 4 * (pat.nnz // 9) + 8 * (pat.V + 1) + 16 * F
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): spmv()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"spmv ms {ms:.4f}  {bytes_ / ms / 1e6:.0f} GB/s  checksum {float(y.sum()):.6e}")

if os.environ.get("PF_CHECK") == "1":           # against the generic CSR kernel on the expanded pattern (small n only)
    row, col, indptr = pat._expand()
    y2 = torch.empty_like(y)
    _lib.call("feb200_spmv_csr", F, _lib.ptr(indptr), _lib.ptr(col), _lib.ptr(vals), _lib.ptr(x), _lib.ptr(y2))
    spmv()
    torch.cuda.synchronize()
    print("max rel diff vs generic CSR kernel:", float((y - y2).abs().max() / y2.abs().max()))
