"""Small driver used for ncu captures of the fused assembly kernel: python profiles/profile_fused.py [n] [M|EM|EQM]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fea-in-jax_b200")]
os.environ.setdefault("FEB200_QUIET", "1")
import torch
import fe_jax as fj
from fe_jax.assembly import _assemble_values, _plan
from bench import make_problem

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
layout = sys.argv[2] if len(sys.argv) > 2 else "M"
Vx, cells, mp = make_problem(n, n, layout)
fe = fj.FiniteElementType(fj.CellType.hexahedron, fj.ElementFamily.P, 1, fj.LagrangeVariant.equispaced, fj.QuadratureType.default, 2)
batch = fj.ElementBatch(fe_type=fe, n_dofs_per_basis=3, connectivity_en=cells, constitutive_model=fj.elastic_isotropic, material_params=mp)
ebc = fj.batch_to_collection(Vx, [batch]); maps = [fj.mesh_to_sparse_assembly_map(Vx.shape[0], cells)]
plan = _plan(ebc, maps)
if os.environ.get("FEB200_FUSED_IMPL", "sweep") == "sweep":
    print("sweep plan stats", plan.sweep_plan(ebc, 0).stats)
else:
    fp = plan.fused_plan(ebc, 0)
    print("plan stats", fp.stats, "tn_cap", fp.tn_cap)
u = torch.zeros(3 * Vx.shape[0], dtype=torch.float64, device="cuda")
mask = torch.zeros_like(u); mask[:300] = 1.0
out = torch.empty(plan.pattern.nnz, dtype=torch.float64, device="cuda")
vals = torch.zeros_like(u)
reps = int(os.environ.get("PF_REPS", "3"))
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        _assemble_values(fj.linear_elasticity_residual, ebc, maps, u, mask, vals, out=out,
                         eliminate=os.environ.get("PF_ELIMINATE", "1") != "0")
    e1.record(); torch.cuda.synchronize(); print("fused assemble ms", e0.elapsed_time(e1) / reps)
