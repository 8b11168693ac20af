"""Residual kernel timing driver: python profiles/profile_residual.py [n] [M|EM|EQM]
times feb200_element_residual with the Dirichlet blend on the fly (mask, values given) and on a pre-blended vector."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fea-in-jax_b200")]
os.environ.setdefault("FEB200_QUIET", "1")
import torch
import fe_jax as fj
from fe_jax import _lib
from bench import make_problem

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
layout = sys.argv[2] if len(sys.argv) > 2 else "EQM"
Vx, cells, mp = make_problem(n, n, layout)
fe = fj.FiniteElementType(fj.CellType.hexahedron, fj.ElementFamily.P, 1, fj.LagrangeVariant.equispaced, fj.QuadratureType.default, 2)
batch = fj.ElementBatch(fe_type=fe, n_dofs_per_basis=3, connectivity_en=cells, constitutive_model=fj.elastic_isotropic, material_params=mp)
ebc = fj.batch_to_collection(Vx, [batch])
res = fj.physics.resolve_residual(fj.linear_elasticity_residual)
b = ebc._batch_struct(0, res.kind)
F = 3 * Vx.shape[0]
u = torch.rand(F, dtype=torch.float64, device="cuda") * 1e-3
mask = torch.zeros_like(u); mask[:3000] = 1.0
vals = torch.zeros_like(u)
uw = torch.empty_like(u)
R_e = ebc._buffer("R_e", ebc.E[0] * 24)


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for rep in range(2):
    print("residual, blend on the fly ms", timed(lambda: _lib.call("feb200_element_residual", b, _lib.ptr(ebc._vertices), _lib.ptr(u), _lib.ptr(mask), _lib.ptr(vals), _lib.ptr(R_e))))
    print("residual, no mask          ms", timed(lambda: _lib.call("feb200_element_residual", b, _lib.ptr(ebc._vertices), _lib.ptr(u), None, None, _lib.ptr(R_e))))
    print("dirichlet_blend            ms", timed(lambda: _lib.call("feb200_dirichlet_blend", F, _lib.ptr(u), _lib.ptr(vals), _lib.ptr(mask), _lib.ptr(uw))))
