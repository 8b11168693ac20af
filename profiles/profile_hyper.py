"""Hyperelastic two-pass assembly timing driver: python profiles/profile_hyper.py [n] [svk|mr] [reps]
times feb200_element_tangent (tensor-core kernel; FEB200_TANGENT_SCALAR=1 selects the scalar kernel), feb200_assemble_csr and
the element residual on the hex8 n^3 grid of the bench, and checks the two tangent kernels against each other."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fea-in-jax_b200")]
os.environ.setdefault("FEB200_QUIET", "1")
import numpy as np
import torch
import fe_jax as fj
from fe_jax import _lib
from fe_jax.assembly import _plan
from bench import make_problem

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
model = sys.argv[2] if len(sys.argv) > 2 else "svk"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
Vx, cells, _ = make_problem(n, n, "M")
mp = np.array([70e9, 0.25]) if model == "svk" else np.array([0.5e6, 0.1e6, 2e-7])
fe = fj.FiniteElementType(fj.CellType.hexahedron, fj.ElementFamily.P, 1, fj.LagrangeVariant.equispaced, fj.QuadratureType.default, 2)
batch = fj.ElementBatch(fe_type=fe, n_dofs_per_basis=3, connectivity_en=cells,
                        constitutive_model=fj.st_venant_kirchhoff if model == "svk" else fj.mooney_rivlin, material_params=mp)
ebc = fj.batch_to_collection(Vx, [batch])
maps = [fj.mesh_to_sparse_assembly_map(Vx.shape[0], cells)]
res = fj.physics.resolve_residual(fj.hyperelasticity_residual)
b = ebc._batch_struct(0, res.kind)
plan = _plan(ebc, maps)
pat = plan.pattern
F = 3 * Vx.shape[0]
u = torch.from_numpy(1e-5 * np.random.default_rng(1).uniform(size=F)).cuda()
K_e = ebc._buffer("K_e", ebc.E[0] * 576)
vals = torch.empty(pat.nnz, dtype=torch.float64, device="cuda")
R_e = ebc._buffer("R_e", ebc.E[0] * 24)


def timed(fn, reps=reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


tangent = lambda: _lib.call("feb200_element_tangent", b, _lib.ptr(ebc._vertices), _lib.ptr(u), None, None, _lib.ptr(K_e))
gather = lambda: _lib.call("feb200_assemble_csr", plan.maps[0], pat.V, pat.U, _lib.ptr(K_e), _lib.ptr(plan.slot_rank[0]),
                           _lib.ptr(pat.nindptr), pat.max_deg, _lib.ptr(vals), 0, None, None)
mask = torch.zeros_like(u); mask[:3000] = 1.0
dvals = torch.zeros_like(u)
tangent_m = lambda: _lib.call("feb200_element_tangent", b, _lib.ptr(ebc._vertices), _lib.ptr(u), _lib.ptr(mask), _lib.ptr(dvals), _lib.ptr(K_e))
gather_m = lambda: _lib.call("feb200_assemble_csr", plan.maps[0], pat.V, pat.U, _lib.ptr(K_e), _lib.ptr(plan.slot_rank[0]),
                             _lib.ptr(pat.nindptr), pat.max_deg, _lib.ptr(vals), 0, _lib.ptr(pat.nindices), _lib.ptr(mask))
from fe_jax.assembly import _vertex_flags
gather_f = lambda: _lib.call("feb200_assemble_csr_flags", plan.maps[0], pat.V, pat.U, _lib.ptr(K_e), _lib.ptr(plan.slot_rank[0]),
                             _lib.ptr(pat.nindptr), pat.max_deg, _lib.ptr(vals), 0, _lib.ptr(pat.nindices), _lib.ptr(_vertex_flags(ebc, mask)))
which = "scalar" if os.environ.get("FEB200_TANGENT_SCALAR") else "tensor-core"
E = ebc.E[0]
for rep in range(2):
    t = timed(tangent)
    print(f"{model} n={n} element_tangent ({which}) {t:8.3f} ms   {E / t / 1e3:9.1f} M elements/s   K_e write {E * 4608 / t / 1e6:7.0f} GB/s")
    t = timed(tangent_m)
    print(f"{model} n={n} element_tangent, Dirichlet blend on the fly {t:8.3f} ms")
    t = timed(gather)
    print(f"{model} n={n} assemble_csr               {t:8.3f} ms   {(E * 4608 + pat.nnz * 8) / t / 1e6:7.0f} GB/s")
    t = timed(gather_m)
    print(f"{model} n={n} assemble_csr, Dirichlet elimination fused (mask doubles) {t:8.3f} ms")
    t = timed(gather_f)
    print(f"{model} n={n} assemble_csr_flags, Dirichlet elimination fused (flag bytes) {t:8.3f} ms")
    t = timed(lambda: _lib.call("feb200_element_residual", b, _lib.ptr(ebc._vertices), _lib.ptr(u), None, None, _lib.ptr(R_e)))
    print(f"{model} n={n} element_residual           {t:8.3f} ms")
if os.environ.get("FEB200_OVERLAP"):
    # feasibility of running the two passes concurrently (FP64-bound tangent kernel, HBM-bound gather): both on all elements /
    # vertices at once on two streams (the values are garbage; only the time counts)
    sb = torch.cuda.Stream()
    def both():
        ev = torch.cuda.Event(); ev.record()
        tangent()
        with torch.cuda.stream(sb):
            sb.wait_event(ev)
            gather()
        torch.cuda.current_stream().wait_stream(sb)
    for rep in range(2):
        print(f"{model} n={n} tangent || gather on two streams {timed(both):8.3f} ms  (FEB200_TG_WARPS={os.environ.get('FEB200_TG_WARPS')})")
if not os.environ.get("FEB200_TANGENT_SCALAR") and n <= 64:
    tangent(); torch.cuda.synchronize()
    ref = subprocess.run([sys.executable, __file__, str(n), model, "1"], env=dict(os.environ, FEB200_TANGENT_SCALAR="1", FEB200_DUMP="1"),
                         capture_output=True, text=True)
    print(ref.stdout[-400:], ref.stderr[-400:])
    k_ref = torch.load("/tmp/feb200_ke_scalar.pt").cuda()
    err = (K_e[: k_ref.numel()] - k_ref).abs().max().item() / k_ref.abs().max().item()
    print(f"tensor-core vs scalar K_e: max rel diff {err:.3e}")
if os.environ.get("FEB200_DUMP"):
    tangent(); torch.cuda.synchronize()
    torch.save(K_e[: E * 576].cpu(), "/tmp/feb200_ke_scalar.pt")
