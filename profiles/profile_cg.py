"""CG loop timing driver: python profiles/profile_cg.py [n] [iters]
hex8 n^3 linear-elastic tangent (no constraints: semi-definite, CG still runs); times, each over `iters` back-to-back
repetitions with CUDA events: (a) the plain SpMV, (b) the CG SpMV stage (fused p.Ap), (c) the two vector stages,
(d) whole iterations through feb200_cg_iterate -- to see where the time of an iteration goes under sustained load."""
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
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
Vx, cells, mp = make_problem(n, n, "M")
fe = fj.FiniteElementType(fj.CellType.hexahedron, fj.ElementFamily.P, 1, fj.LagrangeVariant.equispaced, fj.QuadratureType.default, 2)
batch = fj.ElementBatch(fe_type=fe, n_dofs_per_basis=3, connectivity_en=cells, constitutive_model=fj.elastic_isotropic, material_params=mp)
ebc = fj.batch_to_collection(Vx, [batch]); maps = [fj.mesh_to_sparse_assembly_map(Vx.shape[0], cells)]
pat = _plan(ebc, maps).pattern
lib = _lib.load()
F = 3 * pat.V
dev = "cuda"
from fe_jax.assembly import _assemble_values
vals, _ = _assemble_values(fj.linear_elasticity_residual, ebc, maps, torch.zeros(F, dtype=torch.float64, device=dev))   # the real tangent
b = torch.rand(F, dtype=torch.float64, device=dev)
x = torch.zeros(F, dtype=torch.float64, device=dev)
r, p, Ap = (torch.empty(F, dtype=torch.float64, device=dev) for _ in range(3))
hist = torch.zeros(100000, dtype=torch.float64, device=dev)
scalars = torch.zeros(8, dtype=torch.float64, device=dev)
counters = torch.zeros(4, dtype=torch.int32, device=dev)
ws = _lib.workspace(lib.feb200_reduce_workspace_bytes(), dev)
st = _lib.CGState(V=pat.V, U=pat.U, maxiter=100000, nindptr=_lib.ptr(pat.nindptr), nindices=_lib.ptr(pat.nindices),
                  values=_lib.ptr(vals), inv_diag=None, b=_lib.ptr(b), x=_lib.ptr(x), r=_lib.ptr(r), p=_lib.ptr(p),
                  Ap=_lib.ptr(Ap), hist=_lib.ptr(hist), scalars=_lib.ptr(scalars), counters=_lib.ptr(counters),
                  ws=_lib.ptr(ws), ws_bytes=ws.numel())
_lib.call("feb200_cg_init", st, 0.0, 0.0)


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


plain = lambda: _lib.call("feb200_spmv", pat.V, pat.U, _lib.ptr(pat.nindptr), _lib.ptr(pat.nindices), _lib.ptr(vals), _lib.ptr(p), _lib.ptr(Ap), None, None, 0)
stage = lambda s: _lib.call("feb200_cg_stage", st, s, 0.0, 0.0)
for rep in range(2):
    print(f"cg_iterate (1 call)   {timed(lambda: _lib.call('feb200_cg_iterate', st, iters), 1) / iters:.4f} ms / iteration", "done", int(counters[1].item()), "k", int(counters[0].item()))
for rep in range(2):
    print(f"plain spmv            {timed(plain, iters):.4f} ms")
    print(f"cg spmv stage         {timed(lambda: stage(3), iters):.4f} ms")
    print(f"xr + p                {timed(lambda: (stage(5), stage(6)), iters):.4f} ms", "done", int(counters[1].item()))
