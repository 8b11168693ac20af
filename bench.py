#!/usr/bin/env python
"""Benchmark of the feb200 hot path (contract: see the task statement / DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--grid 200] [--impl feb200|reference]

Workload (BASELINE.json configs[1]): 3-D hex8 linear elasticity on a uniform n^3 grid (n=200: 8 M elements, 24.4 M DoFs,
1.95e9 non-zeros), interior nodes jittered, isotropic material with a Young's modulus per quadrature point (--layout EQM,
the default; M / EM = one material for the batch / per element).  One "step" = one assembly pass of the whole mesh
through the fe_jax call surface: calculate_residual_w_dirichlet + calculate_jacobian_w_dirichlet (CSR values of J_w at u_w
with the Dirichlet elimination fused).  `value` = elements / s with all inputs resident in HBM (CUDA events around the
timed region); `e2e` = the same pass with the solution vector coming from pinned host memory and the residual + Jacobian
diagonal read back to the host inside the timed region.  `newton_cg` = one Newton iteration with 100 CG iterations through
linear_solve (assembly + CG + update + residual); `newton_cg_hyperelastic` = the same on the hyperelastic problem
(BASELINE configs[2]).
N > 1 (torchrun): STRONG scaling -- the same global n^3 grid cut into N x-slabs of n/N element layers, one per GPU; the
residual's interface halo-sum (NCCL) is inside the timed assembly step, the CG loop runs through the C data plane
(csrc/dist.cu: halo-sum after every local SpMV, all-reduced dot products).  `parity` compares a converged partitioned
solve with the single-GPU solve; `weak` is the secondary block with n^3 elements per GPU.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "fea-in-jax_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ.setdefault("FEB200_QUIET", "1")

import numpy as np

METRIC = "hex8_assembly_elements_per_s"
UNIT = "elements/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--grid", dest="n", type=int, default=200, help="elements per axis of the hex8 grid (per rank)")
    ap.add_argument("--impl", default="feb200", choices=["feb200", "reference"])
    ap.add_argument("--cg-iters", type=int, default=100)
    ap.add_argument("--cpu-n", type=int, default=20, help="grid size of the bounded CPU-baseline sample")
    ap.add_argument("--layout", default="EQM", choices=["M", "EM", "EQM"],
                    help="material layout: EQM = per-point Young's modulus (SURVEY 8d primary input), M = batch constant")
    ap.add_argument("--physics", default="linear", choices=["linear", "hyper"],
                    help="hyper = BASELINE configs[2]: hyperelastic (St. Venant-Kirchhoff) residual and tangent, generic two-pass "
                         "assembly; the CPU baseline legs stay on the linear workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-solve", action="store_true", help="N = 1: skip the converged Jacobi-CG / multigrid-CG solves")
    ap.add_argument("--no-hyper", action="store_true", help="skip the hyperelastic (BASELINE configs[2]) Newton-CG block")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the partitioned-vs-single-GPU parity solve")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the secondary weak-scaling block")
    ap.add_argument("--two-pass", action="store_true", help="also time the generic two-pass assembly kernels")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def make_problem(n_x, n, layout, x0=0.0, x1=1.0):
    """hex8 grid of n_x x n x n elements on [x0,x1] x [0,1]^2 with jittered interior nodes."""
    import fe_jax as fj
    Vx, cells = fj.uniform_hex_grid(n_x, n, n, bbox=((x0, x1), (0.0, 1.0), (0.0, 1.0)))
    h = 1.0 / n
    rng = np.random.default_rng(0)
    interior = np.all((Vx[:, 1:] > 1e-12) & (Vx[:, 1:] < 1 - 1e-12), axis=1) & (Vx[:, 0] > x0 + 1e-12) & (Vx[:, 0] < x1 - 1e-12)
    Vx = Vx + 0.2 * h * rng.uniform(-1.0, 1.0, size=Vx.shape) * interior[:, None]
    E = cells.shape[0]
    rm = np.random.default_rng(2)
    if layout == "EQM":
        mp = np.empty((E, 8, 2))
        mp[..., 0] = 10e9 + 90e9 * rm.uniform(size=(E, 8))
        mp[..., 1] = 0.25
    elif layout == "EM":
        mp = np.empty((E, 2))
        mp[:, 0] = 10e9 + 90e9 * rm.uniform(size=E)
        mp[:, 1] = 0.25
    else:
        mp = np.array([70e9, 0.25])
    return Vx, cells, mp


def cpu_reference_sample(n, steps=1, warmup=0, layout="EQM"):
    """The oracle's reference-style assembly (element residual + K_e by 24 residual evaluations + COO triples +
    lexsort + duplicate sum, fea.py:613-660) on an n^3 hex8 grid; returns elements/s."""
    from oracle import fe_oracle as orc
    Vx, cells, u, mp, bcs, vals = orc.synthetic_hex_problem(n, material_layout=layout)
    xi, W = orc.get_quadrature("hexahedron", 2)
    _, dphi = orc.eval_basis_and_derivatives("hexahedron", 1, xi)
    prob = orc.Problem(Vx, [orc.Batch(cells, "linear", "elastic_isotropic", mp, dphi, W)])
    dofs, values_g, mask_g = orc.dirichlet_arrays(prob.F, 3, bcs, vals)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        R = prob.residual_w_dirichlet(u, values_g, mask_g)
        d, r, c = prob.jacobian_wo_dirichlet((1 - mask_g) * u + mask_g * values_g)
        dw = orc.apply_dirichlet_bcs_lhs(d, r, c, dofs)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return cells.shape[0] / float(np.mean(times)), float(np.mean(times)), cells.shape[0]


def _cpu_worker(a):
    os.environ["OMP_NUM_THREADS"] = "1"
    return cpu_reference_sample(*a)


def cpu_reference_parallel(n, steps=1, warmup=0, procs=None, layout="EQM"):
    """The same sample on every host core at once (independent copies, one process per core): aggregate
    elements/s of the CPU restatement with all the host threads it can use."""
    import multiprocessing as mp
    procs = procs or max(1, min(os.cpu_count() or 1, 64))
    if procs == 1:
        eps, sec, E = cpu_reference_sample(n, steps, warmup, layout)
        return eps, sec, E, 1
    with mp.get_context("spawn").Pool(procs) as pool:
        out = pool.map(_cpu_worker, [(n, steps, warmup, layout)] * procs)
    return sum(o[0] for o in out), float(np.mean([o[1] for o in out])), out[0][2], procs


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    eps, sec, E, cores = cpu_reference_parallel(args.cpu_n, steps=max(1, args.steps), warmup=min(1, args.warmup),
                                                layout=args.layout)
    line = {"impl": "reference", "metric": METRIC, "value": eps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"hex8 linear elasticity, uniform grid, residual+Jacobian assembly (BASELINE configs[1]), "
                                   f"material layout ({args.layout}); CPU sample {args.cpu_n}^3 elements of the same workload"},
            "cpu_baseline": {"value": eps, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"oracle (NumPy restatement of the reference algorithm: per-element K_e + COO lexsort + "
                                       f"duplicate sum) on a {args.cpu_n}^3 hex8 grid ({E} elements) run concurrently on "
                                       f"{cores} host cores (one process per core, aggregate throughput); "
                                       "the reference's own JAX-CPU path cannot run (no jax in the image)"},
            "e2e": {"value": eps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_REAL_STDOUT = None


def _protect_stdout():
    """Native libraries (NCCL prints its version banner) must not pollute the single JSON line on stdout:
    route fd 1 to stderr for the whole run and keep the real stdout for emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


class Problem:
    """One rank's problem: mesh slab, batch, Dirichlet data, the collection / maps / plan of the fe_jax surface."""

    def __init__(self, fj, torch, dev, nx, n, layout, hyper, x0, x1, x_hi_global, right_val):
        self.Vx, self.cells, self.mp = make_problem(nx, n, layout, x0=x0, x1=x1)
        self.E, self.V = self.cells.shape[0], self.Vx.shape[0]
        self.F = 3 * self.V
        fe = fj.FiniteElementType(fj.CellType.hexahedron, fj.ElementFamily.P, 1, fj.LagrangeVariant.equispaced,
                                  fj.QuadratureType.default, 2)
        self.token = fj.hyperelasticity_residual if hyper else fj.linear_elasticity_residual
        self.batch = fj.ElementBatch(fe_type=fe, n_dofs_per_basis=3, connectivity_en=self.cells,
                                     constitutive_model=fj.st_venant_kirchhoff if hyper else fj.elastic_isotropic,
                                     material_params=self.mp)
        # Dirichlet: global face x = 0 fixed, global face x = x_hi_global: u_x = right_val
        xs = self.Vx[:, 0]
        self.left = np.nonzero(xs < 1e-12)[0]
        self.right = np.nonzero(xs > x_hi_global - 1e-12)[0]
        self.bc_v = np.concatenate([np.repeat(self.left, 3), self.right])
        self.bc_c = np.concatenate([np.tile(np.arange(3), self.left.size), np.zeros(self.right.size, dtype=np.int64)])
        self.bc_val = np.concatenate([np.zeros(3 * self.left.size), np.full(self.right.size, right_val)])
        # hyperelastic runs keep the random state in the small-strain range so the tangent stays positive definite
        self.u_host = torch.from_numpy((1e-5 if hyper else 1e-3) * np.random.default_rng(1).uniform(size=self.F)).pin_memory()
        self.u = self.u_host.to(dev)


def hyper_kernel_block(fj, _lib, _vertex_flags, timed, args, runh, prh):
    """Per-kernel times of the hyperelastic two-pass assembly (N = 1) and the traffic they move."""
    nb = {}
    eh, ph = runh.ebc, runh.plan
    bh = eh._batch_struct(0, fj.physics.resolve_residual(prh.token).kind)
    Eh = eh.E[0]
    K_eh, R_eh = eh._buffer("K_e", Eh * 576), eh._buffer("R_e", Eh * 24)
    hk = {"element_residual": lambda: _lib.call("feb200_element_residual", bh, _lib.ptr(eh._vertices), _lib.ptr(prh.u),
                                                _lib.ptr(runh.mask), _lib.ptr(runh.values), _lib.ptr(R_eh)),
          "element_tangent": lambda: _lib.call("feb200_element_tangent", bh, _lib.ptr(eh._vertices), _lib.ptr(prh.u),
                                               _lib.ptr(runh.mask), _lib.ptr(runh.values), _lib.ptr(K_eh)),
          "assemble_csr": lambda: _lib.call("feb200_assemble_csr_flags", ph.maps[0], runh.pat.V, runh.pat.U, _lib.ptr(K_eh),
                                            _lib.ptr(ph.slot_rank[0]), _lib.ptr(runh.pat.nindptr), runh.pat.max_deg,
                                            _lib.ptr(runh.J_vals), 0, _lib.ptr(runh.pat.nindices),
                                            _lib.ptr(_vertex_flags(eh, runh.mask)))}
    nb["kernel_ms"] = {k: timed(f, max(2, args.steps // 2), 1)[0] for k, f in hk.items()}
    # two-pass traffic per rank: K_e written once and read once, CSR values written, slot ranks (2 bytes per node pair)
    nnz_h = int(runh.pat.nnz)
    nb["two_pass_bytes"] = 2 * 4608 * Eh + 8 * nnz_h + 2 * 64 * Eh
    nb["two_pass_gbs"] = nb["two_pass_bytes"] / ((nb["kernel_ms"]["element_tangent"] + nb["kernel_ms"]["assemble_csr"]) * 1e-3) / 1e9
    del hk, K_eh, R_eh
    return nb


def main():
    args = parse()
    _protect_stdout()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    import __graft_entry__ as graft

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if not os.path.exists(graft.LIB):
        if rank == 0:
            graft.build()
        if world > 1:
            dist.barrier()
    import fe_jax as fj
    from fe_jax import _lib
    from fe_jax import partition as part
    from fe_jax.assembly import _plan, csr_diagonal

    dev = torch.device("cuda", local_rank)
    n = args.n
    assert n % world == 0, "the grid size must be divisible by the number of GPUs (x-slabs of n / N element layers)"
    hyper_main = args.physics == "hyper"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launches
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, _lib.launches - l0

    class Runner:
        """Assembly / Newton-CG steps of one problem through the fe_jax call surface: the single-GPU functions for N = 1,
        the element-partitioned driver (fe_jax.partition, NCCL halo-sum + all-reduced CG scalars) for N > 1."""

        def __init__(self, pr: Problem):
            self.pr = pr
            t0 = time.perf_counter()
            if world == 1:
                self.ebc = fj.batch_to_collection(pr.Vx, [pr.batch])
                self.maps = [fj.mesh_to_sparse_assembly_map(pr.V, pr.cells)]
                self.plan = _plan(self.ebc, self.maps)
                self.pat = self.plan.pattern
                self.mask = torch.zeros(pr.F, dtype=torch.float64, device=dev)
                self.values = torch.zeros(pr.F, dtype=torch.float64, device=dev)
                dofs = torch.from_numpy(3 * pr.bc_v + pr.bc_c).to(dev)
                self.mask[dofs] = 1.0
                self.values[dofs] = torch.from_numpy(pr.bc_val).to(dev)
                self.dofs, self.dvals = dofs, torch.from_numpy(pr.bc_val).to(dev)
                self.dp = None
            else:
                nx = pr.cells.shape[0] // (n * n)
                ids = np.arange(pr.V).reshape(nx + 1, n + 1, n + 1)
                lm = part.slab_local_mesh(rank, world, pr.Vx, pr.cells, ids[0].ravel(), ids[nx].ravel())
                self.dp = part.DistributedProblem(fj, part.TorchDistComm(), [lm], lambda m: [pr.batch],
                                                  lambda m: (np.stack([pr.bc_v, pr.bc_c], axis=1), pr.bc_val), [dev], pr.token)
                rp = self.dp.rp[0]
                self.ebc, self.maps, self.plan, self.pat = rp.ebc, rp.maps, rp.plan, rp.plan.pattern
                self.mask, self.values, self.dofs, self.dvals = rp.mask_g, rp.values_g, rp.dofs, rp.values
                self.lm = lm
            self.J_vals = torch.empty(self.pat.nnz, dtype=torch.float64, device=dev)
            torch.cuda.synchronize()
            self.setup_s = time.perf_counter() - t0

        # one assembly pass: R_w(u) (N > 1: with the interface halo-sum) and the CSR values of J_w(u)
        def assembly_step(self):
            pr = self.pr
            if self.dp is None:
                R, _ = fj.calculate_residual_w_dirichlet(pr.token, self.ebc, self.maps, pr.u, self.values, self.mask)
                fj.calculate_jacobian_w_dirichlet(pr.token, self.ebc, self.maps, pr.u, self.values, self.mask, None,
                                                  out=self.J_vals)
                return R
            R = self.dp.residual_w_dirichlet([pr.u])[0]
            self.dp.jacobian_w_dirichlet([pr.u], outs=[self.J_vals])
            return R

        # one Newton iteration with a fixed number of CG iterations: assembly, solve, update, residual re-evaluation
        def newton_cg_step(self, cg_iters):
            pr = self.pr
            if self.dp is None:
                opts = fj.SolverOptions(linear_solve_type=fj.LinearSolverType.CG_JAX_SCIPY_W_INFO,
                                        linear_precond_type=fj.PreconditionerType.NONE, linear_relative_tol=0.0,
                                        linear_absolute_tol=0.0, linear_max_iter=cg_iters)
                res = fj.Residual(function=lambda x: fj.calculate_residual_w_dirichlet(
                    pr.token, self.ebc, self.maps, x, self.values, self.mask)[0], dirichlet_bcs_builtin=True)
                jac = fj.Jacobian(function=lambda x: fj.calculate_jacobian_w_dirichlet(
                    pr.token, self.ebc, self.maps, x, self.values, self.mask, None, out=self.J_vals), dirichlet_bcs_builtin=True)
                info0 = fj.init_solver_info(opts, device=dev)
                dx, info = fj.linear_solve(res, jac, None, self.dofs, self.dvals, opts, info0, False, pr.u)
                R2, _ = fj.calculate_residual_w_dirichlet(pr.token, self.ebc, self.maps, pr.u + dx, self.values, self.mask)
                hist = info.cumulative_residual_norm_history
                return int(info.cumulative_linear_iterations), hist, R2
            rp = self.dp.rp[0]
            Rs = self.dp.residual_w_dirichlet([pr.u])
            J = self.dp.jacobian_w_dirichlet([pr.u], outs=[self.J_vals])
            b = (-Rs[0]).contiguous()
            dx0 = torch.zeros_like(b)
            dx0[rp.dofs] = rp.values - pr.u[rp.dofs]
            dxs, k, hist = self.dp.cg(J, [b], [dx0], 0.0, 0.0, cg_iters, jacobi=False, poll_every=cg_iters)
            dxs[0][rp.dofs] = rp.values - pr.u[rp.dofs]
            R2 = self.dp.residual_w_dirichlet([pr.u + dxs[0]])
            return k, hist, R2[0]

        def newton_block(self, cg_iters, steps, ms_assembly, note):
            ms, _ = timed(lambda: self.newton_cg_step(cg_iters), max(1, steps // 2), 1)
            k, hist, _ = self.newton_cg_step(cg_iters)
            out = {"step_ms": ms, "cg_iters": cg_iters, "ms_per_cg_iter": (ms - ms_assembly) / max(cg_iters, 1),
                   "residual_norm_first": float(hist[0]), "residual_norm_last": float(hist[min(k, cg_iters) - 1]), "note": note}
            if self.dp is not None:
                out["interface_dofs_per_rank"] = int(3 * self.lm.n_shared)
                out["data_plane"] = "C (feb200_cg_iterate_dist: own NCCL communicator)" if self.dp.plane is not None else "torch.distributed"
            return out

    # ---- main workload: global n^3 grid (N > 1: cut into N x-slabs of n / N element layers, strong scaling)
    nx = n // world
    pr = Problem(fj, torch, dev, nx, n, args.layout, hyper_main, rank / world, (rank + 1) / world, 1.0,
                 0.0005 if hyper_main else 0.01)
    run = Runner(pr)
    E, V, F, nnz, pat, plan, ebc, maps = pr.E, pr.V, pr.F, run.pat.nnz, run.pat, run.plan, run.ebc, run.maps
    token, u, mask, values, J_vals = pr.token, pr.u, run.mask, run.values, run.J_vals
    total_E = n * n * n

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step, launches = timed(run.assembly_step, args.steps, args.warmup)

    # ---- per-kernel durations (same inputs, each kernel timed on its own over `steps` launches)
    res = fj.physics.resolve_residual(token)
    bstruct = ebc._batch_struct(0, res.kind)
    R_e = ebc._buffer("R_e", E * 24)
    Rtmp = torch.empty(F, dtype=torch.float64, device=dev)
    row = maps[0].indptr.reshape(-1).view(torch.int64)
    col = maps[0].indices.reshape(-1).view(torch.int64)
    from fe_jax.assembly import _vertex_flags, _block_flags, _use_tensor_core_kernel, _use_sweep_kernel, _material_hints
    vflags = _vertex_flags(ebc, mask)
    kernels = {
        "element_residual": lambda: _lib.call("feb200_element_residual", bstruct, _lib.ptr(ebc._vertices), _lib.ptr(u),
                                              _lib.ptr(mask), _lib.ptr(values), _lib.ptr(R_e)),
        "scatter_sum": lambda: _lib.call("feb200_scatter_sum", _lib.ptr(row), _lib.ptr(col), V, 3, _lib.ptr(R_e),
                                         _lib.ptr(Rtmp), 0, _lib.ptr(mask), _lib.ptr(values), _lib.ptr(u)),
    }
    sweep = plan.sweep_plan(ebc, 0) if (not hyper_main and _use_sweep_kernel(ebc, 0, bstruct)) else None
    fused = None if (hyper_main or sweep is not None) else plan.fused_plan(ebc, 0)
    use_mma = fused is not None and _use_tensor_core_kernel(ebc, 0, bstruct, fused)
    if sweep is not None:
        bflags = _block_flags(ebc, pat, mask)
        kernels["assemble_fused_hex8"] = lambda: _lib.call(
            "feb200_assemble_fused_hex8_sweep", bstruct, sweep.struct, _lib.ptr(ebc._vertices), _lib.ptr(sweep.rank8),
            _lib.ptr(bflags), _material_hints(ebc, 0, bstruct), _lib.ptr(J_vals), 0)
    elif use_mma:
        kernels["assemble_fused_hex8"] = lambda: _lib.call(
            "feb200_assemble_fused_hex8_mma", bstruct, fused.struct, _lib.ptr(ebc._vertices), _lib.ptr(plan.slot_rank[0]),
            _lib.ptr(pat.nindices), _lib.ptr(vflags), _material_hints(ebc, 0, bstruct), _lib.ptr(J_vals))
    elif fused is not None:
        kernels["assemble_fused_hex8"] = lambda: _lib.call(
            "feb200_assemble_fused_hex8", bstruct, fused.struct, _lib.ptr(ebc._vertices), _lib.ptr(row),
            _lib.ptr(plan.slot_rank[0]), _lib.ptr(pat.nindptr), _lib.ptr(pat.nindices), _lib.ptr(vflags), _lib.ptr(J_vals))
    if "assemble_fused_hex8" not in kernels or args.two_pass:
        K_e = ebc._buffer("K_e", E * 8 * 8 * 9)
        kernels["element_tangent"] = lambda: _lib.call("feb200_element_tangent", bstruct, _lib.ptr(ebc._vertices), _lib.ptr(u),
                                                       _lib.ptr(mask), _lib.ptr(values), _lib.ptr(K_e))
        kernels["assemble_csr"] = lambda: _lib.call("feb200_assemble_csr_flags", plan.maps[0], pat.V, pat.U, _lib.ptr(K_e),
                                                    _lib.ptr(plan.slot_rank[0]), _lib.ptr(pat.nindptr), pat.max_deg,
                                                    _lib.ptr(J_vals), 0, _lib.ptr(pat.nindices), _lib.ptr(_vertex_flags(ebc, mask)))
    kernel_ms = {k: timed(f, args.steps, 1)[0] for k, f in kernels.items()}
    run.assembly_step()                          # leave J_vals / R consistent for the solve below
    xv = torch.rand(F, dtype=torch.float64, device=dev)
    yv = torch.empty_like(xv)
    spmv = lambda: _lib.call("feb200_spmv", pat.V, pat.U, _lib.ptr(pat.nindptr), _lib.ptr(pat.nindices), _lib.ptr(J_vals),
                             _lib.ptr(xv), _lib.ptr(yv), None, None, 0)
    kernel_ms["spmv"] = timed(spmv, max(args.steps, 10), 3)[0]

    # ---- Newton-CG step of the main workload (fixed CG iteration count: tol = 0)
    newton = run.newton_block(args.cg_iters, args.steps, ms_step,
                              "fixed CG iteration count (tol=0): assembly + CG + update + residual re-evaluation, through "
                              + ("linear_solve" if world == 1 else "the element-partitioned driver (halo-sum after each local SpMV / "
                                 "scatter, all-reduced CG scalars)"))

    # ---- converged solve of the Newton system (relative residual 1e-8): the reference's Jacobi-CG configuration against CG
    #      preconditioned by fe_jax.GeometricMultigrid (iterations and time to solution; N = 1, linear workload)
    solve_block = None
    if world == 1 and not hyper_main and not args.no_solve:
        R0 = run.assembly_step()
        Jc = fj.COO((J_vals,), shape=(F, F), rows_sorted=True, pattern=pat)
        b_rhs = (-R0).contiguous()
        x0 = torch.zeros_like(b_rhs)
        x0[run.dofs] = run.dvals - u[run.dofs]

        def wall(fn):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = fn()
            torch.cuda.synchronize()
            return out, time.perf_counter() - t0

        (xj, ij), tj = wall(lambda: fj.cg(Jc, b_rhs, x0, tol=1e-8, maxiter=20000,
                                          M=fj.JacobiPreconditioner(csr_diagonal(pat, J_vals))))
        young = pr.mp[..., 0] if pr.mp.ndim > 1 else float(pr.mp[0])
        mg, t_setup_mg = wall(lambda: fj.GeometricMultigrid(pr.Vx, (n, n, n), young, 0.25, mask, pat, J_vals))
        (xm, im), tm = wall(lambda: fj.cg(Jc, b_rhs, x0, tol=1e-8, maxiter=500, M=mg))
        solve_block = {"relative_tolerance": 1e-8,
                       "jacobi_cg": {"iterations": int(ij["iterations"]), "solve_s": tj},
                       "multigrid_cg": {"iterations": int(im["iterations"]), "solve_s": tm, "setup_s": t_setup_mg,
                                        "levels": [list(g) for g in mg.shape_summary]},
                       "solutions_max_rel_diff": float(((xj - xm).abs().max() / xj.abs().max()).item()),
                       "note": "Jacobi = the reference's only preconditioner (sparse_linear_solve.py:301-306); multigrid = "
                               "fe_jax.GeometricMultigrid (re-discretised coarse levels, Chebyshev smoothing)"}
        del mg, xj, xm, Jc

    # ---- end to end through the public API with host buffers
    copy_stream = torch.cuda.Stream(device=dev)
    r_ready, u_ready, j_ready = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
    R_host = torch.empty(F, dtype=torch.float64).pin_memory()
    diag_host = torch.empty(F, dtype=torch.float64).pin_memory()
    u_host = pr.u_host

    def jacobian_only():
        if run.dp is None:
            fj.calculate_jacobian_w_dirichlet(token, ebc, maps, u, values, mask, None, out=J_vals)
        else:
            run.dp.jacobian_w_dirichlet([u], outs=[J_vals])

    def residual_only():
        if run.dp is None:
            return fj.calculate_residual_w_dirichlet(token, ebc, maps, u, values, mask)[0]
        return run.dp.residual_w_dirichlet([u])[0]

    def e2e_step():
        # Both copies finish inside the timed region.  Linear elasticity: the tangent does not depend on u, so the
        # Jacobian kernel starts at once and the host -> device copy of u runs under it on the copy stream; the
        # diagonal's device -> host copy then runs under the residual kernels and only the residual's copy is exposed.
        # Hyperelastic: u first, residual, its copy under the tangent + assembly kernels, then the diagonal's copy.
        # The CSR values themselves stay device-resident by design (they feed the SpMV of the solve).
        if not hyper_main:
            with torch.cuda.stream(copy_stream):
                u.copy_(u_host, non_blocking=True)
                u_ready.record()
            jacobian_only()
            d = csr_diagonal(pat, J_vals)
            j_ready.record()
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(j_ready)
                diag_host.copy_(d, non_blocking=True)
            torch.cuda.current_stream().wait_event(u_ready)
            R = residual_only()
            R_host.copy_(R, non_blocking=True)
        else:
            u.copy_(u_host, non_blocking=True)
            R = residual_only()
            r_ready.record()
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(r_ready)
                R_host.copy_(R, non_blocking=True)
            jacobian_only()
            d = csr_diagonal(pat, J_vals)
            diag_host.copy_(d, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        copy_stream.synchronize()
    e2e_ms, _ = timed(e2e_step, max(1, args.steps), 1)
    clocks = sampler.stop() if rank == 0 else None

    # ---- BASELINE configs[2]: the same mesh with the hyperelastic (St. Venant-Kirchhoff) residual and tangent, Newton-CG step
    extra = {}
    if not hyper_main and not args.no_hyper:
        del kernels
        run.J_vals = J_vals = None
        ebc._scratch.clear()
        torch.cuda.empty_cache()
        prh = Problem(fj, torch, dev, nx, n, args.layout, True, rank / world, (rank + 1) / world, 1.0, 0.0005)
        runh = Runner(prh)
        ms_h, _ = timed(runh.assembly_step, max(2, args.steps // 2), 2)
        nb = runh.newton_block(args.cg_iters, args.steps, ms_h,
                               "BASELINE configs[2]: hyperelastic (St. Venant-Kirchhoff) Newton-CG step on the same global mesh, "
                               "two-pass assembly (tensor-core element tangent k_tangent_general_mma -> K_e in HBM -> row-panel gather), fixed CG iteration count")
        nb["assembly_ms"] = ms_h
        nb["assembly_elements_per_s"] = total_E / (ms_h * 1e-3)
        # per-kernel times of the hyperelastic pass (the same C-ABI calls the builders make), live with CUDA events
        if world == 1:
            nb.update(hyper_kernel_block(fj, _lib, _vertex_flags, timed, args, runh, prh))
        extra["newton_cg_hyperelastic"] = nb
        del runh, prh
        torch.cuda.empty_cache()

    # ---- N > 1: parity of the partitioned solve against one GPU (32^3 global mesh, converged Jacobi-CG solves)
    if world > 1 and not args.no_parity:
        extra["parity"] = partition_parity(fj, part, torch, dist, dev, rank, world, 32)
    # ---- N > 1: secondary weak-scaling block (n^3 elements PER GPU, the round-1 headline)
    if world > 1 and not args.no_weak:
        torch.cuda.empty_cache()
        prw = Problem(fj, torch, dev, n, n, args.layout, False, float(rank), float(rank + 1), float(world), 0.01)
        runw = Runner(prw)
        ms_w, _ = timed(runw.assembly_step, max(2, args.steps // 2), 2)
        nbw = runw.newton_block(args.cg_iters, args.steps, ms_w, f"weak scaling: {n}^3 elements per GPU, interface halo-sum in the assembly step")
        extra["weak"] = {"elements_per_gpu": prw.E, "assembly_ms": ms_w, "assembly_elements_per_s": prw.E * world / (ms_w * 1e-3),
                         "newton_cg": nbw}
        del runw, prw

    if rank == 0:
        hbm, hbm_src = peaks()
        value = total_E / (ms_step * 1e-3)
        # algorithmic bytes per rank (the slot map is 2 bytes per node pair here):
        #   Jacobian assembly kernel  B_J = 8 nnz + 2 N^2 E + 4 N E + 8 D V + 8 n_mp (material parameters)
        #   whole assembly pass       B_A = B_J + 8 F (u) + 8 F (R) + 16 F (mask, values) + 8 n_mp
        n_mp = {"M": 2, "EM": 2 * E, "EQM": 16 * E}[args.layout]
        bytes_J = 8 * nnz + 2 * 64 * E + 4 * 8 * E + 8 * 3 * V + 8 * n_mp
        alg_bytes = bytes_J + 8 * F + 8 * F + 16 * F + 8 * n_mp
        if "assemble_fused_hex8" in kernel_ms:
            dom = ("feb200_assemble_fused_hex8_sweep (k_fused_hex8_sweep)" if sweep is not None
                   else "feb200_assemble_fused_hex8_mma (k_fused_hex8_mma)" if use_mma
                   else "feb200_assemble_fused_hex8 (k_fused_hex8_iso)")
            dom_ms = kernel_ms["assemble_fused_hex8"]
        else:
            dom = "element_tangent + assemble_csr (two-pass assembly)"
            dom_ms = kernel_ms["element_tangent"] + kernel_ms["assemble_csr"]
        achieved = bytes_J / (dom_ms * 1e-3) / 1e9
        traffic = None
        try:        # dram bytes of the dominant kernel from the committed ncu --set full capture of this configuration
            with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
                key = ("sweep" if sweep is not None else "fused_mma" if use_mma else "fused") + f"_n{n}_{args.layout}"
                traffic = json.load(f).get(key) if ("assemble_fused_hex8" in kernel_ms and world == 1) else None
        except Exception:
            pass
        spmv_bytes = 8 * nnz + 4 * (nnz // 9) + 8 * (V + 1) + 16 * F
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"3D hex8 linear elasticity, uniform grid {n}^3 elements (BASELINE configs[1])"
                                   + (f" cut into {world} x-slabs of {nx} element layers, one per GPU; interface halo-sum of the "
                                      f"residual (NCCL) inside the timed step" if world > 1 else "")
                                   + f", residual + Jacobian CSR assembly with fused Dirichlet elimination through "
                                     f"calculate_residual_w_dirichlet + calculate_jacobian_w_dirichlet; material layout ({args.layout})"
                                   + ("; HYPERELASTIC (St. Venant-Kirchhoff) residual and tangent, generic two-pass assembly "
                                      "(BASELINE configs[2])" if hyper_main else ""),
                       "elements_total": total_E, "elements_per_gpu": E, "dofs_per_gpu": F, "nnz_per_gpu": nnz,
                       "l2_policy": "inputs larger than L2", "setup_s": run.setup_s},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm, "unit": "GB/s",
                         "frac": achieved / hbm, "traffic": traffic, "peak_source": hbm_src, "algorithmic_bytes": bytes_J,
                         "kernel_ms": kernel_ms,
                         "assembly_pass": {"algorithmic_bytes": alg_bytes, "achieved": alg_bytes / (ms_step * 1e-3) / 1e9,
                                           "frac": alg_bytes / (ms_step * 1e-3) / 1e9 / hbm},
                         "spmv": {"achieved": spmv_bytes / (kernel_ms["spmv"] * 1e-3) / 1e9, "peak": hbm,
                                  "frac": spmv_bytes / (kernel_ms["spmv"] * 1e-3) / 1e9 / hbm, "algorithmic_bytes": spmv_bytes}},
            "newton_cg": newton,
            "e2e": {"value": total_E / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * F,
                    "d2h_bytes_per_step": 16 * F, "ms_per_step": e2e_ms,
                    "note": "per rank: u from pinned host memory, R and diag(J) back; the CSR values stay device-resident by design"},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        line.update(extra)
        if solve_block is not None:
            line["solve_to_tolerance"] = solve_block
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_block(args, fj, _lib, torch, dev)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def partition_parity(fj, part, torch, dist, dev, rank, world, n):
    """A converged linear-elasticity solve (Jacobi-CG to a relative residual of 1e-12) of the jittered n^3 grid cut into
    `world` element ranges and solved over NCCL, against the same solve on one GPU (every rank runs the single-GPU solve of
    this small problem itself): largest relative difference of the solutions over all ranks and the CG iteration counts."""
    Vx, cells = fj.uniform_hex_grid(n, n, n)
    rng = np.random.default_rng(0)
    interior = np.all((Vx > 1e-12) & (Vx < 1 - 1e-12), axis=1)
    Vx = Vx + 0.2 / n * rng.uniform(-1.0, 1.0, size=Vx.shape) * interior[:, None]
    mp = np.array([70e9, 0.25])
    fe = fj.FiniteElementType(fj.CellType.hexahedron, fj.ElementFamily.P, 1, fj.LagrangeVariant.equispaced,
                              fj.QuadratureType.default, 2)
    left = np.nonzero(Vx[:, 0] < 1e-12)[0]
    right = np.nonzero(Vx[:, 0] > 1 - 1e-12)[0]
    bcs = np.concatenate([np.stack([np.repeat(left, 3), np.tile(np.arange(3), left.size)], axis=1),
                          np.stack([right, np.zeros(right.size, dtype=np.int64)], axis=1)])
    vals = np.concatenate([np.zeros(3 * left.size), np.full(right.size, 0.01)])
    opts = fj.SolverOptions(linear_precond_type=fj.PreconditionerType.JACOBI, linear_relative_tol=1e-12,
                            linear_absolute_tol=0.0, linear_max_iter=2000, nonlinear_max_iter=1)
    token = fj.linear_elasticity_residual
    make = lambda c: fj.ElementBatch(fe_type=fe, n_dofs_per_basis=3, connectivity_en=c, constitutive_model=fj.elastic_isotropic,
                                     material_params=mp)
    u_ref, _, _ = fj.solve_bvp(Vx, [make(cells)], token, None, bcs, vals, solver_options=opts)
    it_ref = int(fj.solve_bvp.last_info.cumulative_linear_iterations)
    lm = part.partition_mesh(Vx, cells, world)[rank]
    g = lm.global_vertex_ids

    def dirichlet_local(m):
        keep = np.isin(bcs[:, 0], g)
        loc = np.stack([np.searchsorted(g, bcs[keep, 0]), bcs[keep, 1]], axis=1)
        return loc, vals[keep]

    dp = part.DistributedProblem(fj, part.TorchDistComm(), [lm], lambda m: [make(m.cells)], dirichlet_local, [dev], token)
    rp = dp.rp[0]
    us, info = dp.newton([torch.zeros(rp.F, dtype=torch.float64, device=dev)], opts)
    gd = torch.from_numpy(g).to(dev)
    ref_loc = u_ref.view(-1, 3)[gd].reshape(-1)
    err = ((us[0] - ref_loc).abs().max() / u_ref.abs().max()).reshape(1)
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    it_par = int(info["linear_iterations"][0]) if info["linear_iterations"] else -1
    return {"mesh": f"hex8 {n}^3 jittered, linear elasticity, Jacobi-CG rtol 1e-12, {world} element ranges over NCCL vs one GPU",
            "max_rel": float(err.item()), "iterations_partitioned": it_par, "iterations_single_gpu": it_ref,
            "iterations_equal": it_par == it_ref,
            "data_plane": "C (feb200_halo_sum / feb200_cg_iterate_dist)" if dp.plane is not None else "torch.distributed"}


def cpu_baseline_block(args, fj, _lib, torch, dev):
    """CPU legs beside the GPU numbers, bounded to tens of seconds: the oracle's reference-style
    assembly at 32^3 on every host core at once, and a SciPy-CSR CG iteration (SpMV + the vector updates of
    solve_cg.py:20-63) on the 64^3 matrix -- assembled on the GPU and copied to the host, because the reference-style CPU
    assembly of that size alone would take minutes."""
    n_a = max(args.cpu_n, 32) if args.cpu_n >= 20 else args.cpu_n
    eps, sec, Ec, cores = cpu_reference_parallel(n_a, layout=args.layout)
    out = {"value": eps, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": f"oracle reference-style assembly (per-element K_e + COO lexsort + duplicate sum) on a {n_a}^3 hex8 grid "
                     f"({Ec} elements, {sec:.1f} s) on {cores} host cores concurrently (aggregate)"}
    try:
        import scipy.sparse as sp
        m = 64 if args.cpu_n >= 20 else 8
        Vx, cells, mp = make_problem(m, m, "M")
        fe = fj.FiniteElementType(fj.CellType.hexahedron, fj.ElementFamily.P, 1, fj.LagrangeVariant.equispaced,
                                  fj.QuadratureType.default, 2)
        batch = fj.ElementBatch(fe_type=fe, n_dofs_per_basis=3, connectivity_en=cells, constitutive_model=fj.elastic_isotropic,
                                material_params=mp)
        ebc = fj.batch_to_collection(Vx, [batch])
        maps = [fj.mesh_to_sparse_assembly_map(Vx.shape[0], cells)]
        F = 3 * Vx.shape[0]
        mask = torch.zeros(F, dtype=torch.float64, device=dev)
        mask[:3 * (m + 1) ** 2] = 1.0
        J = fj.calculate_jacobian_w_dirichlet(fj.linear_elasticity_residual, ebc, maps, torch.zeros(F, dtype=torch.float64, device=dev),
                                              torch.zeros_like(mask), mask)
        csr = fj.coo_to_csr(J)
        A = sp.csr_matrix((csr.data.cpu().numpy(), csr.indices.cpu().numpy(), csr.indptr.cpu().numpy()), shape=(F, F))
        rng = np.random.default_rng(0)
        b = rng.normal(size=F)
        x, r = np.zeros(F), b.copy()
        p, gamma = r.copy(), float(r @ r)
        its, t0 = 20, time.perf_counter()
        for _ in range(its):
            Ap = A @ p
            alpha = gamma / float(p @ Ap)
            x += alpha * p
            r -= alpha * Ap
            g2 = float(r @ r)
            p = r + (g2 / gamma) * p
            gamma = g2
        dt = (time.perf_counter() - t0) / its
        out["cg"] = {"ms_per_iteration": dt * 1e3, "grid": f"{m}^3 hex8 ({F} DoFs, {A.nnz} non-zeros)", "cores": 1,
                     "kind": "scipy.sparse CSR SpMV + NumPy vector updates (solve_cg.py:20-63 recurrence), matrix assembled "
                             "on the GPU and copied to the host"}
    except Exception as exc:                       # the CG leg is additional context; the assembly leg is the baseline
        out["cg"] = {"error": repr(exc)}
    return out


if __name__ == "__main__":
    main()
