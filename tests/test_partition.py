"""Element-partitioned multi-GPU mode.
CPU: partition bookkeeping and the halo-sum / weighted-dot collectives, in-process for P = 2..4 and with
real torch.distributed (gloo, world_size 2).  GPU (one device): the whole partitioned Newton-CG run lock-step
through InProcessComm against the single-GPU solve."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import fe_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup_path():
    for p in (ROOT, os.path.join(ROOT, "fea-in-jax_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)


_setup_path()
from fe_jax import partition as part          # noqa: E402


def _mesh(shuffle):
    V, F = orc.uniform_tri_grid(7, 6)
    if shuffle:
        F = F[np.random.default_rng(3).permutation(F.shape[0])]
    return V[:, :2], F


@pytest.mark.parametrize("P", [2, 3, 4])
@pytest.mark.parametrize("shuffle", [False, True])
def test_partition_bookkeeping(P, shuffle):
    V, F = _mesh(shuffle)
    lms = part.partition_mesh(V, F, P)
    assert sum(lm.cells.shape[0] for lm in lms) == F.shape[0]
    owned_total = np.zeros(V.shape[0], dtype=int)
    for lm in lms:
        a, b = lm.element_range
        assert np.array_equal(lm.global_vertex_ids[lm.cells], F[a:b])          # local numbering round-trips
        assert np.all(np.diff(lm.global_vertex_ids) > 0)
        owned_total[lm.global_vertex_ids[lm.owned]] += 1
        for p, idx in lm.neighbours.items():                                    # symmetric, same order on both sides
            other = lms[p].neighbours[lm.rank]
            assert np.array_equal(lm.global_vertex_ids[idx], lms[p].global_vertex_ids[other])
    assert np.all(owned_total == 1)                                             # every vertex owned exactly once


@pytest.mark.parametrize("P", [2, 3, 4])
@pytest.mark.parametrize("shuffle", [False, True])
def test_halo_sum_and_weighted_dot_in_process(P, shuffle):
    V, F = _mesh(shuffle)
    U = 2
    lms = part.partition_mesh(V, F, P)
    rng = np.random.default_rng(0)
    # per-element contributions scattered locally, then halo-summed == global scatter
    contrib = rng.normal(size=(F.shape[0], 3, U))
    glob = np.zeros((V.shape[0], U))
    np.add.at(glob, F, contrib)
    vecs, idx = [], []
    for lm in lms:
        a, b = lm.element_range
        loc = np.zeros((lm.vertices.shape[0], U))
        np.add.at(loc, lm.cells, contrib[a:b])
        vecs.append(torch.from_numpy(loc.reshape(-1).copy()))
        idx.append((lm.rank, {p: torch.from_numpy(i) for p, i in lm.neighbours.items()}))
    comm = part.InProcessComm(P)
    part.halo_sum(comm, idx, vecs, U)
    copies = {}
    for lm, v in zip(lms, vecs):
        got = v.numpy().reshape(-1, U)
        assert np.allclose(got, glob[lm.global_vertex_ids], rtol=1e-13, atol=1e-13)
        for g, row in zip(lm.global_vertex_ids, got):
            copies.setdefault(int(g), []).append(row.copy())
    for g, rows in copies.items():                     # all sharers hold bit-identical values
        for r in rows[1:]:
            assert np.array_equal(rows[0], r)
    # ownership-weighted dot + all-reduce == global dot
    x = rng.normal(size=V.shape[0] * U)
    parts = []
    for lm in lms:
        w = np.repeat(lm.owned.astype(float), U)
        xl = x.reshape(-1, U)[lm.global_vertex_ids].reshape(-1)
        parts.append(torch.tensor([float((w * xl * xl).sum())], dtype=torch.float64))
    comm.all_reduce_sum(parts)
    assert all(abs(float(p) - float(x @ x)) <= 1e-12 * float(x @ x) for p in parts)


def _gloo_worker(rank, world, port, q):
    _setup_path()
    import torch.distributed as dist
    from fe_jax import partition as part2
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        V, F = _mesh(True)
        U = 2
        lm = part2.partition_mesh(V, F, world)[rank]
        contrib = np.random.default_rng(0).normal(size=(F.shape[0], 3, U))
        glob = np.zeros((V.shape[0], U))
        np.add.at(glob, F, contrib)
        a, b = lm.element_range
        loc = np.zeros((lm.vertices.shape[0], U))
        np.add.at(loc, lm.cells, contrib[a:b])
        v = torch.from_numpy(loc.reshape(-1).copy())
        comm = part2.TorchDistComm()
        part2.halo_sum(comm, [(rank, {p: torch.from_numpy(i) for p, i in lm.neighbours.items()})], [v], U)
        ok = np.allclose(v.numpy().reshape(-1, U), glob[lm.global_vertex_ids], rtol=1e-13, atol=1e-13)
        w = np.repeat(lm.owned.astype(float), U)
        t = torch.tensor([float((w * v.numpy() * v.numpy()).sum())], dtype=torch.float64)
        comm.all_reduce_sum([t])
        ok = ok and abs(float(t) - float((glob * glob).sum())) <= 1e-12 * float((glob * glob).sum())
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_halo_sum_torch_distributed_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]


# ------------------------------------------------------------------------------------------ GPU: lock-step ranks on one device
@pytest.mark.gpu
@pytest.mark.parametrize("P,model,prec", [(2, "elastic_isotropic", "NONE"), (3, "elastic_isotropic", "JACOBI"),
                                          (2, "st_venant_kirchhoff", "JACOBI")])
def test_partitioned_newton_cg_matches_single_gpu(P, model, prec):
    os.environ.setdefault("FEB200_QUIET", "1")
    import __graft_entry__ as graft
    graft.build()
    import fe_jax as fj
    n = 6
    Vx, cells, u, mp_, bcs, vals = orc.synthetic_hex_problem(n, material_layout="M")
    if model != "elastic_isotropic":
        mp_ = np.array([1.0e6, 0.3])
        vals = vals * 5.0
    token = fj.linear_elasticity_residual if model == "elastic_isotropic" else fj.hyperelasticity_residual
    fe = fj.FiniteElementType(fj.CellType.hexahedron, fj.ElementFamily.P, 1, fj.LagrangeVariant.equispaced,
                              fj.QuadratureType.default, 2)
    mk = lambda conn: [fj.ElementBatch(fe_type=fe, n_dofs_per_basis=3, connectivity_en=conn,
                                       constitutive_model=getattr(fj, model), material_params=mp_)]
    opts = fj.SolverOptions(linear_precond_type=getattr(fj.PreconditionerType, prec), linear_max_iter=3000,
                            linear_relative_tol=1e-12, nonlinear_max_iter=12)
    u_ref, _, _ = fj.solve_bvp(Vx, mk(cells), token, None, bcs, vals, solver_options=opts)
    info_ref = fj.solve_bvp.last_info

    lms = fj.partition.partition_mesh(Vx, cells, P)
    bc_vertex, bc_comp = bcs[:, 0], bcs[:, 1]

    def dirichlet_local(lm):
        pos = np.searchsorted(lm.global_vertex_ids, bc_vertex)
        pos = np.clip(pos, 0, lm.global_vertex_ids.shape[0] - 1)
        here = lm.global_vertex_ids[pos] == bc_vertex
        return np.stack([pos[here], bc_comp[here]], axis=1), vals[here]

    dev = torch.device("cuda", 0)
    dp = fj.partition.DistributedProblem(fj, fj.partition.InProcessComm(P), lms, lambda lm: mk(lm.cells), dirichlet_local,
                                         [dev] * P, token)
    u0 = [torch.zeros(rp.F, dtype=torch.float64, device=dev) for rp in dp.rp]
    us, info = dp.newton(u0, opts)
    assert info["nonlinear_iterations"] == info_ref.nonlinear_iterations
    assert abs(sum(info["linear_iterations"]) - info_ref.cumulative_linear_iterations) <= 6
    ur = u_ref.cpu().numpy().reshape(-1, 3)
    for lm, ul in zip(lms, us):
        got = ul.cpu().numpy().reshape(-1, 3)
        assert np.abs(got - ur[lm.global_vertex_ids]).max() <= 1e-8 * np.abs(ur).max()
    # interface copies are bit-identical across ranks
    for lm, ul in zip(lms, us):
        for p, idx in lm.neighbours.items():
            mine = ul.cpu().numpy().reshape(-1, 3)[idx]
            theirs = us[p].cpu().numpy().reshape(-1, 3)[lms[p].neighbours[lm.rank]]
            assert np.array_equal(mine, theirs)
