"""Element-partitioned multi-GPU mode of the hot path.

The reference has no distributed solver (only mpi4py demos under tests/test_mpi4py_cuda and the remark in
/root/reference/src/fe_jax/fea.py:331-332 that rank-vs-global numbering is future work).  Design here:

* elements are cut into P contiguous ranges (uniform grid: x-slabs, which are contiguous in both element
  and vertex numbering); each rank holds its elements, the closure of their vertices in a local numbering
  that preserves the global order, and the sub-assembled local CSR (interface rows hold partial sums);
* the only data-path communication is (1) the **halo-sum** of interface DoFs after a local scatter or SpMV
  (point-to-point with the ranks that share vertices, partial values added in ascending rank order on every
  sharer, so all copies are bit-identical) and (2) scalar all-reduces of the CG dot products, computed over
  owned DoFs only (ownership weight = 1 on the lowest rank touching a vertex);
* a constrained interface DoF gets its unit diagonal on the owner rank only, so that the halo-sum of the
  eliminated operator is still (I-M) J (I-M) + M.

`Comm` abstracts the transport: `TorchDistComm` (one process per GPU, NCCL; gloo on CPU for tests) or
`InProcessComm`, which runs all ranks lock-step inside one process -- the way to exercise the multi-rank
numerics on a single GPU (never several spinning processes on one device).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib


# ------------------------------------------------------------------------------------------ partition (host)
@dataclass
class LocalMesh:
    rank: int
    world: int
    vertices: np.ndarray            # (V_loc, D)
    cells: np.ndarray               # (E_loc, N) local vertex ids, int64
    global_vertex_ids: np.ndarray   # (V_loc,) ascending
    element_range: tuple            # (e0, e1) in the global element numbering
    neighbours: dict = field(default_factory=dict)   # rank p -> local vertex ids shared with p (ascending global id)
    owned: np.ndarray = None        # (V_loc,) bool: this rank is the lowest one touching the vertex

    @property
    def n_shared(self):
        return sum(len(v) for v in self.neighbours.values())


def element_ranges(E: int, P: int):
    """P contiguous, nearly equal element ranges."""
    cuts = [(E * r) // P for r in range(P + 1)]
    return [(cuts[r], cuts[r + 1]) for r in range(P)]


def partition_mesh(vertices, cells, P: int, order=None):
    """Cut (vertices, cells) into P LocalMesh objects.  `order` optionally permutes the elements first
    (e.g. np.argsort of centroids along the longest axis for unstructured meshes)."""
    vertices = np.asarray(vertices)
    cells = np.asarray(cells).astype(np.int64)
    if order is not None:
        cells = cells[order]
    ranges = element_ranges(cells.shape[0], P)
    gids = [np.unique(cells[a:b]) for a, b in ranges]
    touch_min = np.full(vertices.shape[0], P, dtype=np.int64)        # lowest rank touching each vertex
    for r in reversed(range(P)):
        touch_min[gids[r]] = r
    out = []
    for r, (a, b) in enumerate(ranges):
        g = gids[r]
        lm = LocalMesh(rank=r, world=P, vertices=vertices[g], cells=np.searchsorted(g, cells[a:b]),
                       global_vertex_ids=g, element_range=(a, b))
        for p in range(P):
            if p == r:
                continue
            shared = np.intersect1d(g, gids[p], assume_unique=True)
            if shared.size:
                lm.neighbours[p] = np.searchsorted(g, shared)
        lm.owned = touch_min[g] == r
        out.append(lm)
    return out


def slab_local_mesh(rank: int, world: int, vertices, cells, lo_plane, hi_plane):
    """LocalMesh of one x-slab built from LOCAL data only (used by bench.py at sizes where no rank can hold the
    global mesh): `lo_plane` / `hi_plane` are the local vertex ids on the faces shared with rank-1 / rank+1,
    listed in the same (global) order on both sides."""
    lm = LocalMesh(rank=rank, world=world, vertices=np.asarray(vertices), cells=np.asarray(cells).astype(np.int64),
                   global_vertex_ids=None, element_range=(None, None))
    lm.owned = np.ones(lm.vertices.shape[0], dtype=bool)
    if rank > 0:
        lm.neighbours[rank - 1] = np.asarray(lo_plane, dtype=np.int64)
        lm.owned[lo_plane] = False
    if rank < world - 1:
        lm.neighbours[rank + 1] = np.asarray(hi_plane, dtype=np.int64)
    return lm


# ------------------------------------------------------------------------------------------ transport
class InProcessComm:
    """All ranks live in this process; collectives are evaluated lock-step over the list of per-rank tensors."""
    world = None

    def __init__(self, world):
        self.world = world
        self.ranks = list(range(world))

    def exchange(self, send):
        """send[r][p] = tensor rank r sends to p  ->  recv[r][p] = tensor r receives from p"""
        return [{p: send[p][r].to(send[r][p].device) for p in send[r]} for r in range(self.world)]

    def all_reduce_sum(self, tensors):
        total = tensors[0].clone()
        for t in tensors[1:]:
            total = total + t.to(total.device)
        for t in tensors:
            t.copy_(total.to(t.device))


class TorchDistComm:
    """One process per rank (torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests)."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = dist.get_world_size()
        self.ranks = [dist.get_rank()]

    def exchange(self, send):
        dist = self.dist
        (mine,) = send
        recv = {p: torch.empty_like(t) for p, t in mine.items()}
        ops = []
        for p in sorted(mine):
            ops.append(dist.P2POp(dist.isend, mine[p].contiguous(), p))
            ops.append(dist.P2POp(dist.irecv, recv[p], p))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return [recv]

    def all_reduce_sum(self, tensors):
        (t,) = tensors
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)


class NcclDataPlane:
    """The C data plane of the partitioned mode (csrc/dist.cu): feb200's own NCCL communicator, created from a unique id
    that rank 0 draws and torch.distributed broadcasts, plus one halo descriptor per RankProblem.  With it the halo-sum and
    the whole CG loop (local SpMV, halo-sum, all-reduced dot products, vector updates) are enqueued from C on the current
    stream; Python only polls the `done` flag between chunks of iterations."""

    def __init__(self, device):
        import ctypes
        import torch.distributed as dist
        lib = _lib.load()
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        idt = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (ctypes.c_uint8 * 128)()
            if lib.feb200_nccl_unique_id(buf) != 0:
                raise _lib.Feb200Error("feb200_nccl_unique_id failed: " + lib.feb200_last_error().decode())
            idt = torch.tensor(list(buf), dtype=torch.uint8)
        idt = idt.to(device)
        dist.broadcast(idt, 0)
        raw = bytes(idt.cpu().tolist())
        comm = ctypes.c_void_p()
        with torch.cuda.device(device):
            if lib.feb200_nccl_comm_init(self.rank, self.world, raw, ctypes.byref(comm)) != 0:
                raise _lib.Feb200Error("feb200_nccl_comm_init failed: " + lib.feb200_last_error().decode())
        self.comm = comm
        self._keep = []

    @staticmethod
    def usable(comm, devices):
        if not isinstance(comm, TorchDistComm) or comm.world < 2 or len(devices) != 1:
            return False
        if not torch.cuda.is_available() or torch.device(devices[0]).type != "cuda":
            return False
        return bool(_lib.load().feb200_nccl_available())

    def halo(self, nb, U, device):
        """feb200_halo for one rank's neighbour index sets {peer: local vertex ids (int64 tensor)}"""
        assert len(nb) <= _lib.MAX_NEIGHBOURS, "too many neighbour ranks for the C halo descriptor"
        h = _lib.Halo(n_neighbours=len(nb), U=U)
        for k, p in enumerate(sorted(nb)):
            idx = nb[p].to(torch.int64).contiguous()
            send = torch.empty(idx.numel() * U, dtype=torch.float64, device=device)
            recv = torch.empty_like(send)
            self._keep.append((idx, send, recv))
            h.peer[k], h.count[k] = p, idx.numel()
            h.idx[k], h.send[k], h.recv[k] = _lib.ptr(idx), _lib.ptr(send), _lib.ptr(recv)
        return h

    def halo_sum(self, h, v):
        _lib.call("feb200_halo_sum", self.comm, h, _lib.ptr(v))

    def all_reduce_sum(self, t):
        _lib.call("feb200_dot_allreduce", self.comm, _lib.ptr(t), t.numel())


def halo_meta(nb):
    """(touched vertex ids, True if some vertex is shared with more than one neighbour) for one rank's index sets"""
    if not nb:
        return None, False
    allidx = torch.cat([idx for idx in nb.values()])
    touched = allidx.unique()
    return touched, bool(touched.shape[0] != allidx.shape[0])


def halo_sum(comm, meshes_idx, vectors, U):
    """In place: every interface DoF becomes the sum of the partial values of all ranks sharing it, added in
    ascending rank order (identical bits on every sharer).  meshes_idx[i] = (rank, {p: local vertex ids (tensor)}
    [, halo_meta(...)]): the optional third entry avoids recomputing the interface bookkeeping on every call.
    When every interface vertex has exactly two sharers (slab partitions) the sum is own + received, which is
    already order independent, so the received values are simply added."""
    send = []
    for entry, v in zip(meshes_idx, vectors):
        nb = entry[1]
        v3 = v.view(-1, U)
        send.append({p: v3[idx] for p, idx in nb.items()})
    recv = comm.exchange(send)
    for entry, v, rcv, snd in zip(meshes_idx, vectors, recv, send):
        rank, nb = entry[0], entry[1]
        if not nb:
            continue
        v3 = v.view(-1, U)
        touched, multi = entry[2] if len(entry) > 2 else halo_meta(nb)
        if not multi:
            for p in sorted(nb):
                v3.index_add_(0, nb[p], rcv[p])
            continue
        own = v3[touched].clone()
        v3[touched] = 0.0
        placed = False
        for p in sorted(nb):
            if p > rank and not placed:
                v3[touched] += own
                placed = True
            v3.index_add_(0, nb[p], rcv[p])
        if not placed:
            v3[touched] += own


# ------------------------------------------------------------------------------------------ per-rank problem state
class RankProblem:
    """Everything one rank holds: local collection, maps, pattern, Dirichlet data, ownership weights."""

    def __init__(self, fj, lm: LocalMesh, make_batches, dirichlet_bcs_local, dirichlet_values_local, device):
        self.lm = lm
        self.device = device
        with torch.cuda.device(device):
            batches = make_batches(lm)
            self.batches = batches
            self.ebc = fj.batch_to_collection(lm.vertices, batches)
            V = lm.vertices.shape[0]
            self.maps = [fj.mesh_to_sparse_assembly_map(V, b.connectivity_en) for b in batches]
            from .assembly import _plan
            self.plan = _plan(self.ebc, self.maps)
            self.U = self.ebc.U[0]
            self.F = V * self.U
            bcs = np.asarray(dirichlet_bcs_local).astype(np.int64).reshape(-1, 2)
            dofs = self.ebc.D * bcs[:, 0] + bcs[:, 1]
            self.dofs = torch.from_numpy(dofs).to(device)
            self.values = torch.from_numpy(np.asarray(dirichlet_values_local, dtype=np.float64)).to(device)
            self.values_g = torch.zeros(self.F, dtype=torch.float64, device=device)
            self.values_g[self.dofs] = self.values
            self.mask_g = torch.zeros(self.F, dtype=torch.float64, device=device)
            self.mask_g[self.dofs] = 1.0
            owned = torch.from_numpy(lm.owned).to(device)
            self.weights = owned.to(torch.float64).repeat_interleave(self.U).contiguous()
            self.nb = {p: torch.from_numpy(np.asarray(idx)).to(device) for p, idx in lm.neighbours.items()}
            self.idx = (lm.rank, self.nb, halo_meta(self.nb))
            self.halo = None                      # feb200_halo, set by DistributedProblem when the C data plane is used


class DistributedProblem:
    """SPMD driver over the ranks visible to this process (1 under torch.distributed, P under InProcessComm)."""

    def __init__(self, fj, comm, local_meshes, make_batches, dirichlet_local, devices, element_residual_func):
        self.fj, self.comm = fj, comm
        self.token = element_residual_func
        self.rp = [RankProblem(fj, lm, make_batches, *dirichlet_local(lm), dev) for lm, dev in zip(local_meshes, devices)]
        # C data plane (own NCCL communicator): one rank per process on a GPU, every interface vertex shared by two ranks
        self.plane = None
        if NcclDataPlane.usable(comm, devices) and not self.rp[0].idx[2][1]:
            self.plane = NcclDataPlane(devices[0])
            rp = self.rp[0]
            rp.halo = self.plane.halo(rp.nb, rp.U, rp.device)

    def _halo_sum(self, vectors):
        if self.plane is not None:
            self.plane.halo_sum(self.rp[0].halo, vectors[0])
        else:
            halo_sum(self.comm, [rp.idx for rp in self.rp], vectors, self.rp[0].U)

    def _each(self, fn):
        out = []
        for rp in self.rp:
            with torch.cuda.device(rp.device):
                out.append(fn(rp))
        return out

    # R_w(u) with the interface halo-sum between the local scatter and the Dirichlet wrapper
    def residual_w_dirichlet(self, us):
        from .assembly import _residual
        Rs = []
        for rp, u in zip(self.rp, us):
            with torch.cuda.device(rp.device):
                # blend fused in the element kernel; the Dirichlet residual formula must wait for the halo-sum
                R, _ = _residual(self.token, rp.ebc, rp.maps, u, rp.mask_g, rp.values_g, fuse_dirichlet=False)
                Rs.append(R)
        self._halo_sum(Rs)
        for rp, u, R in zip(self.rp, us, Rs):
            with torch.cuda.device(rp.device):
                _lib.call("feb200_dirichlet_residual", rp.F, _lib.ptr(u), _lib.ptr(rp.values_g), _lib.ptr(rp.mask_g), _lib.ptr(R))
        return Rs

    # local J_w: assembled at u_w, eliminated; unit diagonal of constrained interface DoFs kept on the owner only
    def jacobian_w_dirichlet(self, us, outs=None):
        from .assembly import _assemble_values
        out = []
        for i, (rp, u) in enumerate(zip(self.rp, us)):
            with torch.cuda.device(rp.device):
                vals, pat = _assemble_values(self.token, rp.ebc, rp.maps, u, rp.mask_g, rp.values_g,
                                             out=None if outs is None else outs[i], eliminate=True)
                pos = self._not_owned_diagonal(rp, pat)
                if pos is not None:
                    vals[pos] = 0.0
                out.append(vals)
        return out

    @staticmethod
    def _not_owned_diagonal(rp, pat):
        """CSR positions of the unit diagonals of constrained interface DoFs this rank does not own (computed once per
        pattern: no host synchronisation on the assembly path)."""
        hit = getattr(rp, "_nod", None)
        if hit is None or hit[0] is not pat:
            dofs = rp.dofs[rp.weights[rp.dofs] == 0.0]
            pos = None
            if dofs.numel():
                a = dofs // pat.U
                i = dofs - a * pat.U
                p0 = pat.nindptr[a]
                deg = pat.nindptr[a + 1] - p0
                pos = (pat.U * pat.U * p0 + i * pat.U * deg + pat.U * pat.diag_rank[a].to(torch.int64) + i).contiguous()
            rp._nod = hit = (pat, pos)
        return hit[1]

    def norm(self, vs):
        parts = []
        for rp, v in zip(self.rp, vs):
            with torch.cuda.device(rp.device):
                parts.append((rp.weights * v * v).sum().reshape(1))
        self.comm.all_reduce_sum(parts)
        return float(parts[0].item()) ** 0.5

    def cg(self, Jvals, bs, x0s, tol, atol, maxiter, jacobi=False, poll_every=25):
        """Distributed CG on the sub-assembled matrices; returns ([x], iterations, history)."""
        lib = _lib.load()
        comm = self.comm
        sts, keep = [], []
        for rp, vals, b, x0 in zip(self.rp, Jvals, bs, x0s):
            with torch.cuda.device(rp.device):
                dev, n, pat = rp.device, rp.F, rp.plan.pattern
                x = x0.clone()
                r, p, Ap = (torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3))
                hist = torch.zeros(max(maxiter, 1), dtype=torch.float64, device=dev)
                scalars = torch.zeros(8, dtype=torch.float64, device=dev)
                counters = torch.zeros(4, dtype=torch.int32, device=dev)
                ws = _lib.workspace(lib.feb200_reduce_workspace_bytes(), dev)
                inv_diag = None
                if jacobi:
                    from .assembly import csr_diagonal
                    d = csr_diagonal(pat, vals)
                    keep.append(d)
                    inv_diag = d
                st = _lib.CGState(V=pat.V, U=pat.U, maxiter=maxiter, nindptr=_lib.ptr(pat.nindptr),
                                  nindices=_lib.ptr(pat.nindices), values=_lib.ptr(vals), inv_diag=None,
                                  weights=_lib.ptr(rp.weights), b=_lib.ptr(b), x=_lib.ptr(x), r=_lib.ptr(r), p=_lib.ptr(p),
                                  Ap=_lib.ptr(Ap), hist=_lib.ptr(hist), scalars=_lib.ptr(scalars),
                                  counters=_lib.ptr(counters), ws=_lib.ptr(ws), ws_bytes=ws.numel())
                sts.append(dict(st=st, x=x, r=r, p=p, Ap=Ap, hist=hist, scalars=scalars, counters=counters, ws=ws,
                                inv_diag=inv_diag, b=b, vals=vals))
        idx = [rp.idx for rp in self.rp]
        U = self.rp[0].U
        if jacobi:      # diag(J_w) = halo-sum of the local diagonals
            self._halo_sum([s["inv_diag"] for s in sts])
            for s in sts:
                s["inv_diag"] = (1.0 / s["inv_diag"]).contiguous()
                s["st"].inv_diag = _lib.ptr(s["inv_diag"])

        if self.plane is not None:      # the whole loop enqueued from C (csrc/dist.cu); one host sync per chunk of iterations
            s0, rp = sts[0], self.rp[0]
            _lib.call("feb200_cg_init_dist", self.plane.comm, s0["st"], rp.halo, float(tol), float(atol))
            it = 0
            while it < maxiter:
                nit = min(poll_every, maxiter - it)
                _lib.call("feb200_cg_iterate_dist", self.plane.comm, s0["st"], rp.halo, nit)
                it += nit
                if int(s0["counters"][1].item()) != 0:
                    break
            return [s0["x"]], int(s0["counters"][0].item()), s0["hist"][:maxiter]

        def stage(k, tol_=0.0, atol_=0.0):
            for rp, s in zip(self.rp, sts):
                with torch.cuda.device(rp.device):
                    _lib.call("feb200_cg_stage", s["st"], k, float(tol_), float(atol_))

        def reduce(lo, hi):
            comm.all_reduce_sum([s["scalars"][lo:hi] for s in sts])

        stage(_lib.CG_INIT_SPMV)
        halo_sum(comm, idx, [s["Ap"] for s in sts], U)
        stage(_lib.CG_INIT_VECTORS, tol, atol)
        reduce(2, 4)
        reduce(5, 6)
        stage(_lib.CG_INIT_FINALIZE)
        it = 0
        while it < maxiter:
            for _ in range(min(poll_every, maxiter - it)):     # (finished solves: the stage kernels leave zero partials)
                stage(_lib.CG_SPMV)                            # Ap = A_loc p and the local p.(A_loc p)
                halo_sum(comm, idx, [s["Ap"] for s in sts], U)
                reduce(1, 2)                                   # global p.Ap = sum over ranks of the local ones
                stage(_lib.CG_UPDATE_XR)
                reduce(2, 4)
                stage(_lib.CG_UPDATE_P)
                it += 1
            if int(sts[0]["counters"][1].item()) != 0:        # identical on every rank (all-reduced scalars)
                break
        k = int(sts[0]["counters"][0].item())
        return [s["x"] for s in sts], k, sts[0]["hist"][:maxiter]

    def newton(self, u0s, options):
        """solve_nonlinear_step (fea.py:968-1129) on the partitioned problem; returns ([u], info dict)."""
        us = [u.clone() for u in u0s]
        Rs = self.residual_w_dirichlet(us)
        r0 = self.norm(Rs)
        it, lin = 0, []
        while True:
            a = self.norm(Rs)
            rel = a / r0 if r0 != 0.0 else float("nan")
            if not (it < options.nonlinear_max_iter and rel > options.nonlinear_relative_tol
                    and a > options.nonlinear_absolute_tol):
                break
            J = self.jacobian_w_dirichlet(us)
            bs, x0s = [], []
            for rp, u, R in zip(self.rp, us, Rs):
                with torch.cuda.device(rp.device):
                    bs.append((-R).contiguous())
                    dx = torch.zeros_like(R)
                    dx[rp.dofs] = rp.values - u[rp.dofs]
                    x0s.append(dx)
            from .solver import PreconditionerType
            dxs, k, hist = self.cg(J, bs, x0s, options.linear_relative_tol, options.linear_absolute_tol,
                                   options.linear_max_iter, jacobi=options.linear_precond_type == PreconditionerType.JACOBI)
            lin.append(k)
            for rp, u, dx in zip(self.rp, us, dxs):
                with torch.cuda.device(rp.device):
                    dx[rp.dofs] = rp.values - u[rp.dofs]
                    u += dx
            Rs = self.residual_w_dirichlet(us)
            it += 1
        return us, dict(nonlinear_iterations=it, linear_iterations=lin, residual_norm=self.norm(Rs))
