"""Streaming plan of the fused hex8 assembly kernel feb200_assemble_fused_hex8_sweep (csrc/fused_sweep.cu).

Built once per mesh with torch index operations (setup plumbing, not on the hot path; runs on the device the mesh lives
on, and on the CPU for the host-side tests).  Vertices are binned on a grid of the mean vertex spacing; a *strip* is a
2 x 2 column of bins along the axis the vertex numbering runs fastest in, its vertices sorted along that axis and cut
into *chunks* of 4 (on a structured grid: one chunk = one layer of the column).  Strips are cut into *segments* of about
`seg_chunks` chunks; a segment is the unit of work of one warp.  Within a segment the elements touching its vertices are
turned into *visits*: a visit at step c covers the element's nodes that live in chunks c and c+1 (the window of 8 panel
slots the warp holds at that step), so every (vertex, incident element) pair is covered by exactly one visit and chunk
c is complete after the last visit of step c.  Works for any hex8 connectivity with vertex valence <= 8 and degree
<= 32 (elements whose in-segment nodes spread over non-adjacent chunks are simply visited more than once); returns None
when the mesh is not eligible (the caller then uses the tile kernel or the two-pass path).

Record layout (include/feb200.h, feb200_sweep_plan): visits (n, 12) int32 = element, slotw, perm, flags, the element's 8
vertex ids; chunk records (n, 16) int32 = v[4], deg[4], p0[4] as int64.
"""
from __future__ import annotations

import torch

from . import _lib

MAXVAL = 8
MAX_DEG = 32
SEG_CHUNKS = 48


def _quantise(vertices: torch.Tensor) -> torch.Tensor:
    """integer grid coordinates of every vertex on the lattice a uniform grid with this bounding box and vertex count
    would have: the spacing h solves prod_a (ext_a / h + 1) = V, which is exact for (jittered) uniform grids of any aspect
    ratio (jitter below half a cell) and a reasonable binning for anything else"""
    V, D = vertices.shape
    lo = vertices.min(dim=0).values
    hi = vertices.max(dim=0).values
    ext = [max(float(e), 1e-300) for e in (hi - lo).tolist()]
    a, b = 1e-12 * max(ext), 2.0 * max(ext)                   # bisection on h: the product decreases monotonically in h
    for _ in range(200):
        h = 0.5 * (a + b)
        prod = 1.0
        for e in ext:
            prod *= e / h + 1.0
        if prod > V:
            a = h
        else:
            b = h
    h = 0.5 * (a + b)
    counts = torch.tensor([max(round(e / h) + 1, 2) for e in ext], dtype=torch.float64, device=vertices.device)
    spacing = torch.tensor(ext, dtype=torch.float64, device=vertices.device) / (counts - 1.0)
    return torch.floor((vertices - lo) / spacing + 0.5).to(torch.int64).clamp_(0, (1 << 20) - 1)


def _sweep_axis(q: torch.Tensor) -> int:
    """the axis along which consecutive vertex ids advance most often (the fastest axis of the numbering)"""
    if q.shape[0] < 2:
        return q.shape[1] - 1
    dq = q[1:] - q[:-1]
    best, best_n = q.shape[1] - 1, -1
    for a in range(q.shape[1]):
        others = [c for c in range(q.shape[1]) if c != a]
        hit = (dq[:, a] == 1)
        for c in others:
            hit &= dq[:, c] == 0
        n = int(hit.sum().item())
        if n > best_n:
            best, best_n = a, n
    return best


_PERM_LUT = None


def _perm_lut(device):
    """mask (8 bits) -> (perm word: nibble i = i-th set bit, count)"""
    global _PERM_LUT
    if _PERM_LUT is None:
        perm, cnt = [], []
        for m in range(256):
            w, c = 0, 0
            for n in range(8):
                if (m >> n) & 1:
                    w |= n << (4 * c)
                    c += 1
            perm.append(w)
            cnt.append(c)
        _PERM_LUT = (torch.tensor(perm, dtype=torch.int64), torch.tensor(cnt, dtype=torch.int64))
    return _PERM_LUT[0].to(device), _PERM_LUT[1].to(device)


class SweepPlan:
    def __init__(self, struct, keep, stats):
        self.struct, self._keep, self.stats = struct, keep, stats
        (self.seg_visit_ptr, self.seg_chunk_ptr, self.visits, self.chunk_rec) = keep

    @property
    def chunk_v(self):
        return self.chunk_rec[:, 0:4]

    @property
    def chunk_deg(self):
        return self.chunk_rec[:, 4:8]

    @property
    def chunk_p0(self):
        return self.chunk_rec[:, 8:16].contiguous().view(torch.int64)


def build_sweep_plan(vertices: torch.Tensor, conn32: torch.Tensor, row_ind: torch.Tensor, col_ind: torch.Tensor,
                     nindptr: torch.Tensor, seg_chunks: int = SEG_CHUNKS):
    """vertices (V,3) f64, conn32 (E,8) i32, row_ind (V+1) / col_ind (E*8) int64 views of the assembly map,
    nindptr (V+1) int64 node pattern offsets."""
    import os
    seg_chunks = int(os.environ.get("FEB200_SWEEP_SEG", seg_chunks))
    if vertices.is_cuda and "FEB200_SWEEP_SEG" not in os.environ:
        # small meshes (one slab of a partitioned run): shorter segments so that every worker (warp) still gets about a
        # dozen of them -- with 5 segments per warp the round-robin assignment left 18 % of the workers one segment short
        workers = torch.cuda.get_device_properties(vertices.device).multi_processor_count * 8
        seg_chunks = max(8, min(seg_chunks, int(vertices.shape[0] / 4 / (12 * workers)) or 8))
    dev = vertices.device
    V = int(vertices.shape[0])
    E, N = int(conn32.shape[0]), int(conn32.shape[1])
    if N != 8 or vertices.shape[1] != 3 or V == 0 or E == 0:
        return None
    val = row_ind[1:] - row_ind[:-1]
    deg = nindptr[1:] - nindptr[:-1]
    max_deg = int(deg.max().item())
    if int(val.max().item()) > MAXVAL or max_deg > MAX_DEG or max_deg < 1:
        return None
    # ---- strips, chunks, segments
    q = _quantise(vertices)
    s_ax = _sweep_axis(q)
    c_ax = [a for a in range(3) if a != s_ax]
    k1, k2 = q[:, c_ax[0]] >> 1, q[:, c_ax[1]] >> 1
    n2 = int(k2.max().item()) + 1
    ns = int(q[:, s_ax].max().item()) + 1
    strip = k1 * n2 + k2
    key = ((strip * ns + q[:, s_ax]) << 2) | ((q[:, c_ax[0]] & 1) << 1) | (q[:, c_ax[1]] & 1)
    active = torch.nonzero(val > 0).reshape(-1)             # vertices without elements have no row to assemble
    if active.numel() == 0:
        return None
    order = active[torch.argsort(key[active], stable=True)]
    n_act = int(order.shape[0])
    strip_s = strip[order]
    new_strip = torch.ones(n_act, dtype=torch.bool, device=dev)
    new_strip[1:] = strip_s[1:] != strip_s[:-1]
    strip_id = torch.cumsum(new_strip.to(torch.int64), 0) - 1
    strip_start = torch.nonzero(new_strip).reshape(-1)
    n_strips = int(strip_start.shape[0])
    strip_len = torch.diff(torch.cat([strip_start, torch.tensor([n_act], device=dev)]))
    pos = torch.arange(n_act, device=dev) - strip_start[strip_id]
    strip_chunks = (strip_len + 3) // 4
    strip_segs = (strip_chunks + seg_chunks - 1) // seg_chunks
    cps = (strip_chunks + strip_segs - 1) // strip_segs                     # chunks per segment of the strip (last may be shorter)
    strip_segs = (strip_chunks + cps - 1) // cps
    seg_off = torch.cumsum(strip_segs, 0) - strip_segs
    chunk_off = torch.cumsum(strip_chunks, 0) - strip_chunks
    n_seg = int(strip_segs.sum().item())
    n_chunks = int(strip_chunks.sum().item())
    cis = pos // 4                                                           # chunk within the strip
    sis = cis // cps[strip_id]                                               # segment within the strip
    seg_s = seg_off[strip_id] + sis
    c_in_seg = cis - sis * cps[strip_id]
    chunk_s = chunk_off[strip_id] + cis
    slot_s = 4 * (c_in_seg & 1) + (pos & 3)
    # per-vertex lookups
    seg_of = torch.full((V,), -1, dtype=torch.int64, device=dev)
    cin_of = torch.zeros(V, dtype=torch.int64, device=dev)
    slot_of = torch.zeros(V, dtype=torch.int64, device=dev)
    seg_of[order], cin_of[order], slot_of[order] = seg_s, c_in_seg, slot_s
    # chunk records
    chunk_v = torch.full((n_chunks * 4,), -1, dtype=torch.int32, device=dev)
    chunk_p0 = torch.zeros(n_chunks * 4, dtype=torch.int64, device=dev)
    chunk_deg = torch.zeros(n_chunks * 4, dtype=torch.int32, device=dev)
    rec = chunk_s * 4 + (pos & 3)
    chunk_v[rec] = order.to(torch.int32)
    chunk_p0[rec] = nindptr[:-1][order]
    chunk_deg[rec] = deg[order].to(torch.int32)
    seg_nchunks = torch.zeros(n_seg, dtype=torch.int64, device=dev)
    seg_nchunks.scatter_reduce_(0, seg_s, c_in_seg + 1, reduce="amax", include_self=True)
    seg_chunk_ptr = torch.zeros(n_seg + 1, dtype=torch.int64, device=dev)
    seg_chunk_ptr[1:] = torch.cumsum(seg_nchunks, 0)
    # ---- (vertex, incident slot) entries -> visits
    n_ent = int(col_ind.shape[0])
    node_of_entry = torch.repeat_interleave(torch.arange(V, device=dev), val)
    elem = torch.div(col_ind, N, rounding_mode="floor")
    n_a = col_ind - elem * N
    e_seg, e_cin, e_slot = seg_of[node_of_entry], cin_of[node_of_entry], slot_of[node_of_entry]
    max_steps = int(seg_nchunks.max().item())
    step = torch.full((n_ent,), -1, dtype=torch.int64, device=dev)
    todo = torch.arange(n_ent, device=dev)
    while todo.numel() > 0:                                  # one pass on structured grids; more when an element's nodes
        gkey = e_seg[todo] * E + elem[todo]                  # spread over non-adjacent chunks of a segment
        ug, inv = torch.unique(gkey, return_inverse=True)
        cmin = torch.full((ug.shape[0],), max_steps + 2, dtype=torch.int64, device=dev)
        cmin.scatter_reduce_(0, inv, e_cin[todo], reduce="amin", include_self=True)
        first = cmin[inv]
        take = e_cin[todo] <= first + 1
        step[todo[take]] = first[take]
        todo = todo[~take]
    vkey = (e_seg * (max_steps + 1) + step) * E + elem
    uv, vinv = torch.unique(vkey, return_inverse=True)       # sorted: segment, step, element
    n_vis = int(uv.shape[0])
    if n_vis >= (1 << 31) - 16:
        return None
    v_elem = uv % E
    v_ss = uv // E
    v_seg = v_ss // (max_steps + 1)
    v_step = v_ss - v_seg * (max_steps + 1)
    absent = torch.zeros(n_vis, dtype=torch.int64, device=dev)
    absent.scatter_add_(0, vinv, (15 - e_slot) << (4 * n_a))
    slotw = 0xFFFFFFFF - absent                              # nibble n = slot of local node n, 0xF when not in the window
    mask = torch.zeros(n_vis, dtype=torch.int64, device=dev)
    mask.scatter_add_(0, vinv, torch.ones_like(n_a) << n_a)
    lut_perm, lut_cnt = _perm_lut(dev)
    perm, cnt = lut_perm[mask], lut_cnt[mask]
    # ---- order within a step: visits that touch disjoint panel slots next to each other, so that the kernel can run them
    #      as one interleaved instruction stream (two independent DMMA / accumulator chains).  A pair occupies positions
    #      (2m, 2m+1) of its segment (groups of four never split a pair); visits with more than four in-window nodes run alone
    #      and go to the front of steps that start at an odd position, to the back otherwise.
    smask = torch.zeros(n_vis, dtype=torch.int64, device=dev)
    smask.scatter_add_(0, vinv, torch.ones_like(e_slot) << e_slot)
    seg_first = torch.searchsorted(v_seg.contiguous(), torch.arange(n_seg, device=dev))
    pos = torch.arange(n_vis, device=dev) - seg_first[v_seg]
    new_step = torch.ones(n_vis, dtype=torch.bool, device=dev)
    new_step[1:] = (v_seg[1:] != v_seg[:-1]) | (v_step[1:] != v_step[:-1])
    gid = torch.cumsum(new_step.to(torch.int64), 0) - 1
    start_odd = (pos[new_step] & 1)[gid]
    popc = lut_cnt[smask]
    alone = cnt > 4
    cls = torch.where(alone, torch.where(start_odd == 1, torch.zeros_like(cnt), torch.full_like(cnt, 2)), torch.ones_like(cnt))
    comp = torch.minimum(smask, 0xFF ^ smask)
    okey = (gid << 24) | (cls << 20) | (popc << 16) | (comp << 8) | smask
    order_v = torch.argsort(okey, stable=True)
    v_elem, v_seg, v_step, slotw, perm, cnt, smask = (t[order_v] for t in (v_elem, v_seg, v_step, slotw, perm, cnt, smask))
    pair_first = torch.zeros(n_vis, dtype=torch.bool, device=dev)
    if n_vis > 1:
        pair_first[:-1] = ((pos[:-1] & 1) == 0) & (v_seg[1:] == v_seg[:-1]) & (v_step[1:] == v_step[:-1]) & \
                          ((smask[1:] & smask[:-1]) == 0) & (cnt[1:] <= 4) & (cnt[:-1] <= 4)
    # chunks completed by a visit: the last visit of step c retires chunks c .. (next step with visits) - 1
    last = torch.ones(n_vis, dtype=torch.bool, device=dev)
    last[:-1] = (v_seg[1:] != v_seg[:-1]) | (v_step[1:] != v_step[:-1])
    nxt_step = torch.empty(n_vis, dtype=torch.int64, device=dev)
    nxt_step[:-1] = torch.where(v_seg[1:] == v_seg[:-1], v_step[1:], seg_nchunks[v_seg[:-1]])
    nxt_step[-1] = seg_nchunks[v_seg[-1]]
    nret = torch.where(last, nxt_step - v_step, torch.zeros_like(v_step))
    first_of_seg = torch.ones(n_vis, dtype=torch.bool, device=dev)
    first_of_seg[1:] = v_seg[1:] != v_seg[:-1]
    if int(nret.max().item()) > 15 or bool((v_step[first_of_seg] != 0).any().item()):
        return None
    flags = cnt | (nret << 4) | (pair_first.to(torch.int64) << 8)
    head = torch.stack([v_elem, slotw, perm, flags], dim=1)
    head = torch.where(head >= (1 << 31), head - (1 << 32), head).to(torch.int32)
    visits = torch.cat([head, conn32[v_elem].to(torch.int32)], dim=1).contiguous()
    seg_visit_ptr = torch.searchsorted(v_seg.contiguous(), torch.arange(n_seg + 1, device=dev)).to(torch.int32)
    if bool((seg_visit_ptr[1:] == seg_visit_ptr[:-1]).any().item()):
        return None                                          # every segment has visits (its vertices have elements)
    chunk_rec = torch.cat([chunk_v.reshape(-1, 4), chunk_deg.reshape(-1, 4),
                           chunk_p0.reshape(-1, 4).contiguous().view(torch.int32).reshape(-1, 8)], dim=1).contiguous()
    keep = [seg_visit_ptr.contiguous(), seg_chunk_ptr.to(torch.int32).contiguous(), visits, chunk_rec]
    stats = dict(n_segments=n_seg, n_chunks=n_chunks, n_visits=n_vis, n_strips=n_strips, sweep_axis=s_ax,
                 visits_per_element=n_vis / E, max_deg=max_deg, paired_fraction=2.0 * float(pair_first.sum().item()) / n_vis,
                 row_tiles_per_vertex=float(((cnt + 1) // 2).sum().item()) / max(n_act, 1))   # two nodes per row tile
    struct = None
    if vertices.is_cuda:
        if int(_lib.load().feb200_fused_sweep_warps(max_deg)) < 1:
            return None
        struct = _lib.SweepPlan(n_segments=n_seg, max_deg=max_deg, seg_visit_ptr=_lib.ptr(keep[0]),
                                seg_chunk_ptr=_lib.ptr(keep[1]), visits=_lib.ptr(keep[2]), chunk_rec=_lib.ptr(keep[3]))
    return SweepPlan(struct, keep, stats)
