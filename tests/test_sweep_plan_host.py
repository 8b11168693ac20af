"""The streaming plan of feb200_assemble_fused_hex8_sweep (fe_jax/sweep_plan.py), checked on the CPU: a NumPy emulation
of the kernel's data flow on the plan -- per segment a window of 8 panel slots, per visit the row tiles of the
in-window nodes accumulated into panel[slot][rank], chunks stored and cleared when the plan retires them -- must
reproduce the oracle's Jacobian (reference path fea.py:519-572 + :613-660).  The CUDA kernel itself is covered by
tests/test_gpu_parity.py."""
import numpy as np
import pytest
import torch

import __graft_entry__ as graft
from oracle import fe_oracle as orc


def _mesh(dims, shuffle, seed=1):
    Vx, cells = orc.uniform_hex_grid(*dims)
    rng = np.random.default_rng(seed)
    Vx = Vx + 0.1 / max(dims) * rng.uniform(-1, 1, size=Vx.shape)
    if shuffle:
        cells = cells[rng.permutation(cells.shape[0])]
    return Vx, cells, rng


def _pattern(V, cells):
    nbrs = [set() for _ in range(V)]
    for c in cells:
        for a in c:
            nbrs[a].update(c.tolist())
    lists = [sorted(s) for s in nbrs]
    nindptr = np.zeros(V + 1, dtype=np.int64)
    nindptr[1:] = np.cumsum([len(l) for l in lists])
    return lists, nindptr


def _build(Vx, cells, seg_chunks):
    graft.build()
    from fe_jax import sweep_plan
    V = Vx.shape[0]
    ri = orc.build_row_ind(V, cells).astype(np.int64)
    ci = orc.build_col_ind_and_data(ri.astype(np.uint64), cells)[0].astype(np.int64)
    lists, nindptr = _pattern(V, cells)
    plan = sweep_plan.build_sweep_plan(torch.from_numpy(Vx), torch.from_numpy(cells.astype(np.int32)), torch.from_numpy(ri),
                                       torch.from_numpy(ci), torch.from_numpy(nindptr), seg_chunks=seg_chunks)
    return plan, lists, nindptr, ri


@pytest.mark.parametrize("dims,shuffle,seg_chunks,layout", [((4, 5, 3), False, 48, "M"), ((6, 3, 7), True, 3, "EM"),
                                                            ((5, 4, 9), True, 2, "EQM"), ((2, 2, 2), False, 48, "M"),
                                                            ((1, 1, 1), False, 48, "M")])
def test_sweep_plan_data_flow_reproduces_oracle_jacobian(dims, shuffle, seg_chunks, layout):
    Vx, cells, rng = _mesh(dims, shuffle)
    V, E = Vx.shape[0], cells.shape[0]
    plan, lists, nindptr, ri = _build(Vx, cells, seg_chunks)
    assert plan is not None
    seg_vptr, seg_cptr, visits = (t.numpy() for t in plan._keep[:3])
    chunk_v, chunk_p0, chunk_deg = plan.chunk_v.numpy(), plan.chunk_p0.numpy(), plan.chunk_deg.numpy()
    assert visits.shape[1] == 12 and np.array_equal(visits[:, 4:], cells[visits[:, 0]])   # the records carry the vertex ids
    max_deg = plan.stats["max_deg"]
    # every vertex in exactly one chunk slot, with its pattern offset and degree
    live = chunk_v[chunk_v >= 0]
    assert sorted(live.tolist()) == list(range(V))
    assert np.array_equal(chunk_p0[chunk_v >= 0], nindptr[:-1][live])
    assert np.array_equal(chunk_deg[chunk_v >= 0], np.diff(nindptr)[live])

    xi, W = orc.get_quadrature("hexahedron", 2)
    _, dphi = orc.eval_basis_and_derivatives("hexahedron", 1, xi)
    Wr = orc.weights_as_reference(W)
    nu0 = 0.3
    if layout == "M":
        mp = np.array([70e9, nu0])
    elif layout == "EM":
        mp = np.stack([10e9 + 90e9 * rng.uniform(size=E), np.full(E, nu0)], axis=-1)
    else:
        mp = np.stack([10e9 + 90e9 * rng.uniform(size=(E, 8)), np.full((E, 8), nu0)], axis=-1)
    mpe = orc.expand_material(mp, E, 8)
    mu_q = 0.5 * mpe[..., 0] / (1.0 + nu0)
    rho = 2.0 * nu0 / (1.0 - 2.0 * nu0)
    g, detJ = orc._geometry(Vx[cells], dphi)                         # g (E,Q,N,D)
    prob = orc.Problem(Vx, [orc.Batch(cells, "linear", "elastic_isotropic", mp, dphi, W)])
    d_o, r_o, c_o = prob.jacobian_wo_dirichlet(np.zeros(prob.F))
    dense = np.zeros((prob.F, prob.F))
    covered = np.zeros((E, 8), dtype=np.int64)
    stored = np.zeros(V, dtype=np.int64)
    for seg in range(seg_vptr.shape[0] - 1):
        panel = np.zeros((8, max_deg, 3, 3))
        c0, c1 = seg_cptr[seg], seg_cptr[seg + 1]
        c_next = c0
        for vi in range(seg_vptr[seg], seg_vptr[seg + 1]):
            e, slotw, perm, flags = (int(x) & 0xFFFFFFFF for x in visits[vi, :4])
            cnt, nret = flags & 0xF, (flags >> 4) & 0xF
            if flags & 0x100:                                        # first of a pair the kernel runs as one instruction stream:
                e2, slotw2, _, flags2 = (int(x) & 0xFFFFFFFF for x in visits[vi + 1, :4])
                s1 = {(slotw >> (4 * n)) & 0xF for n in range(8)} - {0xF}
                s2 = {(slotw2 >> (4 * n)) & 0xF for n in range(8)} - {0xF}
                assert (vi - seg_vptr[seg]) % 2 == 0 and vi + 1 < seg_vptr[seg + 1]      # positions (2m, 2m+1) of the segment
                assert not (s1 & s2) and nret == 0 and cnt <= 4 and (flags2 & 0xF) <= 4 and not (flags2 & 0x100)
            Gm = g[e].transpose(1, 2, 0).reshape(24, 8)              # rows (n, d), columns q
            S = (Gm * (detJ[e] * Wr * mu_q[e])) @ Gm.T               # the material enters through the weights
            in_window = [n for n in range(8) if (slotw >> (4 * n)) & 0xF != 0xF]
            assert cnt == len(in_window) >= 1 and [(perm >> (4 * i)) & 7 for i in range(cnt)] == in_window
            step = c_next - c0
            for mt in range((cnt + 1) // 2):                         # the kernel's row tiles: two nodes each,
                for gid in range(8):                                 # rows 0-2 and 4-6 (3 and 7 stay empty)
                    s, d = 2 * mt + gid // 4, gid % 4
                    if d == 3 or s >= cnt:
                        continue
                    na = (perm >> (4 * s)) & 7
                    slot = (slotw >> (4 * na)) & 0xF
                    a = cells[e, na]
                    # the slot belongs to the vertex: chunk step or step + 1 of the segment, position slot & 3
                    ch = c_next + (((slot >> 2) - step) & 1)
                    assert ch < c1 and chunk_v[ch, slot & 3] == a
                    if d == 0:
                        covered[e, na] += 1
                    for b in range(8):
                        r = lists[a].index(cells[e, b])
                        panel[slot, r, d, :] += S[3 * na + d, 3 * b:3 * b + 3]
            for _ in range(nret):
                h = (c_next - c0) & 1
                for s in range(4):
                    a = chunk_v[c_next, s]
                    if a < 0:
                        continue
                    for r, nb in enumerate(lists[a]):
                        Sb = panel[4 * h + s, r]
                        dense[3 * a:3 * a + 3, 3 * nb:3 * nb + 3] = rho * Sb + Sb.T + np.trace(Sb) * np.eye(3)
                    stored[a] += 1
                    panel[4 * h + s] = 0.0
                c_next += 1
        assert c_next == c1 and np.abs(panel).max() == 0.0           # every chunk of the segment retired, window left clear
    assert (covered == 1).all() and (stored == 1).all()              # every (vertex, incident element) pair exactly once
    assert np.abs(dense[r_o, c_o] - d_o).max() <= 1e-12 * np.abs(d_o).max()
    dense[r_o, c_o] = 0.0
    assert np.abs(dense).max() == 0.0


def test_sweep_strips_are_columns_on_a_structured_grid():
    n = 12
    Vx, cells, _ = _mesh((n, n, n), False)
    plan, _, _, _ = _build(Vx, cells, 48)
    st = plan.stats
    assert st["sweep_axis"] == 2                                     # vertex ids run fastest along z
    # 2 x 2 columns: 9 elements per layer of 4 vertices in the interior -> ~2.25 visits per element, 3.75 row tiles per vertex
    assert st["visits_per_element"] < 2.8 and st["row_tiles_per_vertex"] < 4.6, st


def test_sweep_plan_rejects_ineligible_meshes():
    graft.build()
    from fe_jax import sweep_plan
    Vx, cells, _ = _mesh((2, 2, 2), False)
    V = Vx.shape[0]
    # nine elements on one vertex: valence > 8
    cells9 = np.concatenate([cells, cells[:1]], axis=0)
    ri = orc.build_row_ind(V, cells9).astype(np.int64)
    ci = orc.build_col_ind_and_data(ri.astype(np.uint64), cells9)[0].astype(np.int64)
    _, nindptr = _pattern(V, cells9)
    assert sweep_plan.build_sweep_plan(torch.from_numpy(Vx), torch.from_numpy(cells9.astype(np.int32)), torch.from_numpy(ri),
                                       torch.from_numpy(ci), torch.from_numpy(nindptr)) is None
