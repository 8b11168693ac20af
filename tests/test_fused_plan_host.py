"""The tile plan of the fused hex8 kernel, checked on the CPU: a NumPy emulation of what the kernel does
with the plan (per tile: block rows of every (vertex, slot) item, contributions added per CSR block in
ascending slot order) must reproduce the oracle's Jacobian.  The CUDA kernel itself is covered by
tests/test_gpu_parity.py::test_fused_hex8_assembly_matches_two_pass_and_oracle."""
import numpy as np
import pytest
import torch

import __graft_entry__ as graft
from oracle import fe_oracle as orc


@pytest.mark.parametrize("dims,shuffle", [((4, 5, 3), False), ((6, 3, 7), True), ((9, 9, 2), True)])
def test_tile_plan_reproduces_oracle_jacobian(dims, shuffle):
    graft.build()
    from fe_jax import fused_plan
    Vx, cells = orc.uniform_hex_grid(*dims)
    rng = np.random.default_rng(1)
    Vx = Vx + 0.1 / max(dims) * rng.uniform(-1, 1, size=Vx.shape)
    if shuffle:
        cells = cells[rng.permutation(cells.shape[0])]
    V, E = Vx.shape[0], cells.shape[0]
    ri = orc.build_row_ind(V, cells).astype(np.int64)
    ci = orc.build_col_ind_and_data(ri.astype(np.uint64), cells)[0].astype(np.int64)
    plan = fused_plan.build_fused_plan(torch.from_numpy(Vx), torch.from_numpy(cells.astype(np.int32)),
                                       torch.from_numpy(ri), torch.from_numpy(ci))
    assert plan is not None
    node_ptr, nodes, elem_ptr, elems, item_ptr, items = (t.numpy() for t in plan._keep)
    # every vertex in exactly one tile, tiles within capacity, every (vertex, slot) item exactly once
    assert sorted(nodes.tolist()) == list(range(V))
    assert (np.diff(node_ptr) <= plan.tn_cap).all() and (np.diff(elem_ptr) <= plan.te_cap).all()
    assert items.shape[0] == 8 * E and item_ptr[-1] == 8 * E

    xi, W = orc.get_quadrature("hexahedron", 2)
    _, dphi = orc.eval_basis_and_derivatives("hexahedron", 1, xi)
    mp = np.array([70e9, 0.25])
    prob = orc.Problem(Vx, [orc.Batch(cells, "linear", "elastic_isotropic", mp, dphi, W)])
    K = orc.element_tangents("linear", "elastic_isotropic", np.zeros((E, 8, 3)), Vx[cells], dphi,
                             orc.weights_as_reference(W), orc.expand_material(mp, E, 8))
    d_o, r_o, c_o = prob.jacobian_wo_dirichlet(np.zeros(prob.F))
    dense = np.zeros((prob.F, prob.F))
    for t in range(plan.n_tiles):
        tn = nodes[node_ptr[t]:node_ptr[t + 1]]
        te = elems[elem_ptr[t]:elem_ptr[t + 1]]
        assert np.all(np.diff(te) > 0)
        its = items[item_ptr[t]:item_ptr[t + 1]].astype(np.int64) & 0xFFFFFFFF
        le, na, al, k = its & 0xFFF, (its >> 12) & 0xF, (its >> 16) & 0xFF, (its >> 24) & 0xFF
        assert np.all(np.diff(le * 8 + na) > 0)                       # sorted by (element, local node)
        seen = set()
        for j in np.lexsort((k, al)):                                 # per vertex in ascending slot order
            a = tn[al[j]]
            e = te[le[j]]
            assert cells[e, na[j]] == a                                # the item is a slot of its vertex
            assert ci[ri[a] + k[j]] == e * 8 + na[j]                   # and it is the k-th one
            assert (al[j], k[j]) not in seen
            seen.add((al[j], k[j]))
            for b in range(8):
                cb = cells[e, b]
                dense[3 * a:3 * a + 3, 3 * cb:3 * cb + 3] += K[e, 3 * na[j]:3 * na[j] + 3, 3 * b:3 * b + 3]
        assert len(seen) == sum(ri[a + 1] - ri[a] for a in tn)
    assert np.abs(dense[r_o, c_o] - d_o).max() <= 1e-13 * np.abs(d_o).max()
    dense[r_o, c_o] = 0.0
    assert np.abs(dense).max() == 0.0


def test_morton_tiles_are_compact():
    graft.build()
    from fe_jax import fused_plan
    n = 16
    Vx, cells = orc.uniform_hex_grid(n, n, n)
    ri = orc.build_row_ind(Vx.shape[0], cells).astype(np.int64)
    ci = orc.build_col_ind_and_data(ri.astype(np.uint64), cells)[0].astype(np.int64)
    plan = fused_plan.build_fused_plan(torch.from_numpy(Vx), torch.from_numpy(cells.astype(np.int32)),
                                       torch.from_numpy(ri), torch.from_numpy(ci))
    visits = plan._keep[3].shape[0] / cells.shape[0]
    assert plan.tn_cap == 32 and visits < 3.2, visits        # 4x4x2 bricks: (5*5*3)/(4*4*2) = 2.34 in the interior


@pytest.mark.parametrize("dims,layout", [((4, 5, 3), "M"), ((6, 3, 7), "EM"), ((5, 4, 4), "EQM")])
def test_tensor_core_kernel_data_flow_reproduces_oracle_jacobian(dims, layout):
    """NumPy emulation of csrc/fused_mma.cu on the tile plan: per element one 24 x 24 product S = (w G) G^T whose rows
    are taken in the kernel's row-tile order (row R of tile mt = slot R // 3 of the element's in-tile vertices,
    component R % 3), staged per item as [i][b*3+j], then per CSR block the staged 3x3 blocks added in ascending
    slot order and combined as K = rho S' + S'^T + tr(S') I (mu, or w*mu_q per point, folded into the weights).
    Checks the index logic of the kernel (item ranges per element, row tiles, source table) without a GPU."""
    graft.build()
    from fe_jax import fused_plan
    Vx, cells = orc.uniform_hex_grid(*dims)
    rng = np.random.default_rng(4)
    Vx = Vx + 0.1 / max(dims) * rng.uniform(-1, 1, size=Vx.shape)
    cells = cells[rng.permutation(cells.shape[0])]
    V, E = Vx.shape[0], cells.shape[0]
    ri = orc.build_row_ind(V, cells).astype(np.int64)
    ci = orc.build_col_ind_and_data(ri.astype(np.uint64), cells)[0].astype(np.int64)
    plan = fused_plan.build_fused_plan(torch.from_numpy(Vx), torch.from_numpy(cells.astype(np.int32)),
                                       torch.from_numpy(ri), torch.from_numpy(ci))
    node_ptr, nodes, elem_ptr, elems, item_ptr, items = (t.numpy() for t in plan._keep)
    xi, W = orc.get_quadrature("hexahedron", 2)
    _, dphi = orc.eval_basis_and_derivatives("hexahedron", 1, xi)
    Wr = orc.weights_as_reference(W)
    if layout == "M":
        mp = np.array([70e9, 0.25])
    elif layout == "EM":
        mp = np.stack([10e9 + 90e9 * rng.uniform(size=E), 0.2 + 0.2 * rng.uniform(size=E)], axis=-1)
    else:
        mp = np.stack([10e9 + 90e9 * rng.uniform(size=(E, 8)), np.repeat(0.2 + 0.2 * rng.uniform(size=(E, 1)), 8, axis=1)],
                      axis=-1)
    mpe = orc.expand_material(mp, E, 8)
    Em, nu = mpe[..., 0], mpe[..., 1]
    mu_q = 0.5 * Em / (1.0 + nu)                                     # (E, Q)
    rho_e = (2.0 * nu / (1.0 - 2.0 * nu))[:, 0]                      # lambda / mu, constant within an element
    g, detJ = orc._geometry(Vx[cells], dphi)                         # g (E,Q,N,D)
    # neighbour lists (sorted) and rank of local node b in the row of vertex conn[e, a]
    nbrs = [set() for _ in range(V)]
    for c in cells:
        for a in c:
            nbrs[a].update(c.tolist())
    lists = [sorted(s) for s in nbrs]
    prob = orc.Problem(Vx, [orc.Batch(cells, "linear", "elastic_isotropic", mp, dphi, W)])
    d_o, r_o, c_o = prob.jacobian_wo_dirichlet(np.zeros(prob.F))
    dense = np.zeros((prob.F, prob.F))
    for t in range(plan.n_tiles):
        tn = nodes[node_ptr[t]:node_ptr[t + 1]]
        te = elems[elem_ptr[t]:elem_ptr[t + 1]]
        its = items[item_ptr[t]:item_ptr[t + 1]].astype(np.int64) & 0xFFFFFFFF
        le, na, al, k = its & 0xFFF, (its >> 12) & 0xF, (its >> 16) & 0xFF, (its >> 24) & 0xFF
        NI = its.shape[0]
        # P0: item range of every tile element from the first / last item with that element id
        e0 = {int(le[i]): i for i in range(NI) if i == 0 or le[i - 1] != le[i]}
        e1 = {int(le[i]): i + 1 for i in range(NI) if i == NI - 1 or le[i + 1] != le[i]}
        item_of = {(int(al[i]), int(k[i])): i for i in range(NI)}
        stage = np.full((NI, 3, 24), np.nan)
        # phase A
        for l, e in enumerate(te):
            Gm = g[e].transpose(1, 2, 0).reshape(24, 8)              # rows (n, d), columns q
            w = detJ[e] * Wr * mu_q[e]                               # the material enters through the weights
            S = (Gm * w) @ Gm.T                                      # 24 x 24
            item0, cnt = e0[l], e1[l] - e0[l]
            assert 1 <= cnt <= 8
            nmt = (3 * cnt + 7) // 8
            for mt in range(nmt):
                for gid in range(8):
                    R = 8 * mt + gid
                    s, i = R // 3, R % 3
                    if s < cnt:
                        n_a = na[item0 + s]
                        assert le[item0 + s] == l
                        stage[item0 + s, i, :] = S[3 * n_a + i, :]
        assert not np.isnan(stage).any()                             # every item row was written by exactly its element
        # phase B
        for a_loc, a in enumerate(tn):
            val = ri[a + 1] - ri[a]
            for r, nb in enumerate(lists[a]):
                acc, accl = np.zeros((3, 3)), np.zeros((3, 3))
                for kk in range(val):                                # ascending slot order
                    it = item_of[(a_loc, kk)]
                    e = te[le[it]]
                    hit = np.nonzero(cells[e] == nb)[0]              # source table: local node b with rank r, if any
                    if hit.size:
                        blk = stage[it][:, 3 * hit[0]:3 * hit[0] + 3]
                        acc += blk
                        accl += rho_e[e] * blk
                K = accl + acc.T + np.trace(acc) * np.eye(3)
                dense[3 * a:3 * a + 3, 3 * nb:3 * nb + 3] = K
    assert np.abs(dense[r_o, c_o] - d_o).max() <= 1e-12 * np.abs(d_o).max()
    dense[r_o, c_o] = 0.0
    assert np.abs(dense).max() == 0.0
