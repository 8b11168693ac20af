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
