"""Randomised ragged hex8 meshes through the fused assembly (tensor-core kernel for every material layout it takes):
grids of 1..8 cells per axis with jittered vertices, 30-100 % of the elements kept in shuffled order (vertices with
1..8 incident elements, some with none), random Dirichlet sets; values against the oracle at 1e-12 of the largest entry
and bit-reproducible across launches."""
import numpy as np
import pytest
import torch

from helpers import build_pair, material_params
from oracle import fe_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fj():
    import __graft_entry__ as graft
    graft.build()
    import fe_jax
    return fe_jax


def test_fused_assembly_on_random_ragged_meshes(fj):
    from fe_jax.assembly import _assemble_values, _plan
    rng = np.random.default_rng(12345)
    n_fused = 0
    for case in range(24):
        dims = tuple(int(v) for v in rng.integers(1, 9, size=3))
        Vx, cells = orc.uniform_hex_grid(*dims)
        Vx = Vx + 0.12 / max(dims) * rng.uniform(-1, 1, size=Vx.shape)
        sel = rng.uniform(size=cells.shape[0]) < rng.uniform(0.3, 1.0)
        if sel.sum() == 0:
            sel[0] = True
        cells = cells[sel][rng.permutation(int(sel.sum()))]
        E = cells.shape[0]
        layout = ["M", "EM", "EQM", "EQM_uniform_nu"][case % 4]
        mp = material_params("elastic_isotropic", 3, {"M": (), "EM": (E,)}.get(layout, (E, 8)), seed=case)
        if layout == "EQM":
            mp[..., 1] = mp[:, :1, 1]
        if layout == "EQM_uniform_nu":
            mp[..., 1] = 0.3
        ebc, maps, token, prob, _ = build_pair(fj, Vx, [dict(cell="hexahedron", degree=1, qdeg=2, cells=cells,
                                                            model="elastic_isotropic", mp=mp)])
        u = torch.zeros(Vx.size, dtype=torch.float64, device="cuda")
        mask = torch.zeros(prob.F, dtype=torch.float64, device="cuda")
        dofs = np.unique(rng.integers(0, prob.F, size=max(1, prob.F // 6)))
        mask[torch.tensor(dofs, device="cuda")] = 1.0
        n_fused += _plan(ebc, maps).fused_plan(ebc, 0) is not None
        v, _ = _assemble_values(token, ebc, maps, u, mask, torch.zeros_like(mask), eliminate=True)
        v2, _ = _assemble_values(token, ebc, maps, u, mask, torch.zeros_like(mask), eliminate=True)
        assert torch.equal(v, v2), (case, "not reproducible")
        d_o, r_o, c_o = prob.jacobian_wo_dirichlet(np.zeros(Vx.size))
        dw = orc.apply_dirichlet_bcs_lhs(d_o, r_o, c_o, dofs)
        err = float(np.abs(v.cpu().numpy() - dw).max() / np.abs(dw).max())
        assert err <= 1e-12, (case, dims, layout, err)
    assert n_fused == 24          # every case is eligible for the tile plan
