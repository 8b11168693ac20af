"""CPU-side checks of the product: the C-ABI library builds, loads and exports every symbol that
include/feb200.h declares (no compute calls -- there is no GPU here), the host tables agree with the
oracle's, and the host logic refuses to run without CUDA instead of falling back."""
import os
import re

import numpy as np
import pytest
import torch

import __graft_entry__ as graft
from oracle import fe_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    graft.build()
    from fe_jax import _lib
    return _lib


def test_library_exports_every_declared_symbol(built):
    header = open(os.path.join(ROOT, "include", "feb200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(feb200_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    lib = built.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in feb200.h but not exported"
    assert declared == set(built.SIGNATURES), (declared ^ set(built.SIGNATURES))
    assert lib.feb200_version() >= 100
    assert lib.feb200_reduce_workspace_bytes() > 0


def test_bad_arguments_are_reported_not_crashed(built):
    lib = built.load()
    rc = lib.feb200_gather(None, None, 4, 3, 2, None, None)       # null pointers -> status + message, no launch
    assert rc == -1 and b"gather" in lib.feb200_last_error()
    rc = lib.feb200_spmv(None, 10, 3, None, None, None, None, None, None, None, 0)
    assert rc == -1


def test_sass_is_sm100a(built):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", os.path.abspath(built.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.parametrize("cell,deg,qdeg", [("triangle", 1, 3), ("triangle", 2, 4), ("quadrilateral", 1, 2),
                                           ("quadrilateral", 2, 4), ("tetrahedron", 1, 2), ("tetrahedron", 2, 4),
                                           ("hexahedron", 1, 2), ("hexahedron", 2, 4), ("triangle", 1, 2)])
def test_host_tables_match_oracle(cell, deg, qdeg):
    import fe_jax as fj
    fe = fj.FiniteElementType(getattr(fj.CellType, cell), fj.ElementFamily.P, deg, fj.LagrangeVariant.equispaced,
                              fj.QuadratureType.default, qdeg)
    xi, W = fj.get_quadrature(fe)
    xi_o, W_o = orc.get_quadrature(cell, qdeg)
    assert np.allclose(xi, xi_o, atol=1e-15) and np.allclose(W, W_o, atol=1e-16)
    phi, dphi = fj.eval_basis_and_derivatives(fe, xi)
    phi_o, dphi_o = orc.eval_basis_and_derivatives(cell, deg, xi_o)
    assert np.allclose(phi, phi_o, atol=1e-14) and np.allclose(dphi, dphi_o, atol=1e-13)


def test_host_grids_match_reference_outputs(golden):
    import fe_jax as fj
    V, F = fj.uniform_quad_grid(5, 7, np.array([[0.0, 2.0], [-1.0, 1.0]]))
    assert np.array_equal(F, golden["quad_grid_F"]) and np.array_equal(V, golden["quad_grid_V"])
    V, F = fj.uniform_tri_grid(7, 10)
    assert np.array_equal(F, golden["tri_grid_F"]) and np.array_equal(V, golden["tri_grid_V"])
    V, F = fj.uniform_hex_grid(3, 4, 2)
    Vo, Fo = orc.uniform_hex_grid(3, 4, 2)
    assert np.array_equal(F, Fo) and np.array_equal(V, Vo)


def test_solver_options_surface():
    import fe_jax as fj
    o = fj.SolverOptions()
    assert (o.linear_max_iter, o.linear_relative_tol, o.linear_absolute_tol) == (1000, 1e-14, 1e-10)
    assert (o.nonlinear_max_iter, o.nonlinear_relative_tol, o.nonlinear_absolute_tol) == (10, 1e-10, 1e-8)
    assert o.linear_solve_type == fj.LinearSolverType.CG_JAX_SCIPY_W_INFO
    with pytest.raises(AssertionError):
        fj.SolverOptions(linear_solve_type=fj.LinearSolverType.AMGX)
    info = fj.init_solver_info(o, device="cpu")
    assert info.cumulative_residual_norm_history.shape[0] == 1000 * 10 + 1
    assert info.increment_nl_iteration().nonlinear_iterations == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    import fe_jax as fj
    from fe_jax._lib import Feb200Error
    V, F = fj.uniform_tri_grid(3, 3)
    with pytest.raises(Feb200Error):
        fj.build_row_ind(V.shape[0], F)
    with pytest.raises(Feb200Error):
        fj.mesh_to_sparse_assembly_map(V.shape[0], F)


def test_unknown_callables_are_rejected():
    import fe_jax as fj
    from fe_jax._lib import Feb200Error
    from fe_jax.physics import resolve_model, resolve_residual
    import functools
    assert resolve_model(functools.partial(fj.elastic_isotropic)) is fj.elastic_isotropic
    assert resolve_residual(fj.hyperelasticity_residual).kind == 1
    with pytest.raises(Feb200Error):
        resolve_model(lambda eps_qdd, material_params_qm: eps_qdd)
    with pytest.raises(Feb200Error):
        resolve_residual(print)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "fea-in-jax_b200", "fe_jax")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "oracle" not in src.replace("# oracle", ""), f"{f} mentions the oracle"


def test_transposed_warp_reduction_adds_the_same_pairs_as_the_butterfly():
    """csrc/krylov.cu warp_sum_2x3: the six row sums of a 2-vertex chunk go through one transposed reduction (8 shuffles).
    Emulated lane by lane here: every total must be bit-identical to the xor-butterfly warp_sum of the same value, which
    is what keeps k_spmv3_bulk bit-identical to k_spmv (GPU test test_spmv_bulk_copy_kernel_is_bit_identical...)."""
    import numpy as np
    rng = np.random.default_rng(5)
    lanes = np.arange(32)

    def shfl_xor(v, o):
        return v[lanes ^ o]

    def butterfly(v):
        v = v.copy()
        for o in (16, 8, 4, 2, 1):
            v = v + shfl_xor(v, o)
        return v

    for trial in range(20):
        A = rng.normal(size=(3, 32)) * 10.0 ** rng.integers(-8, 8, size=(3, 32))
        B = rng.normal(size=(3, 32)) * 10.0 ** rng.integers(-8, 8, size=(3, 32))
        if trial % 3 == 0:                       # lanes beyond the degree hold zeros
            A[:, 27:] = 0.0
            B[:, 18:] = 0.0
        up, hi, q = (lanes & 16) != 0, (lanes & 8) != 0, (lanes & 4) != 0
        k = [np.where(up, B[i], A[i]) + shfl_xor(np.where(up, A[i], B[i]), 16) for i in range(3)]
        m0 = np.where(hi, k[2], k[0]) + shfl_xor(np.where(hi, k[0], k[2]), 8)
        m1 = np.where(hi, 0.0, k[1]) + shfl_xor(np.where(hi, k[1], 0.0), 8)
        r = np.where(q, m1, m0) + shfl_xor(np.where(q, m0, m1), 4)
        r = r + shfl_xor(r, 2)
        r = r + shfl_xor(r, 1)
        want = [butterfly(A[0]), butterfly(A[1]), butterfly(A[2]), None, butterfly(B[0]), butterfly(B[1]), butterfly(B[2]), None]
        for s, w in enumerate(want):
            if w is None:
                continue
            got = r[4 * s:4 * s + 4]
            assert np.array_equal(got, np.full(4, w[0])), (trial, s)
            assert np.array_equal(w, np.full(32, w[0]))          # the butterfly leaves the same total on every lane


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): one JSON line with the metric, unit,
    config and the cpu_baseline / e2e objects of the task contract, on a small bounded sample."""
    import json, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-n", "6"], check=True, capture_output=True, text=True, timeout=600, cwd=root).stdout
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "elements/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["dtype"] == "f64"
    assert "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
