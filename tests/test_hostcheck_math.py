"""The element math the CUDA kernels are built from (csrc/element_math.cuh), compiled for the host
with g++ (tests/hostcheck) and compared with the oracle.  CPU only; the kernels themselves are
covered by the -m gpu tests."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import fe_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
dp = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def hc(tmp_path_factory):
    out = tmp_path_factory.mktemp("hostcheck") / "libhostcheck.so"
    src = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", str(out)])
    lib = ctypes.CDLL(str(out))
    for f in (lib.hc_residual, lib.hc_tangent):
        f.argtypes = [ctypes.c_int] * 5 + [dp] * 5 + [ctypes.c_int, dp]
        f.restype = ctypes.c_int
    return lib


def _p(a):
    return a.ctypes.data_as(dp)


CASES = [("triangle", 1, 3, 2), ("quadrilateral", 1, 2, 2), ("hexahedron", 1, 2, 3), ("tetrahedron", 1, 2, 3),
         ("triangle", 2, 4, 2), ("tetrahedron", 2, 4, 3), ("hexahedron", 2, 4, 3)]
MATS = [("linear", "elastic_isotropic", 0, 0), ("linear", "elastic_orthotropic", 0, 1),
        ("hyper", "st_venant_kirchhoff", 1, 2), ("hyper", "mooney_rivlin", 1, 3)]


def _params(model, D, Q, rng):
    if model in ("elastic_isotropic", "st_venant_kirchhoff"):
        return np.stack([10e9 + 90e9 * rng.uniform(size=Q), 0.2 + 0.2 * rng.uniform(size=Q)], axis=1)
    if model == "mooney_rivlin":
        return np.stack([0.5e6 + 1e5 * rng.uniform(size=Q), 0.1e6 + 1e4 * rng.uniform(size=Q),
                         2e-7 + 1e-8 * rng.uniform(size=Q)], axis=1)
    if D == 2:
        return np.stack([40e9 + 5e9 * rng.uniform(size=Q), 10e9 + 1e9 * rng.uniform(size=Q),
                         0.25 + 0.05 * rng.uniform(size=Q), 4e9 + 1e9 * rng.uniform(size=Q)], axis=1)
    return np.stack([40e9 + 5e9 * rng.uniform(size=Q), 10e9 + 1e9 * rng.uniform(size=Q),
                     12e9 + 1e9 * rng.uniform(size=Q), 0.25 + 0.05 * rng.uniform(size=Q),
                     0.30 + 0.05 * rng.uniform(size=Q), 0.20 + 0.05 * rng.uniform(size=Q),
                     4e9 + 1e9 * rng.uniform(size=Q), 3e9 + 1e9 * rng.uniform(size=Q),
                     5e9 + 1e9 * rng.uniform(size=Q)], axis=1)


@pytest.mark.parametrize("cell,deg,qdeg,D", CASES)
@pytest.mark.parametrize("kind,model,kcode,mcode", MATS)
def test_element_math_matches_oracle(hc, cell, deg, qdeg, D, kind, model, kcode, mcode):
    rng = np.random.default_rng(7)
    xi, W = orc.get_quadrature(cell, qdeg)
    _, dphi = orc.eval_basis_and_derivatives(cell, deg, xi)
    nodes = orc.reference_nodes(cell, deg)
    N, Q = nodes.shape[0], xi.shape[0]
    x = np.ascontiguousarray(nodes * 1.2 + 0.03 * rng.normal(size=nodes.shape))
    u = np.ascontiguousarray((1e-3 if kind == "linear" else 3e-2) * rng.normal(size=nodes.shape))
    mp = np.ascontiguousarray(_params(model, D, Q, rng))
    dphi = np.ascontiguousarray(dphi)
    W = np.ascontiguousarray(W)
    R = np.zeros(N * D)
    K = np.zeros((N * D, N * D))
    assert hc.hc_residual(D, kcode, mcode, N, Q, _p(x), _p(u), _p(dphi), _p(W), _p(mp), mp.shape[1], _p(R)) == 0
    assert hc.hc_tangent(D, kcode, mcode, N, Q, _p(x), _p(u), _p(dphi), _p(W), _p(mp), mp.shape[1], _p(K)) == 0
    R_o = orc.element_residuals(kind, model, u[None], x[None], dphi, W, mp[None])[0].ravel()
    K_o = orc.element_tangents(kind, model, u[None], x[None], dphi, W, mp[None])[0]
    assert np.abs(R - R_o).max() <= 1e-12 * np.abs(R_o).max()
    assert np.abs(K - K_o).max() <= 1e-11 * np.abs(K_o).max()
