"""Residual functions and constitutive models of the fe_jax call surface, as kernel selectors.

In the reference these are JAX functions that get vmapped / differentiated
(/root/reference/src/fe_jax/linear_elasticity.py:9-206, hyperelasticity.py:9-139).  Here each name is a
token that selects a compiled CUDA variant (libfeb200, csrc/element_math.cuh); arbitrary user callables
cannot run on the CUDA path and are rejected (there is no CPU fallback).

The two residual tokens are also callable with the reference's per-element signature
`f(u_nd, x_nd, dphi_dxi_qnp, W_q, material_params_qm, [internal_state_qi,] constitutive_model)`; the
call runs the batched kernel on a one-element batch.
"""
from __future__ import annotations

import torch

from . import _lib


class ConstitutiveModel:
    def __init__(self, name, material, kind, n_params):
        self.__name__ = name
        self.material, self.kind, self._n_params = material, kind, n_params

    def n_params(self, D):
        return self._n_params[D]

    def __call__(self, *a, **k):
        raise _lib.Feb200Error(f"{self.__name__} is a kernel selector: it is evaluated inside the CUDA element kernels "
                               "(pass it as `constitutive_model`); it cannot be called on its own")

    def __repr__(self):
        return f"<feb200 constitutive model {self.__name__}>"


elastic_isotropic = ConstitutiveModel("elastic_isotropic", _lib.MAT_ISOTROPIC, _lib.KIND_LINEAR, {2: 2, 3: 2})
elastic_orthotropic = ConstitutiveModel("elastic_orthotropic", _lib.MAT_ORTHOTROPIC, _lib.KIND_LINEAR, {2: 4, 3: 9})
st_venant_kirchhoff = ConstitutiveModel("st_venant_kirchhoff", _lib.MAT_SVK, _lib.KIND_HYPER, {2: 2, 3: 2})
mooney_rivlin = ConstitutiveModel("mooney_rivlin", _lib.MAT_MOONEY_RIVLIN, _lib.KIND_HYPER, {2: 3, 3: 3})


class ElementResidual:
    def __init__(self, name, kind, has_state_arg):
        self.__name__ = name
        self.kind, self.has_state_arg = kind, has_state_arg

    def __repr__(self):
        return f"<feb200 element residual {self.__name__}>"

    def __call__(self, u_nd, x_nd, dphi_dxi_qnp, W_q, material_params_qm, *rest, **kw):
        from .mesh import as_device
        rest = list(rest)
        model = kw.get("constitutive_model", rest[-1] if rest else None)
        model = resolve_model(model)
        u = as_device(u_nd, torch.float64)
        x = as_device(x_nd, torch.float64)
        N, D = x.shape
        dphi = as_device(dphi_dxi_qnp, torch.float64)
        Q = dphi.shape[0]
        W = as_device(W_q, torch.float64).reshape(-1)
        if W.numel() == 1:
            W = W.expand(Q).contiguous()     # einsum("q,q->q") broadcast of the reference
        mp = as_device(material_params_qm, torch.float64).reshape(Q, -1).contiguous()
        conn = torch.arange(N, dtype=torch.int32, device=x.device).reshape(1, N)
        b = _lib.Batch(E=1, N=N, Q=Q, D=D, kind=self.kind, material=model.material, layout=_lib.LAYOUT_EQM,
                       M=mp.shape[1], conn=_lib.ptr(conn), dphi=_lib.ptr(dphi), W=_lib.ptr(W), mp=_lib.ptr(mp))
        R = torch.empty((N, D), dtype=torch.float64, device=x.device)
        _lib.call("feb200_element_residual", b, _lib.ptr(x), _lib.ptr(u.reshape(-1)), None, None, _lib.ptr(R))
        if self.has_state_arg:
            return R, torch.zeros((Q, 0), dtype=torch.float64, device=x.device)
        return R


linear_elasticity_residual = ElementResidual("linear_elasticity_residual", _lib.KIND_LINEAR, True)
hyperelasticity_residual = ElementResidual("hyperelasticity_residual", _lib.KIND_HYPER, False)

_MODELS = {m.__name__: m for m in (elastic_isotropic, elastic_orthotropic, st_venant_kirchhoff, mooney_rivlin)}
_RESIDUALS = {r.__name__: r for r in (linear_elasticity_residual, hyperelasticity_residual)}


def _unwrap(f):
    seen = 0
    while hasattr(f, "func") and seen < 8:      # functools.partial / jax.tree_util.Partial
        f = f.func
        seen += 1
    return f


def resolve_model(f) -> ConstitutiveModel:
    f = _unwrap(f)
    if isinstance(f, ConstitutiveModel):
        return f
    name = getattr(f, "__name__", None)
    if name in _MODELS:
        return _MODELS[name]
    raise _lib.Feb200Error(f"constitutive model {f!r} has no CUDA kernel variant (supported: {sorted(_MODELS)}); "
                           "arbitrary callables cannot run on the feb200 path and there is no CPU fallback")


def resolve_residual(f) -> ElementResidual:
    f = _unwrap(f)
    if isinstance(f, ElementResidual):
        return f
    name = getattr(f, "__name__", None)
    if name in _RESIDUALS:
        return _RESIDUALS[name]
    raise _lib.Feb200Error(f"element residual {f!r} has no CUDA kernel variant (supported: {sorted(_RESIDUALS)}); "
                           "arbitrary callables cannot run on the feb200 path and there is no CPU fallback")
