"""Conjugate gradients with iteration count and residual-norm history.

`cg(A, b, x0=None, *, tol=1e-5, atol=0.0, maxiter=None, M=None) -> (x, {"iterations",
"residual_norm_history"})` as /root/reference/src/fe_jax/solve_cg.py:126-137 (recurrence :20-63).

Fast path: A is a matrix assembled on a NodePattern (COO / CSR from calculate_jacobian_wo_dirichlet,
apply_dirichlet_bcs_lhs, coo_to_csr) and M is None or a diagonal given as the vector 1/diag: the whole
loop runs on the device (feb200_cg_init / feb200_cg_iterate); the host only polls the `done` flag every
few iterations.  General A (dense tensor, generic CSR / COO, callable) and general M use the same
kernels for SpMV / dot / axpy with a host-driven loop.
"""
from __future__ import annotations

import torch

from . import _lib
from .containers import COO, CSR, NodePattern
from .mesh import as_device
from .sparse import coo_matvec, csr_matvec

POLL_EVERY = 25       # iterations enqueued between two polls of the device-side `done` flag


class JacobiPreconditioner:
    """M = diag(inv_diag): what `lambda x: x / diag_J(x_0)` is in sparse_linear_solve.py:301-306."""

    def __init__(self, diag):
        self.inv_diag = (1.0 / as_device(diag, torch.float64)).contiguous()

    def __call__(self, v):
        return v * self.inv_diag


def cg_on_pattern(pat: NodePattern, values, b, x0, tol, atol, maxiter, inv_diag=None, poll_every=POLL_EVERY,
                  max_launch_iters=None):
    """Device-resident CG.  Returns (x, iterations:int, history (maxiter,) tensor)."""
    lib = _lib.load()
    dev = values.device
    n = pat.U * pat.V
    x = x0.clone() if x0 is not None else torch.zeros(n, dtype=torch.float64, device=dev)
    r = torch.empty(n, dtype=torch.float64, device=dev)
    p = torch.empty(n, dtype=torch.float64, device=dev)
    Ap = torch.empty(n, dtype=torch.float64, device=dev)
    hist = torch.zeros(max(maxiter, 1), dtype=torch.float64, device=dev)
    scalars = torch.zeros(8, dtype=torch.float64, device=dev)
    counters = torch.zeros(4, dtype=torch.int32, device=dev)
    ws = _lib.workspace(lib.feb200_reduce_workspace_bytes(), dev)
    st = _lib.CGState(V=pat.V, U=pat.U, maxiter=maxiter, nindptr=_lib.ptr(pat.nindptr), nindices=_lib.ptr(pat.nindices),
                      values=_lib.ptr(values), inv_diag=_lib.ptr(inv_diag), b=_lib.ptr(b), x=_lib.ptr(x), r=_lib.ptr(r),
                      p=_lib.ptr(p), Ap=_lib.ptr(Ap), hist=_lib.ptr(hist), scalars=_lib.ptr(scalars),
                      counters=_lib.ptr(counters), ws=_lib.ptr(ws), ws_bytes=ws.numel())
    _lib.call("feb200_cg_init", st, float(tol), float(atol))
    launched = 0
    limit = maxiter if max_launch_iters is None else min(maxiter, max_launch_iters)
    while launched < limit:
        nit = min(poll_every, limit - launched)
        _lib.call("feb200_cg_iterate", st, nit)
        launched += nit
        if int(counters[1].item()) != 0:          # the only host<->device sync of the loop
            break
    k = int(counters[0].item())
    return x, k, hist[:maxiter]


def cg_on_pattern_batched(pat: NodePattern, values, b, x0, tol, atol, maxiter, n_copies, inv_diag=None,
                          poll_every=POLL_EVERY):
    """CG on `n_copies` independent systems laid side by side on one node pattern (block-diagonal matrix; copy c owns the
    DoFs [c*len, (c+1)*len)): every copy has its own alpha / beta / stopping test / iteration count, as vmap(cg) of the
    reference (solve_cg.py:20-137).  Per iteration: one SpMV over all copies + one launch with a CTA per copy.
    Returns (x, iterations (n_copies,) int32 tensor, history (n_copies, maxiter) tensor)."""
    dev = values.device
    n = pat.U * pat.V
    assert n % n_copies == 0, "the copies must have equal size"
    ln = n // n_copies
    x = x0.clone() if x0 is not None else torch.zeros(n, dtype=torch.float64, device=dev)
    r, p, Ap = (torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3))
    hist = torch.zeros(n_copies * max(maxiter, 1), dtype=torch.float64, device=dev)
    scal = torch.zeros(n_copies * 4, dtype=torch.float64, device=dev)
    cnt = torch.zeros(n_copies * 2, dtype=torch.int32, device=dev)

    def spmv(v, out):
        _lib.call("feb200_spmv", pat.V, pat.U, _lib.ptr(pat.nindptr), _lib.ptr(pat.nindices), _lib.ptr(values),
                  _lib.ptr(v), _lib.ptr(out), None, None, 0)

    spmv(x, Ap)
    _lib.call("feb200_bcg_init", n_copies, ln, _lib.ptr(b), _lib.ptr(Ap), _lib.ptr(inv_diag), _lib.ptr(r), _lib.ptr(p),
              _lib.ptr(scal), _lib.ptr(cnt), _lib.ptr(hist), float(tol), float(atol), int(maxiter))
    launched = 0
    done = cnt.view(n_copies, 2)[:, 1]
    while launched < maxiter:
        for _ in range(min(poll_every, maxiter - launched)):
            spmv(p, Ap)
            _lib.call("feb200_bcg_step", n_copies, ln, _lib.ptr(Ap), _lib.ptr(inv_diag), _lib.ptr(x), _lib.ptr(r),
                      _lib.ptr(p), _lib.ptr(scal), _lib.ptr(cnt), _lib.ptr(hist), int(maxiter))
            launched += 1
        if bool(done.all().item()):               # the only host<->device sync of the loop
            break
    return x, cnt.view(n_copies, 2)[:, 0].clone(), hist.view(n_copies, max(maxiter, 1))[:, :maxiter]


def _as_operator(A):
    if callable(A) and not isinstance(A, (COO, CSR, torch.Tensor)):
        return A
    if isinstance(A, COO):
        return lambda v: coo_matvec(A, v)
    if isinstance(A, CSR):
        return lambda v: csr_matvec(A, v)
    dense = as_device(A, torch.float64)
    nz = dense.to_sparse_csr()
    csr = CSR((nz.values().contiguous(), nz.col_indices().to(torch.int64).contiguous(),
               nz.crow_indices().to(torch.int64).contiguous()), shape=tuple(dense.shape))
    return lambda v: csr_matvec(csr, v)


def _dot(a, b, out, ws):
    _lib.call("feb200_dot", a.shape[0], _lib.ptr(a), _lib.ptr(b), _lib.ptr(out), _lib.ptr(ws), ws.numel())
    return float(out.item())


def cg(A, b, x0=None, *, tol=1e-5, atol=0.0, maxiter=None, M=None):
    """solve_cg.py:126-137."""
    lib = _lib.load()
    b = as_device(b, torch.float64).contiguous()
    n = b.shape[0]
    x0 = torch.zeros_like(b) if x0 is None else as_device(x0, torch.float64).contiguous()
    if maxiter is None:
        maxiter = 10 * n                                              # copied from scipy (solve_cg.py:84-86)
    pat = getattr(A, "_pattern", None)
    if pat is not None and (M is None or isinstance(M, JacobiPreconditioner)):
        x, k, hist = cg_on_pattern(pat, A.data, b, x0, tol, atol, maxiter, None if M is None else M.inv_diag)
        return x, {"iterations": k, "residual_norm_history": hist}

    # general operators: host-driven loop over the same kernels
    matvec = _as_operator(A)
    if M is None:
        prec = None
    elif callable(M) and not isinstance(M, torch.Tensor):
        prec = M
    else:
        prec = _as_operator(M)
    dev = b.device
    ws = _lib.workspace(lib.feb200_reduce_workspace_bytes(), dev)
    ws.zero_()
    out = torch.zeros(1, dtype=torch.float64, device=dev)
    atol2 = max(tol * tol * _dot(b, b, out, ws), atol * atol)
    x = x0.clone()
    r = (b - matvec(x)).contiguous()
    z = r if prec is None else prec(r).contiguous()
    p = z.clone()
    gamma = _dot(r, z, out, ws)
    hist = torch.zeros(maxiter, dtype=torch.float64, device=dev)
    if maxiter > 0:
        hist[0] = _dot(r, r, out, ws) ** 0.5
    k = 0
    while True:
        rs = gamma if prec is None else _dot(r, r, out, ws)
        if not (rs > atol2 and k < maxiter):
            break
        Ap = matvec(p).contiguous()
        alpha = gamma / _dot(p, Ap, out, ws)
        _lib.call("feb200_axpby", n, 1.0, _lib.ptr(x), alpha, _lib.ptr(p), _lib.ptr(x))
        _lib.call("feb200_axpby", n, 1.0, _lib.ptr(r), -alpha, _lib.ptr(Ap), _lib.ptr(r))
        z = r if prec is None else prec(r).contiguous()
        gamma_new = _dot(r, z, out, ws)
        beta = gamma_new / gamma
        pn = torch.empty_like(p)
        _lib.call("feb200_axpby", n, 1.0, _lib.ptr(z), beta, _lib.ptr(p), _lib.ptr(pn))
        p = pn
        if k + 1 < maxiter:
            hist[k + 1] = _dot(r, r, out, ws) ** 0.5
        gamma = gamma_new
        k += 1
    return x, {"iterations": k, "residual_norm_history": hist}
