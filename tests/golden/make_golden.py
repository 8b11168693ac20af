"""
Generate tests/golden/*.npz by executing the REFERENCE'S OWN SOURCE TEXT.

JAX / flax / basix are not installable in this image, so `import fe_jax` fails.
This script therefore parses the reference files with `ast`, drops their import
statements and executes the selected function definitions in a namespace where
`jax.numpy` is replaced by a thin NumPy stand-in (same function names, same
einsum strings, `.at[...]` updates, `lax.while_loop` as a Python loop) and
`numba.njit` is the real numba.  Nothing is copied into this repository: the
sources are read from /root/reference at run time (build container only) and
only their *outputs* on seeded inputs are stored.

The single deliberate patch: `rank2_voigt_to_tensor` tests `voigt.shape[2] == 6`
(src/fe_jax/utils.py:246), which raises on (Q, 6) input; it is rewritten to
`voigt.shape[-1] == 6` so that the 3-D branches of the constitutive laws can
run at all.  3-D goldens are labelled accordingly.

Run:  python tests/golden/make_golden.py      (needs /root/reference)
"""
import ast
import os
import sys
import types

import numpy as np

REF = "/root/reference/src/fe_jax"
OUT = os.path.dirname(os.path.abspath(__file__))


# ------------------------------------------------------------------ numpy stand-in for jax.numpy
class _AtIndexer:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtOp(self.arr, idx)


class _AtOp:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def _copy(self):
        return np.array(self.arr).view(JArr)

    def set(self, v):
        out = self._copy()
        try:
            out[self.idx] = v
        except IndexError:          # JAX drops out-of-bounds updates
            pass
        return out

    def add(self, v):
        out = self._copy()
        np.add.at(out, self.idx, v)
        return out

    def multiply(self, v):
        out = self._copy()
        out[self.idx] = out[self.idx] * v
        return out


class JArr(np.ndarray):
    @property
    def at(self):
        return _AtIndexer(self)


def _wrap(x):
    if isinstance(x, np.ndarray):
        return x.view(JArr)
    if isinstance(x, tuple):
        return tuple(_wrap(v) for v in x)
    if isinstance(x, (np.floating, np.integer, np.bool_)):
        return np.asarray(x).view(JArr)
    return x


def _wrapfn(f):
    def g(*a, **k):
        return _wrap(f(*a, **k))
    return g


def _bincount(x, weights=None, minlength=0, length=None):
    n = length if length is not None else minlength
    out = np.bincount(np.asarray(x), weights=weights, minlength=n)
    return out[:n] if length is not None else out


jnp = types.SimpleNamespace()
for _n in ("einsum array zeros zeros_like ones ones_like eye lexsort vstack hstack diff append cumsum sort "
           "where max isin maximum square concatenate multiply sum sqrt abs arange").split():
    setattr(jnp, _n, _wrapfn(getattr(np, _n)))
jnp.bincount = _wrapfn(_bincount)
jnp.newaxis = None
jnp.ndarray = np.ndarray
jnp.float64 = np.float64
jnp.int64 = np.int64
jnp.linalg = types.SimpleNamespace(inv=_wrapfn(np.linalg.inv), det=_wrapfn(np.linalg.det),
                                   norm=_wrapfn(np.linalg.norm))
jnp.result_type = lambda *a: np.result_type(*a)

jax = types.SimpleNamespace()
jax.jit = lambda f=None, **k: (f if f is not None else (lambda g: g))
jax.numpy = jnp
jax.Array = np.ndarray


def _while_loop(cond_fun, body_fun, init_val):
    v = init_val
    while bool(cond_fun(v)):
        v = body_fun(v)
    return v


lax = types.SimpleNamespace(while_loop=_while_loop)


class COO:  # stand-in for jax.experimental.sparse.COO (attribute names only)
    def __init__(self, args, shape, rows_sorted=False, cols_sorted=False):
        self.data, self.row, self.col = (_wrap(np.asarray(a)) for a in args)
        self.shape = shape
        self._rows_sorted, self._cols_sorted = rows_sorted, cols_sorted


jsparse = types.SimpleNamespace(COO=COO)


def load_defs(filename, names=None, patches=()):
    """exec the (import-stripped) function definitions of a reference file."""
    src = open(os.path.join(REF, filename)).read()
    for old, new in patches:
        assert old in src, (filename, old)
        src = src.replace(old, new)
    tree = ast.parse(src)
    body = []
    for node in tree.body:
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            continue
        if isinstance(node, ast.FunctionDef) and (names is None or node.name in names):
            body.append(node)
    return ast.Module(body=body, type_ignores=[])


def main():
    import functools
    import inspect
    import numba

    ns = dict(jax=jax, jnp=jnp, np=np, lax=lax, jsparse=jsparse, partial=functools.partial,
              Callable=object, Any=object, njit=numba.njit, inspect=inspect,
              debug_print=lambda *a, **k: None)
    ns["_identity"] = lambda x: x
    ns["_vdot_real_tree"] = lambda a, b: _wrap(np.vdot(a, b).real)
    ns["_add"] = lambda a, b: a + b
    ns["_sub"] = lambda a, b: a - b
    ns["_mul"] = lambda s, a: s * a
    ns["tree_leaves"] = lambda x: [x]

    for fname, names, patches in [
        ("utils.py", {"rank2_tensor_to_voigt", "rank2_voigt_to_tensor", "is_required"},
         [("elif voigt.shape[2] == 6:", "elif voigt.shape[-1] == 6:")]),
        ("linear_elasticity.py", None, []),
        ("hyperelasticity.py", None, []),
        ("setup.py", {"uniform_quad_grid", "uniform_tri_grid", "build_row_ind", "build_col_ind_and_data",
                      "mesh_to_jax_helper"}, []),
        ("sparse_matrix.py", {"coo_arrays_sum_duplicates", "coo_arrays_sum_duplicates_jit",
                              "apply_dirichlet_bcs_lhs"}, []),
        ("solve_cg.py", {"_cg_solve"}, []),
    ]:
        mod = load_defs(fname, names, patches)
        exec(compile(mod, os.path.join(REF, fname), "exec"), ns)

    sys.path.insert(0, os.path.join(OUT, "..", ".."))
    from oracle import fe_oracle as orc   # only for the basis tables fed to the reference functions

    rng = np.random.default_rng(1234)
    G = {}

    # ---- grids and index maps (real numba execution of the reference functions)
    Vq, Fq = ns["uniform_quad_grid"](5, 7, np.array([[0.0, 2.0], [-1.0, 1.0]]))
    Vt, Ft = ns["uniform_tri_grid"](7, 10)
    G["quad_grid_V"], G["quad_grid_F"], G["tri_grid_V"], G["tri_grid_F"] = Vq, Fq, Vt, Ft
    for tag, (nv, cells) in {"tri": (Vt.shape[0], Ft), "quad": (Vq.shape[0], Fq)}.items():
        ri = ns["build_row_ind"](nv, cells.astype(np.uint64))
        ci, dat = ns["build_col_ind_and_data"](ri, cells.astype(np.uint64))
        G[f"maps_{tag}_cells"], G[f"maps_{tag}_row_ind"] = cells, ri
        G[f"maps_{tag}_col_ind"], G[f"maps_{tag}_data"] = ci, dat
    _, Fh = orc.uniform_hex_grid(3, 4, 2)
    perm_cells = Fh[rng.permutation(Fh.shape[0])]       # unordered mesh, irregular valence order
    ri = ns["build_row_ind"](int(Fh.max()) + 1, perm_cells.astype(np.uint64))
    ci, dat = ns["build_col_ind_and_data"](ri, perm_cells.astype(np.uint64))
    G["maps_hex_cells"], G["maps_hex_row_ind"], G["maps_hex_col_ind"] = perm_cells, ri, ci
    G["mesh_to_jax_tri"] = ns["mesh_to_jax_helper"](Vt, Ft.astype(np.uint64))

    # ---- Voigt
    t2, t3 = rng.normal(size=(5, 2, 2)), rng.normal(size=(5, 3, 3))
    G["voigt_t2"], G["voigt_t3"] = t2, t3
    G["voigt_v2"] = ns["rank2_tensor_to_voigt"](_wrap(t2))
    G["voigt_v3"] = ns["rank2_tensor_to_voigt"](_wrap(t3))
    G["voigt_back2"] = ns["rank2_voigt_to_tensor"](_wrap(np.array(G["voigt_v2"])))
    G["voigt_back3"] = ns["rank2_voigt_to_tensor"](_wrap(np.array(G["voigt_v3"])))

    # ---- constitutive laws
    Q = 8
    for D in (2, 3):
        eps = rng.normal(size=(Q, D, D)) * 1e-3
        eps = 0.5 * (eps + eps.transpose(0, 2, 1))
        Fd = np.eye(D)[None] + 0.05 * rng.normal(size=(Q, D, D))
        iso = np.stack([10e9 + 90e9 * rng.uniform(size=Q), 0.2 + 0.2 * rng.uniform(size=Q)], axis=1)
        if D == 2:
            orth = np.stack([40e9 + 5e9 * rng.uniform(size=Q), 10e9 + 1e9 * rng.uniform(size=Q),
                             0.25 + 0.05 * rng.uniform(size=Q), 4e9 + 1e9 * rng.uniform(size=Q)], axis=1)
        else:
            orth = np.stack([40e9 + 5e9 * rng.uniform(size=Q), 10e9 + 1e9 * rng.uniform(size=Q),
                             12e9 + 1e9 * rng.uniform(size=Q), 0.25 + 0.05 * rng.uniform(size=Q),
                             0.30 + 0.05 * rng.uniform(size=Q), 0.20 + 0.05 * rng.uniform(size=Q),
                             4e9 + 1e9 * rng.uniform(size=Q), 3e9 + 1e9 * rng.uniform(size=Q),
                             5e9 + 1e9 * rng.uniform(size=Q)], axis=1)
        mr = np.stack([0.5e6 + 1e5 * rng.uniform(size=Q), 0.1e6 + 1e4 * rng.uniform(size=Q),
                       2e-7 + 1e-8 * rng.uniform(size=Q)], axis=1)
        G[f"mat{D}_eps"], G[f"mat{D}_F"] = eps, Fd
        G[f"mat{D}_iso_mp"], G[f"mat{D}_orth_mp"], G[f"mat{D}_mr_mp"] = iso, orth, mr
        G[f"mat{D}_iso_sig"] = ns["elastic_isotropic"](_wrap(eps), _wrap(iso))[0]
        G[f"mat{D}_orth_sig"] = ns["elastic_orthotropic"](_wrap(eps), _wrap(orth))[0]
        G[f"mat{D}_svk_S"] = ns["st_venant_kirchhoff"](_wrap(Fd), _wrap(iso))
        G[f"mat{D}_mr_S"] = ns["mooney_rivlin"](_wrap(Fd), _wrap(mr))

    # ---- element residuals (tables from the oracle's basix stand-in; the reference
    #      function receives them as plain arrays, exactly as fea.py:824-845 does)
    cases = {"tri3": ("triangle", 1, 3, 2), "quad4": ("quadrilateral", 1, 2, 2),
             "hex8": ("hexahedron", 1, 2, 3), "tet4": ("tetrahedron", 1, 2, 3)}
    for tag, (cell, deg, qdeg, D) in cases.items():
        xi, W = orc.get_quadrature(cell, qdeg)
        _, dphi = orc.eval_basis_and_derivatives(cell, deg, xi)
        nodes = orc.reference_nodes(cell, deg)
        N, Qn = nodes.shape[0], xi.shape[0]
        x = nodes * np.array([1.3, 0.9, 1.1][:D]) + 0.08 * rng.normal(size=(N, D))
        u = 1e-3 * rng.normal(size=(N, D))
        iso = np.stack([10e9 + 90e9 * rng.uniform(size=Qn), np.full(Qn, 0.25)], axis=1)
        mr = np.tile(np.array([0.5e6, 0.1e6, 2e-7]), (Qn, 1))
        W1 = W[0:1]                                 # what fea.py:285-297 actually passes
        G[f"res_{tag}_x"], G[f"res_{tag}_u"], G[f"res_{tag}_dphi"], G[f"res_{tag}_W"] = x, u, dphi, W
        G[f"res_{tag}_iso_mp"], G[f"res_{tag}_mr_mp"] = iso, mr
        R, st = ns["linear_elasticity_residual"](_wrap(u), _wrap(x), _wrap(dphi), _wrap(W1), _wrap(iso),
                                                 _wrap(np.zeros((Qn, 0))), ns["elastic_isotropic"])
        G[f"res_{tag}_lin_iso_R"] = R
        uh = 0.05 * rng.normal(size=(N, D))
        G[f"res_{tag}_uh"] = uh
        G[f"res_{tag}_hyp_svk_R"] = ns["hyperelasticity_residual"](
            _wrap(uh), _wrap(x), _wrap(dphi), _wrap(W1), _wrap(iso), ns["st_venant_kirchhoff"])
        G[f"res_{tag}_hyp_mr_R"] = ns["hyperelasticity_residual"](
            _wrap(uh), _wrap(x), _wrap(dphi), _wrap(W1), _wrap(mr), ns["mooney_rivlin"])

    # ---- COO duplicate summation and Dirichlet elimination
    n, nse = 40, 300
    row = rng.integers(0, n, size=nse)
    col = rng.integers(0, n, size=nse)
    dat = rng.normal(size=nse)
    perm = np.lexsort((col, row))                   # _sort_indices() before the call (fea.py:651-654)
    A = COO((dat[perm], row[perm], col[perm]), shape=(n, n), rows_sorted=True)
    d1, r1, c1 = ns["coo_arrays_sum_duplicates"](A)
    nnz = d1.shape[0]
    d2, r2, c2 = ns["coo_arrays_sum_duplicates_jit"](A, nnz)
    G["coo_in_data"], G["coo_in_row"], G["coo_in_col"] = dat[perm], row[perm], col[perm]
    G["coo_out_data"], G["coo_out_row"], G["coo_out_col"] = d1, r1, c1
    G["coo_jit_data"], G["coo_jit_row"], G["coo_jit_col"] = d2, r2, c2
    ddofs = np.array([0, 3, 17, 39])
    B = ns["apply_dirichlet_bcs_lhs"](COO((d1, r1, c1), shape=(n, n), rows_sorted=True), _wrap(ddofs))
    G["dirichlet_dofs"], G["dirichlet_lhs_data"] = ddofs, B.data

    # ---- CG recurrence (reference _cg_solve) on the 3x3 system of tests/test_solve_cg.py
    #      and on a random SPD system, with and without a Jacobi preconditioner
    A3 = np.array([[2.0, -1.0, 0.0], [-1.0, 2.0, 0.0], [0.0, 0.0, 1.0]])
    b3 = np.array([2.0, 5.0, 1.0])
    Mrand = rng.normal(size=(30, 30))
    Aspd = Mrand @ Mrand.T + 30 * np.eye(30)
    bspd = rng.normal(size=30)
    for tag, (Am, bv, maxiter) in {"cg3": (A3, b3, 30), "cgspd": (Aspd, bspd, 12)}.items():
        for ptag, Mfun in (("none", ns["_identity"]), ("jacobi", (lambda d: (lambda v: _wrap(v / d)))(np.diag(Am)))):
            x, info = ns["_cg_solve"](lambda v: _wrap(Am @ v), _wrap(bv), x0=_wrap(np.zeros_like(bv)),
                                      maxiter=maxiter, tol=1e-12, atol=0.0, M=Mfun)
            G[f"{tag}_{ptag}_x"] = x
            G[f"{tag}_{ptag}_iters"] = np.array(int(info["iterations"]))
            G[f"{tag}_{ptag}_hist"] = info["residual_norm_history"]
        G[f"{tag}_A"], G[f"{tag}_b"] = Am, bv

    np.savez_compressed(os.path.join(OUT, "reference_outputs.npz"),
                        **{k: np.asarray(v) for k, v in G.items()})
    print("wrote", os.path.join(OUT, "reference_outputs.npz"), len(G), "arrays")


if __name__ == "__main__":
    main()
