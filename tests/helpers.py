"""Shared builders for the parity tests: the same problem as an fe_jax (CUDA) collection and as an
oracle Problem."""
import numpy as np

from oracle import fe_oracle as orc

CELL_DIM = {"triangle": 2, "quadrilateral": 2, "tetrahedron": 3, "hexahedron": 3}
KIND_OF = {"elastic_isotropic": "linear", "elastic_orthotropic": "linear", "st_venant_kirchhoff": "hyper",
           "mooney_rivlin": "hyper"}


def lattice_mesh(cell, degree, n, jitter=0.15, seed=0):
    """Small conforming mesh of `cell` elements of `degree` (1 or 2) on an n^D block of cubes/squares;
    vertices jittered so no two elements are congruent."""
    D = CELL_DIM[cell]
    L = degree * n + 1                                   # lattice points per axis
    grid = np.stack(np.meshgrid(*([np.arange(L)] * D), indexing="ij"), axis=-1).reshape(-1, D)
    pts = grid / float(degree * n)
    strides = np.array([L ** (D - 1 - d) for d in range(D)])
    ref = orc.reference_nodes(cell if cell in ("quadrilateral", "hexahedron") else
                              ("quadrilateral" if D == 2 else "hexahedron"), 1)
    cells = []
    rn = np.rint(orc.reference_nodes(cell, degree) * degree).astype(int)   # lattice offsets of the element nodes
    if cell in ("quadrilateral", "hexahedron"):
        for c in np.ndindex(*([n] * D)):
            base = np.array(c) * degree
            cells.append([(base + o) @ strides for o in rn])
    else:
        # split every square / cube into simplices; simplex vertices are cube corners
        if D == 2:
            splits = [((0, 0), (1, 0), (0, 1)), ((1, 0), (1, 1), (0, 1))]
        else:
            import itertools
            splits = []
            for perm in itertools.permutations(range(3)):      # Kuhn triangulation: 6 tets per cube
                v = [np.zeros(3, dtype=int)]
                for ax in perm:
                    w = v[-1].copy()
                    w[ax] = 1
                    v.append(w)
                # orient positively
                M = np.array([v[1] - v[0], v[2] - v[0], v[3] - v[0]])
                if np.linalg.det(M) < 0:
                    v[2], v[3] = v[3], v[2]
                splits.append(tuple(tuple(w) for w in v))
        bary_nodes = orc.reference_nodes(cell, degree)           # in reference coordinates
        for c in np.ndindex(*([n] * D)):
            base = np.array(c) * degree
            for sv in splits:
                sv = np.array(sv) * degree                        # lattice offsets of simplex vertices
                lam = np.concatenate([1 - bary_nodes.sum(1, keepdims=True), bary_nodes], axis=1)
                offs = np.rint(lam @ sv).astype(int)
                cells.append([(base + o) @ strides for o in offs])
    cells = np.array(cells, dtype=np.int64)
    rng = np.random.default_rng(seed)
    h = 1.0 / (degree * n)
    pts = pts + jitter * h * rng.uniform(-1, 1, size=pts.shape) * (degree == 1)
    if degree == 2:     # keep mid-side nodes consistent: jitter the whole lattice smoothly instead
        pts = pts + 0.05 * np.sin(3.0 * pts[:, ::-1] + 0.3)
    return pts, cells


def material_params(model, D, shape, seed=0):
    """random but sane parameters with leading `shape` (e.g. (E, Q), (E,), ())"""
    rng = np.random.default_rng(seed)
    u = lambda: rng.uniform(size=shape)
    if model in ("elastic_isotropic", "st_venant_kirchhoff"):
        cols = [10e9 + 90e9 * u(), 0.2 + 0.2 * u()]
    elif model == "mooney_rivlin":
        cols = [0.5e6 + 1e5 * u(), 0.1e6 + 1e4 * u(), 2e-7 + 1e-8 * u()]
    elif D == 2:
        cols = [40e9 + 5e9 * u(), 10e9 + 1e9 * u(), 0.25 + 0.05 * u(), 4e9 + 1e9 * u()]
    else:
        cols = [40e9 + 5e9 * u(), 10e9 + 1e9 * u(), 12e9 + 1e9 * u(), 0.25 + 0.05 * u(), 0.30 + 0.05 * u(),
                0.20 + 0.05 * u(), 4e9 + 1e9 * u(), 3e9 + 1e9 * u(), 5e9 + 1e9 * u()]
    return np.stack(cols, axis=-1)


def fe_type(fj, cell, degree, qdeg):
    return fj.FiniteElementType(getattr(fj.CellType, cell), fj.ElementFamily.P, degree, fj.LagrangeVariant.equispaced,
                                fj.QuadratureType.default, qdeg)


def build_pair(fj, pts, batch_specs):
    """batch_specs: list of dict(cell, degree, qdeg, cells, model, mp).  Returns (ebc, maps, residual_token,
    oracle Problem)."""
    fbatches, obatches = [], []
    kind = None
    for s in batch_specs:
        fe = fe_type(fj, s["cell"], s["degree"], s["qdeg"])
        D = pts.shape[1]
        fbatches.append(fj.ElementBatch(fe_type=fe, n_dofs_per_basis=D, connectivity_en=s["cells"],
                                        constitutive_model=getattr(fj, s["model"]), material_params=s["mp"]))
        xi, W = orc.get_quadrature(s["cell"], s["qdeg"])
        _, dphi = orc.eval_basis_and_derivatives(s["cell"], s["degree"], xi)
        kind = KIND_OF[s["model"]]
        obatches.append(orc.Batch(s["cells"], kind, s["model"], s["mp"], dphi, W))
    ebc = fj.batch_to_collection(pts, fbatches)
    maps = [fj.mesh_to_sparse_assembly_map(pts.shape[0], s["cells"]) for s in batch_specs]
    token = fj.linear_elasticity_residual if kind == "linear" else fj.hyperelasticity_residual
    return ebc, maps, token, orc.Problem(pts, obatches), fbatches


def rel_err(a, b):
    a = a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
    b = np.asarray(b)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
