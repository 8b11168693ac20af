"""CPU emulation of the fragment data flow of csrc/tangent_general_mma.cu (phase 2): the lane <-> matrix-entry maps of
mma.sync.m8n8k4.f64, the rotating ownership of the contraction index (point q, gradient component i) and the per-lane
select tables are restated in NumPy and checked against the plain contraction
    K[(a,d),(b,k)] = sum_q sum_{i,l} g_q[a,i] wA_q[(i,d),(k,l)] g_q[b,l]
(fea.py:519-572 with the hand-derived tangent of csrc/element.cu).  No GPU needed; the kernel itself is covered by the -m gpu
parity tests (hex8 x {orthotropic, St. Venant-Kirchhoff, Mooney-Rivlin})."""
import numpy as np

LANES = np.arange(32)
GID, TIG = LANES >> 2, LANES & 3


def dmma(C, a, b):
    """C (32, 2) += A (8 x 4: lane holds [gid][tig]) @ B (4 x 8: lane holds [tig][gid]); lane holds C[gid][2 tig + j]."""
    A = a.reshape(8, 4)
    B = b.reshape(8, 4).T
    P = A @ B
    for j in range(2):
        C[:, j] += P[GID, 2 * TIG + j]


def phase2(G, A):
    """G: (8 points, 8 nodes, 3), A: (8 points, 81) = w A_q[(i*3+d)*9 + k*3 + l].  Returns K_e[(a)][b][d][k] (8, 8, 3, 3)."""
    Gf = G.reshape(8, 24)
    C = np.zeros((9, 32, 2))                                    # [dk] accumulators
    tig_o, jcol = GID >> 1, GID & 1                             # product 1: owner and slot parity of this lane's B column
    for h in range(2):
        Y = np.zeros((4, 5, 32, 2))                             # [r][nt] C fragments of product 1
        for r in range(4):
            q = 4 * h + r
            a1 = np.where(TIG < 3, Gf[q][GID * 3 + np.minimum(TIG, 2)], 0.0)          # g_q[b = gid][l = tig]
            io = np.minimum((tig_o - r) & 3, 2)
            for nt in range(5):
                dk = np.minimum(2 * nt + jcol, 8)
                b1 = A[q][io * 27 + dk * 3 + TIG - (TIG == 3)]  # tig = 3: any finite value (its A column is zero)
                dmma(Y[r, nt], a1, b1)
        for st in range(3):
            if st == 0:
                rsel = np.where(TIG == 3, 1, 0)
            elif st == 1:
                rsel = np.where((TIG == 0) | (TIG == 3), 2, 1)
            else:
                rsel = np.where(TIG == 2, 2, 3)
            qsel = 4 * h + rsel
            isel = (TIG - rsel) & 3
            assert (isel < 3).all()
            aop = Gf[qsel, GID * 3 + isel]
            for dk in range(9):
                y = Y[rsel, dk >> 1, LANES, dk & 1]
                dmma(C[dk], aop, y)
    K = np.zeros((8, 8, 3, 3))
    for dk in range(9):
        for j in range(2):
            K[GID, 2 * TIG + j, dk // 3, dk % 3] = C[dk][:, j]
    return K


def test_rotating_ownership_covers_every_contraction_index_once():
    for dk in range(9):
        seen = set()
        for h in range(2):
            for st, table in enumerate([(0, 0, 0, 1), (2, 1, 1, 2), (3, 3, 2, 3)]):
                for t in range(4):
                    q = 4 * h + table[t]
                    i = (t - table[t]) & 3
                    assert i < 3
                    seen.add((q, i))
        assert len(seen) == 24


def test_fragment_data_flow_matches_plain_contraction():
    rng = np.random.default_rng(11)
    G = rng.normal(size=(8, 8, 3))
    A = rng.normal(size=(8, 81))
    K = phase2(G, A)
    A5 = A.reshape(8, 3, 3, 3, 3)                               # [q][i][d][k][l]
    K_ref = np.einsum("qai,qidkl,qbl->abdk", G, A5, G)
    assert np.abs(K - K_ref).max() <= 1e-12 * np.abs(K_ref).max()
