"""Search for a bank-conflict-free layout of the sweep kernel's accumulator panels (csrc/fused_sweep.cu): emulates the
64-bit shared-memory accesses of the accumulator loads / stores of every row tile of a real plan (8 x 8 x 24 grid), 16 lanes
per half-warp = 4 rows x 4 column pairs, address = base[slot] + rank * BS + d * DS, and anneals over the per-slot base
offsets for every block stride BS, component stride DS and assignment of the 8 local nodes to the two column sets.
python profiles/panel_layout_search.py [1 = component-major rows | 0 = node-major rows]     (CPU only, minutes)
Result (profiles/r02_panel_layout_search.txt): nothing below 1.59 wavefronts per ideal wavefront with three nodes per row
tile -> the kernel uses two-node row tiles instead (one node per half-warp: conflict-free by construction)."""
import sys, itertools
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'fea-in-jax_b200')]
import numpy as np, torch
from oracle import fe_oracle as orc
from fe_jax import sweep_plan
dims = (8, 8, 24)
Vx, cells = orc.uniform_hex_grid(*dims)
V = Vx.shape[0]
ri = orc.build_row_ind(V, cells).astype(np.int64)
ci = orc.build_col_ind_and_data(ri.astype(np.uint64), cells)[0].astype(np.int64)
nb = [set() for _ in range(V)]
for c in cells:
    for a in c: nb[a].update(c.tolist())
lists = [sorted(s) for s in nb]
nindptr = np.zeros(V + 1, dtype=np.int64); nindptr[1:] = np.cumsum([len(l) for l in lists])
plan = sweep_plan.build_sweep_plan(torch.from_numpy(Vx), torch.from_numpy(cells.astype(np.int32)), torch.from_numpy(ri), torch.from_numpy(ci), torch.from_numpy(nindptr))
visits = plan.visits.numpy()[:, :4].astype(np.int64) & 0xFFFFFFFF
rank = np.zeros((cells.shape[0], 8, 8), dtype=np.int64)
for e, c in enumerate(cells):
    for a in range(8):
        for b in range(8):
            rank[e, a, b] = lists[c[a]].index(c[b])
rng = np.random.default_rng(5)
DM = int(sys.argv[1]) if len(sys.argv) > 1 else 1          # 1 = d-major rows
rows_list = []
for (e, slotw, perm, flags) in visits:
    cnt = int(flags & 0xF)
    for mt in range((3 * cnt + 7) // 8):
        for half in (0, 1):
            rows = []
            for g in range(4):
                R = 8 * mt + 4 * half + g
                if R < 3 * cnt:
                    d, s = (R // cnt, R % cnt) if DM else (R % 3, R // 3)
                    na = (perm >> (4 * s)) & 7
                    rows.append((int((slotw >> (4 * na)) & 0xF), int(d)) + tuple(int(x) for x in rank[e, na]))
            if rows:
                while len(rows) < 4: rows.append((-1, 0) + (0,) * 8)
                rows_list.append(tuple(rows))
from collections import Counter
cnts = Counter(rows_list)
pats = list(cnts.keys()); wts = np.array([cnts[p] for p in pats], dtype=np.float64)
P = len(pats)
print('unique half-warp patterns', P, 'of', len(rows_list), flush=True)
slot = np.array([[r[0] for r in p] for p in pats]); dd = np.array([[r[1] for r in p] for p in pats])
rk8 = np.array([[r[2:] for r in p] for p in pats])          # (P,4,8)
valid = slot >= 0
def cost(base, BS, DS, set0, set1):
    tot = 0.0
    for cs in (set0, set1):
        addr = base[np.maximum(slot, 0)][:, :, None] + rk8[:, :, list(cs)] * BS + dd[:, :, None] * DS
        bank = addr % 16
        oh = np.zeros((P, 16), dtype=np.int32)
        v = np.broadcast_to(valid[:, :, None], bank.shape)
        np.add.at(oh, (np.broadcast_to(np.arange(P)[:, None, None], bank.shape)[v], bank[v]), 1)
        tot += (oh.max(axis=1) * wts).sum()
    return tot / (2 * wts.sum())
nodes = list(range(8))
parts = [(s0, tuple(n for n in nodes if n not in s0)) for s0 in itertools.combinations(nodes, 4) if 0 in s0]
res = []
for BS in (9, 11, 13, 15):
    for DS in (3, 1):
        for (set0, set1) in parts:
            best = (9e9, None)
            for restart in range(3):
                base = rng.integers(0, 16, size=8)
                cur = cost(base, BS, DS, set0, set1)
                T = 0.05
                for it in range(400):
                    s = rng.integers(0, 8); old = base[s]; base[s] = rng.integers(0, 16)
                    c = cost(base, BS, DS, set0, set1)
                    if c <= cur or rng.random() < np.exp((cur - c) / T): cur = c
                    else: base[s] = old
                    T *= 0.99
                if cur < best[0]: best = (cur, base.copy())
            res.append((round(float(best[0]), 4), BS, DS, set0, set1, best[1].tolist()))
    res.sort()
    print(BS, res[:4], flush=True)
print('BEST', res[:10])
