// Fused hex8 assembly: element tangents are computed and summed into the CSR row panels inside one
// kernel, so K_e never goes to HBM (the two-pass path of element.cu + assemble.cu moves ~5x the
// algorithmic bytes).  Replaces, for linear isotropic hex8 batches,
//   vmap(jacfwd(residual_kernel))            /root/reference/src/fe_jax/fea.py:519-572
//   COO concat + sort + coo_sum_duplicates   fea.py:613-660, src/fe_jax/sparse_matrix.py:44-110
//   apply_dirichlet_bcs_lhs                  sparse_matrix.py:157-184 (optional, fused into the store)
//
// Work decomposition ("tile plan", built once per mesh by fe_jax/fused_plan.py, any connectivity):
//   tile  = up to TN vertices that are close in space (Morton cell) -> one CTA, 8 threads per vertex
//   P1    geometry of every element touching the tile, once per tile: g[le][n][q][3], detJ*W   -> smem
//   P2    one thread per (vertex a, incident slot k): the block row K_e[a, 0..8) of that element from the
//         staged gradients: S_ab = sum_q w g_a (x) g_b (72 FMA), K_ab = lam S + mu S^T + mu tr(S) I
//   P3    block rows to shared memory (aliases the gradients)
//   P4    one lane per CSR block (a, r): adds the <= 8 contributions in ascending slot order -- the order
//         a stable (row, col) sort + bincount gives -- applies the Dirichlet elimination and stores the
//         panel entries (coalesced).  No atomics; bit-reproducible.
#include "common.cuh"
#include "element_math.cuh"

namespace feb {

constexpr int FUSED_Q = 8;
constexpr int FUSED_MAXVAL = 8;          // incident slots per vertex handled (lane = 4*k + b/2)
constexpr int FUSED_MAXDEG = 32;         // neighbour blocks per vertex (one lane each; hex8: 27)
#ifndef FEB_FUSED_WARPS
#define FEB_FUSED_WARPS 12
#endif
#ifndef FEB_FUSED_MINBLOCKS
#define FEB_FUSED_MINBLOCKS 1
#endif
constexpr int FUSED_WARPS = FEB_FUSED_WARPS;
constexpr int FUSED_THREADS = FUSED_WARPS * 32;
constexpr int G_STRIDE = 26;             // doubles per (element, node) gradient record (24 used): 16-byte aligned
constexpr int EL_STRIDE = 8 * G_STRIDE + 2;   // doubles per element in G: records of consecutive elements are 4 banks apart
constexpr int XS_STRIDE = 25;            // doubles per element coordinate record (24 used)
constexpr int DPHI_STRIDE = 25;          // doubles per quadrature point in the basis table (24 used)
constexpr int WS_DOUBLES = FUSED_MAXVAL * 8 * 9 + 4;  // per-warp staging: block (k, b) at (k*8+b)*9 + (k>>1): the per-k pad makes
                                                      // both the writes (lane = 4k + b/2) and the reads (8 b's of one k) conflict-free
constexpr int WD_STRIDE = 9;             // doubles per element in wdet (8 used, odd)
constexpr int PF = 2;                    // element-node prefetch registers per thread: te_cap*8 <= PF*THREADS (te_cap <= 96)
constexpr int PN = (FUSED_MAXDEG * 32 + FUSED_THREADS - 1) / FUSED_THREADS;   // neighbour-table entries per thread

struct FusedSmem {
    int tn_cap, te_cap;
    size_t off_g, off_ws, off_dphi, off_lm, off_ints, total;
};
static inline FusedSmem fused_layout(int tn_cap, int te_cap) {
    FusedSmem L;
    L.tn_cap = tn_cap;
    L.te_cap = te_cap;
    L.off_g = 0;
    size_t g_doubles = (size_t)te_cap * EL_STRIDE + 2 * (size_t)te_cap * WD_STRIDE;    // G + wdet + lambda/mu ratio
    g_doubles = (g_doubles + 1) & ~(size_t)1;
    L.off_ws = g_doubles * 8;
    size_t ws_doubles = (size_t)FUSED_WARPS * WS_DOUBLES;                          // xs (te*25) aliases the start
    size_t xs_doubles = (size_t)te_cap * XS_STRIDE;
    if (ws_doubles < xs_doubles) ws_doubles = xs_doubles;
    ws_doubles = (ws_doubles + 1) & ~(size_t)1;
    L.off_dphi = L.off_ws + ws_doubles * 8;
    L.off_lm = L.off_dphi + (size_t)FUSED_Q * DPHI_STRIDE * 8;
    L.off_ints = L.off_lm + (size_t)te_cap * 2 * 8;
    L.off_ints = (L.off_ints + 15) & ~(size_t)15;
    // tbl[tn*32] u64 | p0[tn] i64 | items[tn*8] u32 | nbr[tn*32] i32 | node[tn], meta[tn] i32 | flg[tn*32] u8
    L.total = L.off_ints + (size_t)tn_cap * FUSED_MAXDEG * 8 + (size_t)tn_cap * 8 + (size_t)tn_cap * FUSED_MAXVAL * 4 +
              (size_t)tn_cap * FUSED_MAXDEG * 4 + (size_t)tn_cap * 4 * 2 + (size_t)tn_cap * FUSED_MAXDEG + 64;
    return L;
}

// Everything a tile reads from global memory is fetched one tile ahead, in three dependency levels that are issued
// at different points of the previous tile's computation:
//   A (plan arrays)          element ids, the thread's item word, vertex id / CSR offset / degree
//   B (needs A)              vertex ids of the elements, neighbour vertex ids, element id of the item
//   C (needs B)              coordinates, constraint flags of the neighbours, slot ranks of the item
struct TilePrefetch {
    int32_t el[PF];          // A
    uint32_t item;           // A
    int32_t node, meta;      // A   (meta = degree | valence << 8 | row flags << 16), valid for tid < TN
    int64_t p0;              // A
    int64_t p0n[PN];         // A   CSR offset of the vertex owning neighbour-table entry tid + j*THREADS
    int32_t degn[PN];        // A
    int32_t vtx[PF];         // B
    int32_t nbr[PN];         // B
    int32_t item_el;         // B
    double mpq[PF][2];       // B   PERQ: (E, nu) of quadrature point (tid + j*THREADS) & 7 of element el[j]
    double x[PF][3];         // C
    uint32_t flg;            // C   PN flags packed
    uint4 rk8;               // C
};

// P0/P1 are CTA-wide (one barrier each); after that every warp works alone on one vertex at a time:
//   A  lane = 4*k + b/2 computes the two blocks K_e[a, b], K_e[a, b+1] of incident slot k  -> warp-private staging
//   B  lane = neighbour rank r adds its <= 8 contributions in ascending slot order, eliminates, stores
// UNIFORM (material layout (M,): one lambda, mu for the batch): the blocks S_ab are summed first and
// K = lam S + mu S^T + mu tr(S) I is applied once per CSR block instead of once per contribution.
// PERQ (material layout (E,Q,M): lambda, mu per quadrature point): with m_q = detJ W mu g_a and l_q = (lambda/mu) m_q,
// K_ab = S_l + S_m^T + tr(S_m) I where S_l = sum_q l_q (x) g_b and S_m = sum_q m_q (x) g_b.
template <bool ELIM, bool UNIFORM, bool PERQ>
__global__ void __launch_bounds__(FUSED_THREADS, FEB_FUSED_MINBLOCKS)
k_fused_hex8_iso(feb200_batch b, feb200_fused_plan plan, FusedSmem L, const double* __restrict__ x_vd,
                 const uint16_t* __restrict__ slot_rank, const int32_t* __restrict__ nindices,
                 const uint8_t* __restrict__ vflags, double* __restrict__ values) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* G = reinterpret_cast<double*>(smem_raw + L.off_g);              // [le] EL_STRIDE : [n] G_STRIDE : g[q][3]
    double* wdet = G + (size_t)L.te_cap * EL_STRIDE;                        // [le][q]  (PERQ: detJ*W*mu)
    double* rlm = wdet + (size_t)L.te_cap * WD_STRIDE;                      // [le][q]  PERQ only: lambda / mu
    double* wstage = reinterpret_cast<double*>(smem_raw + L.off_ws);        // [warp][k][b][9]
    double* xs = wstage;                                                    // [le] XS_STRIDE : [n][3]   (P0-P1 only)
    double* s_dphi = reinterpret_cast<double*>(smem_raw + L.off_dphi);      // [q] DPHI_STRIDE : [n][3]
    double* lm = reinterpret_cast<double*>(smem_raw + L.off_lm);            // [le][2] lambda, mu
    unsigned long long* tbl = reinterpret_cast<unsigned long long*>(smem_raw + L.off_ints);  // [a][r]: byte k = b or 0xff
    int64_t* p0_s = reinterpret_cast<int64_t*>(tbl + (size_t)L.tn_cap * FUSED_MAXDEG);       // [a]
    uint32_t* items_s = reinterpret_cast<uint32_t*>(p0_s + L.tn_cap);       // [a][k] = le | na << 12
    int32_t* nbr_s = reinterpret_cast<int32_t*>(items_s + (size_t)L.tn_cap * FUSED_MAXVAL);  // [a][r] neighbour vertex
    int32_t* node_s = nbr_s + (size_t)L.tn_cap * FUSED_MAXDEG;              // [a]
    int32_t* meta_s = node_s + L.tn_cap;                                    // [a]
    uint8_t* flg_s = reinterpret_cast<uint8_t*>(meta_s + L.tn_cap);         // [a][r] constrained components of the neighbour

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < FUSED_Q * 24; i += FUSED_THREADS) s_dphi[(i / 24) * DPHI_STRIDE + (i % 24)] = b.dphi[i];

    TilePrefetch pf;
    auto level_a = [&](int tile) {
        const int n0 = plan.tile_node_ptr[tile], TN = plan.tile_node_ptr[tile + 1] - n0;
        const int e0 = plan.tile_elem_ptr[tile], TE = plan.tile_elem_ptr[tile + 1] - e0;
        const int i0 = plan.tile_item_ptr[tile], NI = plan.tile_item_ptr[tile + 1] - i0;
#pragma unroll
        for (int j = 0; j < PF; ++j) {
            const int i = tid + j * FUSED_THREADS;
            pf.el[j] = (i < TE * 8) ? plan.tile_elems[e0 + (i >> 3)] : -1;
        }
        pf.item = (tid < NI) ? plan.tile_items[i0 + tid] : 0xffffffffu;
        pf.node = -1;
        if (tid < TN) {
            pf.node = plan.tile_nodes[n0 + tid];
            pf.meta = plan.tile_node_meta[n0 + tid];
            pf.p0 = plan.tile_node_p0[n0 + tid];
        }
#pragma unroll
        for (int j = 0; j < PN; ++j) {
            const int a = (tid + j * FUSED_THREADS) / FUSED_MAXDEG;
            pf.degn[j] = 0;
            if (a < TN) {
                pf.p0n[j] = plan.tile_node_p0[n0 + a];
                pf.degn[j] = plan.tile_node_meta[n0 + a] & 0xff;
            }
        }
    };
    auto level_b = [&](int tile) {
        const int e0 = plan.tile_elem_ptr[tile];
#pragma unroll
        for (int j = 0; j < PF; ++j) {
            pf.vtx[j] = (pf.el[j] >= 0) ? b.conn[(int64_t)pf.el[j] * 8 + ((tid + j * FUSED_THREADS) & 7)] : -1;
            if (PERQ && pf.el[j] >= 0) {
                const double* mp = b.mp + ((int64_t)pf.el[j] * 8 + ((tid + j * FUSED_THREADS) & 7)) * b.M;
                pf.mpq[j][0] = mp[0];
                pf.mpq[j][1] = mp[1];
            }
        }
#pragma unroll
        for (int j = 0; j < PN; ++j) {
            const int r = (tid + j * FUSED_THREADS) % FUSED_MAXDEG;
            pf.nbr[j] = (r < pf.degn[j]) ? nindices[pf.p0n[j] + r] : -1;
        }
        pf.item_el = (pf.item != 0xffffffffu) ? plan.tile_elems[e0 + (pf.item & 0xfff)] : -1;
        if (ELIM && pf.node >= 0) pf.meta |= (int32_t)vflags[pf.node] << 16;      // constrained components of the row vertex
    };
    auto level_c = [&]() {
#pragma unroll
        for (int j = 0; j < PF; ++j)
            if (pf.vtx[j] >= 0) {
#pragma unroll
                for (int d = 0; d < 3; ++d) pf.x[j][d] = x_vd[(int64_t)pf.vtx[j] * 3 + d];
            }
        pf.flg = 0;
        if (ELIM) {
#pragma unroll
            for (int j = 0; j < PN; ++j)
                if (pf.nbr[j] >= 0) pf.flg |= (uint32_t)vflags[pf.nbr[j]] << (8 * j);
        }
        if (pf.item_el >= 0)
            pf.rk8 = *reinterpret_cast<const uint4*>(slot_rank + ((int64_t)pf.item_el * 8 + ((pf.item >> 12) & 0xf)) * 8);
    };

    int tile = blockIdx.x;
    if (tile < plan.n_tiles) { level_a(tile); level_b(tile); level_c(); }

    for (; tile < plan.n_tiles; tile += gridDim.x) {
        const int TN = plan.tile_node_ptr[tile + 1] - plan.tile_node_ptr[tile];
        const int TE = plan.tile_elem_ptr[tile + 1] - plan.tile_elem_ptr[tile];
        const int next = tile + gridDim.x;
        const bool has_next = next < plan.n_tiles;
        __syncthreads();                                     // previous tile fully consumed
        // ---- P0: prefetched data -> shared memory
#pragma unroll
        for (int j = 0; j < PF; ++j) {
            const int i = tid + j * FUSED_THREADS;
            if (i < TE * 8) {
                const int le = i >> 3, n = i & 7;
#pragma unroll
                for (int d = 0; d < 3; ++d) xs[le * XS_STRIDE + n * 3 + d] = pf.x[j][d];
                if (!PERQ && n == 0) {
                    int qs;
                    const double* mp = mp_of(b, pf.el[j], qs);
                    double lam, mu;
                    isotropic_lame<3>(mp, lam, mu);
                    lm[le * 2] = lam;
                    lm[le * 2 + 1] = mu;
                }
            }
        }
        if (tid < TN) { node_s[tid] = pf.node; meta_s[tid] = pf.meta; p0_s[tid] = pf.p0; }
#pragma unroll
        for (int j = 0; j < PN; ++j) {
            const int i = tid + j * FUSED_THREADS;
            if (i < TN * FUSED_MAXDEG) {
                nbr_s[i] = pf.nbr[j];
                flg_s[i] = (uint8_t)((pf.flg >> (8 * j)) & 0xffu);
                tbl[i] = ~0ull;
            }
        }
        const uint32_t item = pf.item;
        const uint4 rk8 = pf.rk8;
        if (item != 0xffffffffu) items_s[((item >> 16) & 0xff) * FUSED_MAXVAL + ((item >> 24) & 0xff)] = item & 0xffffu;
        if (has_next) level_a(next);
        __syncthreads();
        // source table: tbl[a][rank of local node b] byte k = b   (items of this tile, after the table was cleared)
        if (item != 0xffffffffu) {
            const unsigned int w4[4] = {rk8.x, rk8.y, rk8.z, rk8.w};
            uint8_t* tb = reinterpret_cast<uint8_t*>(tbl + (size_t)((item >> 16) & 0xff) * FUSED_MAXDEG) + ((item >> 24) & 0xff);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                tb[(w4[j] & 0xffffu) * 8] = (uint8_t)(2 * j);
                tb[(w4[j] >> 16) * 8] = (uint8_t)(2 * j + 1);
            }
        }
        // ---- P1: geometry, one thread per (element, quadrature point)
#pragma unroll
        for (int jj = 0; jj < PF; ++jj) {
            const int i = tid + jj * FUSED_THREADS;
            if (i >= TE * 8) break;
            const int le = i >> 3, q = i & 7;
            double g[24];
            const double dt = shape_gradients<8, 3>(xs + le * XS_STRIDE, s_dphi + q * DPHI_STRIDE, g);
            if (PERQ) {
                double lam, mu;
                isotropic_lame<3>(pf.mpq[jj], lam, mu);
                wdet[le * WD_STRIDE + q] = dt * b.W[q] * mu;
                rlm[le * WD_STRIDE + q] = lam / mu;
            } else {
                wdet[le * WD_STRIDE + q] = dt * b.W[q];
            }
#pragma unroll
            for (int n = 0; n < 8; ++n)
#pragma unroll
                for (int d = 0; d < 3; ++d) G[(size_t)le * EL_STRIDE + n * G_STRIDE + q * 3 + d] = g[n * 3 + d];
        }
        if (has_next) level_b(next);
        __syncthreads();                                     // gradients complete; xs dead -> warp staging may be written
        // ---- P2: every warp takes vertices a = warp, warp + WARPS, ...
        double* ws = wstage + (size_t)warp * WS_DOUBLES;
        bool fetched_c = false;
        for (int a = warp; a < TN; a += FUSED_WARPS) {
            const int meta = meta_s[a];
            const int deg = meta & 0xff, val = (meta >> 8) & 0xff;
            // A: lane = 4*k + b/2
            {
                const int k = lane >> 2, bp = lane & 3;
                if (k < val) {
                    const uint32_t itw = items_s[a * FUSED_MAXVAL + k];
                    const int le = itw & 0xfff, na = (itw >> 12) & 0xf;
                    double wga[8][3];
                    {
                        const double2* rec = reinterpret_cast<const double2*>(G + (size_t)le * EL_STRIDE + na * G_STRIDE);
                        double tmp[24];
#pragma unroll
                        for (int j = 0; j < 12; ++j) { const double2 v = rec[j]; tmp[2 * j] = v.x; tmp[2 * j + 1] = v.y; }
                        const double* w = wdet + le * WD_STRIDE;
#pragma unroll
                        for (int q = 0; q < 8; ++q)
#pragma unroll
                            for (int d = 0; d < 3; ++d) wga[q][d] = w[q] * tmp[q * 3 + d];
                    }
                    const double lam = PERQ ? 0.0 : lm[le * 2], mu = PERQ ? 0.0 : lm[le * 2 + 1];
                    const double* rl = rlm + le * WD_STRIDE;
#pragma unroll
                    for (int bi = 0; bi < 2; ++bi) {
                        const int bn = 2 * bp + bi;
                        const double2* rec = reinterpret_cast<const double2*>(G + (size_t)le * EL_STRIDE + bn * G_STRIDE);
                        double* dst = ws + (k * 8 + bn) * 9 + (k >> 1);
                        if (PERQ) {
                            double Sl[9], Sm[9];
#pragma unroll
                            for (int c = 0; c < 9; ++c) { Sl[c] = 0.0; Sm[c] = 0.0; }
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const double2 v0 = rec[3 * j], v1 = rec[3 * j + 1], v2 = rec[3 * j + 2];
                                const double gb0[3] = {v0.x, v0.y, v1.x}, gb1[3] = {v1.y, v2.x, v2.y};
#pragma unroll
                                for (int i = 0; i < 3; ++i) {
                                    const double m0 = wga[2 * j][i], m1 = wga[2 * j + 1][i];
                                    const double l0 = rl[2 * j] * m0, l1 = rl[2 * j + 1] * m1;
#pragma unroll
                                    for (int l = 0; l < 3; ++l) {
                                        Sm[i * 3 + l] += m0 * gb0[l];
                                        Sm[i * 3 + l] += m1 * gb1[l];
                                        Sl[i * 3 + l] += l0 * gb0[l];
                                        Sl[i * 3 + l] += l1 * gb1[l];
                                    }
                                }
                            }
                            const double tr = Sm[0] + Sm[4] + Sm[8];
#pragma unroll
                            for (int d = 0; d < 3; ++d)
#pragma unroll
                                for (int kk = 0; kk < 3; ++kk) dst[d * 3 + kk] = Sl[d * 3 + kk] + Sm[kk * 3 + d] + (d == kk ? tr : 0.0);
                        } else {
                        double S[9];
#pragma unroll
                        for (int c = 0; c < 9; ++c) S[c] = 0.0;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {          // 3 double2 = 6 doubles = 2 quadrature points
                            const double2 v0 = rec[3 * j], v1 = rec[3 * j + 1], v2 = rec[3 * j + 2];
                            const double gb0[3] = {v0.x, v0.y, v1.x}, gb1[3] = {v1.y, v2.x, v2.y};
#pragma unroll
                            for (int i = 0; i < 3; ++i)
#pragma unroll
                                for (int l = 0; l < 3; ++l) {
                                    S[i * 3 + l] += wga[2 * j][i] * gb0[l];
                                    S[i * 3 + l] += wga[2 * j + 1][i] * gb1[l];
                                }
                        }
                        if (UNIFORM) {
#pragma unroll
                            for (int c = 0; c < 9; ++c) dst[c] = S[c];
                        } else {
                            const double tr = mu * (S[0] + S[4] + S[8]);
#pragma unroll
                            for (int d = 0; d < 3; ++d)
#pragma unroll
                                for (int kk = 0; kk < 3; ++kk)
                                    dst[d * 3 + kk] = lam * S[d * 3 + kk] + mu * S[kk * 3 + d] + (d == kk ? tr : 0.0);
                        }
                        }
                    }
                }
            }
            __syncwarp();
            // B: lane = neighbour rank
            if (lane < deg) {
                const int r = lane;
                const int64_t node = node_s[a], p0 = p0_s[a];
                const int row_len = 3 * deg;
                double acc[9];
#pragma unroll
                for (int c = 0; c < 9; ++c) acc[c] = 0.0;
                unsigned long long src = tbl[a * FUSED_MAXDEG + r];
#pragma unroll
                for (int k = 0; k < FUSED_MAXVAL; ++k) {
                    const unsigned int bsel = (unsigned int)(src & 0xffull);
                    src >>= 8;
                    if (bsel != 0xffu) {
                        const double* sp = ws + (k * 8 + bsel) * 9 + (k >> 1);
#pragma unroll
                        for (int c = 0; c < 9; ++c) acc[c] += sp[c];
                    }
                }
                if (UNIFORM) {
                    const double lam = lm[0], mu = lm[1];
                    double S[9];
#pragma unroll
                    for (int c = 0; c < 9; ++c) S[c] = acc[c];
                    const double tr = mu * (S[0] + S[4] + S[8]);
#pragma unroll
                    for (int d = 0; d < 3; ++d)
#pragma unroll
                        for (int kk = 0; kk < 3; ++kk)
                            acc[d * 3 + kk] = lam * S[d * 3 + kk] + mu * S[kk * 3 + d] + (d == kk ? tr : 0.0);
                }
                if (ELIM) {
                    const int32_t colv = nbr_s[a * FUSED_MAXDEG + r];
                    const unsigned int cf = flg_s[a * FUSED_MAXDEG + r];
                    const unsigned int rf = ((unsigned int)meta >> 16) & 7u;
                    if (cf | rf) {
#pragma unroll
                        for (int i = 0; i < 3; ++i)
#pragma unroll
                            for (int j = 0; j < 3; ++j)
                                if (((rf >> i) & 1u) || ((cf >> j) & 1u))
                                    acc[i * 3 + j] = (((rf >> i) & 1u) && colv == node && i == j) ? 1.0 : 0.0;
                    }
                }
                double* dst = values + 9 * p0 + 3 * r;
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) __stcs(dst + (int64_t)i * row_len + j, acc[i * 3 + j]);
            }
            __syncwarp();
            if (has_next && !fetched_c) { level_c(); fetched_c = true; }
        }
        if (has_next && !fetched_c) level_c();
    }
}

}  // namespace feb

using namespace feb;

extern "C" {

size_t feb200_fused_smem_bytes(int32_t tn_cap, int32_t te_cap) { return fused_layout(tn_cap, te_cap).total; }

int feb200_assemble_fused_hex8(feb200_stream s, const feb200_batch* bp, const feb200_fused_plan* plan, const double* x_vd,
                               const uint64_t* row_ind, const uint16_t* slot_rank, const int64_t* nindptr,
                               const int32_t* nindices, const uint8_t* vertex_flags, double* values) {
    FEB_REQUIRE(bp && plan && x_vd && row_ind && slot_rank && nindptr && nindices && values, FEB200_ERR_INVALID,
                "assemble_fused_hex8: null pointer");
    const feb200_batch& b = *bp;
    FEB_REQUIRE(b.N == 8 && b.Q == 8 && b.D == 3, FEB200_ERR_UNSUPPORTED, "assemble_fused_hex8: needs hex8 with 8 quadrature points");
    FEB_REQUIRE(b.kind == FEB200_KIND_LINEAR && b.material == FEB200_MAT_ISOTROPIC, FEB200_ERR_UNSUPPORTED,
                "assemble_fused_hex8: needs the linear isotropic material");
    FEB_REQUIRE(b.conn && b.dphi && b.W && b.mp && b.M >= 2, FEB200_ERR_INVALID, "assemble_fused_hex8: bad batch");
    FEB_REQUIRE(plan->tile_node_ptr && plan->tile_nodes && plan->tile_elem_ptr && plan->tile_elems && plan->tile_item_ptr &&
                    plan->tile_items && plan->tile_node_p0 && plan->tile_node_meta && plan->n_tiles >= 0,
                FEB200_ERR_INVALID, "assemble_fused_hex8: bad plan");
    FEB_REQUIRE(plan->tn_cap > 0 && plan->tn_cap % 4 == 0 && plan->tn_cap <= 32 && plan->te_cap > 0 &&
                    plan->te_cap * 8 <= PF * FUSED_THREADS && plan->tn_cap * FUSED_MAXVAL <= FUSED_THREADS,
                FEB200_ERR_INVALID, "assemble_fused_hex8: plan capacities tn=%d te=%d out of range", plan->tn_cap, plan->te_cap);
    if (plan->n_tiles == 0) return FEB200_OK;
    const FusedSmem L = fused_layout(plan->tn_cap, plan->te_cap);
    FEB_REQUIRE(L.total <= 227 * 1024, FEB200_ERR_CAPACITY, "assemble_fused_hex8: tile needs %zu bytes of shared memory", L.total);
    cudaStream_t st = as_stream(s);
    (void)row_ind;
    (void)nindptr;
    int grid = plan->n_tiles < sm_count() * FEB_FUSED_MINBLOCKS ? plan->n_tiles : sm_count() * FEB_FUSED_MINBLOCKS;
    const bool uniform = (b.layout == FEB200_LAYOUT_M), perq = (b.layout == FEB200_LAYOUT_EQM);
#define FEB_FUSED_LAUNCH(EL, UN)                                                                                       \
    if (perq) {                                                                                                        \
        auto kern = k_fused_hex8_iso<EL, false, true>;                                                                 \
        FEB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));               \
        kern<<<grid, FUSED_THREADS, L.total, st>>>(b, *plan, L, x_vd, slot_rank, nindices, vertex_flags, values);      \
    } else {                                                                                                           \
        auto kern = k_fused_hex8_iso<EL, UN, false>;                                                                   \
        FEB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));               \
        kern<<<grid, FUSED_THREADS, L.total, st>>>(b, *plan, L, x_vd, slot_rank, nindices, vertex_flags, values);      \
    }
    if (vertex_flags) {
        if (uniform) FEB_FUSED_LAUNCH(true, true) else FEB_FUSED_LAUNCH(true, false)
    } else {
        if (uniform) FEB_FUSED_LAUNCH(false, true) else FEB_FUSED_LAUNCH(false, false)
    }
#undef FEB_FUSED_LAUNCH
    FEB_CHECK_LAUNCH("assemble_fused_hex8");
    return FEB200_OK;
}

}  // extern "C"
