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
constexpr int FUSED_MAXVAL = 8;          // threads per vertex = max incident slots handled
constexpr int FUSED_MAXDEG = 32;         // neighbour blocks per vertex (one lane each; hex8: 27)
constexpr int FUSED_THREADS = 256;
constexpr int G_STRIDE = 26;             // doubles per (element, node) gradient record (24 used): 16-byte aligned
constexpr int EL_STRIDE = 8 * G_STRIDE + 2;   // doubles per element in G: records of consecutive elements are 4 banks apart
constexpr int XS_STRIDE = 25;            // doubles per element coordinate record (24 used)
constexpr int DPHI_STRIDE = 25;          // doubles per quadrature point in the basis table (24 used)
constexpr int BLK_STRIDE = 73;           // doubles per item block row in the staging area (72 used, odd: conflict-free)
constexpr int PF = 3;                    // element-node prefetch registers per thread: te_cap*8 <= PF*256 (te_cap <= 96)
constexpr int PN = FUSED_MAXDEG * 32 / FUSED_THREADS;   // neighbour-table entries per thread (tn_cap <= 32)

struct FusedSmem {
    int tn_cap, te_cap;
    size_t off_stage, off_xs, off_dphi, off_lm, off_ints, total;
};
static inline FusedSmem fused_layout(int tn_cap, int te_cap) {
    FusedSmem L;
    L.tn_cap = tn_cap;
    L.te_cap = te_cap;
    size_t g_doubles = (size_t)te_cap * EL_STRIDE + (size_t)te_cap * FUSED_Q;      // G + wdet
    size_t st_doubles = (size_t)FUSED_THREADS * BLK_STRIDE;
    size_t a = g_doubles > st_doubles ? g_doubles : st_doubles;
    a = (a + 1) & ~(size_t)1;
    L.off_stage = 0;
    L.off_xs = a * 8;
    L.off_dphi = L.off_xs + (size_t)te_cap * XS_STRIDE * 8;
    L.off_lm = L.off_dphi + (size_t)FUSED_Q * DPHI_STRIDE * 8;
    L.off_ints = L.off_lm + (size_t)te_cap * 2 * 8;
    L.off_ints = (L.off_ints + 15) & ~(size_t)15;
    // tbl[tn*32][8] u16 (source table, 16 B rows) | p0[tn] i64 | nbr[tn*32] i32 | node[tn], meta[tn] i32 | flg[tn*32] u8
    L.total = L.off_ints + (size_t)tn_cap * FUSED_MAXDEG * 16 + (size_t)tn_cap * 8 + (size_t)tn_cap * FUSED_MAXDEG * 4 +
              (size_t)tn_cap * 4 * 2 + (size_t)tn_cap * FUSED_MAXDEG + 64;
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
    int32_t node, meta;      // A   (meta = degree | valence << 8), valid for tid < TN
    int64_t p0;              // A
    int64_t p0n[PN];         // A   CSR offset of the vertex owning neighbour-table entry tid + j*256
    int32_t degn[PN];        // A
    int32_t vtx[PF];         // B
    int32_t nbr[PN];         // B
    int32_t item_el;         // B
    double x[PF][3];         // C
    uint32_t flg;            // C   PN flags packed
    uint4 rk8;               // C
};

template <bool ELIM>
__global__ void __launch_bounds__(FUSED_THREADS, 1)
k_fused_hex8_iso(feb200_batch b, feb200_fused_plan plan, FusedSmem L, const double* __restrict__ x_vd,
                 const uint16_t* __restrict__ slot_rank, const int32_t* __restrict__ nindices,
                 const uint8_t* __restrict__ vflags, double* __restrict__ values) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* G = reinterpret_cast<double*>(smem_raw + L.off_stage);          // [le] EL_STRIDE : [n] G_STRIDE : g[q][3]
    double* wdet = G + (size_t)L.te_cap * EL_STRIDE;                        // [le][q]
    double* stage = reinterpret_cast<double*>(smem_raw + L.off_stage);      // [item] BLK_STRIDE  (aliases G)
    double* xs = reinterpret_cast<double*>(smem_raw + L.off_xs);            // [le] XS_STRIDE : [n][3]
    double* s_dphi = reinterpret_cast<double*>(smem_raw + L.off_dphi);      // [q] DPHI_STRIDE : [n][3]
    double* lm = reinterpret_cast<double*>(smem_raw + L.off_lm);            // [le][2] lambda, mu
    uint16_t* tbl = reinterpret_cast<uint16_t*>(smem_raw + L.off_ints);     // [a][r][k] = item << 3 | b, or 0xffff
    int64_t* p0_s = reinterpret_cast<int64_t*>(tbl + (size_t)L.tn_cap * FUSED_MAXDEG * 8);   // [a]
    int32_t* nbr_s = reinterpret_cast<int32_t*>(p0_s + L.tn_cap);           // [a][r] neighbour vertex
    int32_t* node_s = nbr_s + (size_t)L.tn_cap * FUSED_MAXDEG;              // [a]
    int32_t* meta_s = node_s + L.tn_cap;                                    // [a] degree | valence << 8
    uint8_t* flg_s = reinterpret_cast<uint8_t*>(meta_s + L.tn_cap);         // [a][r] constrained components of the neighbour

    const int tid = threadIdx.x;
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
        for (int j = 0; j < PF; ++j)
            pf.vtx[j] = (pf.el[j] >= 0) ? b.conn[(int64_t)pf.el[j] * 8 + ((tid + j * FUSED_THREADS) & 7)] : -1;
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
                if (n == 0) {
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
                uint4 ff;
                ff.x = ff.y = ff.z = ff.w = 0xffffffffu;
                *reinterpret_cast<uint4*>(tbl + (size_t)i * 8) = ff;
            }
        }
        // the item of this thread for this tile (registers), then start fetching the next tile
        const uint32_t item = pf.item;
        const uint4 rk8 = pf.rk8;
        if (has_next) level_a(next);
        __syncthreads();
        // ---- P1: geometry, one thread per (element, quadrature point)
        for (int i = tid; i < TE * 8; i += FUSED_THREADS) {
            const int le = i >> 3, q = i & 7;
            double g[24];
            const double dt = shape_gradients<8, 3>(xs + le * XS_STRIDE, s_dphi + q * DPHI_STRIDE, g);
            wdet[i] = dt * b.W[q];
#pragma unroll
            for (int n = 0; n < 8; ++n)
#pragma unroll
                for (int d = 0; d < 3; ++d) G[(size_t)le * EL_STRIDE + n * G_STRIDE + q * 3 + d] = g[n * 3 + d];
        }
        __syncthreads();
        // ---- P2: block rows, one thread per (vertex, incident slot); items are sorted by element so that the
        //          lanes of a warp read few distinct gradient records (broadcast)
        double blk[8][9];
        const bool has_item = item != 0xffffffffu;
        if (has_item) {
            const int le = item & 0xfff, na = (item >> 12) & 0xf;
            const int a_loc = (item >> 16) & 0xff, kk = (item >> 24) & 0xff;
            // source table: tbl[a][rank of local node b][k] = (this item << 3) | b
            const unsigned int w4[4] = {rk8.x, rk8.y, rk8.z, rk8.w};
            uint16_t* tb = tbl + (size_t)a_loc * FUSED_MAXDEG * 8 + kk;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                tb[(w4[j] & 0xffffu) * 8] = (uint16_t)((tid << 3) | (2 * j));
                tb[(w4[j] >> 16) * 8] = (uint16_t)((tid << 3) | (2 * j + 1));
            }
            double wga[8][3];
            {
                const double2* rec = reinterpret_cast<const double2*>(G + (size_t)le * EL_STRIDE + na * G_STRIDE);
                double tmp[24];
#pragma unroll
                for (int j = 0; j < 12; ++j) { const double2 v = rec[j]; tmp[2 * j] = v.x; tmp[2 * j + 1] = v.y; }
                const double* w = wdet + le * 8;
#pragma unroll
                for (int q = 0; q < 8; ++q)
#pragma unroll
                    for (int d = 0; d < 3; ++d) wga[q][d] = w[q] * tmp[q * 3 + d];
            }
            const double lam = lm[le * 2], mu = lm[le * 2 + 1];
#pragma unroll
            for (int bn = 0; bn < 8; ++bn) {
                const double2* rec = reinterpret_cast<const double2*>(G + (size_t)le * EL_STRIDE + bn * G_STRIDE);
                double S[9];
#pragma unroll
                for (int c = 0; c < 9; ++c) S[c] = 0.0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {                  // 3 double2 = 6 doubles = 2 quadrature points
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
                const double tr = mu * (S[0] + S[4] + S[8]);
#pragma unroll
                for (int d = 0; d < 3; ++d)
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        blk[bn][d * 3 + k] = lam * S[d * 3 + k] + mu * S[k * 3 + d] + (d == k ? tr : 0.0);
            }
        }
        if (has_next) level_b(next);
        __syncthreads();                                     // all gradient reads done: staging may overwrite them
        // ---- P3: block rows to shared memory, indexed by item (conflict-free)
        if (has_item) {
            double* dst = stage + (size_t)tid * BLK_STRIDE;
#pragma unroll
            for (int bn = 0; bn < 8; ++bn)
#pragma unroll
                for (int c = 0; c < 9; ++c) dst[bn * 9 + c] = blk[bn][c];
        }
        if (has_next) level_c();
        __syncthreads();
        // ---- P4: one warp per vertex, one lane per neighbour block; contributions added in ascending slot order
        const int lane = tid & 31, warp = tid >> 5;
        for (int a = warp; a < TN; a += FUSED_THREADS / 32) {
            const int deg = meta_s[a] & 0xff;
            if (lane < deg) {
                const int r = lane;
                const int64_t node = node_s[a], p0 = p0_s[a];
                const int row_len = 3 * deg;
                double acc[9];
#pragma unroll
                for (int c = 0; c < 9; ++c) acc[c] = 0.0;
                const uint4 row = *reinterpret_cast<const uint4*>(tbl + (size_t)(a * FUSED_MAXDEG + r) * 8);
                const unsigned int w4[4] = {row.x, row.y, row.z, row.w};
#pragma unroll
                for (int k = 0; k < FUSED_MAXVAL; ++k) {
                    const unsigned int ent = (k & 1) ? (w4[k >> 1] >> 16) : (w4[k >> 1] & 0xffffu);
                    if (ent != 0xffffu) {
                        const double* sp = stage + (size_t)(ent >> 3) * BLK_STRIDE + (ent & 7u) * 9;
#pragma unroll
                        for (int c = 0; c < 9; ++c) acc[c] += sp[c];
                    }
                }
                if (ELIM) {
                    const int32_t colv = nbr_s[a * FUSED_MAXDEG + r];
                    const unsigned int cf = flg_s[a * FUSED_MAXDEG + r];
                    const unsigned int rf = ((unsigned int)meta_s[a] >> 16) & 7u;
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
        }
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
    FEB_REQUIRE(b.kind == FEB200_KIND_LINEAR && b.material == FEB200_MAT_ISOTROPIC && b.layout != FEB200_LAYOUT_EQM,
                FEB200_ERR_UNSUPPORTED, "assemble_fused_hex8: needs linear isotropic material, constant per element");
    FEB_REQUIRE(b.conn && b.dphi && b.W && b.mp && b.M >= 2, FEB200_ERR_INVALID, "assemble_fused_hex8: bad batch");
    FEB_REQUIRE(plan->tile_node_ptr && plan->tile_nodes && plan->tile_elem_ptr && plan->tile_elems && plan->tile_item_ptr &&
                    plan->tile_items && plan->tile_node_p0 && plan->tile_node_meta && plan->n_tiles >= 0,
                FEB200_ERR_INVALID, "assemble_fused_hex8: bad plan");
    FEB_REQUIRE(plan->tn_cap > 0 && plan->tn_cap % 4 == 0 && plan->tn_cap <= 32 && plan->te_cap > 0 &&
                    plan->te_cap * 8 <= PF * 256,
                FEB200_ERR_INVALID, "assemble_fused_hex8: plan capacities tn=%d te=%d out of range", plan->tn_cap, plan->te_cap);
    if (plan->n_tiles == 0) return FEB200_OK;
    const FusedSmem L = fused_layout(plan->tn_cap, plan->te_cap);
    FEB_REQUIRE(L.total <= 227 * 1024, FEB200_ERR_CAPACITY, "assemble_fused_hex8: tile needs %zu bytes of shared memory", L.total);
    cudaStream_t st = as_stream(s);
    (void)row_ind;
    (void)nindptr;
    int grid = plan->n_tiles < sm_count() ? plan->n_tiles : sm_count();
    if (vertex_flags) {
        auto kern = k_fused_hex8_iso<true>;
        FEB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
        kern<<<grid, FUSED_THREADS, L.total, st>>>(b, *plan, L, x_vd, slot_rank, nindices, vertex_flags, values);
    } else {
        auto kern = k_fused_hex8_iso<false>;
        FEB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
        kern<<<grid, FUSED_THREADS, L.total, st>>>(b, *plan, L, x_vd, slot_rank, nindices, vertex_flags, values);
    }
    FEB_CHECK_LAUNCH("assemble_fused_hex8");
    return FEB200_OK;
}

}  // extern "C"
