// Fused hex8 assembly, tensor-core variant: the element blocks S_ab = sum_q w_q g_a,q (x) g_b,q of one
// element are one 24 x 24 x 8 FP64 product  S^e = (w G)(G)^T,  G[(n,d)][q] = d phi_n / d x_d at point q,
// evaluated with mma.sync.m8n8k4.f64 (DMMA).  DMMA and DFMA share the FP64 pipe on B200 (same peak, see
// profiles/r01_fp64_dmma_peak.txt); what the tensor-core form buys is operand traffic: every gradient is
// read from shared memory once per element visit instead of once per 3x3 block, which is what bounded the
// scalar kernel of fused.cu (shared-memory wavefronts, profiles/r01_fused_v4_ncu_summary.txt).
//
// Same tile plan and same contract as feb200_assemble_fused_hex8 (fused.cu; reference path
// fea.py:519-572 + fea.py:613-660 + sparse_matrix.py:157-184).  One persistent CTA per SM, 8 warps; per tile
// of <= 32 vertices (shared memory: 144 KB staged block rows + 53 KB fragment buffers + coordinates + tables):
//   P0   data prefetched into registers one tile ahead (three dependency levels) -> shared memory; the
//        (vertex, rank) -> source table is built with integer and-masks
//   A    warps pull steps of 4 elements from a tile-local counter (the last round in steps of 2, so the
//        round not every warp takes part in is short): lane = (element, quadrature point) computes the
//        gradients into a warp-private buffer laid out as DMMA fragments; then per element only the row
//        tiles that hold vertices of this tile are multiplied (1-3 row tiles x 3 n-tiles x 2 k-steps) and the
//        rows go to the item's staging record  stage[item][i][b*3+j]
//   B    lane = CSR block rank r, two vertices per warp at a time: adds the <= 8 staged blocks in ascending
//        slot order (absent ones are predicated-off loads), applies K = lam S + mu S^T + mu tr(S) I and the
//        Dirichlet elimination, stores the panel rows (streaming stores).
// No floating-point atomics; the summation order is fixed by the plan, so results are bit-reproducible.
// The work counter only decides which warp computes which element group, never the order of a sum.
// Measured alternatives (profiles/r01_fused_mma_ncu_summary.txt): 10/12/16 warps with 2-element
// steps, two CTAs per SM on 16-vertex tiles, panel stores transposed through shared memory, list-walking
// or slot-pairing B phases -- all slower than this configuration.
#include "common.cuh"
#include "element_math.cuh"
#include "hex8_mma.cuh"

#ifndef FEB_MM_LEVELC_AFTER
#define FEB_MM_LEVELC_AFTER 1        /* B phase: issue the coordinate prefetch after this many vertices of the warp */
#endif
#ifndef FEB_MM_WARPS
#define FEB_MM_WARPS 8
#endif
#ifndef FEB_MM_GROUP
#define FEB_MM_GROUP 4               /* elements per warp step: 4 = lane (element, point); 2 = lane (element, point, node half) */
#endif
#ifndef FEB_MM_MINBLOCKS
#define FEB_MM_MINBLOCKS 1           /* CTAs per SM (2 needs 16-vertex tiles: FEB200_FUSED_TN=16) */
#endif
#ifndef FEB_MM_TAIL2
#define FEB_MM_TAIL2 1               /* finish phase A with 2-element steps */
#endif
#ifndef FEB_MM_BV
#define FEB_MM_BV 2                  /* vertices a warp of phase B works on at a time */
#endif

namespace feb {

constexpr int MM_BV = FEB_MM_BV;
constexpr int MM_WARPS = FEB_MM_WARPS;
constexpr int MM_THREADS = MM_WARPS * 32;
constexpr int MM_MAXVAL = 8;             // incident slots per vertex
constexpr int MM_MAXDEG = 32;            // CSR blocks per vertex row (lane each)
constexpr int MM_GROUP = FEB_MM_GROUP;     // elements per warp step
constexpr int MM_KS = 104;               // doubles per k-step half of an element's fragment buffer: [row 0..23][4 points] + 8 (weights)
#ifndef FEB_MM_EL
#define FEB_MM_EL 212
#endif
constexpr int MM_EL = FEB_MM_EL;               // doubles per element in the fragment buffer (2 * KS + 4): KS = 8, EL = 4 (mod 16) keeps the
                                         // geometry writes (lane = element, point) on distinct banks
constexpr int MM_IS = 72;                // doubles per staged item: [i][b*3+j]
constexpr int MM_XS = 26;                // doubles per element coordinate record: [n][3], mu, lambda/mu
constexpr int MM_DPHI = 26;              // doubles per quadrature point in the basis table (24 used, 16-byte rows)
constexpr int MM_PF = (96 * 8 + MM_THREADS - 1) / MM_THREADS;   // te_cap * 8 <= PF * THREADS   (te_cap <= 96)
constexpr int MM_PN = (MM_MAXDEG * 32 + MM_THREADS - 1) / MM_THREADS;

struct MmaSmem {
    int tn_cap, te_cap;
    size_t off_stage, off_gw, off_xs, off_eq, off_dphi, off_tbl, off_p0, off_itw, off_meta, off_e0, off_e1, off_itof, off_flg,
        off_ctr, total;
};
static inline MmaSmem mma_layout(int tn_cap, int te_cap, bool perq) {
    MmaSmem L;
    L.tn_cap = tn_cap;
    L.te_cap = te_cap;
    size_t o = 0;
    L.off_stage = o; o += (size_t)tn_cap * MM_MAXVAL * MM_IS * 8;
    L.off_gw = o;    o += (size_t)MM_WARPS * MM_GROUP * MM_EL * 8;
    L.off_xs = o;    o += (size_t)te_cap * MM_XS * 8;
    L.off_eq = o;    o += perq ? (size_t)te_cap * 8 * 8 : 0;              // [le][q] Young's modulus per point (layout (E,Q,M))
    L.off_dphi = o;  o += (size_t)(8 * MM_DPHI + 8) * 8;                  // basis table + quadrature weights
    L.off_p0 = o;    o += (size_t)tn_cap * 8;
    L.off_tbl = o;   o += (size_t)tn_cap * MM_MAXDEG * 4;                 // [a][r] nibble k = local node b or 0xf
    L.off_itw = o;   o += (size_t)tn_cap * MM_MAXVAL * 4;                 // item words, plan order
    L.off_meta = o;  o += (size_t)tn_cap * 4;
    L.off_ctr = o;   o += 16;
    L.off_e0 = o;    o += (size_t)te_cap * 2;                             // first item of the element
    o = (o + 3) & ~(size_t)3;
    L.off_e1 = o;    o += (size_t)te_cap * 2;                             // one past its last item
    o = (o + 3) & ~(size_t)3;
    o = (o + 7) & ~(size_t)7;
    L.off_itof = o;  o += (size_t)tn_cap * MM_MAXVAL;                     // [a][k] -> item index
    L.off_flg = o;   o += (size_t)tn_cap * MM_MAXDEG;                     // [a][r] constrained components of the neighbour | diag << 3
    L.total = (o + 15) & ~(size_t)15;
    return L;
}

// Everything a tile reads from global memory is fetched one tile ahead into registers, in three dependency
// levels (A: plan arrays, B: ids that need A, C: data that need B).  Loaded values are only *used* in P0 of the
// tile they belong to (or as addresses of the next level, long after they were issued): a use right after the
// issue, or a use of an old load while newer loads share its scoreboard, stalls for a full memory latency.
struct MmaPrefetch {
    int32_t el[MM_PF];       // A   element of (tid + j*THREADS) >> 3
    int32_t el_own;          // A   element tid (material parameters)
    uint32_t item, item_prev, item_next;   // A
    int32_t node, meta;      // A
    int64_t p0;              // A
    int64_t p0n[MM_PN];      // A
    int32_t degn[MM_PN];     // A
    int32_t noden[MM_PN];    // A
    int32_t vtx[MM_PF];      // B
    int32_t nbr[MM_PN];      // B
    int32_t item_el;         // B
    uint32_t rowflag;        // B
    double mp[2];            // B   (E, nu) of element tid, layouts with per-element material
    double eq[MM_PF];        // B   per-point layout: E of point (tid + j*THREADS) & 7 of element el[j]
    double x[MM_PF][3];      // C
    uint32_t flgb[MM_PN];    // C
    uint4 rk8;               // C
};

// The same for the four nodes 4h .. 4h+3 only (lane = (element, point, node half) when a warp step covers two elements).
__device__ __forceinline__ double hex8_gradients_half(const double* __restrict__ xe, const double* __restrict__ dq, int h,
                                                      double (&g)[12]) {
    const double2* x2 = reinterpret_cast<const double2*>(xe);
    const double2* d2 = reinterpret_cast<const double2*>(dq);
    double J[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
#pragma unroll
    for (int hh = 0; hh < 4; ++hh) {
        const double2 xa = x2[3 * hh], xb = x2[3 * hh + 1], xc = x2[3 * hh + 2];
        const double2 da = d2[3 * hh], db = d2[3 * hh + 1], dc = d2[3 * hh + 2];
        const double x[6] = {xa.x, xa.y, xb.x, xb.y, xc.x, xc.y};
        const double d[6] = {da.x, da.y, db.x, db.y, dc.x, dc.y};
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int pp = 0; pp < 3; ++pp)
#pragma unroll
                for (int dd = 0; dd < 3; ++dd) J[pp][dd] += x[3 * m + dd] * d[3 * m + pp];
    }
    double Ji[3][3];
    const double dt = inverse(J, Ji);
    const double2* dh = d2 + 6 * h;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        const double2 da = dh[3 * hh], db = dh[3 * hh + 1], dc = dh[3 * hh + 2];
        const double d[6] = {da.x, da.y, db.x, db.y, dc.x, dc.y};
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int dd = 0; dd < 3; ++dd) {
                double sacc = 0.0;
#pragma unroll
                for (int pp = 0; pp < 3; ++pp) sacc += Ji[dd][pp] * d[3 * m + pp];
                g[(2 * hh + m) * 3 + dd] = sacc;
            }
    }
    return dt;
}

// Row tiles 0..NMT-1 of one element: S rows of the in-tile vertices (slot s = row / 3, component i = row % 3).
template <int NMT>
__device__ __forceinline__ void element_rows(const double* __restrict__ Ge, const double (&Bf)[3][2], double w0, double w1,
                                             const uint32_t* __restrict__ itw_e, int cnt, double* __restrict__ stage_e,
                                             int gid, int tig) {
    double A[NMT][2];
    int off[NMT];
#pragma unroll
    for (int mt = 0; mt < NMT; ++mt) {
        const int R = 8 * mt + gid, s = R / 3, i = R - 3 * s;
        const bool valid = s < cnt;
        const int na = valid ? (int)((itw_e[s] >> 12) & 7u) : 0;
        const double g0 = Ge[(na * 3 + i) * 4 + tig], g1 = Ge[MM_KS + (na * 3 + i) * 4 + tig];
        A[mt][0] = valid ? w0 * g0 : 0.0;
        A[mt][1] = valid ? w1 * g1 : 0.0;
        off[mt] = valid ? s * MM_IS + i * 24 + 2 * tig : -1;
    }
    double C[NMT][3][2];
#pragma unroll
    for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
            C[mt][nt][0] = 0.0;
            C[mt][nt][1] = 0.0;
            dmma884(C[mt][nt][0], C[mt][nt][1], A[mt][0], Bf[nt][0]);
        }
#pragma unroll
    for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) dmma884(C[mt][nt][0], C[mt][nt][1], A[mt][1], Bf[nt][1]);
#pragma unroll
    for (int mt = 0; mt < NMT; ++mt)
        if (off[mt] >= 0) {
#pragma unroll
            for (int nt = 0; nt < 3; ++nt)
                *reinterpret_cast<double2*>(stage_e + off[mt] + 8 * nt) = make_double2(C[mt][nt][0], C[mt][nt][1]);
        }
}

// WM, how the material enters the quadrature weight of the product:
//   0  not at all: one (lambda, mu) for the batch, applied in phase B                          layout (M,)
//   1  w_q * mu_e                                                                               layout (E,M)
//   2  w_q * E_q / (2 (1 + nu_e)): per-point Young's modulus, Poisson ratio per element        layout (E,Q,M), checked by the host
// For WM > 0 phase B needs rho = lambda / mu = 2 nu / (1 - 2 nu): K = sum rho_k S'_k + (sum S'_k)^T + tr(sum S'_k) I.
// UNIFORM: rho is one number for the batch (WM = 0, or the caller's hint that every element has the same Poisson
// ratio), so the blocks are summed first and combined once; otherwise rho is applied per contribution.
template <bool ELIM, bool UNIFORM, int WM>
__global__ void __launch_bounds__(MM_THREADS, FEB_MM_MINBLOCKS)
k_fused_hex8_mma(feb200_batch b, feb200_fused_plan plan, MmaSmem L, const double* __restrict__ x_vd,
                 const uint16_t* __restrict__ slot_rank, const int32_t* __restrict__ nindices,
                 const uint8_t* __restrict__ vflags, double* __restrict__ values) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* stage = reinterpret_cast<double*>(smem_raw + L.off_stage);       // [item] IS : [i][b*3+j]
    double* gw = reinterpret_cast<double*>(smem_raw + L.off_gw);             // [warp][4] EL : [ks] KS : [row][4 points], weights at 96
    double* xs = reinterpret_cast<double*>(smem_raw + L.off_xs);             // [le] XS
    double* eq_s = reinterpret_cast<double*>(smem_raw + L.off_eq);           // [le][q]  (WM == 2)
    double* s_dphi = reinterpret_cast<double*>(smem_raw + L.off_dphi);       // [q] DPHI : [n][3]
    double* s_w = s_dphi + 8 * MM_DPHI;                                      // [q]
    int64_t* p0_s = reinterpret_cast<int64_t*>(smem_raw + L.off_p0);
    unsigned int* tbl = reinterpret_cast<unsigned int*>(smem_raw + L.off_tbl);
    uint32_t* itw_s = reinterpret_cast<uint32_t*>(smem_raw + L.off_itw);
    int32_t* meta_s = reinterpret_cast<int32_t*>(smem_raw + L.off_meta);
    int* ctr = reinterpret_cast<int*>(smem_raw + L.off_ctr);
    uint16_t* e0_s = reinterpret_cast<uint16_t*>(smem_raw + L.off_e0);
    uint16_t* e1_s = reinterpret_cast<uint16_t*>(smem_raw + L.off_e1);
    uint8_t* itof_s = smem_raw + L.off_itof;
    uint8_t* flg_s = smem_raw + L.off_flg;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    for (int i = tid; i < 8 * 24; i += MM_THREADS) s_dphi[(i / 24) * MM_DPHI + (i % 24)] = b.dphi[i];
    if (tid < 8) s_w[tid] = b.W[tid];
    for (int i = tid; i < L.tn_cap * MM_MAXDEG; i += MM_THREADS) tbl[i] = 0xffffffffu;
    double u_lam = 0.0, u_mu = 0.0;
    if (UNIFORM) {
        isotropic_lame<3>(b.mp, u_lam, u_mu);
        if (WM > 0) { u_lam = u_lam / u_mu; u_mu = 1.0; }                     // mu is already in the weights: K = rho S' + S'^T + tr(S') I
    }

    struct TilePtr { int n0, n1, e0, e1, i0, i1; };
    auto load_ptr = [&](int tile) {
        TilePtr t;
        t.n0 = plan.tile_node_ptr[tile]; t.n1 = plan.tile_node_ptr[tile + 1];
        t.e0 = plan.tile_elem_ptr[tile]; t.e1 = plan.tile_elem_ptr[tile + 1];
        t.i0 = plan.tile_item_ptr[tile]; t.i1 = plan.tile_item_ptr[tile + 1];
        return t;
    };
    MmaPrefetch pf;
    auto level_a = [&](const TilePtr& tp) {
        const int n0 = tp.n0, TN = tp.n1 - n0;
        const int e0 = tp.e0, TE = tp.e1 - e0;
        const int i0 = tp.i0, NI = tp.i1 - i0;
#pragma unroll
        for (int j = 0; j < MM_PF; ++j) {
            const int i = tid + j * MM_THREADS;
            pf.el[j] = (i < TE * 8) ? plan.tile_elems[e0 + (i >> 3)] : -1;
        }
        pf.el_own = (WM > 0 && tid < TE) ? plan.tile_elems[e0 + tid] : -1;
        pf.item = pf.item_prev = pf.item_next = 0xffffffffu;
        if (tid < NI) {
            pf.item = plan.tile_items[i0 + tid];
            if (tid > 0) pf.item_prev = plan.tile_items[i0 + tid - 1];
            if (tid < NI - 1) pf.item_next = plan.tile_items[i0 + tid + 1];
        }
        pf.node = -1;
        if (tid < TN) {
            pf.node = plan.tile_nodes[n0 + tid];
            pf.meta = plan.tile_node_meta[n0 + tid];
            pf.p0 = plan.tile_node_p0[n0 + tid];
        }
#pragma unroll
        for (int j = 0; j < MM_PN; ++j) {
            const int a = (tid + j * MM_THREADS) / MM_MAXDEG;
            pf.degn[j] = -1;
            if (a < TN) {
                pf.p0n[j] = plan.tile_node_p0[n0 + a];
                pf.degn[j] = plan.tile_node_meta[n0 + a];
                pf.noden[j] = plan.tile_nodes[n0 + a];
            }
        }
    };
    auto level_b = [&](const TilePtr& tp) {
        const int e0 = tp.e0;
#pragma unroll
        for (int j = 0; j < MM_PF; ++j)
            pf.vtx[j] = (pf.el[j] >= 0) ? b.conn[(int64_t)pf.el[j] * 8 + ((tid + j * MM_THREADS) & 7)] : -1;
        if (WM == 2) {
#pragma unroll
            for (int j = 0; j < MM_PF; ++j)
                if (pf.el[j] >= 0) pf.eq[j] = b.mp[((int64_t)pf.el[j] * 8 + ((tid + j * MM_THREADS) & 7)) * b.M];
        }
#pragma unroll
        for (int j = 0; j < MM_PN; ++j) {
            const int r = (tid + j * MM_THREADS) % MM_MAXDEG;
            pf.nbr[j] = (pf.degn[j] >= 0 && r < (pf.degn[j] & 0xff)) ? nindices[pf.p0n[j] + r] : -1;
        }
        pf.item_el = (pf.item != 0xffffffffu) ? plan.tile_elems[e0 + (pf.item & 0xfff)] : -1;
        pf.rowflag = (ELIM && pf.node >= 0) ? (uint32_t)vflags[pf.node] : 0u;
        if (WM > 0 && pf.el_own >= 0) {
            int qs;
            const double* mp = mp_of(b, pf.el_own, qs);
            pf.mp[0] = mp[0];
            pf.mp[1] = mp[1];
        }
    };
    auto level_c = [&]() {
#pragma unroll
        for (int j = 0; j < MM_PF; ++j)
            if (pf.vtx[j] >= 0) {
#pragma unroll
                for (int d = 0; d < 3; ++d) pf.x[j][d] = x_vd[(int64_t)pf.vtx[j] * 3 + d];
            }
#pragma unroll
        for (int j = 0; j < MM_PN; ++j) pf.flgb[j] = (ELIM && pf.nbr[j] >= 0) ? (uint32_t)vflags[pf.nbr[j]] : 0u;
        if (pf.item_el >= 0)
            pf.rk8 = *reinterpret_cast<const uint4*>(slot_rank + ((int64_t)pf.item_el * 8 + ((pf.item >> 12) & 0xf)) * 8);
    };

    int tile = blockIdx.x;
    TilePtr cur = {0, 0, 0, 0, 0, 0}, nxt = {0, 0, 0, 0, 0, 0}, nxt2 = {0, 0, 0, 0, 0, 0};
    if (tile < plan.n_tiles) {
        cur = load_ptr(tile);
        if (tile + (int)gridDim.x < plan.n_tiles) nxt = load_ptr(tile + gridDim.x);
        level_a(cur); level_b(cur); level_c();
    }

    for (; tile < plan.n_tiles; tile += gridDim.x) {
        const int TN = cur.n1 - cur.n0;
        const int TE = cur.e1 - cur.e0;
        const int next = tile + gridDim.x;
        const bool has_next = next < plan.n_tiles;
        if (next + (int)gridDim.x < plan.n_tiles) nxt2 = load_ptr(next + gridDim.x);   // pointers two tiles ahead
        __syncthreads();                                     // previous tile fully consumed (its B phase left tbl cleared)
        // ---- P0: prefetched data -> shared memory
#pragma unroll
        for (int j = 0; j < MM_PF; ++j) {
            const int i = tid + j * MM_THREADS;
            if (i < TE * 8) {
                const int le = i >> 3, n = i & 7;
#pragma unroll
                for (int d = 0; d < 3; ++d) xs[le * MM_XS + n * 3 + d] = pf.x[j][d];
                if (WM == 2) eq_s[i] = pf.eq[j];
            }
        }
        if (WM > 0 && tid < TE) {
            double lam, mu;
            isotropic_lame<3>(pf.mp, lam, mu);
            xs[tid * MM_XS + 24] = (WM == 2) ? 0.5 / (1.0 + pf.mp[1]) : mu;   // WM 2: mu_q = E_q * this
            xs[tid * MM_XS + 25] = lam / mu;
        }
        if (tid < TN) { meta_s[tid] = pf.meta | (int32_t)(pf.rowflag << 16); p0_s[tid] = pf.p0; }
#pragma unroll
        for (int j = 0; j < MM_PN; ++j) {
            const int i = tid + j * MM_THREADS;
            if (i < TN * MM_MAXDEG) flg_s[i] = (uint8_t)((pf.flgb[j] & 7u) | (pf.nbr[j] == pf.noden[j] ? 8u : 0u));
        }
        if (pf.item != 0xffffffffu) {
            const uint32_t item = pf.item, le = item & 0xfffu;
            const int a = (item >> 16) & 0xff, k = (item >> 24) & 0xff;
            itw_s[tid] = item;
            itof_s[a * MM_MAXVAL + k] = (uint8_t)tid;
            if (pf.item_prev == 0xffffffffu || (pf.item_prev & 0xfffu) != le) e0_s[le] = (uint16_t)tid;
            if (pf.item_next == 0xffffffffu || (pf.item_next & 0xfffu) != le) e1_s[le] = (uint16_t)(tid + 1);
            // source table: tbl[a][rank of local node b] nibble k = b   (integer and-masks: order-independent)
            const unsigned int w4[4] = {pf.rk8.x, pf.rk8.y, pf.rk8.z, pf.rk8.w};
            unsigned int* tb = tbl + (size_t)a * MM_MAXDEG;
            const unsigned int sh = 4u * k;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                atomicAnd(tb + (w4[j] & 0xffffu), ~(0xfu << sh) | ((unsigned int)(2 * j) << sh));
                atomicAnd(tb + (w4[j] >> 16), ~(0xfu << sh) | ((unsigned int)(2 * j + 1) << sh));
            }
        }
        if (tid == 0) *ctr = 0;
        if (has_next) level_a(nxt);
        __syncthreads();
        // ---- A: element groups -> staged block rows
        {
            double* Gw = gw + (size_t)warp * MM_GROUP * MM_EL;
            // work units: full rounds of 4-element steps, then the remainder in 2-element steps (lane = element, point,
            // node half) so that the last round, which not every warp takes part in, is short
#if FEB_MM_TAIL2
            const int n4 = (MM_GROUP == 4) ? ((TE / 4) / MM_WARPS) * MM_WARPS : 0;
#else
            const int n4 = (MM_GROUP == 4) ? (TE + 3) / 4 : 0;
#endif
            const int rest = TE - 4 * n4 > 0 ? TE - 4 * n4 : 0;
            const int n_groups = n4 + (rest + 1) / 2;
            int g_next = 0;
            if (lane == 0) g_next = atomicAdd(ctr, 1);
            for (;;) {
                const int g = __shfl_sync(0xffffffffu, g_next, 0);
                if (g >= n_groups) break;
                if (lane == 0) g_next = atomicAdd(ctr, 1);   // claimed one step ahead: the counter's latency is off the path
                const bool four = g < n4;
                const int le0 = four ? 4 * g : 4 * n4 + 2 * (g - n4);
                const int n_el = min(four ? 4 : 2, TE - le0);
                int my0 = 0, my1 = 0;                        // item range of element le0 + (lane & 3)
                {
                    const int le = le0 + (lane & 3);
                    if ((lane & 3) < n_el) { my0 = e0_s[le]; my1 = e1_s[le]; }
                }
                if (four) {            // geometry: lane = (element j, point q)
                    const int j = lane >> 3, q = lane & 7, le = le0 + j;
                    if (j < n_el) {
                        double gr[24];
                        const double* xe = xs + le * MM_XS;
                        const double dt = hex8_gradients(xe, s_dphi + q * MM_DPHI, gr);
                        double* dst = Gw + j * MM_EL + (q >> 2) * MM_KS + (q & 3);
#pragma unroll
                        for (int r = 0; r < 24; ++r) dst[r * 4] = gr[r];
                        dst[96] = (WM == 0) ? dt * s_w[q] : (WM == 2 ? dt * s_w[q] * xe[24] * eq_s[le * 8 + q] : dt * s_w[q] * xe[24]);
                    }
                } else {               // lane = (element j, point q, node half h)
                    const int j = lane >> 4, q = (lane >> 1) & 7, h = lane & 1, le = le0 + j;
                    if (j < n_el) {
                        double gr[12];
                        const double* xe = xs + le * MM_XS;
                        const double dt = hex8_gradients_half(xe, s_dphi + q * MM_DPHI, h, gr);
                        double* dst = Gw + j * MM_EL + (q >> 2) * MM_KS + (q & 3);
#pragma unroll
                        for (int r = 0; r < 12; ++r) dst[(12 * h + r) * 4] = gr[r];
                        if (h == 0) dst[96] = (WM == 0) ? dt * s_w[q] : (WM == 2 ? dt * s_w[q] * xe[24] * eq_s[le * 8 + q] : dt * s_w[q] * xe[24]);
                    }
                }
                __syncwarp();
#pragma unroll 1
                for (int j = 0; j < n_el; ++j) {
                    const int item0 = __shfl_sync(0xffffffffu, my0, j), cnt = __shfl_sync(0xffffffffu, my1, j) - item0;
                    const double* Ge = Gw + j * MM_EL;
                    double Bf[3][2];
#pragma unroll
                    for (int nt = 0; nt < 3; ++nt) {
                        Bf[nt][0] = Ge[(8 * nt + gid) * 4 + tig];
                        Bf[nt][1] = Ge[MM_KS + (8 * nt + gid) * 4 + tig];
                    }
                    const double w0 = Ge[96 + tig], w1 = Ge[MM_KS + 96 + tig];
                    const uint32_t* itw_e = itw_s + item0;
                    double* stage_e = stage + (size_t)item0 * MM_IS;
                    if (cnt > 5) element_rows<3>(Ge, Bf, w0, w1, itw_e, cnt, stage_e, gid, tig);
                    else if (cnt > 2) element_rows<2>(Ge, Bf, w0, w1, itw_e, cnt, stage_e, gid, tig);
                    else element_rows<1>(Ge, Bf, w0, w1, itw_e, cnt, stage_e, gid, tig);
                }
                __syncwarp();                                // fragment buffer free for the next group
            }
            if (has_next) level_b(nxt);                      // after the last group: issued earlier, these loads hold scoreboards the group loop needs
        }
        __syncthreads();                                     // all block rows staged
        // ---- B: lane = CSR block rank; a warp works on MM_BV vertices at a time (independent instruction streams:
        //         the phase is bound by dependent-issue latency, not by any pipe)
        bool fetched_c = false;
        int n_done = 0;
        if (FEB_MM_LEVELC_AFTER == 0 && has_next) { level_c(); fetched_c = true; }   // before the first vertex group
        for (int a0 = warp; a0 < TN; a0 += MM_BV * MM_WARPS) {
            const int r = lane;
            int av[MM_BV], meta[MM_BV];
            unsigned int src[MM_BV];
            uint2 tk[MM_BV];
            double acc[MM_BV][9], accl[UNIFORM ? 1 : MM_BV][9];
#pragma unroll
            for (int v = 0; v < MM_BV; ++v) {
                const int a = a0 + v * MM_WARPS;
                av[v] = a < TN ? a : -1;
                const int as = a < TN ? a : a0;
                meta[v] = meta_s[as];
                src[v] = tbl[as * MM_MAXDEG + r];
                tk[v] = *reinterpret_cast<const uint2*>(itof_s + as * MM_MAXVAL);
#pragma unroll
                for (int c = 0; c < 9; ++c) {
                    acc[v][c] = 0.0;
                    if (!UNIFORM) accl[v][c] = 0.0;
                }
            }
            __syncwarp();
#pragma unroll
            for (int v = 0; v < MM_BV; ++v)
                if (av[v] >= 0) tbl[av[v] * MM_MAXDEG + r] = 0xffffffffu;      // leave the table cleared for the next tile
#pragma unroll
            for (int k = 0; k < MM_MAXVAL; ++k) {
#pragma unroll
                for (int v = 0; v < MM_BV; ++v) {
                    const unsigned int bsel = (src[v] >> (4 * k)) & 0xfu;
                    const int t = (int)(((k < 4 ? tk[v].x : tk[v].y) >> (8 * (k & 3))) & 0xffu);
                    const bool on = (bsel != 0xfu) && (av[v] >= 0);
                    // absent contributions: load and addition are both predicated off.  (The select form
                    // `acc += on ? x : 0.0` compiled to a zero-fill per load plus a copy of the accumulator on the FP64 pipe.)
                    const double* sp = stage + (size_t)t * MM_IS + bsel * 3;
                    if (UNIFORM) {
#pragma unroll
                        for (int i = 0; i < 3; ++i)
#pragma unroll
                            for (int j = 0; j < 3; ++j)
                                if (on) acc[v][i * 3 + j] += sp[i * 24 + j];
                    } else {
                        const double rho = on ? xs[(itw_s[t] & 0xfffu) * MM_XS + 25] : 0.0;
#pragma unroll
                        for (int i = 0; i < 3; ++i)
#pragma unroll
                            for (int j = 0; j < 3; ++j)
                                if (on) {
                                    const double val = sp[i * 24 + j];
                                    acc[v][i * 3 + j] += val;
                                    accl[UNIFORM ? 0 : v][i * 3 + j] += rho * val;
                                }
                    }
                }
            }
#pragma unroll
            for (int v = 0; v < MM_BV; ++v) {
                const int a = av[v];
                const int deg = meta[v] & 0xff;
                if (a >= 0 && r < deg) {
                    const int64_t p0 = p0_s[a];
                    const int row_len = 3 * deg;
                    double K[9];
                    if (UNIFORM) {
                        const double tr = u_mu * (acc[v][0] + acc[v][4] + acc[v][8]);
#pragma unroll
                        for (int d = 0; d < 3; ++d)
#pragma unroll
                            for (int kk = 0; kk < 3; ++kk)
                                K[d * 3 + kk] = u_lam * acc[v][d * 3 + kk] + u_mu * acc[v][kk * 3 + d] + (d == kk ? tr : 0.0);
                    } else {
                        const double tr = acc[v][0] + acc[v][4] + acc[v][8];
#pragma unroll
                        for (int d = 0; d < 3; ++d)
#pragma unroll
                            for (int kk = 0; kk < 3; ++kk)
                                K[d * 3 + kk] = accl[UNIFORM ? 0 : v][d * 3 + kk] + acc[v][kk * 3 + d] + (d == kk ? tr : 0.0);
                    }
                    if (ELIM) {
                        const unsigned int cfd = flg_s[a * MM_MAXDEG + r];
                        const unsigned int cf = cfd & 7u, diag = cfd & 8u;
                        const unsigned int rf = ((unsigned int)meta[v] >> 16) & 7u;
                        if (cf | rf) {
#pragma unroll
                            for (int i = 0; i < 3; ++i)
#pragma unroll
                                for (int j = 0; j < 3; ++j)
                                    if (((rf >> i) & 1u) || ((cf >> j) & 1u))
                                        K[i * 3 + j] = (((rf >> i) & 1u) && diag && i == j) ? 1.0 : 0.0;
                        }
                    }
                    double* dst = values + 9 * p0 + 3 * r;
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int j = 0; j < 3; ++j) __stcs(dst + (int64_t)i * row_len + j, K[i * 3 + j]);
                }
            }
            if (has_next && !fetched_c && ++n_done >= FEB_MM_LEVELC_AFTER) { level_c(); fetched_c = true; }
        }
        if (has_next && !fetched_c) level_c();
        cur = nxt;
        nxt = nxt2;
    }
}

}  // namespace feb

using namespace feb;

extern "C" {

size_t feb200_fused_mma_smem_bytes(int32_t tn_cap, int32_t te_cap, int32_t material_layout) {
    return mma_layout(tn_cap, te_cap, material_layout == FEB200_LAYOUT_EQM).total;
}

int feb200_assemble_fused_hex8_mma(feb200_stream s, const feb200_batch* bp, const feb200_fused_plan* plan, const double* x_vd,
                                   const uint16_t* slot_rank, const int32_t* nindices, const uint8_t* vertex_flags,
                                   int32_t material_hints, double* values) {
    FEB_REQUIRE(bp && plan && x_vd && slot_rank && nindices && values, FEB200_ERR_INVALID, "assemble_fused_hex8_mma: null pointer");
    const feb200_batch& b = *bp;
    FEB_REQUIRE(b.N == 8 && b.Q == 8 && b.D == 3, FEB200_ERR_UNSUPPORTED,
                "assemble_fused_hex8_mma: needs hex8 with 8 quadrature points");
    FEB_REQUIRE(b.kind == FEB200_KIND_LINEAR && b.material == FEB200_MAT_ISOTROPIC, FEB200_ERR_UNSUPPORTED,
                "assemble_fused_hex8_mma: needs the linear isotropic material");
    FEB_REQUIRE(b.conn && b.dphi && b.W && b.mp && b.M >= 2, FEB200_ERR_INVALID, "assemble_fused_hex8_mma: bad batch");
    FEB_REQUIRE(plan->tile_node_ptr && plan->tile_nodes && plan->tile_elem_ptr && plan->tile_elems && plan->tile_item_ptr &&
                    plan->tile_items && plan->tile_node_p0 && plan->tile_node_meta && plan->n_tiles >= 0,
                FEB200_ERR_INVALID, "assemble_fused_hex8_mma: bad plan");
    FEB_REQUIRE(plan->tn_cap > 0 && plan->tn_cap <= 32 && plan->te_cap > 0 && plan->te_cap * 8 <= MM_PF * MM_THREADS &&
                    plan->tn_cap * MM_MAXVAL <= MM_THREADS,
                FEB200_ERR_INVALID, "assemble_fused_hex8_mma: plan capacities tn=%d te=%d out of range", plan->tn_cap, plan->te_cap);
    if (plan->n_tiles == 0) return FEB200_OK;
    const bool perq = (b.layout == FEB200_LAYOUT_EQM);
    const MmaSmem L = mma_layout(plan->tn_cap, plan->te_cap, perq);
    FEB_REQUIRE(L.total <= 227 * 1024, FEB200_ERR_CAPACITY, "assemble_fused_hex8_mma: tile needs %zu bytes of shared memory", L.total);
    cudaStream_t st = as_stream(s);
    const int max_grid = sm_count() * (L.total <= 113 * 1024 ? FEB_MM_MINBLOCKS : 1);
    const int grid = plan->n_tiles < max_grid ? plan->n_tiles : max_grid;
    const int wm = perq ? 2 : (b.layout == FEB200_LAYOUT_EM ? 1 : 0);
    const bool uniform = (wm == 0) || (material_hints & FEB200_HINT_UNIFORM_POISSON);
#define FEB_MMA_LAUNCH(EL, UN, WMODE)                                                                                  \
    {                                                                                                                  \
        auto kern = k_fused_hex8_mma<EL, UN, WMODE>;                                                                   \
        FEB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));               \
        kern<<<grid, MM_THREADS, L.total, st>>>(b, *plan, L, x_vd, slot_rank, nindices, vertex_flags, values);         \
    }
#define FEB_MMA_DISPATCH(EL)                                                                                           \
    if (wm == 0) FEB_MMA_LAUNCH(EL, true, 0)                                                                           \
    else if (wm == 1) { if (uniform) FEB_MMA_LAUNCH(EL, true, 1) else FEB_MMA_LAUNCH(EL, false, 1) }                   \
    else { if (uniform) FEB_MMA_LAUNCH(EL, true, 2) else FEB_MMA_LAUNCH(EL, false, 2) }
    if (vertex_flags) { FEB_MMA_DISPATCH(true) } else { FEB_MMA_DISPATCH(false) }
#undef FEB_MMA_DISPATCH
#undef FEB_MMA_LAUNCH
    FEB_CHECK_LAUNCH("assemble_fused_hex8_mma");
    return FEB200_OK;
}

}  // extern "C"
