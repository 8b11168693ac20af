// Fused hex8 assembly, streaming ("sweep") variant: every warp is an independent worker that walks a strip of
// vertices chunk by chunk (4 vertices per chunk; on a structured grid a 2 x 2 column marching along the fastest
// vertex axis) and keeps the CSR row panels of the two live chunks (8 vertices, 27 x 9 doubles each) in its own
// shared memory.  Reference path: fea.py:519-572 (element tangent) + fea.py:613-660 (COO assembly) +
// sparse_matrix.py:157-184 (Dirichlet elimination); K_e never reaches HBM.
//
// Per warp, per segment of a strip (plan: fe_jax/sweep_plan.py):
//   * visits are taken in groups of 4: lane = (element, quadrature point) computes the shape-function gradients into a
//     warp-private buffer laid out as DMMA fragments;
//   * per visit, S^e = (w G) G^T restricted to the rows of the element's nodes that live in the window
//     (1-3 row tiles of mma.sync.m8n8k4.f64; rows component-major over the in-window nodes, columns component-major
//     over the 8 nodes, so a lane ends up with the three components of two 3x3 block rows): the accumulator fragments
//     are LOADED from the panels, the DMMAs add the element's contribution, and the fragments are stored back -- the
//     sum over elements happens in the panels, warp-privately, in plan order (bit-reproducible, no atomics, no
//     block-wide barrier anywhere);
//   * when the plan says a chunk is complete, lane = CSR block rank reads the 3x3 sums, applies
//     K = lam S + mu S^T + mu tr(S) I and the Dirichlet elimination, streams the panel rows to HBM and clears the slot.
// Everything a warp reads from global memory arrives through cp.async one or two groups (chunks) ahead: visit records
// (with the element's vertex ids), coordinates, slot ranks, material, chunk records and Dirichlet flag rows.  No load
// with a register destination is in flight across the visit loop (the compiler would wait for its scoreboard at the
// loop head: 18 % of all stall samples in the first version, profiles/r02_sweep_v1_ncu_summary.txt).
// Each element is visited by ~2.3 warps on a regular grid (2.34 tiles in fused_mma.cu) with 3.7 row tiles per vertex
// (3.84), but the 144 KB staging round trip, its second pass over shared memory and all block barriers are gone.
// Handles batches whose Poisson ratio is one number (layout (M,), or the caller's FEB200_HINT_UNIFORM_POISSON);
// everything else stays on feb200_assemble_fused_hex8_mma / feb200_assemble_fused_hex8.
#include "common.cuh"
#include "element_math.cuh"
#include "hex8_mma.cuh"

#ifndef FEB_SW_MAXWARPS
#define FEB_SW_MAXWARPS 8            /* measured at 200^3: 7 / 8 / 9 warps (168 registers) = 13.0 / 11.9 / 12.7 ms */
#endif
#ifndef FEB_SW_DMAJOR
#define FEB_SW_DMAJOR 0            /* rows of a visit: 1 = component-major over the in-window nodes, 0 = node-major */
#endif
#ifndef FEB_SW_ROWS6
#define FEB_SW_ROWS6 1             /* 1 = a row tile holds two nodes (rows 0-2 and 4-6; rows 3 and 7 stay empty): the 16 lanes of a
                                      half-warp then touch ONE node's three panel rows and the accumulator loads / stores are free of
                                      bank conflicts (with three nodes per tile every such access was a 2-way conflict, 70 of 330
                                      wavefronts per vertex); costs one more row tile for the elements with all 8 nodes in the window */
#endif
#ifndef FEB_SW_DPHI_REGS
#define FEB_SW_DPHI_REGS 1         /* 1 = every lane keeps the basis table of its quadrature point in registers */
#endif
// Panel layout: value (rank r, row component d, column component k) of slot s at  sbase[s] + r * BS + d * DS + k * KK.
// COLS: local node of accumulator column 0..7 (a lane holds columns 2 * tig and 2 * tig + 1 of every column tile);
// PAD: start of slot s modulo 16 doubles (-1: wherever the previous slot ends) -- both chosen so that the 16 lanes of a half-warp (4 rows x 4 column
// pairs) spread over the 16 eight-byte banks on structured grids (profiles/r02_panel_layout_search.txt).
#ifndef FEB_SW_ROT
#define FEB_SW_ROT 1              /* rotation of the node order per component in the fragment buffer */
#endif
#ifndef FEB_SW_EXCH
#define FEB_SW_EXCH 1             /* 1 = read-and-clear of a retiring panel entry as one shared-memory exchange */
#endif
#ifndef FEB_SW_BS
#define FEB_SW_BS 9
#endif
#ifndef FEB_SW_DS
#define FEB_SW_DS 3
#endif
#ifndef FEB_SW_KK
#define FEB_SW_KK 1
#endif
#ifndef FEB_SW_COLS
#define FEB_SW_COLS {0, 1, 2, 3, 4, 5, 6, 7}
#endif
#ifndef FEB_SW_PAD
#define FEB_SW_PAD {-1, -1, -1, -1, -1, -1, -1, -1}
#endif

namespace feb {

constexpr int SW_KS = 104;     // doubles per k-step half of an element's fragment buffer: [row 0..23][4 points] + 8 (weights)
constexpr int SW_EL = 212;     // doubles per element in the fragment buffer (2 * KS + 4): geometry stores hit distinct banks
constexpr int SW_XS = 24;      // doubles per element coordinate record [n][3]
constexpr int SW_DPHI = 26;    // doubles per quadrature point of the basis table (24 used, 16-byte rows)
constexpr int SW_SLOTS = 8;    // panel slots per warp: two chunks of four vertices
constexpr int SW_REC = 12;     // int32 per visit record: element, slotw, perm, flags, 8 vertex ids
constexpr size_t SW_SMEM_MAX = 227 * 1024;

struct SwSmem {
    int warps, ss, panel_doubles;
    int sbase[8];                  // doubles: start of panel slot s within the warp's panel area
    int off_rk, off_rec, off_info, off_flg, off_eq, off_x, off_g, off_panel;
    size_t off_warps, per_warp, total;
};
static inline SwSmem sweep_layout(int max_deg) {
    SwSmem L;
    L.ss = FEB_SW_BS * max_deg;                          // doubles per panel slot: [rank] BS
    const int pad[8] = FEB_SW_PAD;
    int panel_doubles = 0;
    for (int sl = 0; sl < SW_SLOTS; ++sl) {
        int start = panel_doubles;
        if (pad[sl] >= 0)
            while ((start & 15) != (pad[sl] & 15)) ++start;
        L.sbase[sl] = start;
        panel_doubles = start + L.ss;
    }
    size_t o = (size_t)(8 * SW_DPHI + 8) * 8 + 8 * 4;    // basis table + weights + slot starts (shared by the block)
    L.off_warps = (o + 15) & ~(size_t)15;
    int w = 0;
    L.off_rk = w;    w += 2 * 4 * 64;                    // [2][4 elements][8][8] uint8 slot ranks (double buffered)
    L.off_rec = w;   w += 3 * 4 * SW_REC * 4;            // [3][4] visit records (groups g, g+1, g+2)
    L.off_info = w;  w += 4 * 64;                        // [4] chunk records {v[4], deg[4], p0[4]}
    L.off_flg = w;   w += 3 * 4 * 32;                    // [3][4 vertices][32 ranks] Dirichlet flag bytes
    L.off_eq = w;    w += 4 * 8 * 8;                     // Young's modulus per (element, point) of the next group
    L.off_x = w;     w += 4 * SW_XS * 8;                 // coordinates of the next group
    L.off_g = w;     w += 4 * SW_EL * 8;                 // fragment buffers of the current group
    L.panel_doubles = panel_doubles;
    L.off_panel = w; w += panel_doubles * 8;
    L.per_warp = ((size_t)w + 15) & ~(size_t)15;
    int warps = (int)((SW_SMEM_MAX - L.off_warps) / L.per_warp);
    if (warps > FEB_SW_MAXWARPS) warps = FEB_SW_MAXWARPS;
    L.warps = warps;
    L.total = L.off_warps + (size_t)warps * L.per_warp;
    return L;
}

__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Row of (node n, component d) in an element's fragment buffer: component-major, the nodes of a component permuted so
// that antipodal nodes share a residue mod 4 and the residues rotate by FEB_SW_ROT per component.  With two-node row tiles a
// half-warp loads the three component rows of ONE node at once: rotation 1 puts them into distinct 4-double bank groups
// (rotation 2, chosen for the older three-node tiles, made components 0 and 2 collide: 4 instead of 2 wavefronts per A-fragment
// load, 10.77 -> 10.70 ms at 200^3).  The 8 rows of a component (B fragments) are a permutation of 8 consecutive rows either way.
__device__ __forceinline__ int sw_grow(int n, int d) { return 8 * d + (((n < 4 ? n : 11 - n) + FEB_SW_ROT * d) & 7); }

// Row tiles 0..NMT-1 of one visit, in three stages so that two independent visits can be interleaved (sweep_pair).
// Rows: the visit's in-window nodes in ascending local order, two per row tile (rows 0-2 / 4-6 = the three components of a
// node: FEB_SW_ROWS6); columns: component-major (column tile nt = component, column in tile = local node).
template <int NMT>
struct SwRows {
    double A[NMT][2];
    double* base0[NMT];
    double* base1[NMT];
    bool valid[NMT];
    double C[NMT][3][2];

    // addresses, weighted A fragments, accumulators loaded from the panels
    __device__ __forceinline__ void load(const double* __restrict__ Ge, double w0, double w1, const uint8_t* __restrict__ rk,
                                         unsigned slotw, unsigned perm, int cnt, double* __restrict__ panel,
                                         const int* __restrict__ sbase, int gid, int tig, int node0, int node1) {
#pragma unroll
        for (int mt = 0; mt < NMT; ++mt) {
#if FEB_SW_ROWS6
            const int s = 2 * mt + (gid >> 2), d = gid & 3;
            valid[mt] = d < 3 && s < cnt;
#elif FEB_SW_DMAJOR
            const int R = 8 * mt + gid;
            const int d = (R >= cnt ? 1 : 0) + (R >= 2 * cnt ? 1 : 0), s = R - d * cnt;
            valid[mt] = R < 3 * cnt;
#else
            const int R = 8 * mt + gid;
            const int s = R / 3, d = R - 3 * s;
            valid[mt] = s < cnt;
#endif
            const int na = valid[mt] ? (int)((perm >> (4 * s)) & 7u) : 0;
            const int dd = valid[mt] ? d : 0;
            const int slot = (int)((slotw >> (4 * na)) & 7u);
            double* pb = panel + sbase[slot] + dd * FEB_SW_DS;
            base0[mt] = pb + (unsigned)rk[na * 8 + node0] * (unsigned)FEB_SW_BS;
            base1[mt] = pb + (unsigned)rk[na * 8 + node1] * (unsigned)FEB_SW_BS;
            const int gr = sw_grow(na, dd) * 4 + tig;
            const double g0 = Ge[gr], g1 = Ge[SW_KS + gr];
            A[mt][0] = valid[mt] ? w0 * g0 : 0.0;
            A[mt][1] = valid[mt] ? w1 * g1 : 0.0;
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
                C[mt][nt][0] = 0.0;
                C[mt][nt][1] = 0.0;
                if (valid[mt]) {
                    C[mt][nt][0] = base0[mt][nt * FEB_SW_KK];
                    C[mt][nt][1] = base1[mt][nt * FEB_SW_KK];
                }
            }
        }
    }
    __device__ __forceinline__ void mma(const double (&Bf)[3][2], int ks) {
#pragma unroll
        for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) dmma884(C[mt][nt][0], C[mt][nt][1], A[mt][ks], Bf[nt][ks]);
    }
    __device__ __forceinline__ void store() {
#pragma unroll
        for (int mt = 0; mt < NMT; ++mt)
            if (valid[mt]) {
#pragma unroll
                for (int nt = 0; nt < 3; ++nt) {
                    base0[mt][nt * FEB_SW_KK] = C[mt][nt][0];
                    base1[mt][nt * FEB_SW_KK] = C[mt][nt][1];
                }
            }
    }
};

struct SwVisit {
    const double* Ge;
    const uint8_t* rk;
    unsigned slotw, perm;
    int cnt;
    double Bf[3][2];
    double w0, w1;
};

template <int NMT>
__device__ __forceinline__ void sweep_rows(const SwVisit& v, double* __restrict__ panel, const int* __restrict__ sbase, int gid,
                                           int tig, int node0, int node1) {
    SwRows<NMT> r;
    r.load(v.Ge, v.w0, v.w1, v.rk, v.slotw, v.perm, v.cnt, panel, sbase, gid, tig, node0, node1);
    r.mma(v.Bf, 0);
    r.mma(v.Bf, 1);
    r.store();
}

// Two visits that touch disjoint panel slots (the plan pairs them): one instruction stream with two independent chains.
template <int NA, int NB>
__device__ __forceinline__ void sweep_pair(const SwVisit& va, const SwVisit& vb, double* __restrict__ panel,
                                           const int* __restrict__ sbase, int gid, int tig, int node0, int node1) {
    SwRows<NA> ra;
    SwRows<NB> rb;
    ra.load(va.Ge, va.w0, va.w1, va.rk, va.slotw, va.perm, va.cnt, panel, sbase, gid, tig, node0, node1);
    rb.load(vb.Ge, vb.w0, vb.w1, vb.rk, vb.slotw, vb.perm, vb.cnt, panel, sbase, gid, tig, node0, node1);
    ra.mma(va.Bf, 0);
    rb.mma(vb.Bf, 0);
    ra.mma(va.Bf, 1);
    rb.mma(vb.Bf, 1);
    ra.store();
    rb.store();
}

// WM, how the material enters the quadrature weight (the Poisson ratio is one number for the batch):
//   0  not at all: one (lambda, mu) for the batch, applied when a chunk is stored               layout (M,)
//   1  w_q * mu_e                                                                               layout (E,M)
//   2  w_q * E_q / (2 (1 + nu))                                                                 layout (E,Q,M)
template <bool ELIM, int WM, bool ACC>
__global__ void __launch_bounds__(FEB_SW_MAXWARPS * 32, 1)
k_fused_hex8_sweep(feb200_batch b, feb200_sweep_plan plan, SwSmem L, const double* __restrict__ x_vd,
                   const uint8_t* __restrict__ slot_rank, const uint8_t* __restrict__ block_flags,
                   double* __restrict__ values) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_dphi = reinterpret_cast<double*>(smem_raw);                    // [q] DPHI : [n][3]
    double* s_w = s_dphi + 8 * SW_DPHI;                                      // [q]
    int* s_sbase = reinterpret_cast<int*>(s_w + 8);                          // [slot] start of the slot's panel (doubles)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    unsigned char* mine = smem_raw + L.off_warps + (size_t)warp * L.per_warp;
    uint8_t* rkbuf = mine + L.off_rk;                                        // [2][4][64]
    int* recs = reinterpret_cast<int*>(mine + L.off_rec);                    // [3][4] REC
    int* info = reinterpret_cast<int*>(mine + L.off_info);                   // [4] 16 ints: v[4], deg[4], p0[4] (int64)
    uint8_t* flg = mine + L.off_flg;                                         // [3][4][32]
    double* eqbuf = reinterpret_cast<double*>(mine + L.off_eq);              // [4][8]
    double* xbuf = reinterpret_cast<double*>(mine + L.off_x);                // [4] XS
    double* Gw = reinterpret_cast<double*>(mine + L.off_g);                  // [4] EL : [ks] KS : [row][4 points], weights at 96
    double* panel = reinterpret_cast<double*>(mine + L.off_panel);           // slot s at sbase[s] : [rank] BS : (d, k) at d * DS + k * KK
    constexpr int cols[8] = FEB_SW_COLS;
    int node0 = 0, node1 = 0, bnode = 0;                                     // local nodes of this lane's two accumulator columns / its B column
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        if (c == 2 * tig) node0 = cols[c];
        if (c == 2 * tig + 1) node1 = cols[c];
        if (c == gid) bnode = cols[c];
    }

    for (int i = tid; i < 8 * 24; i += blockDim.x) s_dphi[(i / 24) * SW_DPHI + (i % 24)] = b.dphi[i];
    if (tid < 8) { s_w[tid] = b.W[tid]; s_sbase[tid] = L.sbase[tid]; }
    for (int i = lane; i < L.panel_doubles; i += 32) panel[i] = 0.0;
#if FEB_SW_DPHI_REGS
    double dq[24];                                       // basis table of this lane's quadrature point
#pragma unroll
    for (int i = 0; i < 24; ++i) dq[i] = b.dphi[(lane & 7) * 24 + i];
#endif
    double u_lam, u_mu;
    isotropic_lame<3>(b.mp, u_lam, u_mu);
    const double half_inv = 0.5 / (1.0 + b.mp[1]);      // mu = E * half_inv
    if (WM > 0) { u_lam = u_lam / u_mu; u_mu = 1.0; }    // mu is in the weights: K = (lam/mu) S' + S'^T + tr(S') I
    __syncthreads();                                     // the only block-wide barrier: basis table ready

    const int n_workers = gridDim.x * (blockDim.x >> 5);
    const int jj = lane >> 3;                            // element of the group this lane copies / computes for
    const int qq = lane & 7;                             // its quadrature point (geometry) or local node (copies)
    const int* visits = plan.visits;

    for (int seg = blockIdx.x * (blockDim.x >> 5) + warp; seg < plan.n_segments; seg += n_workers) {
        const int v0 = plan.seg_visit_ptr[seg], v1 = plan.seg_visit_ptr[seg + 1];
        const int c0 = plan.seg_chunk_ptr[seg], c1 = plan.seg_chunk_ptr[seg + 1];
        const int n_groups = (v1 - v0 + 3) >> 2;

        // records of group gi -> ring slot gi % 3 (entries past the end of the segment: element = -1)
        auto issue_rec = [&](int gi) {
            if (lane < 12) {
                const int j = lane / 3, p = lane - 3 * j;
                const int v = v0 + 4 * gi + j;
                int* dst = recs + ((gi % 3) * 4 + j) * SW_REC + 4 * p;
                if (v < v1) cp_async16(dst, visits + (int64_t)v * SW_REC + 4 * p);
                else if (p == 0) dst[0] = -1;
            }
        };
        // coordinates, slot ranks and material of group gi (its records have arrived)
        auto issue_data = [&](int gi) {
            const int* r = recs + ((gi % 3) * 4 + jj) * SW_REC;
            const int el = r[0];
            if (el >= 0) {
                const int vid = r[4 + qq];
                const double* src = x_vd + (int64_t)vid * 3;
                double* dst = xbuf + jj * SW_XS + qq * 3;
                cp_async8(dst, src);
                cp_async8(dst + 1, src + 1);
                cp_async8(dst + 2, src + 2);
                if (qq < 4) cp_async16(rkbuf + (gi & 1) * 256 + jj * 64 + qq * 16, slot_rank + (int64_t)el * 64 + qq * 16);
                if (WM == 2) cp_async8(eqbuf + jj * 8 + qq, b.mp + ((int64_t)el * 8 + qq) * b.M);
                if (WM == 1 && qq == 0) cp_async8(eqbuf + jj * 8, b.mp + (int64_t)el * b.M);
            }
        };
        // chunk record c -> ring slot (c - c0) % 4; flag rows of its vertices -> ring slot (c - c0) % 3
        auto issue_info = [&](int c) {
            if (lane < 4) {
                int* dst = info + ((c - c0) & 3) * 16 + 4 * lane;
                if (c < c1) cp_async16(dst, plan.chunk_rec + (int64_t)c * 16 + 4 * lane);
                else if (lane == 0) { dst[0] = -1; dst[1] = -1; dst[2] = -1; dst[3] = -1; }
            }
        };
        auto issue_flags = [&](int c) {
            if (ELIM && lane < 8 && c < c1) {
                const int s = lane >> 1, h = lane & 1;
                const int v = info[((c - c0) & 3) * 16 + s];
                if (v >= 0) cp_async16(flg + (((c - c0) % 3) * 4 + s) * 32 + 16 * h, block_flags + (int64_t)v * 32 + 16 * h);
            }
        };

        issue_rec(0);
        issue_rec(1);
        issue_info(c0);
        issue_info(c0 + 1);
        issue_info(c0 + 2);
        cp_async_commit();
        cp_async_wait_all();
        __syncwarp();
        issue_data(0);
        issue_flags(c0);
        issue_flags(c0 + 1);
        cp_async_commit();
        int c_next = c0;                                  // the next chunk to retire; it lives in slots 4 * ((c - c0) & 1) ..

        for (int gi = 0; gi < n_groups; ++gi) {
            cp_async_wait_all();
            __syncwarp();
            const int* rg = recs + (gi % 3) * 4 * SW_REC;
            // ---- geometry of the group: lane = (element jj, point qq)
            if (rg[jj * SW_REC] >= 0) {
                double gr[24];
#if FEB_SW_DPHI_REGS
                const double dt = hex8_gradients_rd(xbuf + jj * SW_XS, dq, gr);
#else
                const double dt = hex8_gradients(xbuf + jj * SW_XS, s_dphi + qq * SW_DPHI, gr);
#endif
                double* dst = Gw + jj * SW_EL + (qq >> 2) * SW_KS + (qq & 3);
#pragma unroll
                for (int n = 0; n < 8; ++n)
#pragma unroll
                    for (int d = 0; d < 3; ++d) dst[sw_grow(n, d) * 4] = gr[n * 3 + d];
                double w = dt * s_w[qq];
                if (WM == 1) w *= eqbuf[jj * 8] * half_inv;
                if (WM == 2) w *= eqbuf[jj * 8 + qq] * half_inv;
                dst[96] = w;
            }
            __syncwarp();
            // ---- prefetch: data of the next group (its records arrived with the wait above), records two groups ahead
            issue_data(gi + 1);
            issue_rec(gi + 2);
            cp_async_commit();
            // ---- visits
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                int4 hd = *reinterpret_cast<const int4*>(rg + j * SW_REC);
                if (hd.x < 0) break;
                auto fetch = [&](SwVisit& v, const int4& h, int jv) {
                    v.slotw = (unsigned)h.y;
                    v.perm = (unsigned)h.z;
                    v.cnt = h.w & 0xf;
                    v.Ge = Gw + jv * SW_EL;
                    v.rk = rkbuf + (gi & 1) * 256 + jv * 64;
#pragma unroll
                    for (int nt = 0; nt < 3; ++nt) {
                        v.Bf[nt][0] = v.Ge[sw_grow(bnode, nt) * 4 + tig];
                        v.Bf[nt][1] = v.Ge[SW_KS + sw_grow(bnode, nt) * 4 + tig];
                    }
                    v.w0 = v.Ge[96 + tig];
                    v.w1 = v.Ge[SW_KS + 96 + tig];
                };
                SwVisit va;
                fetch(va, hd, j);
#if FEB_SW_ROWS6
                if (hd.w & 0x100) {          // paired with the next visit (disjoint panel slots, both with <= 4 in-window nodes)
                    const int4 hb = *reinterpret_cast<const int4*>(rg + (j + 1) * SW_REC);
                    SwVisit vb;
                    fetch(vb, hb, j + 1);
                    const bool a2 = va.cnt > 2, b2 = vb.cnt > 2;
                    if (a2 && b2) sweep_pair<2, 2>(va, vb, panel, s_sbase, gid, tig, node0, node1);
                    else if (a2) sweep_pair<2, 1>(va, vb, panel, s_sbase, gid, tig, node0, node1);
                    else if (b2) sweep_pair<1, 2>(va, vb, panel, s_sbase, gid, tig, node0, node1);
                    else sweep_pair<1, 1>(va, vb, panel, s_sbase, gid, tig, node0, node1);
                    hd = hb;                 // the chunks completed by the pair are recorded on its second visit
                    ++j;
                } else if (va.cnt > 6) sweep_rows<4>(va, panel, s_sbase, gid, tig, node0, node1);
                else if (va.cnt > 4) sweep_rows<3>(va, panel, s_sbase, gid, tig, node0, node1);
                else if (va.cnt > 2) sweep_rows<2>(va, panel, s_sbase, gid, tig, node0, node1);
                else sweep_rows<1>(va, panel, s_sbase, gid, tig, node0, node1);
#else
                if (va.cnt > 5) sweep_rows<3>(va, panel, s_sbase, gid, tig, node0, node1);
                else if (va.cnt > 2) sweep_rows<2>(va, panel, s_sbase, gid, tig, node0, node1);
                else sweep_rows<1>(va, panel, s_sbase, gid, tig, node0, node1);
#endif
                __syncwarp();
                // ---- chunks completed by this visit: combine, eliminate, store, clear
                for (int nret = (hd.w >> 4) & 0xf; nret > 0; --nret) {
                    cp_async_wait_all();                  // chunk record and flag rows of c_next (issued two chunks ago)
                    __syncwarp();
                    const int cr = c_next - c0;
                    const int* ci = info + (cr & 3) * 16;
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const int v = ci[s];
                        const int deg = ci[4 + s];
                        if (v >= 0 && lane < deg) {
                            const int64_t p0 = *reinterpret_cast<const int64_t*>(ci + 8 + 2 * s);
                            double* src = panel + s_sbase[4 * (cr & 1) + s] + lane * FEB_SW_BS;
                            double S[9], K[9];
#pragma unroll
                            for (int d = 0; d < 3; ++d)
#pragma unroll
                                for (int kk = 0; kk < 3; ++kk) {
#if FEB_SW_EXCH
                                    S[d * 3 + kk] = __longlong_as_double((long long)atomicExch(
                                        reinterpret_cast<unsigned long long*>(src + d * FEB_SW_DS + kk * FEB_SW_KK), 0ull));
#else
                                    S[d * 3 + kk] = src[d * FEB_SW_DS + kk * FEB_SW_KK];
                                    src[d * FEB_SW_DS + kk * FEB_SW_KK] = 0.0;
#endif
                                }
                            const double tr = u_mu * (S[0] + S[4] + S[8]);
#pragma unroll
                            for (int d = 0; d < 3; ++d)
#pragma unroll
                                for (int kk = 0; kk < 3; ++kk)
                                    K[d * 3 + kk] = u_lam * S[d * 3 + kk] + u_mu * S[kk * 3 + d] + (d == kk ? tr : 0.0);
                            double* dst = values + 9 * p0 + 3 * lane;
                            const int row_len = 3 * deg;
                            if (ACC) {           // a later batch of a multi-batch mesh: add to what the earlier batches stored
#pragma unroll
                                for (int i = 0; i < 3; ++i)
#pragma unroll
                                    for (int jx = 0; jx < 3; ++jx) K[i * 3 + jx] += __ldcs(dst + (int64_t)i * row_len + jx);
                            }
                            if (ELIM) {          // (on the sum over all batches when accumulating: pass the flags with the last batch)
                                const unsigned f = flg[((cr % 3) * 4 + s) * 32 + lane];
                                const unsigned cf = f & 7u, diag = f & 8u, rf = (f >> 4) & 7u;
                                if (cf | rf) {
#pragma unroll
                                    for (int i = 0; i < 3; ++i)
#pragma unroll
                                        for (int jx = 0; jx < 3; ++jx)
                                            if (((rf >> i) & 1u) || ((cf >> jx) & 1u))
                                                K[i * 3 + jx] = (((rf >> i) & 1u) && diag && i == jx) ? 1.0 : 0.0;
                                }
                            }
#pragma unroll
                            for (int i = 0; i < 3; ++i)
#pragma unroll
                                for (int jx = 0; jx < 3; ++jx) __stcs(dst + (int64_t)i * row_len + jx, K[i * 3 + jx]);
                        }
                    }
                    __syncwarp();
                    ++c_next;
                    issue_info(c_next + 2);               // into the ring slot of the chunk just stored
                    issue_flags(c_next + 1);              // its record arrived with the wait above
                    cp_async_commit();
                }
            }
        }
        cp_async_wait_all();                              // nothing may still be landing in the rings when the next segment starts
        __syncwarp();
    }
}

// One byte per (vertex, CSR block rank): bits 0-2 constrained components of the neighbour, bit 3 diagonal block,
// bits 4-6 constrained components of the vertex itself; 32 bytes per vertex (degree <= 32).
__global__ void k_sweep_block_flags(int64_t V, const int64_t* __restrict__ nindptr, const int32_t* __restrict__ nindices,
                                    const uint8_t* __restrict__ vflags, uint8_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t v = w0; v < V; v += nw) {
        const int64_t p0 = nindptr[v];
        const int deg = (int)(nindptr[v + 1] - p0);
        unsigned f = 0u;
        if (lane < deg) {
            const int nb = nindices[p0 + lane];
            f = ((unsigned)vflags[nb] & 7u) | (nb == v ? 8u : 0u) | (((unsigned)vflags[v] & 7u) << 4);
        }
        out[v * 32 + lane] = (uint8_t)f;
    }
}

}  // namespace feb

using namespace feb;

extern "C" {

int32_t feb200_fused_sweep_warps(int32_t max_deg) {
    if (max_deg < 1) return 0;
    return sweep_layout(max_deg).warps;
}

int feb200_sweep_block_flags(feb200_stream s, int64_t V, const int64_t* nindptr, const int32_t* nindices,
                             const uint8_t* vertex_flags, uint8_t* block_flags) {
    FEB_REQUIRE(nindptr && nindices && vertex_flags && block_flags && V >= 0, FEB200_ERR_INVALID, "sweep_block_flags: bad argument");
    if (V == 0) return FEB200_OK;
    k_sweep_block_flags<<<grid_for(V * 32, 256), 256, 0, as_stream(s)>>>(V, nindptr, nindices, vertex_flags, block_flags);
    FEB_CHECK_LAUNCH("sweep_block_flags");
    return FEB200_OK;
}

int feb200_assemble_fused_hex8_sweep(feb200_stream s, const feb200_batch* bp, const feb200_sweep_plan* plan, const double* x_vd,
                                     const uint8_t* slot_rank8, const uint8_t* block_flags, int32_t material_hints,
                                     double* values, int32_t accumulate) {
    const uint8_t* slot_rank = slot_rank8;
    FEB_REQUIRE(bp && plan && x_vd && slot_rank && values, FEB200_ERR_INVALID, "assemble_fused_hex8_sweep: null pointer");
    const feb200_batch& b = *bp;
    FEB_REQUIRE(b.N == 8 && b.Q == 8 && b.D == 3, FEB200_ERR_UNSUPPORTED,
                "assemble_fused_hex8_sweep: needs hex8 with 8 quadrature points");
    FEB_REQUIRE(b.kind == FEB200_KIND_LINEAR && b.material == FEB200_MAT_ISOTROPIC, FEB200_ERR_UNSUPPORTED,
                "assemble_fused_hex8_sweep: needs the linear isotropic material");
    FEB_REQUIRE(b.conn && b.dphi && b.W && b.mp && b.M >= 2, FEB200_ERR_INVALID, "assemble_fused_hex8_sweep: bad batch");
    FEB_REQUIRE(b.layout == FEB200_LAYOUT_M || (material_hints & FEB200_HINT_UNIFORM_POISSON), FEB200_ERR_UNSUPPORTED,
                "assemble_fused_hex8_sweep: needs one Poisson ratio for the batch (layout (M,) or FEB200_HINT_UNIFORM_POISSON)");
    FEB_REQUIRE(plan->seg_visit_ptr && plan->seg_chunk_ptr && plan->visits && plan->chunk_rec && plan->n_segments >= 0,
                FEB200_ERR_INVALID, "assemble_fused_hex8_sweep: bad plan");
    FEB_REQUIRE(plan->max_deg >= 1 && plan->max_deg <= 32, FEB200_ERR_INVALID,
                "assemble_fused_hex8_sweep: plan max_deg=%d out of range (1..32)", plan->max_deg);
    FEB_REQUIRE((((uintptr_t)plan->visits | (uintptr_t)plan->chunk_rec | (uintptr_t)slot_rank | (uintptr_t)block_flags) & 15) == 0,
                FEB200_ERR_INVALID, "assemble_fused_hex8_sweep: visits, chunk records, slot ranks and block flags must be 16-byte aligned");
    if (plan->n_segments == 0) return FEB200_OK;
    const SwSmem L = sweep_layout(plan->max_deg);
    FEB_REQUIRE(L.warps >= 1, FEB200_ERR_CAPACITY, "assemble_fused_hex8_sweep: a worker needs %zu bytes of shared memory", L.per_warp);
    cudaStream_t st = as_stream(s);
    const int64_t want = ceil_div(plan->n_segments, L.warps);
    const int grid = (int)(want < sm_count() ? want : sm_count());
    const int wm = (b.layout == FEB200_LAYOUT_EQM) ? 2 : (b.layout == FEB200_LAYOUT_EM ? 1 : 0);
    static bool attr_set[2][3][2] = {};
#define FEB_SW_LAUNCH(EL, WMODE, AC)                                                                                   \
    {                                                                                                                  \
        auto kern = k_fused_hex8_sweep<EL, WMODE, AC>;                                                                 \
        if (!attr_set[EL ? 1 : 0][WMODE][AC ? 1 : 0]) {                                                                \
            FEB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SW_SMEM_MAX));       \
            attr_set[EL ? 1 : 0][WMODE][AC ? 1 : 0] = true;                                                            \
        }                                                                                                              \
        kern<<<grid, L.warps * 32, L.total, st>>>(b, *plan, L, x_vd, slot_rank, block_flags, values);                  \
    }
#define FEB_SW_DISPATCH2(EL, AC)                                                                                       \
    if (wm == 0) FEB_SW_LAUNCH(EL, 0, AC) else if (wm == 1) FEB_SW_LAUNCH(EL, 1, AC) else FEB_SW_LAUNCH(EL, 2, AC)
#define FEB_SW_DISPATCH(EL)                                                                                            \
    if (accumulate) { FEB_SW_DISPATCH2(EL, true) } else { FEB_SW_DISPATCH2(EL, false) }
    if (block_flags) { FEB_SW_DISPATCH(true) } else { FEB_SW_DISPATCH(false) }
#undef FEB_SW_DISPATCH2
#undef FEB_SW_DISPATCH
#undef FEB_SW_LAUNCH
    FEB_CHECK_LAUNCH("assemble_fused_hex8_sweep");
    return FEB200_OK;
}

}  // extern "C"
