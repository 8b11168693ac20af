// Fused hex8 assembly, streaming ("sweep") variant: every warp is an independent worker that walks a strip of
// vertices chunk by chunk (4 vertices per chunk; on a structured grid a 2 x 2 column marching along the fastest
// vertex axis) and keeps the CSR row panels of the two live chunks (8 vertices, 27 x 9 doubles each) in its own
// shared memory.  Reference path: fea.py:519-572 (element tangent) + fea.py:613-660 (COO assembly) +
// sparse_matrix.py:157-184 (Dirichlet elimination); K_e never reaches HBM.
//
// Per warp, per segment of a strip (plan: fe_jax/sweep_plan.py):
//   * visits are taken in groups of 4: lane = (element, quadrature point) computes the shape-function gradients into a
//     warp-private buffer laid out as DMMA fragments (coordinates arrive through cp.async one group ahead);
//   * per visit, S^e = (w G) G^T restricted to the rows of the element's nodes that live in the window
//     (1-3 row tiles of mma.sync.m8n8k4.f64, columns ordered component-major so that a lane ends up with the three
//     components of two 3x3 block rows): the accumulator fragments are LOADED from the panels, the DMMAs add the
//     element's contribution, and the fragments are stored back -- the sum over elements happens in the panels,
//     warp-privately, in plan order (bit-reproducible, no atomics, no block-wide barrier anywhere);
//   * when the plan says a chunk is complete, lane = CSR block rank reads the 3x3 sums, applies
//     K = lam S + mu S^T + mu tr(S) I and the Dirichlet elimination, streams the panel rows to HBM and clears the slot.
// Each element is visited by ~2.25 warps on a regular grid (2.34 tiles in fused_mma.cu) with 3.75 row tiles per vertex
// (3.84), but the 144 KB staging round trip, its second pass over shared memory and all block barriers are gone.
// Handles batches whose Poisson ratio is one number (layout (M,), or the caller's FEB200_HINT_UNIFORM_POISSON);
// everything else stays on feb200_assemble_fused_hex8_mma / feb200_assemble_fused_hex8.
#include "common.cuh"
#include "element_math.cuh"
#include "hex8_mma.cuh"

#ifndef FEB_SW_MAXWARPS
#define FEB_SW_MAXWARPS 9
#endif

namespace feb {

constexpr int SW_KS = 104;     // doubles per k-step half of an element's fragment buffer: [row 0..23][4 points] + 8 (weights)
constexpr int SW_EL = 212;     // doubles per element in the fragment buffer (2 * KS + 4): geometry stores hit distinct banks
constexpr int SW_XS = 26;      // doubles per element coordinate record [n][3] (+2: records of a group on distinct banks)
constexpr int SW_DPHI = 26;    // doubles per quadrature point of the basis table (24 used, 16-byte rows)
constexpr int SW_SLOTS = 8;    // panel slots per warp: two chunks of four vertices
constexpr size_t SW_SMEM_MAX = 227 * 1024;

struct SwSmem {
    int warps, ss;
    size_t off_warps, per_warp, off_rk, off_x, off_g, off_panel, total;
};
static inline SwSmem sweep_layout(int max_deg) {
    SwSmem L;
    L.ss = 9 * max_deg;                                  // doubles per panel slot: [rank][3][3]
    size_t o = (size_t)(8 * SW_DPHI + 8) * 8;            // basis table + weights (shared by the block)
    L.off_warps = (o + 15) & ~(size_t)15;
    size_t w = 0;
    L.off_rk = w;    w += 2 * 4 * 64 * 2;                // [2][4 elements][8][8] uint16 slot ranks (double buffered)
    L.off_x = w;     w += (size_t)4 * SW_XS * 8;         // coordinates of the next group
    L.off_g = w;     w += (size_t)4 * SW_EL * 8;         // fragment buffers of the current group
    L.off_panel = w; w += (size_t)SW_SLOTS * L.ss * 8;
    L.per_warp = (w + 15) & ~(size_t)15;
    int warps = (int)((SW_SMEM_MAX - L.off_warps) / L.per_warp);
    if (warps > FEB_SW_MAXWARPS) warps = FEB_SW_MAXWARPS;
    L.warps = warps;
    L.total = L.off_warps + (size_t)warps * L.per_warp;
    return L;
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

// per-chunk data of the retirement pipeline; v / p0 / deg / rowf are meaningful in lanes 0..3 (lane = slot within the chunk)
struct SwChunk {
    int v, deg;
    int64_t p0;
    unsigned rowf;
};

// Row tiles 0..NMT-1 of one visit.  Rows: the visit's in-window nodes in ascending local order, node-major
// (row = 3 * position + component); columns: component-major (column tile nt = component, column in tile = node).
template <int NMT>
__device__ __forceinline__ void sweep_rows(const double* __restrict__ Ge, const double (&Bf)[3][2], double w0, double w1,
                                           const uint16_t* __restrict__ rk, unsigned slotw, unsigned perm, int cnt,
                                           double* __restrict__ panel, int ss, int gid, int tig) {
    double A[NMT][2];
    double* base0[NMT];
    double* base1[NMT];
    bool valid[NMT];
    double C[NMT][3][2];
#pragma unroll
    for (int mt = 0; mt < NMT; ++mt) {
        const int R = 8 * mt + gid, s = R / 3, d = R - 3 * s;
        valid[mt] = s < cnt;
        const int na = valid[mt] ? (int)((perm >> (4 * s)) & 7u) : 0;
        const int slot = (int)((slotw >> (4 * na)) & 7u);
        const unsigned rk2 = *reinterpret_cast<const unsigned*>(rk + na * 8 + 2 * tig);
        double* pb = panel + slot * ss + d * 3;
        base0[mt] = pb + (rk2 & 0xffffu) * 9u;
        base1[mt] = pb + (rk2 >> 16) * 9u;
        const double g0 = Ge[(na * 3 + d) * 4 + tig], g1 = Ge[SW_KS + (na * 3 + d) * 4 + tig];
        A[mt][0] = valid[mt] ? w0 * g0 : 0.0;
        A[mt][1] = valid[mt] ? w1 * g1 : 0.0;
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
            C[mt][nt][0] = 0.0;
            C[mt][nt][1] = 0.0;
            if (valid[mt]) {
                C[mt][nt][0] = base0[mt][nt];
                C[mt][nt][1] = base1[mt][nt];
            }
        }
    }
#pragma unroll
    for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) dmma884(C[mt][nt][0], C[mt][nt][1], A[mt][0], Bf[nt][0]);
#pragma unroll
    for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) dmma884(C[mt][nt][0], C[mt][nt][1], A[mt][1], Bf[nt][1]);
#pragma unroll
    for (int mt = 0; mt < NMT; ++mt)
        if (valid[mt]) {
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
                base0[mt][nt] = C[mt][nt][0];
                base1[mt][nt] = C[mt][nt][1];
            }
        }
}

// WM, how the material enters the quadrature weight (the Poisson ratio is one number for the batch):
//   0  not at all: one (lambda, mu) for the batch, applied when a chunk is stored               layout (M,)
//   1  w_q * mu_e                                                                               layout (E,M)
//   2  w_q * E_q / (2 (1 + nu))                                                                 layout (E,Q,M)
template <bool ELIM, int WM>
__global__ void __launch_bounds__(FEB_SW_MAXWARPS * 32, 1)
k_fused_hex8_sweep(feb200_batch b, feb200_sweep_plan plan, SwSmem L, const double* __restrict__ x_vd,
                   const uint16_t* __restrict__ slot_rank, const int32_t* __restrict__ nindices,
                   const uint8_t* __restrict__ vflags, double* __restrict__ values) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_dphi = reinterpret_cast<double*>(smem_raw);                    // [q] DPHI : [n][3]
    double* s_w = s_dphi + 8 * SW_DPHI;                                      // [q]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    unsigned char* mine = smem_raw + L.off_warps + (size_t)warp * L.per_warp;
    uint16_t* rkbuf = reinterpret_cast<uint16_t*>(mine + L.off_rk);          // [2][4][64]
    double* xbuf = reinterpret_cast<double*>(mine + L.off_x);                // [4] XS
    double* Gw = reinterpret_cast<double*>(mine + L.off_g);                  // [4] EL : [ks] KS : [row][4 points], weights at 96
    double* panel = reinterpret_cast<double*>(mine + L.off_panel);           // [slot] ss : [rank][3][3]
    const int ss = L.ss;

    for (int i = tid; i < 8 * 24; i += blockDim.x) s_dphi[(i / 24) * SW_DPHI + (i % 24)] = b.dphi[i];
    if (tid < 8) s_w[tid] = b.W[tid];
    for (int i = lane; i < SW_SLOTS * ss; i += 32) panel[i] = 0.0;
    double u_lam, u_mu;
    isotropic_lame<3>(b.mp, u_lam, u_mu);
    const double half_inv = 0.5 / (1.0 + b.mp[1]);      // mu = E * half_inv
    if (WM > 0) { u_lam = u_lam / u_mu; u_mu = 1.0; }    // mu is in the weights: K = (lam/mu) S' + S'^T + tr(S') I
    __syncthreads();                                     // the only block-wide barrier: basis table ready

    const int n_workers = gridDim.x * (blockDim.x >> 5);
    const int jj = lane >> 3;                            // element of the group this lane loads / computes for
    const int qq = lane & 7;                             // its quadrature point (geometry) or local node (loads)

    for (int seg = blockIdx.x * (blockDim.x >> 5) + warp; seg < plan.n_segments; seg += n_workers) {
        const int v0 = plan.seg_visit_ptr[seg], v1 = plan.seg_visit_ptr[seg + 1];
        const int c0 = plan.seg_chunk_ptr[seg], c1 = plan.seg_chunk_ptr[seg + 1];

        // ---- retirement pipeline: chunk records one to three chunks ahead of their use
        auto load_l1 = [&](int c) {
            SwChunk k;
            k.v = -1; k.deg = 0; k.p0 = 0; k.rowf = 0u;
            if (lane < 4 && c < c1) {
                k.v = plan.chunk_v[(int64_t)c * 4 + lane];
                k.p0 = plan.chunk_p0[(int64_t)c * 4 + lane];
                k.deg = plan.chunk_deg[(int64_t)c * 4 + lane];
            }
            return k;
        };
        auto load_l2 = [&](SwChunk& k, int (&nbr)[4]) {       // neighbour ids of the four vertices, the row flags
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int64_t p0 = __shfl_sync(0xffffffffu, k.p0, s);
                const int deg = __shfl_sync(0xffffffffu, k.deg, s);
                nbr[s] = (lane < deg) ? nindices[p0 + lane] : -1;
            }
            k.rowf = (lane < 4 && k.v >= 0) ? (unsigned)vflags[k.v] : 0u;
        };
        auto load_l3 = [&](const SwChunk& k, const int (&nbr)[4]) {   // constrained components of the neighbours | diag << 3
            unsigned cfl = 0u;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int v = __shfl_sync(0xffffffffu, k.v, s);
                if (nbr[s] >= 0) cfl |= (((unsigned)vflags[nbr[s]] & 7u) | (nbr[s] == v ? 8u : 0u)) << (8 * s);
            }
            return cfl;
        };
        SwChunk k0 = load_l1(c0), k1 = load_l1(c0 + 1), k2 = load_l1(c0 + 2), k3 = load_l1(c0 + 3);
        int nbr1[4] = {-1, -1, -1, -1}, nbr2[4] = {-1, -1, -1, -1};
        unsigned cfl0 = 0u, cfl1 = 0u;
        if (ELIM) {
            int nbr0[4];
            load_l2(k0, nbr0);
            load_l2(k1, nbr1);
            load_l2(k2, nbr2);
            cfl0 = load_l3(k0, nbr0);
            cfl1 = load_l3(k1, nbr1);
        }
        int c_next = c0;                                  // the next chunk to retire; it lives in slots 4 * ((c - c0) & 1) ..

        // ---- visit pipeline: records three groups ahead, vertex ids two, coordinates / slot ranks / material one
        auto load_rec = [&](int g) {
            const int v = g + jj;
            int4 r = make_int4(-1, 0, 0, 0);
            if (v < v1) r = __ldg(reinterpret_cast<const int4*>(plan.visits) + v);
            return r;
        };
        auto load_vid = [&](const int4& r) { return r.x >= 0 ? __ldg(b.conn + (int64_t)r.x * 8 + qq) : -1; };
        auto load_mat = [&](const int4& r) {
            double e = 0.0;
            if (WM == 1 && r.x >= 0) e = b.mp[(int64_t)r.x * b.M];
            if (WM == 2 && r.x >= 0) e = b.mp[((int64_t)r.x * 8 + qq) * b.M];
            return e;
        };
        auto issue_copies = [&](const int4& r, int vid, int pb) {
            if (r.x >= 0) {
                const double* src = x_vd + (int64_t)vid * 3;
                double* dst = xbuf + jj * SW_XS + qq * 3;
                cp_async8(dst, src);
                cp_async8(dst + 1, src + 1);
                cp_async8(dst + 2, src + 2);
                cp_async16(rkbuf + pb * 256 + jj * 64 + qq * 8, slot_rank + (int64_t)r.x * 64 + qq * 8);
            }
            cp_async_commit();
        };
        int4 rec0 = load_rec(v0), rec1 = load_rec(v0 + 4), rec2 = load_rec(v0 + 8);
        int vid0 = load_vid(rec0), vid1 = load_vid(rec1);
        double eq0 = load_mat(rec0);
        int pb = 0;
        __syncwarp();
        issue_copies(rec0, vid0, pb);

        for (int g = v0; g < v1; g += 4) {
            cp_async_wait_all();
            __syncwarp();
            // ---- geometry of the group: lane = (element jj, point qq)
            if (rec0.x >= 0) {
                double gr[24];
                const double dt = hex8_gradients(xbuf + jj * SW_XS, s_dphi + qq * SW_DPHI, gr);
                double* dst = Gw + jj * SW_EL + (qq >> 2) * SW_KS + (qq & 3);
#pragma unroll
                for (int r = 0; r < 24; ++r) dst[r * 4] = gr[r];
                dst[96] = (WM == 0) ? dt * s_w[qq] : dt * s_w[qq] * (eq0 * half_inv);
            }
            __syncwarp();
            // ---- prefetch: coordinates / ranks of the next group (its vertex ids arrived long ago), ids and records further out
            issue_copies(rec1, vid1, pb ^ 1);
            const double eq1 = load_mat(rec1);
            const int vid2 = load_vid(rec2);
            const int4 rec3 = load_rec(g + 12);
            // ---- visits
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                const int el = __shfl_sync(0xffffffffu, rec0.x, 8 * j);
                if (el < 0) break;
                const unsigned slotw = (unsigned)__shfl_sync(0xffffffffu, rec0.y, 8 * j);
                const unsigned perm = (unsigned)__shfl_sync(0xffffffffu, rec0.z, 8 * j);
                const int flags = __shfl_sync(0xffffffffu, rec0.w, 8 * j);
                const int cnt = flags & 0xf;
                const double* Ge = Gw + j * SW_EL;
                const uint16_t* rk = rkbuf + pb * 256 + j * 64;
                double Bf[3][2];
#pragma unroll
                for (int nt = 0; nt < 3; ++nt) {
                    Bf[nt][0] = Ge[(gid * 3 + nt) * 4 + tig];
                    Bf[nt][1] = Ge[SW_KS + (gid * 3 + nt) * 4 + tig];
                }
                const double w0 = Ge[96 + tig], w1 = Ge[SW_KS + 96 + tig];
                if (cnt > 5) sweep_rows<3>(Ge, Bf, w0, w1, rk, slotw, perm, cnt, panel, ss, gid, tig);
                else if (cnt > 2) sweep_rows<2>(Ge, Bf, w0, w1, rk, slotw, perm, cnt, panel, ss, gid, tig);
                else sweep_rows<1>(Ge, Bf, w0, w1, rk, slotw, perm, cnt, panel, ss, gid, tig);
                __syncwarp();
                // ---- chunks completed by this visit: combine, eliminate, store, clear
                for (int nret = (flags >> 4) & 0xf; nret > 0; --nret) {
                    const int h = (c_next - c0) & 1;
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const int v = __shfl_sync(0xffffffffu, k0.v, s);
                        const int deg = __shfl_sync(0xffffffffu, k0.deg, s);
                        const int64_t p0 = __shfl_sync(0xffffffffu, k0.p0, s);
                        const unsigned rf = __shfl_sync(0xffffffffu, k0.rowf, s) & 7u;
                        if (v >= 0 && lane < deg) {
                            double* src = panel + (4 * h + s) * ss + lane * 9;
                            double S[9], K[9];
#pragma unroll
                            for (int c = 0; c < 9; ++c) { S[c] = src[c]; src[c] = 0.0; }
                            const double tr = u_mu * (S[0] + S[4] + S[8]);
#pragma unroll
                            for (int d = 0; d < 3; ++d)
#pragma unroll
                                for (int kk = 0; kk < 3; ++kk)
                                    K[d * 3 + kk] = u_lam * S[d * 3 + kk] + u_mu * S[kk * 3 + d] + (d == kk ? tr : 0.0);
                            if (ELIM) {
                                const unsigned cfd = (cfl0 >> (8 * s)) & 0xffu;
                                const unsigned cf = cfd & 7u, diag = cfd & 8u;
                                if (cf | rf) {
#pragma unroll
                                    for (int i = 0; i < 3; ++i)
#pragma unroll
                                        for (int jx = 0; jx < 3; ++jx)
                                            if (((rf >> i) & 1u) || ((cf >> jx) & 1u))
                                                K[i * 3 + jx] = (((rf >> i) & 1u) && diag && i == jx) ? 1.0 : 0.0;
                                }
                            }
                            double* dst = values + 9 * p0 + 3 * lane;
                            const int row_len = 3 * deg;
#pragma unroll
                            for (int i = 0; i < 3; ++i)
#pragma unroll
                                for (int jx = 0; jx < 3; ++jx) __stcs(dst + (int64_t)i * row_len + jx, K[i * 3 + jx]);
                        }
                    }
                    __syncwarp();
                    // advance the pipeline: chunk c_next + 1 becomes current
                    ++c_next;
                    k0 = k1; k1 = k2; k2 = k3;
                    k3 = load_l1(c_next + 3);
                    if (ELIM) {
                        cfl0 = cfl1;
                        cfl1 = load_l3(k1, nbr2);           // nbr2 was loaded for the chunk that is now k1
                        load_l2(k2, nbr2);
                    }
                }
            }
            rec0 = rec1; rec1 = rec2; rec2 = rec3;
            vid1 = vid2;
            eq0 = eq1;
            pb ^= 1;
        }
        cp_async_wait_all();                              // the last (empty) prefetch group
        __syncwarp();
    }
}

}  // namespace feb

using namespace feb;

extern "C" {

int32_t feb200_fused_sweep_warps(int32_t max_deg) {
    if (max_deg < 1) return 0;
    return sweep_layout(max_deg).warps;
}

int feb200_assemble_fused_hex8_sweep(feb200_stream s, const feb200_batch* bp, const feb200_sweep_plan* plan, const double* x_vd,
                                     const uint16_t* slot_rank, const int32_t* nindices, const uint8_t* vertex_flags,
                                     int32_t material_hints, double* values) {
    FEB_REQUIRE(bp && plan && x_vd && slot_rank && nindices && values, FEB200_ERR_INVALID, "assemble_fused_hex8_sweep: null pointer");
    const feb200_batch& b = *bp;
    FEB_REQUIRE(b.N == 8 && b.Q == 8 && b.D == 3, FEB200_ERR_UNSUPPORTED,
                "assemble_fused_hex8_sweep: needs hex8 with 8 quadrature points");
    FEB_REQUIRE(b.kind == FEB200_KIND_LINEAR && b.material == FEB200_MAT_ISOTROPIC, FEB200_ERR_UNSUPPORTED,
                "assemble_fused_hex8_sweep: needs the linear isotropic material");
    FEB_REQUIRE(b.conn && b.dphi && b.W && b.mp && b.M >= 2, FEB200_ERR_INVALID, "assemble_fused_hex8_sweep: bad batch");
    FEB_REQUIRE(b.layout == FEB200_LAYOUT_M || (material_hints & FEB200_HINT_UNIFORM_POISSON), FEB200_ERR_UNSUPPORTED,
                "assemble_fused_hex8_sweep: needs one Poisson ratio for the batch (layout (M,) or FEB200_HINT_UNIFORM_POISSON)");
    FEB_REQUIRE(plan->seg_visit_ptr && plan->seg_chunk_ptr && plan->visits && plan->chunk_v && plan->chunk_p0 &&
                    plan->chunk_deg && plan->n_segments >= 0,
                FEB200_ERR_INVALID, "assemble_fused_hex8_sweep: bad plan");
    FEB_REQUIRE(plan->max_deg >= 1 && plan->max_deg <= 32, FEB200_ERR_INVALID,
                "assemble_fused_hex8_sweep: plan max_deg=%d out of range (1..32)", plan->max_deg);
    if (plan->n_segments == 0) return FEB200_OK;
    const SwSmem L = sweep_layout(plan->max_deg);
    FEB_REQUIRE(L.warps >= 1, FEB200_ERR_CAPACITY, "assemble_fused_hex8_sweep: a worker needs %zu bytes of shared memory", L.per_warp);
    cudaStream_t st = as_stream(s);
    const int64_t want = ceil_div(plan->n_segments, L.warps);
    const int grid = (int)(want < sm_count() ? want : sm_count());
    const int wm = (b.layout == FEB200_LAYOUT_EQM) ? 2 : (b.layout == FEB200_LAYOUT_EM ? 1 : 0);
    static bool attr_set[2][3] = {{false, false, false}, {false, false, false}};
#define FEB_SW_LAUNCH(EL, WMODE)                                                                                       \
    {                                                                                                                  \
        auto kern = k_fused_hex8_sweep<EL, WMODE>;                                                                     \
        if (!attr_set[EL ? 1 : 0][WMODE]) {                                                                            \
            FEB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SW_SMEM_MAX));       \
            attr_set[EL ? 1 : 0][WMODE] = true;                                                                        \
        }                                                                                                              \
        kern<<<grid, L.warps * 32, L.total, st>>>(b, *plan, L, x_vd, slot_rank, nindices, vertex_flags, values);       \
    }
#define FEB_SW_DISPATCH(EL)                                                                                            \
    if (wm == 0) FEB_SW_LAUNCH(EL, 0) else if (wm == 1) FEB_SW_LAUNCH(EL, 1) else FEB_SW_LAUNCH(EL, 2)
    if (vertex_flags) { FEB_SW_DISPATCH(true) } else { FEB_SW_DISPATCH(false) }
#undef FEB_SW_DISPATCH
#undef FEB_SW_LAUNCH
    FEB_CHECK_LAUNCH("assemble_fused_hex8_sweep");
    return FEB200_OK;
}

}  // extern "C"
