// Deterministic sparse assembly and the COO / CSR helpers of the reference.
//   calculate_jacobian_wo_dirichlet        /root/reference/src/fe_jax/fea.py:613-660
//   coo_arrays_sum_duplicates_jit          src/fe_jax/sparse_matrix.py:44-92
//   coo_to_csr                             src/fe_jax/sparse_matrix.py:113-153
//   apply_dirichlet_bcs_lhs                src/fe_jax/sparse_matrix.py:157-184
//   calculate_jacobian_diag_wo_dirichlet   src/fe_jax/fea.py:759-797
//
// The reference concatenates E*(N*U)^2 COO triples, sorts them by (row, col) and sums duplicates
// with bincount -- i.e. every CSR entry is the sum of its element contributions in ascending element
// order.  Here every vertex owns its U CSR rows ("row panel") and pulls the same contributions, in
// the same order, through the assembly map (its slots e*N+a, ascending) and the slot-rank map, so
// no atomics are involved and the result is bit-reproducible.
#include <stdlib.h>

#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace feb {

// Final store of one vertex's row panel from the shared-memory accumulators (panel layout), with the
// optional fused Dirichlet elimination (apply_dirichlet_bcs_lhs, sparse_matrix.py:157-184).  The constraint
// flags of the <= deg neighbours are gathered once per vertex into `flags` (per-warp shared memory: bit j =
// component j constrained, bit 7 = the neighbour is the row vertex itself), not once per matrix entry.
template <int U>
__device__ __forceinline__ void store_panel(int lane, int64_t a, int deg, int64_t p0, const double* acc,
                                            double* __restrict__ values, int accumulate,
                                            const int32_t* __restrict__ nindices, const double* __restrict__ mask,
                                            const uint8_t* __restrict__ vflags, unsigned char* flags) {
    constexpr int U2 = U * U;
    const int row_len = U * deg;
    double* dst = values + (int64_t)U2 * p0;
    const bool elim = mask || vflags;
    if (elim) {
        for (int r = lane; r < deg; r += 32) {
            const int64_t colv = nindices[p0 + r];
            unsigned int f = (colv == a) ? 0x80u : 0u;
            if (vflags) f |= (unsigned int)vflags[colv] & ((1u << U) - 1u);
            else {
#pragma unroll
                for (int j = 0; j < U; ++j) f |= (mask[colv * U + j] != 0.0) ? (1u << j) : 0u;
            }
            flags[r] = (unsigned char)f;
        }
        __syncwarp();
    }
#pragma unroll
    for (int i = 0; i < U; ++i) {                       // row by row: no division by the run-time row length
        const bool rm = vflags ? (((unsigned int)vflags[a] >> i) & 1u) : (mask ? (mask[a * U + i] != 0.0) : false);
        for (int c = lane; c < row_len; c += 32) {
            const int t = i * row_len + c;
            double v = acc[t];
            if (accumulate) v += dst[t];
            if (elim) {
                const int r = c / U, j = c - r * U;
                const unsigned int f = flags[r];
                const bool cm = (f >> j) & 1u;
                if (rm || cm) v = (rm && (f & 0x80u) && i == j) ? 1.0 : 0.0;
            }
            dst[t] = v;
        }
    }
}

// one warp per vertex; accumulators in shared memory in PANEL layout: acc[i*(U*deg) + U*r + j].
// Generic variant: any N (chunk = N*U*U doubles per incident slot), slots processed one after the other.
template <int U>
__global__ void k_assemble_csr(const unsigned long long* __restrict__ row_ind, const unsigned long long* __restrict__ col_ind,
                               int64_t V, int N, const double* __restrict__ K_e, const uint16_t* __restrict__ slot_rank,
                               const int64_t* __restrict__ nindptr, int max_deg, double* __restrict__ values, int accumulate,
                               const int32_t* __restrict__ nindices, const double* __restrict__ mask,
                               const uint8_t* __restrict__ vflags) {
    constexpr int U2 = U * U;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    double* acc = smem + (size_t)warp * max_deg * U2;
    unsigned char* flags = reinterpret_cast<unsigned char*>(smem + (size_t)wpb * max_deg * U2) + (size_t)warp * ((max_deg + 7) & ~7);
    const int items = N * U;                                   // (block, row-of-block) items per slot chunk
    for (int64_t a = (int64_t)blockIdx.x * wpb + warp; a < V; a += (int64_t)gridDim.x * wpb) {
        const int64_t p0 = nindptr[a];
        const int deg = (int)(nindptr[a + 1] - p0);
        const int row_len = U * deg;
        for (int t = lane; t < deg * U2; t += 32) acc[t] = 0.0;
        __syncwarp();
        const unsigned long long s0 = row_ind[a], s1 = row_ind[a + 1];
        for (unsigned long long k = s0; k < s1; ++k) {          // ascending slots == ascending elements
            const unsigned long long slot = col_ind[k];
            const double* chunk = K_e + slot * (unsigned long long)(N * U2);
            const uint16_t* rk = slot_rank + slot * (unsigned long long)N;
            for (int t = lane; t < items; t += 32) {
                const int bn = t / U, i = t - bn * U;
                double* dstp = acc + i * row_len + U * (int)rk[bn];   // distinct bn -> distinct rank: no conflicts
#pragma unroll
                for (int j = 0; j < U; ++j) dstp[j] += chunk[bn * U2 + i * U + j];
            }
            __syncwarp();
        }
        store_panel<U>(lane, a, deg, p0, acc, values, accumulate, nindices, mask, vflags, flags);
        __syncwarp();
    }
}

// Fast variant for N*U <= 32 (tri3, quad4, tet4, hex8): lane <-> (local node bn, block row i).  The chunks
// of up to PD incident slots are loaded first (PD*U independent 8-byte loads in flight per lane), then
// added in slot order.
template <int U, int PD>
__global__ void __launch_bounds__(256, 4) k_assemble_csr_pre(const unsigned long long* __restrict__ row_ind,
                                                          const unsigned long long* __restrict__ col_ind, int64_t V, int N,
                                                          const double* __restrict__ K_e,
                                                          const uint16_t* __restrict__ slot_rank,
                                                          const int64_t* __restrict__ nindptr, int max_deg,
                                                          double* __restrict__ values, int accumulate,
                                                          const int32_t* __restrict__ nindices,
                                                          const double* __restrict__ mask,
                                                          const uint8_t* __restrict__ vflags) {
    constexpr int U2 = U * U;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    double* acc = smem + (size_t)warp * max_deg * U2;
    unsigned char* flags = reinterpret_cast<unsigned char*>(smem + (size_t)wpb * max_deg * U2) + (size_t)warp * ((max_deg + 7) & ~7);
    const int items = N * U;
    const bool has_item = lane < items;
    const int bn = lane / U, i = lane - bn * U;
    for (int64_t a = (int64_t)blockIdx.x * wpb + warp; a < V; a += (int64_t)gridDim.x * wpb) {
        const int64_t p0 = nindptr[a];
        const int deg = (int)(nindptr[a + 1] - p0);
        const int row_len = U * deg;
        for (int t = lane; t < deg * U2; t += 32) acc[t] = 0.0;
        const unsigned long long s0 = row_ind[a], s1 = row_ind[a + 1];
        for (unsigned long long kb = s0; kb < s1; kb += PD) {
            const int cnt = (int)((s1 - kb) < (unsigned long long)PD ? (s1 - kb) : PD);
            const unsigned long long my_slot = (lane < cnt) ? col_ind[kb + lane] : 0ull;
            double v[PD][U];
            int rk[PD];
#pragma unroll
            for (int k = 0; k < PD; ++k) {
                const unsigned long long slot = __shfl_sync(0xffffffffu, my_slot, k);
                if (k < cnt && has_item) {
                    const double* src = K_e + slot * (unsigned long long)(N * U2) + bn * U2 + i * U;
#pragma unroll
                    for (int j = 0; j < U; ++j) v[k][j] = __ldcs(src + j);          // streamed once
                    rk[k] = (int)slot_rank[slot * (unsigned long long)N + bn];
                }
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < PD; ++k) {
                if (k < cnt && has_item) {
                    double* dstp = acc + i * row_len + U * rk[k];
#pragma unroll
                    for (int j = 0; j < U; ++j) dstp[j] += v[k][j];
                }
                __syncwarp();
            }
        }
        store_panel<U>(lane, a, deg, p0, acc, values, accumulate, nindices, mask, vflags, flags);
        __syncwarp();
    }
}


// ------------------------------------------------------------------------------------------------
// Streaming variant for hex8 row blocks (N = 8, U = 3: 576-byte blocks K_e[(e*8+a)][0..8][3][3]) and vertex degree <= 32:
// a warp owns a vertex, lane = CSR block rank holds the nine sums of its block in registers.  The row blocks of up to 8
// incident slots (4608 bytes) and their slot-rank rows arrive through cp.async, two vertices ahead (12 warps per SM keep
// ~110 KB in flight); the slot ids and the CSR offsets are prefetched one and two vertices further ahead in registers.
// Contributions are added in ascending slot order: bit-identical to k_assemble_csr_pre and to the reference's stable sort +
// bincount.  Measured at 200^3 (36.9 GB of K_e in, 15.6 GB of values out): 11.4 ms with k_assemble_csr_pre (4.6 TB/s).
// ------------------------------------------------------------------------------------------------
#ifndef FEB_AS_WARPS
#define FEB_AS_WARPS 12
#endif
constexpr int AS_WARPS = FEB_AS_WARPS;
constexpr int AS_NB = 3;                           // buffers per warp: vertices k, k+1, k+2
constexpr int AS_SL = 8;                           // slots per buffer
constexpr int AS_BUF = AS_SL * 576 + AS_SL * 16;   // bytes: row blocks + rank rows

__device__ __forceinline__ void as_cp_async16(void* dst, const void* src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}

template <bool ELIM, bool ACC>
__global__ void __launch_bounds__(AS_WARPS * 32, 1)
k_assemble_hex8_stream(const unsigned long long* __restrict__ row_ind, const unsigned long long* __restrict__ col_ind, int64_t V,
                       const double* __restrict__ K_e, const uint16_t* __restrict__ slot_rank,
                       const int64_t* __restrict__ nindptr, double* __restrict__ values,
                       const int32_t* __restrict__ nindices, const double* __restrict__ mask,
                       const uint8_t* __restrict__ vflags) {
    extern __shared__ __align__(16) unsigned char as_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* my = as_smem + (size_t)warp * AS_NB * AS_BUF;
    const int64_t W = (int64_t)gridDim.x * AS_WARPS;
    const int64_t w0 = (int64_t)blockIdx.x * AS_WARPS + warp;
    const int64_t n_mine = (w0 < V) ? (V - w0 + W - 1) / W : 0;          // vertices w0, w0 + W, ...
    const unsigned char* kbase = reinterpret_cast<const unsigned char*>(K_e);
    const unsigned char* rbase = reinterpret_cast<const unsigned char*>(slot_rank);

    // lanes 0, 1: row_ind[a], row_ind[a + 1]; lanes 2, 3: nindptr[a], nindptr[a + 1]
    auto load_meta = [&](int64_t k) -> int64_t {
        if (k >= n_mine || lane >= 4) return 0;
        const int64_t a = w0 + k * W;
        return lane < 2 ? (int64_t)row_ind[a + lane] : nindptr[a + lane - 2];
    };
    auto load_slot = [&](int64_t k, int64_t meta) -> int64_t {
        const int64_t s0 = __shfl_sync(0xffffffffu, meta, 0), s1 = __shfl_sync(0xffffffffu, meta, 1);
        return (k < n_mine && lane < AS_SL && s0 + lane < s1) ? (int64_t)col_ind[s0 + lane] : 0;
    };
    auto issue = [&](int64_t k, int64_t meta, int64_t my_slot) {
        const int64_t s0 = __shfl_sync(0xffffffffu, meta, 0), s1 = __shfl_sync(0xffffffffu, meta, 1);
        const int cnt = (k < n_mine) ? (int)((s1 - s0) < AS_SL ? (s1 - s0) : AS_SL) : 0;
        unsigned char* buf = my + (size_t)(k % AS_NB) * AS_BUF;
#pragma unroll
        for (int i = 0; i < 9; ++i) {                                      // 8 x 36 sixteen-byte pieces over 32 lanes
            const int t = lane + 32 * i, j = t / 36, c = t - 36 * j;
            const int64_t slot = __shfl_sync(0xffffffffu, my_slot, j);
            if (j < cnt) as_cp_async16(buf + j * 576 + c * 16, kbase + slot * 576 + c * 16);
        }
        if (lane < cnt) as_cp_async16(buf + AS_SL * 576 + lane * 16, rbase + my_slot * 16);
        asm volatile("cp.async.commit_group;" ::: "memory");               // always: the group count stays uniform
    };

    // Dirichlet elimination: neighbour id (two vertices ahead) and the mask entries of the neighbour / of the row vertex
    // (one vertex ahead) are prefetched into registers like the slot ids, so nothing is gathered at store time
    auto load_nb = [&](int64_t k, int64_t meta) -> int {
        const int64_t p0 = __shfl_sync(0xffffffffu, meta, 2), p1 = __shfl_sync(0xffffffffu, meta, 3);
        return (ELIM && k < n_mine && p0 + lane < p1) ? nindices[p0 + lane] : -1;
    };
    auto load_mask = [&](int64_t k, int nb, double (&mk)[3], double& rowm, unsigned& fb, unsigned& frow) {
        mk[0] = mk[1] = mk[2] = 0.0;
        rowm = 0.0;
        fb = frow = 0u;
        if (ELIM && k < n_mine) {
            if (vflags) {                       // one byte per vertex instead of three doubles
                if (nb >= 0) fb = vflags[nb];
                if (lane == 0) frow = vflags[w0 + k * W];
            } else {
                if (nb >= 0) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) mk[j] = mask[(int64_t)nb * 3 + j];
                }
                if (lane < 3) rowm = mask[(w0 + k * W) * 3 + lane];
            }
        }
    };
    int64_t meta0 = load_meta(0), meta1 = load_meta(1), meta2 = load_meta(2), meta3 = load_meta(3);
    int64_t sl0 = load_slot(0, meta0), sl1 = load_slot(1, meta1), sl2 = load_slot(2, meta2);
    int nb0 = load_nb(0, meta0), nb1 = load_nb(1, meta1);
    double mk0[3], rowm0;
    unsigned fb0, frow0;
    load_mask(0, nb0, mk0, rowm0, fb0, frow0);
    issue(0, meta0, sl0);
    issue(1, meta1, sl1);
    for (int64_t k = 0; k < n_mine; ++k) {
        issue(k + 2, meta2, sl2);
        const int64_t sl3 = load_slot(k + 3, meta3);
        const int64_t meta4 = load_meta(k + 4);
        const int nb2 = load_nb(k + 2, meta2);
        double mk1[3], rowm1;
        unsigned fb1, frow1;
        load_mask(k + 1, nb1, mk1, rowm1, fb1, frow1);
        asm volatile("cp.async.wait_group 2;" ::: "memory");              // the group of vertex k has landed
        __syncwarp();
        const int64_t a = w0 + k * W;
        const int64_t s0 = __shfl_sync(0xffffffffu, meta0, 0), s1 = __shfl_sync(0xffffffffu, meta0, 1);
        const int64_t p0 = __shfl_sync(0xffffffffu, meta0, 2), p1 = __shfl_sync(0xffffffffu, meta0, 3);
        const int deg = (int)(p1 - p0);
        const int cnt = (int)((s1 - s0) < AS_SL ? (s1 - s0) : AS_SL);
        const unsigned char* buf = my + (size_t)(k % AS_NB) * AS_BUF;
        double acc[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) acc[c] = 0.0;
        for (int j = 0; j < cnt; ++j) {                                   // ascending slots == ascending elements
            const uint4 rr = *reinterpret_cast<const uint4*>(buf + AS_SL * 576 + j * 16);
            const unsigned rw[4] = {rr.x, rr.y, rr.z, rr.w};
            int src = -1;
#pragma unroll
            for (int bn = 0; bn < 8; ++bn)
                if ((int)((rw[bn >> 1] >> (16 * (bn & 1))) & 0xffffu) == lane) src = bn;
            if (src >= 0) {
                const double* blk = reinterpret_cast<const double*>(buf + j * 576) + src * 9;
#pragma unroll
                for (int c = 0; c < 9; ++c) acc[c] += blk[c];
            }
        }
        for (int64_t sx = s0 + AS_SL; sx < s1; ++sx) {                     // vertices with more than 8 incident slots: plain loads
            const unsigned long long slot = col_ind[sx];
            int src = -1;
#pragma unroll
            for (int bn = 0; bn < 8; ++bn)
                if ((int)slot_rank[slot * 8 + bn] == lane) src = bn;
            if (src >= 0) {
#pragma unroll
                for (int c = 0; c < 9; ++c) acc[c] += __ldcs(K_e + slot * 72 + src * 9 + c);
            }
        }
        bool rm3[3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
            rm3[i] = ELIM && (vflags ? ((__shfl_sync(0xffffffffu, frow0, 0) >> i) & 1u) != 0u
                                                : __shfl_sync(0xffffffffu, rowm0, i) != 0.0);
        if (lane < deg) {
            double* dst = values + 9 * p0 + 3 * lane;
            const int64_t row_len = 3 * (int64_t)deg;
            if (ACC) {
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) acc[i * 3 + j] += __ldcs(dst + i * row_len + j);
            }
            if (ELIM) {
                // (per-block flag rows fetched with the row blocks, as the sweep kernel uses them, were measured: no difference)
                const unsigned cf = vflags ? (fb0 & 7u)
                                           : ((mk0[0] != 0.0 ? 1u : 0u) | (mk0[1] != 0.0 ? 2u : 0u) | (mk0[2] != 0.0 ? 4u : 0u));
                const unsigned rf = (rm3[0] ? 1u : 0u) | (rm3[1] ? 2u : 0u) | (rm3[2] ? 4u : 0u);
                const bool diag = (nb0 == a);
                if (cf | rf) {
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int j = 0; j < 3; ++j)
                            if (((rf >> i) & 1u) || ((cf >> j) & 1u)) acc[i * 3 + j] = (((rf >> i) & 1u) && diag && i == j) ? 1.0 : 0.0;
                }
            }
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) __stcs(dst + i * row_len + j, acc[i * 3 + j]);
        }
        __syncwarp();                                                      // buffer k % NB is free for vertex k + 3
        meta0 = meta1; meta1 = meta2; meta2 = meta3; meta3 = meta4;
        sl0 = sl1; sl1 = sl2; sl2 = sl3;
        nb0 = nb1; nb1 = nb2;
        mk0[0] = mk1[0]; mk0[1] = mk1[1]; mk0[2] = mk1[2]; rowm0 = rowm1; fb0 = fb1; frow0 = frow1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <int U>
__global__ void k_dirichlet_lhs(int64_t V, const int64_t* __restrict__ nindptr, const int32_t* __restrict__ nindices,
                                const double* __restrict__ mask, double* __restrict__ values) {
    constexpr int U2 = U * U;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    for (int64_t a = (int64_t)blockIdx.x * wpb + warp; a < V; a += (int64_t)gridDim.x * wpb) {
        const int64_t p0 = nindptr[a];
        const int deg = (int)(nindptr[a + 1] - p0);
        const int row_len = U * deg;
        double* dst = values + (int64_t)U2 * p0;
        double mrow[U];
        bool any_row = false;
#pragma unroll
        for (int i = 0; i < U; ++i) { mrow[i] = mask[a * U + i]; any_row |= (mrow[i] != 0.0); }
        for (int t = lane; t < U * row_len; t += 32) {
            const int i = t / row_len, c = t - i * row_len;
            const int r = c / U, j = c - r * U;
            const int64_t colv = nindices[p0 + r];
            const bool rm = (mrow[i] != 0.0);
            const bool cm = (mask[colv * U + j] != 0.0);
            if (rm || cm) dst[t] = (rm && colv == a && i == j) ? 1.0 : 0.0;
        }
        (void)any_row;
    }
}

__global__ void k_dirichlet_lhs_coo(int64_t nnz, const int64_t* __restrict__ row, const int64_t* __restrict__ col,
                                    const double* __restrict__ mask, const double* __restrict__ in, double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = row[i], c = col[i];
        const bool rm = mask[r] != 0.0, cm = mask[c] != 0.0;
        double v = (rm || cm) ? 0.0 : in[i];
        if (r == c && rm) v = 1.0;
        out[i] = v;
    }
}

template <int U>
__global__ void k_csr_diagonal(int64_t V, const int64_t* __restrict__ nindptr, const int32_t* __restrict__ diag_rank,
                               const double* __restrict__ values, double* __restrict__ diag) {
    for (int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; a < V; a += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p0 = nindptr[a];
        const int deg = (int)(nindptr[a + 1] - p0);
        const int r = diag_rank[a];
#pragma unroll
        for (int i = 0; i < U; ++i)
            diag[a * U + i] = (r >= 0) ? values[(int64_t)U * U * p0 + (int64_t)i * U * deg + U * r + i] : 0.0;
    }
}

__global__ void k_mark_heads(int64_t n, const int64_t* __restrict__ row, const int64_t* __restrict__ col, int32_t* head) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        head[i] = (i == 0 || row[i] != row[i - 1] || col[i] != col[i - 1]) ? 1 : 0;
}
__global__ void k_minus_one(int64_t n, int64_t* seg) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) seg[i] -= 1;
}
// thread i owning a segment head adds its run sequentially (stable order == bincount on CPU)
__global__ void k_segment_sum(int64_t n, const int64_t* __restrict__ row, const int64_t* __restrict__ col,
                              const double* __restrict__ data, const int32_t* __restrict__ head,
                              const int64_t* __restrict__ seg, int64_t n_out, double* __restrict__ out_data,
                              int64_t* __restrict__ out_row, int64_t* __restrict__ out_col) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (!head[i]) continue;
        const int64_t o = seg[i];
        if (o >= n_out) continue;
        double acc = data[i];
        for (int64_t j = i + 1; j < n && !head[j]; ++j) acc += data[j];
        out_data[o] = acc;
        if (out_row) out_row[o] = row[i];
        if (out_col) out_col[o] = col[i];
    }
}
// indptr[r] = first index with row >= r (rows sorted ascending)
__global__ void k_rows_to_indptr(int64_t nnz, const int64_t* __restrict__ row, int64_t n_rows, int64_t* __restrict__ indptr) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r <= n_rows; r += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = nnz;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (row[mid] < r) lo = mid + 1; else hi = mid;
        }
        indptr[r] = lo;
    }
}

}  // namespace feb

using namespace feb;

extern "C" {

static int assemble_csr_impl(feb200_stream s, const feb200_batch_map* map, int64_t V, int32_t U, const double* K_e,
                             const uint16_t* slot_rank, const int64_t* nindptr, int32_t max_deg, double* values,
                             int32_t accumulate, const int32_t* nindices, const double* mask, const uint8_t* vflags) {
    FEB_REQUIRE(map && map->row_ind && map->col_ind && K_e && slot_rank && nindptr && values, FEB200_ERR_INVALID,
                "assemble_csr: null pointer");
    FEB_REQUIRE(!(mask || vflags) || nindices, FEB200_ERR_INVALID, "assemble_csr: mask / vertex flags given without nindices");
    FEB_REQUIRE(max_deg > 0 && max_deg <= 65535, FEB200_ERR_INVALID, "assemble_csr: max_deg=%d out of range", max_deg);
    if (V == 0) return FEB200_OK;
    cudaStream_t st = as_stream(s);
    const unsigned long long* ri = reinterpret_cast<const unsigned long long*>(map->row_ind);
    const unsigned long long* ci = reinterpret_cast<const unsigned long long*>(map->col_ind);
    int wpb = 8;
    size_t per_warp = (size_t)max_deg * U * U * sizeof(double) + (size_t)((max_deg + 7) & ~7);
    while (wpb > 1 && per_warp * wpb > 64 * 1024) wpb >>= 1;
    const size_t smem = per_warp * wpb;
    FEB_REQUIRE(smem <= 200 * 1024, FEB200_ERR_CAPACITY, "assemble_csr: vertex degree %d needs %zu bytes of shared memory", max_deg, smem);
    int64_t g = ceil_div(V, wpb);
    const int64_t cap = (int64_t)sm_count() * 16;
    if (g > cap) g = cap;
    // hex8 row blocks: streaming kernel (cp.async pipeline); FEB200_ASSEMBLE_STREAM=0 keeps the plain-load kernel (A/B)
    static const bool stream_ok = !(getenv("FEB200_ASSEMBLE_STREAM") && getenv("FEB200_ASSEMBLE_STREAM")[0] == '0');
    if (stream_ok && map->N == 8 && U == 3 && max_deg <= 32 && (((uintptr_t)K_e | (uintptr_t)slot_rank) & 15) == 0) {
        const size_t sm = (size_t)AS_WARPS * AS_NB * AS_BUF;
        int64_t gs = ceil_div(V, AS_WARPS);
        if (gs > sm_count()) gs = sm_count();
#define FEB_AS_STREAM(EL, AC)                                                                                          \
        {                                                                                                              \
            auto kern = k_assemble_hex8_stream<EL, AC>;                                                                \
            static bool attr_set = false;                                                                              \
            if (!attr_set) {                                                                                           \
                FEB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));            \
                attr_set = true;                                                                                       \
            }                                                                                                          \
            kern<<<(int)gs, AS_WARPS * 32, sm, st>>>(ri, ci, V, K_e, slot_rank, nindptr, values, nindices, mask, vflags);      \
        }
        if (mask || vflags) { if (accumulate) FEB_AS_STREAM(true, true) else FEB_AS_STREAM(true, false) }
        else { if (accumulate) FEB_AS_STREAM(false, true) else FEB_AS_STREAM(false, false) }
#undef FEB_AS_STREAM
        FEB_CHECK_LAUNCH("assemble_csr");
        return FEB200_OK;
    }
    const bool fast = (map->N * U <= 32);
#define FEB_ASM(UU)                                                                                                    \
    {                                                                                                                  \
        if (fast) {                                                                                                    \
            auto kern = k_assemble_csr_pre<UU, 4>;                                                                     \
            if (smem > 48 * 1024) FEB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            kern<<<(int)g, wpb * 32, smem, st>>>(ri, ci, V, map->N, K_e, slot_rank, nindptr, max_deg, values, accumulate, \
                                                 nindices, mask, vflags);                                              \
        } else {                                                                                                       \
            auto kern = k_assemble_csr<UU>;                                                                            \
            if (smem > 48 * 1024) FEB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            kern<<<(int)g, wpb * 32, smem, st>>>(ri, ci, V, map->N, K_e, slot_rank, nindptr, max_deg, values, accumulate, \
                                                 nindices, mask, vflags);                                              \
        }                                                                                                              \
    }
    if (U == 2) FEB_ASM(2)
    else if (U == 3) FEB_ASM(3)
    else if (U == 1) FEB_ASM(1)
    else FEB_REQUIRE(false, FEB200_ERR_UNSUPPORTED, "assemble_csr: U=%d not supported", U);
#undef FEB_ASM
    FEB_CHECK_LAUNCH("assemble_csr");
    return FEB200_OK;
}

int feb200_assemble_csr(feb200_stream s, const feb200_batch_map* map, int64_t V, int32_t U, const double* K_e,
                        const uint16_t* slot_rank, const int64_t* nindptr, int32_t max_deg, double* values,
                        int32_t accumulate, const int32_t* nindices, const double* mask) {
    return assemble_csr_impl(s, map, V, U, K_e, slot_rank, nindptr, max_deg, values, accumulate, nindices, mask, nullptr);
}

int feb200_assemble_csr_flags(feb200_stream s, const feb200_batch_map* map, int64_t V, int32_t U, const double* K_e,
                              const uint16_t* slot_rank, const int64_t* nindptr, int32_t max_deg, double* values,
                              int32_t accumulate, const int32_t* nindices, const uint8_t* vertex_flags) {
    FEB_REQUIRE(U <= 7, FEB200_ERR_UNSUPPORTED, "assemble_csr_flags: U=%d does not fit the flag byte", U);
    return assemble_csr_impl(s, map, V, U, K_e, slot_rank, nindptr, max_deg, values, accumulate, nindices, nullptr, vertex_flags);
}

int feb200_dirichlet_lhs(feb200_stream s, int64_t V, int32_t U, const int64_t* nindptr, const int32_t* nindices,
                         const double* mask, double* values) {
    FEB_REQUIRE(nindptr && nindices && mask && values, FEB200_ERR_INVALID, "dirichlet_lhs: null pointer");
    if (V == 0) return FEB200_OK;
    cudaStream_t st = as_stream(s);
    int64_t g = ceil_div(V, 8);
    const int64_t cap = (int64_t)sm_count() * 16;
    if (g > cap) g = cap;
    if (U == 2) k_dirichlet_lhs<2><<<(int)g, 256, 0, st>>>(V, nindptr, nindices, mask, values);
    else if (U == 3) k_dirichlet_lhs<3><<<(int)g, 256, 0, st>>>(V, nindptr, nindices, mask, values);
    else if (U == 1) k_dirichlet_lhs<1><<<(int)g, 256, 0, st>>>(V, nindptr, nindices, mask, values);
    else FEB_REQUIRE(false, FEB200_ERR_UNSUPPORTED, "dirichlet_lhs: U=%d not supported", U);
    FEB_CHECK_LAUNCH("dirichlet_lhs");
    return FEB200_OK;
}

int feb200_dirichlet_lhs_coo(feb200_stream s, int64_t nnz, const int64_t* row, const int64_t* col, const double* mask,
                             const double* data_in, double* data_out) {
    FEB_REQUIRE(row && col && mask && data_in && data_out, FEB200_ERR_INVALID, "dirichlet_lhs_coo: null pointer");
    if (nnz == 0) return FEB200_OK;
    k_dirichlet_lhs_coo<<<grid_for(nnz, 256), 256, 0, as_stream(s)>>>(nnz, row, col, mask, data_in, data_out);
    FEB_CHECK_LAUNCH("dirichlet_lhs_coo");
    return FEB200_OK;
}

int feb200_csr_diagonal(feb200_stream s, int64_t V, int32_t U, const int64_t* nindptr, const int32_t* diag_rank,
                        const double* values, double* diag) {
    FEB_REQUIRE(nindptr && diag_rank && values && diag, FEB200_ERR_INVALID, "csr_diagonal: null pointer");
    if (V == 0) return FEB200_OK;
    cudaStream_t st = as_stream(s);
    const int g = grid_for(V, 128);
    if (U == 2) k_csr_diagonal<2><<<g, 128, 0, st>>>(V, nindptr, diag_rank, values, diag);
    else if (U == 3) k_csr_diagonal<3><<<g, 128, 0, st>>>(V, nindptr, diag_rank, values, diag);
    else if (U == 1) k_csr_diagonal<1><<<g, 128, 0, st>>>(V, nindptr, diag_rank, values, diag);
    else FEB_REQUIRE(false, FEB200_ERR_UNSUPPORTED, "csr_diagonal: U=%d not supported", U);
    FEB_CHECK_LAUNCH("csr_diagonal");
    return FEB200_OK;
}

int feb200_coo_mark_heads(feb200_stream s, int64_t n, const int64_t* row, const int64_t* col, int32_t* head) {
    FEB_REQUIRE(row && col && head, FEB200_ERR_INVALID, "coo_mark_heads: null pointer");
    if (n == 0) return FEB200_OK;
    k_mark_heads<<<grid_for(n, 256), 256, 0, as_stream(s)>>>(n, row, col, head);
    FEB_CHECK_LAUNCH("coo_mark_heads");
    return FEB200_OK;
}

int feb200_coo_scan_heads(feb200_stream s, int64_t n, const int32_t* head, int64_t* seg, void* ws, size_t ws_bytes) {
    FEB_REQUIRE(head && seg, FEB200_ERR_INVALID, "coo_scan_heads: null pointer");
    FEB_REQUIRE(ws_bytes >= feb200_scan_workspace_bytes(n), FEB200_ERR_WORKSPACE, "coo_scan_heads: workspace too small");
    if (n == 0) return FEB200_OK;
    cudaStream_t st = as_stream(s);
    size_t bytes = ws_bytes;
    FEB_CUDA(cub::DeviceScan::InclusiveSum(ws, bytes, head, seg, n, st));
    k_minus_one<<<grid_for(n, 256), 256, 0, st>>>(n, seg);
    FEB_CHECK_LAUNCH("coo_scan_heads");
    return FEB200_OK;
}

int feb200_coo_segment_sum(feb200_stream s, int64_t n, const int64_t* row, const int64_t* col, const double* data,
                           const int32_t* head, const int64_t* seg, int64_t n_out, double* out_data, int64_t* out_row,
                           int64_t* out_col) {
    FEB_REQUIRE(row && col && data && head && seg && out_data, FEB200_ERR_INVALID, "coo_segment_sum: null pointer");
    if (n == 0) return FEB200_OK;
    k_segment_sum<<<grid_for(n, 256), 256, 0, as_stream(s)>>>(n, row, col, data, head, seg, n_out, out_data, out_row, out_col);
    FEB_CHECK_LAUNCH("coo_segment_sum");
    return FEB200_OK;
}

int feb200_coo_rows_to_indptr(feb200_stream s, int64_t nnz, const int64_t* row, int64_t n_rows, int64_t* indptr) {
    FEB_REQUIRE((row || nnz == 0) && indptr, FEB200_ERR_INVALID, "coo_rows_to_indptr: null pointer");
    k_rows_to_indptr<<<grid_for(n_rows + 1, 256), 256, 0, as_stream(s)>>>(nnz, row, n_rows, indptr);
    FEB_CHECK_LAUNCH("coo_rows_to_indptr");
    return FEB200_OK;
}

}  // extern "C"
