// CSR SpMV on the node-pattern matrix and the device-resident conjugate-gradient loop.
//   cg / _cg_solve             /root/reference/src/fe_jax/solve_cg.py:20-63
//   linear_solve (CG branches) src/fe_jax/sparse_linear_solve.py:339-376  (Jacobi :301-306)
//
// The reference applies A(p) as a JVP through the whole assembly in every iteration
// (sparse_linear_solve.py:251-258); here A is the assembled Dirichlet-eliminated CSR and one
// iteration is three launches:
//   1. Ap = A p                       + partial sums of p.Ap      (last block: pAp)
//   2. x += a p ; r -= a Ap ; z = M r + partial sums r.z, r.r     (last block: gamma', loop test, hist)
//   3. p = z + (gamma'/gamma) p
// All scalars live in device memory; kernels return immediately once `done` is set, so a caller can
// enqueue `maxiter` iterations without synchronising.  All reductions use a fixed order
// (common.cuh), so repeated runs are bit-identical.
#include "common.cuh"

namespace feb {

enum { S_GAMMA = 0, S_PAP = 1, S_RR = 2, S_GAMMA_NEW = 3, S_ATOL2 = 4, S_BB = 5, S_TOL = 6, S_ATOL = 7 };
enum { C_K = 0, C_DONE = 1, C_MAXITER = 2 };

struct ReduceWs {
    double* part;
    unsigned int* ticket;
};
static inline ReduceWs reduce_ws(void* ws) {
    ReduceWs r;
    r.part = reinterpret_cast<double*>(ws);
    r.ticket = reinterpret_cast<unsigned int*>(r.part + (size_t)MAX_PARTIALS * 4);
    return r;
}

// y = A x for the vertices handled by this warp; returns sum_i x[U a + i] * y[U a + i] in lane 0
template <int U>
__device__ __forceinline__ double spmv_vertex(int64_t a, int lane, const int64_t* __restrict__ nindptr,
                                              const int32_t* __restrict__ nindices, const double* __restrict__ values,
                                              const double* __restrict__ x, double* __restrict__ y) {
    const int64_t p0 = nindptr[a];
    const int deg = (int)(nindptr[a + 1] - p0);
    const double* val = values + (int64_t)U * U * p0;
    const int row_len = U * deg;
    double acc[U];
#pragma unroll
    for (int i = 0; i < U; ++i) acc[i] = 0.0;
    for (int r = lane; r < deg; r += 32) {
        const int64_t bnode = nindices[p0 + r];
        double xb[U];
#pragma unroll
        for (int j = 0; j < U; ++j) xb[j] = x[bnode * U + j];
#pragma unroll
        for (int i = 0; i < U; ++i)
#pragma unroll
            for (int j = 0; j < U; ++j) acc[i] += __ldcs(val + i * row_len + U * r + j) * xb[j];
    }
    double d = 0.0;
#pragma unroll
    for (int i = 0; i < U; ++i) {
        acc[i] = warp_sum(acc[i]);
        if (lane == 0) {
            y[a * U + i] = acc[i];
            d += x[a * U + i] * acc[i];
        }
    }
    return d;
}

template <int U>
__global__ void __launch_bounds__(256) k_spmv(int64_t V, const int64_t* __restrict__ nindptr,
                                              const int32_t* __restrict__ nindices, const double* __restrict__ values,
                                              const double* __restrict__ x, double* __restrict__ y, double* dot_out,
                                              ReduceWs ws, const int32_t* counters, double* scalars) {
    if (counters && counters[C_DONE]) {       // finished: leave a zero partial (an all-reduce of a stale one would grow it)
        if (scalars && blockIdx.x == 0 && threadIdx.x == 0) scalars[S_PAP] = 0.0;
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    double d[1] = {0.0};
    for (int64_t a = (int64_t)blockIdx.x * wpb + warp; a < V; a += (int64_t)gridDim.x * wpb)
        d[0] += spmv_vertex<U>(a, lane, nindptr, nindices, values, x, y);
    if (!dot_out && !scalars) return;
    __shared__ double red[32];
    if (lane != 0) d[0] = 0.0;
    block_sum<1>(d, red);
    double total[1];
    if (grid_sum_last_block<1>(d, ws.part, ws.ticket, total)) {
        if (dot_out) dot_out[0] = total[0];
        if (scalars) scalars[S_PAP] = total[0];
    }
}

// ---- SpMV fed by bulk asynchronous copies (cp.async.bulk + mbarrier, the TMA engine's 1-D path) ---------------------
// The row panels of consecutive vertices are contiguous in the reference's CSR order, so a warp asks the copy engine for
// the panels of SPB_CH consecutive vertices at once (one ~8 KB copy into one of SPB_NB shared-memory buffers, completion
// on an mbarrier), one chunk ahead of the one it multiplies; neighbour ids and x of the next chunks are gathered into
// registers meanwhile.  The matrix (98 % of the bytes) never goes through L1 / the LSU request path, which is what bounds
// the plain kernel (ncu: L1 request slots 81 % busy, long-scoreboard stalls).  Same arithmetic per row as k_spmv, so the
// results are bit-identical (tested).  Default for U = 3 (FEB200_SPMV_BULK=0 selects k_spmv).  At 200^3: 3.40 ms against
// 3.70 ms for k_spmv on the same box with 2 vertices per copy, 2 buffers per warp and 2 CTAs x 10 warps per SM; a pure read
// stream of the same bytes takes 2.45 ms on this GPU (profiles/read_bandwidth.py).  Other configurations in DESIGN.md.
#ifndef FEB_SPMV_NB
#define FEB_SPMV_NB 2
#endif
// Proxy fence before a buffer is overwritten by the next copy (generic-proxy reads, then an async-proxy write).  It
// compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC; measured with and without it on the same box: 2.74 against 2.75 ms at
// 200^3, so it stays in.
#ifndef FEB_SPMV_WAR_FENCE
#define FEB_SPMV_WAR_FENCE 1
#endif
#ifndef FEB_SPMV_WARPS
#define FEB_SPMV_WARPS 10
#endif
#ifndef FEB_SPMV_CH
#define FEB_SPMV_CH 2
#endif
#ifndef FEB_SPMV_CTAS
#define FEB_SPMV_CTAS 2
#endif
constexpr int SPB_NB = FEB_SPMV_NB;
constexpr int SPB_WARPS = FEB_SPMV_WARPS;
constexpr int SPB_CH = FEB_SPMV_CH;
constexpr int SPB_MAXDEG = 32;
constexpr int SPB_BUF_BYTES = SPB_CH * 9 * SPB_MAXDEG * 8 + 32;      // aligned window around the panels of CH vertices

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}

// Sums over the warp of the six values A[0..2], B[0..2] at once: after the steps with offsets 16, 8 and 4 every group of
// four lanes keeps one value (slot = lane >> 2: A0 A1 A2 - B0 B1 B2 -), after 2 and 1 each lane of the group holds its total.
// Every step adds the same pairs as the butterfly of warp_sum, so each total is bit-identical to warp_sum of that value.
__device__ __forceinline__ double warp_sum_2x3(const double (&A)[3], const double (&B)[3], int lane) {
    const bool up = (lane & 16) != 0, hi = (lane & 8) != 0, q = (lane & 4) != 0;
    double k[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double keep = up ? B[i] : A[i], send = up ? A[i] : B[i];
        k[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    double m0 = hi ? k[2] : k[0], m1 = hi ? 0.0 : k[1];
    const double t0 = hi ? k[0] : k[2], t1 = hi ? k[1] : 0.0;
    m0 += __shfl_xor_sync(0xffffffffu, t0, 8);
    m1 += __shfl_xor_sync(0xffffffffu, t1, 8);
    double r = (q ? m1 : m0) + __shfl_xor_sync(0xffffffffu, q ? m0 : m1, 4);
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;
}

__global__ void __launch_bounds__(SPB_WARPS * 32, FEB_SPMV_CTAS)
k_spmv3_bulk(int64_t V, const int64_t* __restrict__ nindptr, const int32_t* __restrict__ nindices,
             const double* __restrict__ values, const double* __restrict__ x, double* __restrict__ y, double* dot_out,
             ReduceWs ws, const int32_t* counters, double* scalars) {
    if (counters && counters[C_DONE]) {
        if (scalars && blockIdx.x == 0 && threadIdx.x == 0) scalars[S_PAP] = 0.0;
        return;
    }
    extern __shared__ __align__(16) unsigned char spb_smem[];
    __shared__ __align__(8) uint64_t bars[SPB_WARPS * SPB_NB];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* my = spb_smem + (size_t)warp * SPB_NB * SPB_BUF_BYTES;
    uint64_t* mybar = bars + warp * SPB_NB;
    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < SPB_NB; ++b) mbar_init(mybar + b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    const int64_t n_chunks = (V + SPB_CH - 1) / SPB_CH;
    const int64_t W = (int64_t)gridDim.x * SPB_WARPS;
    const int64_t w0 = (int64_t)blockIdx.x * SPB_WARPS + warp;
    const int64_t n_mine = (w0 < n_chunks) ? (n_chunks - w0 + W - 1) / W : 0;        // chunks w0, w0 + W, ...
    const int64_t total_nodes = nindptr[V];
    const unsigned char* vbase = reinterpret_cast<const unsigned char*>(values);

    // CSR offsets of a chunk: lane l <= CH holds nindptr[a0 + l] (clamped at V)
    auto load_ptr = [&](int64_t k) -> int64_t {
        if (k >= n_mine) return 0;
        const int64_t a = (w0 + k * W) * SPB_CH + (lane <= SPB_CH ? lane : SPB_CH);
        return nindptr[a < V ? a : V];
    };
    // copy request for my k-th chunk into buffer k % NB; false when it has to take the plain path
    auto issue = [&](int64_t k, int64_t pl) -> bool {
        const int64_t p0 = __shfl_sync(0xffffffffu, pl, 0), p1 = __shfl_sync(0xffffffffu, pl, SPB_CH);
        int maxdeg = 0;
#pragma unroll
        for (int v = 0; v < SPB_CH; ++v) {
            const int dv = (int)(__shfl_sync(0xffffffffu, pl, v + 1) - __shfl_sync(0xffffffffu, pl, v));
            maxdeg = dv > maxdeg ? dv : maxdeg;
        }
        const int64_t vs = (72 * p0) & ~(int64_t)15, vend = (72 * p1 + 15) & ~(int64_t)15;
        if (maxdeg > SPB_MAXDEG || vend > 72 * total_nodes || p1 == p0) return false;
        if (lane == 0) {
            const int b = (int)(k % SPB_NB);
#if FEB_SPMV_WAR_FENCE
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // (compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC)
#endif
            mbar_expect_tx(mybar + b, (uint32_t)(vend - vs));
            bulk_g2s(my + (size_t)b * SPB_BUF_BYTES, vbase + vs, (uint32_t)(vend - vs), mybar + b);
        }
        return true;
    };
    // neighbour ids (two chunks ahead) and x (one chunk ahead) into registers, lane = neighbour
    auto load_nb = [&](int64_t k, int64_t pl, int32_t (&nb)[SPB_CH]) {
#pragma unroll
        for (int v = 0; v < SPB_CH; ++v) {
            const int64_t p0 = __shfl_sync(0xffffffffu, pl, v), p1 = __shfl_sync(0xffffffffu, pl, v + 1);
            nb[v] = (k < n_mine && lane < p1 - p0 && lane < SPB_MAXDEG) ? nindices[p0 + lane] : -1;
        }
    };
    // x of chunk k at the neighbours (lane = neighbour).  Returns the row's own x for the lane that ends up holding that
    // row's sum after the transposed reduction below: lane 4 s, s = 4 v + i, holds y[3 (a0 + v) + i]
    auto load_x = [&](int64_t k, const int32_t (&nb)[SPB_CH], double (&xv)[SPB_CH][3]) -> double {
#pragma unroll
        for (int v = 0; v < SPB_CH; ++v) {
            xv[v][0] = xv[v][1] = xv[v][2] = 0.0;
            if (nb[v] >= 0) {
#pragma unroll
                for (int j = 0; j < 3; ++j) xv[v][j] = x[(int64_t)nb[v] * 3 + j];
            }
        }
        double xo = 0.0;
        const int sl = lane >> 2;
        const int64_t a = (w0 + k * W) * SPB_CH + (sl >> 2);
        if (k < n_mine && (lane & 3) == 0 && (sl & 3) < 3 && (sl >> 2) < SPB_CH && a < V) xo = x[a * 3 + (sl & 3)];
        return xo;
    };

    double d = 0.0;
    unsigned int phase = 0;                                                  // bit b: parity of buffer b's next completion
    // pipeline registers: offsets of chunks k, k+1, k+2; bulk flags of k, k+1; x of k, k+1
    int64_t pl0 = load_ptr(0), pl1 = load_ptr(1), pl2 = load_ptr(2);
    bool bulk0 = n_mine > 0 && issue(0, pl0);
    bool bulk1 = n_mine > 1 && issue(1, pl1);
    double xc[SPB_CH][3], xn[SPB_CH][3];
    int32_t nb1[SPB_CH], nb2[SPB_CH];
    load_nb(0, pl0, nb1);
    double xoc = load_x(0, nb1, xc);
    load_nb(1, pl1, nb1);
    for (int64_t k = 0; k < n_mine; ++k) {
        const int64_t a0 = (w0 + k * W) * SPB_CH;
        const double xon = load_x(k + 1, nb1, xn);                     // chunk k + 1: ids arrived an iteration ago
        load_nb(k + 2, pl2, nb2);                                            // chunk k + 2
        const int64_t pl3 = load_ptr(k + 3);
        bool bulk2 = false;
        if (SPB_NB >= 3) bulk2 = (k + 2 < n_mine) && issue(k + 2, pl2);      // buffer (k+2) % NB was released two iterations ago
        if (bulk0) {
            const int b = (int)(k % SPB_NB);
            while (!mbar_try_wait(mybar + b, (phase >> b) & 1u)) {}
            phase ^= 1u << b;
            const int64_t pbase = __shfl_sync(0xffffffffu, pl0, 0);
            const unsigned char* buf = my + (size_t)b * SPB_BUF_BYTES + ((72 * pbase) & 15);
            double acc[SPB_CH][3];
#pragma unroll
            for (int v = 0; v < SPB_CH; ++v) {
                const int64_t p0 = __shfl_sync(0xffffffffu, pl0, v), p1 = __shfl_sync(0xffffffffu, pl0, v + 1);
                const int deg = (int)(p1 - p0);                              // 0 beyond the last vertex (offsets are clamped)
                const int row_len = 3 * deg;
                const double* val = reinterpret_cast<const double*>(buf + 72 * (p0 - pbase));
                acc[v][0] = acc[v][1] = acc[v][2] = 0.0;
                if (lane < deg) {
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int j = 0; j < 3; ++j) acc[v][i] += val[i * row_len + 3 * lane + j] * xc[v][j];
                }
            }
            if (SPB_CH == 2) {
                // the six row sums of the chunk in one transposed reduction (8 shuffles instead of 30; same addition tree
                // per row as warp_sum, so the rows stay bit-identical to k_spmv): lane 4 s ends up with row s = 4 v + i
                const double rsum = warp_sum_2x3(acc[0], acc[SPB_CH - 1], lane);
                const int sl = lane >> 2;
                if ((lane & 3) == 0 && (sl & 3) < 3 && a0 + (sl >> 2) < V) {
                    y[(a0 + (sl >> 2)) * 3 + (sl & 3)] = rsum;
                    d += xoc * rsum;
                }
            } else {
#pragma unroll
                for (int v = 0; v < SPB_CH; ++v) {
                    const int64_t a = a0 + v;
                    if (a >= V) break;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        acc[v][i] = warp_sum(acc[v][i]);
                        if (lane == 0) {
                            y[a * 3 + i] = acc[v][i];
                            d += x[a * 3 + i] * acc[v][i];
                        }
                    }
                }
            }
        } else {
            for (int v = 0; v < SPB_CH; ++v)
                if (a0 + v < V) d += spmv_vertex<3>(a0 + v, lane, nindptr, nindices, values, x, y);
        }
        __syncwarp();
        if (SPB_NB < 3) {                                                    // double buffering: the buffer just read is the one chunk k + 2 uses
            bulk2 = (k + 2 < n_mine) && issue(k + 2, pl2);
        }
        pl0 = pl1; pl1 = pl2; pl2 = pl3;
        bulk0 = bulk1; bulk1 = bulk2;
#pragma unroll
        for (int v = 0; v < SPB_CH; ++v) nb1[v] = nb2[v];
#pragma unroll
        for (int v = 0; v < SPB_CH; ++v)
#pragma unroll
            for (int j = 0; j < 3; ++j) xc[v][j] = xn[v][j];
        xoc = xon;
    }
    if (!dot_out && !scalars) return;
    __shared__ double red[32];
    double dd[1] = {d};                                                      // partial sums sit on the lanes 4 s (and lane 0)
    block_sum<1>(dd, red);
    double total[1];
    if (grid_sum_last_block<1>(dd, ws.part, ws.ticket, total)) {
        if (dot_out) dot_out[0] = total[0];
        if (scalars) scalars[S_PAP] = total[0];
    }
}

__global__ void k_spmv_csr(int64_t n_rows, const int64_t* __restrict__ indptr, const int64_t* __restrict__ indices,
                           const double* __restrict__ values, const double* __restrict__ x, double* __restrict__ y) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    for (int64_t r = (int64_t)blockIdx.x * wpb + warp; r < n_rows; r += (int64_t)gridDim.x * wpb) {
        double acc = 0.0;
        for (int64_t k = indptr[r] + lane; k < indptr[r + 1]; k += 32) acc += values[k] * x[indices[k]];
        acc = warp_sum(acc);
        if (lane == 0) y[r] = acc;
    }
}

__global__ void __launch_bounds__(256) k_dot(int64_t n, const double* __restrict__ a, const double* __restrict__ b,
                                             const double* __restrict__ w, double* out, ReduceWs ws,
                                             const int32_t* counters) {
    if (counters && counters[C_DONE]) {
        if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = 0.0;
        return;
    }
    double d[1] = {0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        d[0] += w ? w[i] * (a[i] * b[i]) : a[i] * b[i];
    __shared__ double red[32];
    block_sum<1>(d, red);
    double total[1];
    if (grid_sum_last_block<1>(d, ws.part, ws.ticket, total)) out[0] = total[0];
}

// r = b - Ax ; z = M r ; p = z ; (weighted) partial sums r.z -> S_GAMMA_NEW, r.r -> S_RR, b.b -> S_BB.
// With ownership weights `w` (multi-GPU) the three sums are LOCAL and the caller all-reduces
// scalars[S_RR..S_GAMMA_NEW] and scalars[S_BB] before k_cg_init_finalize.
__global__ void __launch_bounds__(256) k_cg_init_vectors(int64_t n, const double* __restrict__ b,
                                                         const double* __restrict__ Ax,
                                                         const double* __restrict__ inv_diag,
                                                         const double* __restrict__ w, double* __restrict__ r,
                                                         double* __restrict__ p, double* scalars, ReduceWs ws,
                                                         double tol, double atol) {
    double d[3] = {0.0, 0.0, 0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double bi = b[i];
        const double ri = bi - Ax[i];
        const double zi = inv_diag ? ri * inv_diag[i] : ri;
        r[i] = ri;
        p[i] = zi;
        const double wi = w ? w[i] : 1.0;
        d[0] += wi * (ri * zi);
        d[1] += wi * (ri * ri);
        d[2] += wi * (bi * bi);
    }
    __shared__ double red[96];
    block_sum<3>(d, red);
    double t[3];
    if (grid_sum_last_block<3>(d, ws.part, ws.ticket, t)) {
        scalars[S_GAMMA_NEW] = t[0];
        scalars[S_RR] = t[1];
        scalars[S_BB] = t[2];
        scalars[S_PAP] = 0.0;
        scalars[S_TOL] = tol;
        scalars[S_ATOL] = atol;
    }
}
__global__ void k_cg_init_finalize(double* scalars, int32_t* counters, double* hist, int maxiter, int preconditioned) {
    const double tol = scalars[S_TOL], atol = scalars[S_ATOL];
    const double atol2 = fmax(tol * tol * scalars[S_BB], atol * atol);
    scalars[S_ATOL2] = atol2;
    scalars[S_GAMMA] = scalars[S_GAMMA_NEW];
    if (maxiter > 0) hist[0] = sqrt(scalars[S_RR]);
    const double rs = preconditioned ? scalars[S_RR] : scalars[S_GAMMA];
    counters[C_K] = 0;
    counters[C_MAXITER] = maxiter;
    counters[C_DONE] = (rs > atol2 && 0 < maxiter) ? 0 : 1;
}

__global__ void __launch_bounds__(256) k_cg_update_xr(int64_t n, const double* __restrict__ p, const double* __restrict__ Ap,
                                                      const double* __restrict__ inv_diag, const double* __restrict__ w,
                                                      double* __restrict__ x, double* __restrict__ r, double* scalars,
                                                      const int32_t* counters, ReduceWs ws, int vec2) {
    if (counters[C_DONE]) {
        if (w && blockIdx.x == 0 && threadIdx.x == 0) { scalars[S_GAMMA_NEW] = 0.0; scalars[S_RR] = 0.0; }   // partial sums of a partitioned run
        return;
    }
    const double alpha = scalars[S_GAMMA] / scalars[S_PAP];
    double d[2] = {0.0, 0.0};
    auto one = [&](int64_t i) {
        x[i] = __fma_rn(alpha, p[i], x[i]);            // explicit roundings: a DoF must get the same bits whether it
        const double ri = __fma_rn(-alpha, Ap[i], r[i]);  // sits in the .x / .y half of a pair or in the scalar tail (the
        r[i] = ri;                                       // copies of an interface DoF on two ranks do)
        const double zi = inv_diag ? __dmul_rn(ri, inv_diag[i]) : ri;
        const double wi = w ? w[i] : 1.0;
        d[0] += wi * (ri * zi);
        d[1] += wi * (ri * ri);
    };
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    if (vec2) {                                    // 16-byte accesses (all vectors 16-byte aligned, checked by the host side)
        const double2* p2 = reinterpret_cast<const double2*>(p);
        const double2* A2 = reinterpret_cast<const double2*>(Ap);
        const double2* i2 = reinterpret_cast<const double2*>(inv_diag);
        const double2* w2 = reinterpret_cast<const double2*>(w);
        double2* x2 = reinterpret_cast<double2*>(x);
        double2* r2 = reinterpret_cast<double2*>(r);
        const int64_t n2 = n >> 1;
        for (int64_t i = tid; i < n2; i += stride) {
            const double2 pv = p2[i], av = A2[i];
            double2 xv = x2[i], rv = r2[i];
            xv.x = __fma_rn(alpha, pv.x, xv.x); xv.y = __fma_rn(alpha, pv.y, xv.y);
            rv.x = __fma_rn(-alpha, av.x, rv.x); rv.y = __fma_rn(-alpha, av.y, rv.y);
            x2[i] = xv;
            r2[i] = rv;
            double2 zv = rv, wv = make_double2(1.0, 1.0);
            if (inv_diag) { const double2 iv = i2[i]; zv.x = __dmul_rn(zv.x, iv.x); zv.y = __dmul_rn(zv.y, iv.y); }
            if (w) wv = w2[i];
            d[0] += wv.x * (rv.x * zv.x);
            d[1] += wv.x * (rv.x * rv.x);
            d[0] += wv.y * (rv.y * zv.y);
            d[1] += wv.y * (rv.y * rv.y);
        }
        if ((n & 1) && tid == 0) one(n - 1);
    } else {
        for (int64_t i = tid; i < n; i += stride) one(i);
    }
    __shared__ double red[64];
    block_sum<2>(d, red);
    double t[2];
    if (grid_sum_last_block<2>(d, ws.part, ws.ticket, t)) {
        scalars[S_GAMMA_NEW] = t[0];
        scalars[S_RR] = t[1];
    }
}

// p = z + beta p
__global__ void __launch_bounds__(256) k_cg_update_p(int64_t n, const double* __restrict__ r,
                                                     const double* __restrict__ inv_diag, double* __restrict__ p,
                                                     const double* scalars, const int32_t* counters, int vec2) {
    if (counters[C_DONE]) return;
    const double beta = scalars[S_GAMMA_NEW] / scalars[S_GAMMA];
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    if (vec2) {
        const double2* r2 = reinterpret_cast<const double2*>(r);
        const double2* i2 = reinterpret_cast<const double2*>(inv_diag);
        double2* p2 = reinterpret_cast<double2*>(p);
        const int64_t n2 = n >> 1;
        for (int64_t i = tid; i < n2; i += stride) {
            double2 zv = r2[i];
            if (inv_diag) { const double2 iv = i2[i]; zv.x = __dmul_rn(zv.x, iv.x); zv.y = __dmul_rn(zv.y, iv.y); }
            double2 pv = p2[i];
            pv.x = __fma_rn(beta, pv.x, zv.x);         // z rounded first, then one fused step: the same in every path
            pv.y = __fma_rn(beta, pv.y, zv.y);
            p2[i] = pv;
        }
        if ((n & 1) && tid == 0) {
            const int64_t i = n - 1;
            const double zi = inv_diag ? __dmul_rn(r[i], inv_diag[i]) : r[i];
            p[i] = __fma_rn(beta, p[i], zi);
        }
        return;
    }
    for (int64_t i = tid; i < n; i += stride) {
        const double zi = inv_diag ? __dmul_rn(r[i], inv_diag[i]) : r[i];
        p[i] = __fma_rn(beta, p[i], zi);
    }
}
// loop state update (separate tiny kernel: every block of update_p has read gamma before it changes):
// hist[k+1] = |r| (dropped when k+1 == maxiter, solve_cg.py:45), gamma <- gamma', k++, loop test.
__global__ void k_cg_advance(double* scalars, int32_t* counters, double* hist, int preconditioned) {
    if (counters[C_DONE]) return;
    const double g = scalars[S_GAMMA_NEW];
    scalars[S_GAMMA] = g;
    const int k = counters[C_K] + 1;
    if (k < counters[C_MAXITER]) hist[k] = sqrt(scalars[S_RR]);
    counters[C_K] = k;
    const double rs = preconditioned ? scalars[S_RR] : g;
    if (!(rs > scalars[S_ATOL2] && k < counters[C_MAXITER])) counters[C_DONE] = 1;
}

// ---- batched CG: C independent systems laid side by side (block-diagonal matrix, copy c owns DoFs [c*len, (c+1)*len)) ---
// One CTA per copy: every scalar of the iteration is a reduction over that copy's segment only, so a whole CG step
// after the SpMV (alpha, x and r update, z, the two new dot products, beta, p update, loop test) is ONE kernel and each
// copy keeps its own alpha / beta / iteration count / stopping test, exactly as vmap(cg) gives in the reference
// (solve_cg.py:20-137 under tests/test_microscale_bvp.py).  Finished copies are skipped.
// scal[c*4 + {0,1,2,3}] = gamma, rr, bb, atol2 ; cnt[c*2 + {0,1}] = k, done ; hist[c*maxiter + k]
__global__ void __launch_bounds__(256) k_bcg_init(int64_t len, const double* __restrict__ b, const double* __restrict__ Ax,
                                                  const double* __restrict__ inv_diag, double* __restrict__ r,
                                                  double* __restrict__ p, double* scal, int32_t* cnt, double* hist,
                                                  double tol, double atol, int maxiter) {
    const int64_t c = blockIdx.x, o = c * len;
    double d[3] = {0.0, 0.0, 0.0};
    for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
        const double bi = b[o + i];
        const double ri = bi - Ax[o + i];
        const double zi = inv_diag ? ri * inv_diag[o + i] : ri;
        r[o + i] = ri;
        p[o + i] = zi;
        d[0] += ri * zi;
        d[1] += ri * ri;
        d[2] += bi * bi;
    }
    __shared__ double red[96];
    block_sum<3>(d, red);
    if (threadIdx.x == 0) {
        const double atol2 = fmax(tol * tol * d[2], atol * atol);
        scal[c * 4 + 0] = d[0];
        scal[c * 4 + 1] = d[1];
        scal[c * 4 + 2] = d[2];
        scal[c * 4 + 3] = atol2;
        if (maxiter > 0) hist[c * (int64_t)maxiter] = sqrt(d[1]);
        const double rs = inv_diag ? d[1] : d[0];
        cnt[c * 2 + 0] = 0;
        cnt[c * 2 + 1] = (rs > atol2 && 0 < maxiter) ? 0 : 1;
    }
}

__global__ void __launch_bounds__(256) k_bcg_step(int64_t len, const double* __restrict__ Ap,
                                                  const double* __restrict__ inv_diag, double* __restrict__ x,
                                                  double* __restrict__ r, double* __restrict__ p, double* scal, int32_t* cnt,
                                                  double* hist, int maxiter) {
    const int64_t c = blockIdx.x, o = c * len;
    if (cnt[c * 2 + 1]) return;
    __shared__ double red[64];
    __shared__ double bc[2];
    double d1[1] = {0.0};
    for (int64_t i = threadIdx.x; i < len; i += blockDim.x) d1[0] += p[o + i] * Ap[o + i];
    block_sum<1>(d1, red);
    if (threadIdx.x == 0) bc[0] = scal[c * 4 + 0] / d1[0];                   // alpha = gamma / p.Ap
    __syncthreads();
    const double alpha = bc[0];
    double d[2] = {0.0, 0.0};
    for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
        x[o + i] = x[o + i] + alpha * p[o + i];
        const double ri = r[o + i] - alpha * Ap[o + i];
        r[o + i] = ri;
        const double zi = inv_diag ? ri * inv_diag[o + i] : ri;
        d[0] += ri * zi;
        d[1] += ri * ri;
    }
    block_sum<2>(d, red);
    if (threadIdx.x == 0) {
        const double gamma = scal[c * 4 + 0];
        bc[1] = d[0] / gamma;                                                 // beta
        scal[c * 4 + 0] = d[0];
        scal[c * 4 + 1] = d[1];
        const int k = cnt[c * 2 + 0] + 1;
        if (k < maxiter) hist[c * (int64_t)maxiter + k] = sqrt(d[1]);
        cnt[c * 2 + 0] = k;
        const double rs = inv_diag ? d[1] : d[0];
        if (!(rs > scal[c * 4 + 3] && k < maxiter)) cnt[c * 2 + 1] = 1;
    }
    __syncthreads();
    const double beta = bc[1];
    for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
        const double zi = inv_diag ? r[o + i] * inv_diag[o + i] : r[o + i];
        p[o + i] = zi + beta * p[o + i];
    }
}

static inline int reduce_grid(int64_t n, int block) {
    int g = grid_for(n, block);
    const int cap = sm_count() * 4;
    if (g > cap) g = cap;
    if (g > MAX_PARTIALS) g = MAX_PARTIALS;
    return g;
}

template <int U>
static void launch_spmv(cudaStream_t st, int64_t V, const int64_t* nindptr, const int32_t* nindices, const double* values,
                        const double* x, double* y, double* dot_out, ReduceWs ws, const int32_t* counters, double* scalars) {
    int64_t g = ceil_div(V, 8);
    int64_t cap = (int64_t)sm_count() * 8;
    if (dot_out || scalars) cap = MAX_PARTIALS < cap ? MAX_PARTIALS : cap;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    k_spmv<U><<<(int)g, 256, 0, st>>>(V, nindptr, nindices, values, x, y, dot_out, ws, counters, scalars);
}

static bool spmv_bulk_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("FEB200_SPMV_BULK");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

static int spmv_dispatch(cudaStream_t st, int64_t V, int U, const int64_t* nindptr, const int32_t* nindices,
                         const double* values, const double* x, double* y, double* dot_out, ReduceWs ws,
                         const int32_t* counters, double* scalars) {
    // the bulk copies read 16-byte aligned pieces of `values`: any other base address (a tensor view, an XLA sub-buffer)
    // takes the plain-load kernel instead of faulting
    if (U == 3 && spmv_bulk_enabled() && ((uintptr_t)values & 15u) == 0) {
        const size_t smem = (size_t)SPB_WARPS * SPB_NB * SPB_BUF_BYTES;
        static bool attr_set = false;
        if (!attr_set) {
            FEB_CUDA(cudaFuncSetAttribute(k_spmv3_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = true;
        }
        int64_t g = ceil_div(ceil_div(V, SPB_CH), SPB_WARPS);
        int64_t cap = (int64_t)sm_count() * FEB_SPMV_CTAS;
        if (dot_out || scalars) cap = MAX_PARTIALS < cap ? MAX_PARTIALS : cap;
        if (g > cap) g = cap;
        if (g < 1) g = 1;
        k_spmv3_bulk<<<(int)g, SPB_WARPS * 32, smem, st>>>(V, nindptr, nindices, values, x, y, dot_out, ws, counters, scalars);
    } else if (U == 3) launch_spmv<3>(st, V, nindptr, nindices, values, x, y, dot_out, ws, counters, scalars);
    else if (U == 2) launch_spmv<2>(st, V, nindptr, nindices, values, x, y, dot_out, ws, counters, scalars);
    else if (U == 1) launch_spmv<1>(st, V, nindptr, nindices, values, x, y, dot_out, ws, counters, scalars);
    else FEB_REQUIRE(false, FEB200_ERR_UNSUPPORTED, "spmv: U=%d not supported", U);
    return FEB200_OK;
}

static int check_cg(const feb200_cg_state* st) {
    FEB_REQUIRE(st, FEB200_ERR_INVALID, "cg: null state");
    FEB_REQUIRE(st->nindptr && st->nindices && st->values && st->b && st->x && st->r && st->p && st->Ap && st->hist &&
                    st->scalars && st->counters && st->ws,
                FEB200_ERR_INVALID, "cg: null pointer in state");
    FEB_REQUIRE(st->ws_bytes >= REDUCE_WS_BYTES, FEB200_ERR_WORKSPACE, "cg: reduce workspace too small");
    FEB_REQUIRE(st->maxiter >= 0 && st->V >= 0, FEB200_ERR_INVALID, "cg: bad sizes");
    return FEB200_OK;
}

}  // namespace feb

using namespace feb;

extern "C" {

size_t feb200_reduce_workspace_bytes(void) { return REDUCE_WS_BYTES; }

int feb200_spmv(feb200_stream s, int64_t V, int32_t U, const int64_t* nindptr, const int32_t* nindices,
                const double* values, const double* x, double* y, double* dot_out, void* ws, size_t ws_bytes) {
    FEB_REQUIRE(nindptr && nindices && values && x && y, FEB200_ERR_INVALID, "spmv: null pointer");
    FEB_REQUIRE(!dot_out || (ws && ws_bytes >= REDUCE_WS_BYTES), FEB200_ERR_WORKSPACE, "spmv: reduce workspace too small");
    if (V == 0) return FEB200_OK;
    ReduceWs r = ws ? reduce_ws(ws) : ReduceWs{nullptr, nullptr};
    int rc = spmv_dispatch(as_stream(s), V, U, nindptr, nindices, values, x, y, dot_out, r, nullptr, nullptr);
    if (rc) return rc;
    FEB_CHECK_LAUNCH("spmv");
    return FEB200_OK;
}

int feb200_spmv_csr(feb200_stream s, int64_t n_rows, const int64_t* indptr, const int64_t* indices, const double* values,
                    const double* x, double* y) {
    FEB_REQUIRE(indptr && indices && values && x && y, FEB200_ERR_INVALID, "spmv_csr: null pointer");
    if (n_rows == 0) return FEB200_OK;
    int64_t g = ceil_div(n_rows, 8);
    const int64_t cap = (int64_t)sm_count() * 16;
    if (g > cap) g = cap;
    k_spmv_csr<<<(int)g, 256, 0, as_stream(s)>>>(n_rows, indptr, indices, values, x, y);
    FEB_CHECK_LAUNCH("spmv_csr");
    return FEB200_OK;
}

int feb200_dot(feb200_stream s, int64_t n, const double* a, const double* b, double* out, void* ws, size_t ws_bytes) {
    FEB_REQUIRE(a && b && out && ws, FEB200_ERR_INVALID, "dot: null pointer");
    FEB_REQUIRE(ws_bytes >= REDUCE_WS_BYTES, FEB200_ERR_WORKSPACE, "dot: reduce workspace too small");
    k_dot<<<reduce_grid(n, 256), 256, 0, as_stream(s)>>>(n, a, b, nullptr, out, reduce_ws(ws), nullptr);
    FEB_CHECK_LAUNCH("dot");
    return FEB200_OK;
}

}  // extern "C"

static int cg_stage(cudaStream_t cs, const feb200_cg_state* st, int stage, double tol, double atol) {
    const int64_t n = st->V * st->U;
    ReduceWs r = reduce_ws(st->ws);
    const int pre = st->inv_diag ? 1 : 0;
    // 16-byte accesses in the vector kernels when every vector they touch is 16-byte aligned
    const int vec2 = (((uintptr_t)st->x | (uintptr_t)st->r | (uintptr_t)st->p | (uintptr_t)st->Ap | (uintptr_t)st->inv_diag |
                       (uintptr_t)st->weights) & 15u) == 0;
    int rc = FEB200_OK;
    switch (stage) {
        case FEB200_CG_INIT_SPMV:
            FEB_CUDA(cudaMemsetAsync(r.ticket, 0, REDUCE_CHANNELS * sizeof(unsigned int), cs));
            FEB_CUDA(cudaMemsetAsync(st->hist, 0, (size_t)st->maxiter * sizeof(double), cs));
            if (st->V > 0)
                rc = spmv_dispatch(cs, st->V, st->U, st->nindptr, st->nindices, st->values, st->x, st->Ap, nullptr, r,
                                   nullptr, nullptr);
            break;
        case FEB200_CG_INIT_VECTORS:
            k_cg_init_vectors<<<reduce_grid(n, 256), 256, 0, cs>>>(n, st->b, st->Ap, st->inv_diag, st->weights, st->r,
                                                                 st->p, st->scalars, r, tol, atol);
            break;
        case FEB200_CG_INIT_FINALIZE:
            k_cg_init_finalize<<<1, 1, 0, cs>>>(st->scalars, st->counters, st->hist, st->maxiter, pre);
            break;
        case FEB200_CG_SPMV:     // Ap = A p and scalars[1] = p.(A p) with the LOCAL product.  Element-partitioned: the global
                                 // p.Ap is the sum over ranks of p_r.(A_r p_r) (A = sum R_r^T A_r R_r), so the caller all-reduces
                                 // scalars[1] as it is -- no halo-summed product, no ownership weights, no extra pass
            if (st->V > 0)
                rc = spmv_dispatch(cs, st->V, st->U, st->nindptr, st->nindices, st->values, st->p, st->Ap, nullptr, r,
                                   st->counters, st->scalars);
            break;
        case FEB200_CG_PAP:      // weighted p.Ap -> scalars[1] (local; caller all-reduces)
            k_dot<<<reduce_grid(n, 256), 256, 0, cs>>>(n, st->p, st->Ap, st->weights, st->scalars + S_PAP, r, st->counters);
            break;
        case FEB200_CG_UPDATE_XR:
            k_cg_update_xr<<<reduce_grid(n, 256), 256, 0, cs>>>(n, st->p, st->Ap, st->inv_diag, st->weights, st->x, st->r,
                                                              st->scalars, st->counters, r, vec2);
            break;
        case FEB200_CG_UPDATE_P:
            k_cg_update_p<<<reduce_grid(n, 256), 256, 0, cs>>>(n, st->r, st->inv_diag, st->p, st->scalars, st->counters, vec2);
            k_cg_advance<<<1, 1, 0, cs>>>(st->scalars, st->counters, st->hist, pre);
            break;
        default:
            FEB_REQUIRE(false, FEB200_ERR_INVALID, "cg_stage: unknown stage %d", stage);
    }
    return rc;
}

extern "C" {

int feb200_cg_stage(feb200_stream s, const feb200_cg_state* st, int32_t stage, double tol, double atol) {
    int rc = check_cg(st);
    if (rc) return rc;
    rc = cg_stage(as_stream(s), st, stage, tol, atol);
    if (rc) return rc;
    FEB_CHECK_LAUNCH("cg_stage");
    return FEB200_OK;
}

int feb200_bcg_init(feb200_stream s, int64_t n_copies, int64_t len, const double* b, const double* Ax0, const double* inv_diag,
                    double* r, double* p, double* scal, int32_t* cnt, double* hist, double tol, double atol,
                    int32_t maxiter) {
    FEB_REQUIRE(b && Ax0 && r && p && scal && cnt && hist && n_copies >= 0 && len >= 0 && maxiter >= 0, FEB200_ERR_INVALID,
                "bcg_init: bad arguments");
    if (n_copies == 0) return FEB200_OK;
    k_bcg_init<<<(unsigned int)n_copies, 256, 0, as_stream(s)>>>(len, b, Ax0, inv_diag, r, p, scal, cnt, hist, tol, atol, maxiter);
    FEB_CHECK_LAUNCH("bcg_init");
    return FEB200_OK;
}

int feb200_bcg_step(feb200_stream s, int64_t n_copies, int64_t len, const double* Ap, const double* inv_diag, double* x,
                    double* r, double* p, double* scal, int32_t* cnt, double* hist, int32_t maxiter) {
    FEB_REQUIRE(Ap && x && r && p && scal && cnt && hist && n_copies >= 0 && len >= 0, FEB200_ERR_INVALID,
                "bcg_step: bad arguments");
    if (n_copies == 0) return FEB200_OK;
    k_bcg_step<<<(unsigned int)n_copies, 256, 0, as_stream(s)>>>(len, Ap, inv_diag, x, r, p, scal, cnt, hist, maxiter);
    FEB_CHECK_LAUNCH("bcg_step");
    return FEB200_OK;
}

int feb200_cg_init(feb200_stream s, const feb200_cg_state* st, double tol, double atol) {
    int rc = check_cg(st);
    if (rc) return rc;
    FEB_REQUIRE(!st->weights, FEB200_ERR_INVALID, "cg_init: ownership weights need the staged API (all-reduce between stages)");
    cudaStream_t cs = as_stream(s);
    for (int stage = FEB200_CG_INIT_SPMV; stage <= FEB200_CG_INIT_FINALIZE; ++stage) {
        rc = cg_stage(cs, st, stage, tol, atol);
        if (rc) return rc;
    }
    FEB_CHECK_LAUNCH("cg_init");
    return FEB200_OK;
}

int feb200_cg_iterate(feb200_stream s, const feb200_cg_state* st, int32_t n_iters) {
    int rc = check_cg(st);
    if (rc) return rc;
    FEB_REQUIRE(!st->weights, FEB200_ERR_INVALID, "cg_iterate: ownership weights need the staged API");
    cudaStream_t cs = as_stream(s);
    if (st->V * st->U == 0) return FEB200_OK;
    for (int it = 0; it < n_iters; ++it) {
        if ((rc = cg_stage(cs, st, FEB200_CG_SPMV, 0, 0))) return rc;
        if ((rc = cg_stage(cs, st, FEB200_CG_UPDATE_XR, 0, 0))) return rc;
        if ((rc = cg_stage(cs, st, FEB200_CG_UPDATE_P, 0, 0))) return rc;
    }
    FEB_CHECK_LAUNCH("cg_iterate");
    return FEB200_OK;
}

}  // extern "C"
