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

enum { S_GAMMA = 0, S_PAP = 1, S_RR = 2, S_ATOL2 = 3, S_GAMMA_NEW = 4, S_BB = 5 };
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
            for (int j = 0; j < U; ++j) acc[i] += val[i * row_len + U * r + j] * xb[j];
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
    if (counters && counters[C_DONE]) return;
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
                                             double* out, ReduceWs ws) {
    double d[1] = {0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        d[0] += a[i] * b[i];
    __shared__ double red[32];
    block_sum<1>(d, red);
    double total[1];
    if (grid_sum_last_block<1>(d, ws.part, ws.ticket, total)) out[0] = total[0];
}

// r = b - Ap(=A x0) ; z = M r ; p = z ; partial sums (r.z, r.r, b.b); last block initialises the loop state
__global__ void __launch_bounds__(256) k_cg_init(int64_t n, const double* __restrict__ b, const double* __restrict__ Ax,
                                                 const double* __restrict__ inv_diag, double* __restrict__ r,
                                                 double* __restrict__ p, double* hist, double* scalars, int32_t* counters,
                                                 ReduceWs ws, double tol, double atol, int maxiter) {
    double d[3] = {0.0, 0.0, 0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double bi = b[i];
        const double ri = bi - Ax[i];
        const double zi = inv_diag ? ri * inv_diag[i] : ri;
        r[i] = ri;
        p[i] = zi;
        d[0] += ri * zi;
        d[1] += ri * ri;
        d[2] += bi * bi;
    }
    __shared__ double red[96];
    block_sum<3>(d, red);
    double t[3];
    if (grid_sum_last_block<3>(d, ws.part, ws.ticket, t)) {
        const double atol2 = fmax(tol * tol * t[2], atol * atol);
        scalars[S_GAMMA] = t[0];
        scalars[S_RR] = t[1];
        scalars[S_BB] = t[2];
        scalars[S_ATOL2] = atol2;
        scalars[S_GAMMA_NEW] = t[0];
        scalars[S_PAP] = 0.0;
        if (maxiter > 0) hist[0] = sqrt(t[1]);
        const double rs = inv_diag ? t[1] : t[0];
        counters[C_K] = 0;
        counters[C_MAXITER] = maxiter;
        counters[C_DONE] = (rs > atol2 && 0 < maxiter) ? 0 : 1;
    }
}

__global__ void __launch_bounds__(256) k_cg_update_xr(int64_t n, const double* __restrict__ p, const double* __restrict__ Ap,
                                                      const double* __restrict__ inv_diag, double* __restrict__ x,
                                                      double* __restrict__ r, double* hist, double* scalars,
                                                      int32_t* counters, ReduceWs ws) {
    if (counters[C_DONE]) return;
    const double alpha = scalars[S_GAMMA] / scalars[S_PAP];
    double d[2] = {0.0, 0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        x[i] = x[i] + alpha * p[i];
        const double ri = r[i] - alpha * Ap[i];
        r[i] = ri;
        const double zi = inv_diag ? ri * inv_diag[i] : ri;
        d[0] += ri * zi;
        d[1] += ri * ri;
    }
    __shared__ double red[64];
    block_sum<2>(d, red);
    double t[2];
    if (grid_sum_last_block<2>(d, ws.part, ws.ticket, t)) {
        scalars[S_GAMMA_NEW] = t[0];
        scalars[S_RR] = t[1];
        const int k = counters[C_K];
        const int maxiter = counters[C_MAXITER];
        if (k + 1 < maxiter) hist[k + 1] = sqrt(t[1]);     // the write at k+1 == maxiter is dropped (solve_cg.py:45)
    }
}

// p = z + beta p ; thread 0 of block 0 then advances the loop state (after everyone has read it:
// the state update happens in a separate tiny kernel to avoid a grid-wide race)
__global__ void __launch_bounds__(256) k_cg_update_p(int64_t n, const double* __restrict__ r,
                                                     const double* __restrict__ inv_diag, double* __restrict__ p,
                                                     const double* scalars, const int32_t* counters) {
    if (counters[C_DONE]) return;
    const double beta = scalars[S_GAMMA_NEW] / scalars[S_GAMMA];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double zi = inv_diag ? r[i] * inv_diag[i] : r[i];
        p[i] = zi + beta * p[i];
    }
}
__global__ void k_cg_advance(double* scalars, int32_t* counters, int preconditioned) {
    if (counters[C_DONE]) return;
    const double g = scalars[S_GAMMA_NEW];
    scalars[S_GAMMA] = g;
    const int k = counters[C_K] + 1;
    counters[C_K] = k;
    const double rs = preconditioned ? scalars[S_RR] : g;
    if (!(rs > scalars[S_ATOL2] && k < counters[C_MAXITER])) counters[C_DONE] = 1;
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

static int spmv_dispatch(cudaStream_t st, int64_t V, int U, const int64_t* nindptr, const int32_t* nindices,
                         const double* values, const double* x, double* y, double* dot_out, ReduceWs ws,
                         const int32_t* counters, double* scalars) {
    if (U == 3) launch_spmv<3>(st, V, nindptr, nindices, values, x, y, dot_out, ws, counters, scalars);
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
    k_dot<<<reduce_grid(n, 256), 256, 0, as_stream(s)>>>(n, a, b, out, reduce_ws(ws));
    FEB_CHECK_LAUNCH("dot");
    return FEB200_OK;
}

int feb200_cg_init(feb200_stream s, const feb200_cg_state* st, double tol, double atol) {
    int rc = check_cg(st);
    if (rc) return rc;
    cudaStream_t cs = as_stream(s);
    const int64_t n = st->V * st->U;
    ReduceWs r = reduce_ws(st->ws);
    FEB_CUDA(cudaMemsetAsync(r.ticket, 0, REDUCE_CHANNELS * sizeof(unsigned int), cs));
    FEB_CUDA(cudaMemsetAsync(st->hist, 0, (size_t)st->maxiter * sizeof(double), cs));
    if (st->V > 0) {
        rc = spmv_dispatch(cs, st->V, st->U, st->nindptr, st->nindices, st->values, st->x, st->Ap, nullptr, r, nullptr, nullptr);
        if (rc) return rc;
    }
    k_cg_init<<<reduce_grid(n, 256), 256, 0, cs>>>(n, st->b, st->Ap, st->inv_diag, st->r, st->p, st->hist, st->scalars,
                                                 st->counters, r, tol, atol, st->maxiter);
    FEB_CHECK_LAUNCH("cg_init");
    return FEB200_OK;
}

int feb200_cg_iterate(feb200_stream s, const feb200_cg_state* st, int32_t n_iters) {
    int rc = check_cg(st);
    if (rc) return rc;
    cudaStream_t cs = as_stream(s);
    const int64_t n = st->V * st->U;
    if (n == 0) return FEB200_OK;
    ReduceWs r = reduce_ws(st->ws);
    const int gv = reduce_grid(n, 256);
    const int gp = grid_for(n, 256);
    for (int it = 0; it < n_iters; ++it) {
        rc = spmv_dispatch(cs, st->V, st->U, st->nindptr, st->nindices, st->values, st->p, st->Ap, nullptr, r, st->counters,
                           st->scalars);
        if (rc) return rc;
        k_cg_update_xr<<<gv, 256, 0, cs>>>(n, st->p, st->Ap, st->inv_diag, st->x, st->r, st->hist, st->scalars, st->counters, r);
        k_cg_update_p<<<gp, 256, 0, cs>>>(n, st->r, st->inv_diag, st->p, st->scalars, st->counters);
        k_cg_advance<<<1, 1, 0, cs>>>(st->scalars, st->counters, st->inv_diag ? 1 : 0);
    }
    FEB_CHECK_LAUNCH("cg_iterate");
    return FEB200_OK;
}

}  // extern "C"
