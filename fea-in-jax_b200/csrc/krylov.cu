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
                                             const double* __restrict__ w, double* out, ReduceWs ws,
                                             const int32_t* counters) {
    if (counters && counters[C_DONE]) return;
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
                                                      const int32_t* counters, ReduceWs ws) {
    if (counters[C_DONE]) return;
    const double alpha = scalars[S_GAMMA] / scalars[S_PAP];
    double d[2] = {0.0, 0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        x[i] = x[i] + alpha * p[i];
        const double ri = r[i] - alpha * Ap[i];
        r[i] = ri;
        const double zi = inv_diag ? ri * inv_diag[i] : ri;
        const double wi = w ? w[i] : 1.0;
        d[0] += wi * (ri * zi);
        d[1] += wi * (ri * ri);
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
                                                     const double* scalars, const int32_t* counters) {
    if (counters[C_DONE]) return;
    const double beta = scalars[S_GAMMA_NEW] / scalars[S_GAMMA];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double zi = inv_diag ? r[i] * inv_diag[i] : r[i];
        p[i] = zi + beta * p[i];
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
    k_dot<<<reduce_grid(n, 256), 256, 0, as_stream(s)>>>(n, a, b, nullptr, out, reduce_ws(ws), nullptr);
    FEB_CHECK_LAUNCH("dot");
    return FEB200_OK;
}

}  // extern "C"

static int cg_stage(cudaStream_t cs, const feb200_cg_state* st, int stage, double tol, double atol) {
    const int64_t n = st->V * st->U;
    ReduceWs r = reduce_ws(st->ws);
    const int pre = st->inv_diag ? 1 : 0;
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
                                                              st->scalars, st->counters, r);
            break;
        case FEB200_CG_UPDATE_P:
            k_cg_update_p<<<grid_for(n, 256), 256, 0, cs>>>(n, st->r, st->inv_diag, st->p, st->scalars, st->counters);
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
