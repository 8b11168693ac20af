// Assembly-map builders, gather / scatter transforms and the Dirichlet wrappers.
//   build_row_ind / build_col_ind_and_data     /root/reference/src/fe_jax/setup.py:212-246
//   transform_global_unraveled_to_element_node setup.py:304-322   (gather)
//   transform_element_node_to_global_unraveled_sum setup.py:344-356 (row-ordered segmented sum)
//   calculate_residual_w_dirichlet wrappers    src/fe_jax/fea.py:948-963
#include <cub/device/device_scan.cuh>
#include <stdarg.h>

#include "common.cuh"

namespace feb {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            n = 148;
    }
    return n;
}

// ------------------------------------------------------------------------------------ kernels
__global__ void k_valence(const int64_t* __restrict__ cells, int64_t n_slots, int64_t V, unsigned long long* cnt_plus1) {
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < n_slots; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = cells[s];
        if (v >= 0 && v < V) atomicAdd(&cnt_plus1[v + 1], 1ull);   // integer atomics: order-independent result
    }
}

__global__ void k_zero_u64(unsigned long long* p, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = 0ull;
}

// unordered fill (atomic cursor per vertex) ...
__global__ void k_fill_slots(const int64_t* __restrict__ cells, int64_t n_slots, int64_t V,
                             const unsigned long long* __restrict__ row_ind, unsigned int* cursor,
                             unsigned long long* col_ind, double* data) {
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < n_slots; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = cells[s];
        if (v < 0 || v >= V) continue;
        const unsigned int k = atomicAdd(&cursor[v], 1u);
        const unsigned long long pos = row_ind[v] + k;
        col_ind[pos] = (unsigned long long)s;
        if (data) data[pos] = 1.0;
    }
}
// ... then each row is sorted ascending: the result is the unique stable counting sort of the reference.
__global__ void k_sort_rows(int64_t V, const unsigned long long* __restrict__ row_ind, unsigned long long* col_ind) {
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < V; v += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long b = row_ind[v], e = row_ind[v + 1];
        for (unsigned long long i = b + 1; i < e; ++i) {       // insertion sort; valence is small
            const unsigned long long key = col_ind[i];
            unsigned long long j = i;
            while (j > b && col_ind[j - 1] > key) { col_ind[j] = col_ind[j - 1]; --j; }
            col_ind[j] = key;
        }
    }
}

__global__ void k_gather(const int32_t* __restrict__ conn, int64_t n_slots, int U, const double* __restrict__ v,
                         double* __restrict__ out) {
    const int64_t total = n_slots * U;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = i / U;
        const int j = (int)(i - s * U);
        out[i] = v[(int64_t)conn[s] * U + j];
    }
}

template <int U>
__global__ void k_scatter_sum(const unsigned long long* __restrict__ row_ind, const unsigned long long* __restrict__ col_ind,
                              int64_t V, const double* __restrict__ v_en, double* __restrict__ out, int accumulate,
                              const double* __restrict__ mask, const double* __restrict__ values,
                              const double* __restrict__ u) {
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < V; v += (int64_t)gridDim.x * blockDim.x) {
        double acc[U];
#pragma unroll
        for (int j = 0; j < U; ++j) acc[j] = 0.0;
        const unsigned long long b = row_ind[v], e = row_ind[v + 1];
        for (unsigned long long k = b; k < e; ++k) {            // ascending slot order == reference order
            const unsigned long long s = col_ind[k];
#pragma unroll
            for (int j = 0; j < U; ++j) acc[j] += v_en[s * U + j];
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int64_t f = v * U + j;
            double r = accumulate ? out[f] + acc[j] : acc[j];
            if (mask) {
                const double m = mask[f];
                r = (1.0 - m) * r + m * (u[f] - values[f]);
            }
            out[f] = r;
        }
    }
}

__global__ void k_dirichlet_blend(int64_t F, const double* __restrict__ u, const double* __restrict__ g,
                                  const double* __restrict__ m, double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < F; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (1.0 - m[i]) * u[i] + m[i] * g[i];
}
__global__ void k_dirichlet_residual(int64_t F, const double* __restrict__ u, const double* __restrict__ g,
                                     const double* __restrict__ m, double* __restrict__ R) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < F; i += (int64_t)gridDim.x * blockDim.x)
        R[i] = (1.0 - m[i]) * R[i] + m[i] * (u[i] - g[i]);
}
__global__ void k_axpby(int64_t n, double a, const double* __restrict__ x, double b, const double* __restrict__ y,
                        double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = a * x[i] + (y ? b * y[i] : 0.0);
}
__global__ void k_scatter_set(int64_t n, const int64_t* __restrict__ idx, const double* __restrict__ vals, double scale,
                              const double* __restrict__ minus_src, double* __restrict__ dst) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = idx[i];
        dst[f] = scale * vals[i] - (minus_src ? minus_src[f] : 0.0);
    }
}

}  // namespace feb

using namespace feb;

extern "C" {

const char* feb200_last_error(void) { return feb::g_err; }
int feb200_version(void) { return 100; }

size_t feb200_scan_workspace_bytes(int64_t n) {
    size_t bytes = 0;
    cub::DeviceScan::InclusiveSum((void*)nullptr, bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, n);
    size_t b2 = 0;
    cub::DeviceScan::InclusiveSum((void*)nullptr, b2, (const int32_t*)nullptr, (int64_t*)nullptr, n);
    return (bytes > b2 ? bytes : b2) + 256;
}

size_t feb200_build_row_ind_workspace_bytes(int64_t V) { return feb200_scan_workspace_bytes(V + 1); }

int feb200_build_row_ind(feb200_stream s, const int64_t* cells, int64_t E, int32_t N, int64_t V, uint64_t* row_ind,
                         void* ws, size_t ws_bytes) {
    FEB_REQUIRE((cells || E == 0) && row_ind && E >= 0 && N > 0 && V >= 0, FEB200_ERR_INVALID, "build_row_ind: bad arguments");
    FEB_REQUIRE(ws_bytes >= feb200_build_row_ind_workspace_bytes(V), FEB200_ERR_WORKSPACE, "build_row_ind: workspace too small");
    cudaStream_t st = as_stream(s);
    unsigned long long* r = reinterpret_cast<unsigned long long*>(row_ind);
    k_zero_u64<<<grid_for(V + 1, 256), 256, 0, st>>>(r, V + 1);
    if (E > 0) k_valence<<<grid_for(E * N, 256), 256, 0, st>>>(cells, E * (int64_t)N, V, r);
    size_t bytes = ws_bytes;
    FEB_CUDA(cub::DeviceScan::InclusiveSum(ws, bytes, r, r, V + 1, st));   // in place; r[0] stays 0
    FEB_CHECK_LAUNCH("build_row_ind");
    return FEB200_OK;
}

size_t feb200_build_col_ind_workspace_bytes(int64_t V) { return (size_t)(V + 1) * sizeof(unsigned int) + 256; }

int feb200_build_col_ind(feb200_stream s, const int64_t* cells, int64_t E, int32_t N, int64_t V, const uint64_t* row_ind,
                         uint64_t* col_ind, double* data, void* ws, size_t ws_bytes) {
    FEB_REQUIRE((cells || E == 0) && row_ind && (col_ind || E == 0) && N > 0, FEB200_ERR_INVALID, "build_col_ind: bad arguments");
    FEB_REQUIRE(ws_bytes >= feb200_build_col_ind_workspace_bytes(V), FEB200_ERR_WORKSPACE, "build_col_ind: workspace too small");
    cudaStream_t st = as_stream(s);
    FEB_CUDA(cudaMemsetAsync(ws, 0, (size_t)(V + 1) * sizeof(unsigned int), st));
    if (E > 0) {
        k_fill_slots<<<grid_for(E * N, 256), 256, 0, st>>>(cells, E * (int64_t)N, V,
                                                         reinterpret_cast<const unsigned long long*>(row_ind),
                                                         reinterpret_cast<unsigned int*>(ws),
                                                         reinterpret_cast<unsigned long long*>(col_ind), data);
        k_sort_rows<<<grid_for(V, 128), 128, 0, st>>>(V, reinterpret_cast<const unsigned long long*>(row_ind),
                                                    reinterpret_cast<unsigned long long*>(col_ind));
    }
    FEB_CHECK_LAUNCH("build_col_ind");
    return FEB200_OK;
}

int feb200_gather(feb200_stream s, const int32_t* conn, int64_t E, int32_t N, int32_t U, const double* v, double* out) {
    FEB_REQUIRE(conn && v && out && N > 0 && U > 0, FEB200_ERR_INVALID, "gather: bad arguments");
    if (E == 0) return FEB200_OK;
    k_gather<<<grid_for(E * N * U, 256), 256, 0, as_stream(s)>>>(conn, E * (int64_t)N, U, v, out);
    FEB_CHECK_LAUNCH("gather");
    return FEB200_OK;
}

int feb200_scatter_sum(feb200_stream s, const uint64_t* row_ind, const uint64_t* col_ind, int64_t V, int32_t U,
                       const double* v_en, double* out, int32_t accumulate, const double* mask, const double* values,
                       const double* u) {
    FEB_REQUIRE(row_ind && col_ind && out, FEB200_ERR_INVALID, "scatter_sum: bad arguments");
    FEB_REQUIRE(!mask || (values && u), FEB200_ERR_INVALID, "scatter_sum: mask given without values / u");
    if (V == 0) return FEB200_OK;
    const unsigned long long* ri = reinterpret_cast<const unsigned long long*>(row_ind);
    const unsigned long long* ci = reinterpret_cast<const unsigned long long*>(col_ind);
    cudaStream_t st = as_stream(s);
    const int g = grid_for(V, 128);
    switch (U) {
        case 1: k_scatter_sum<1><<<g, 128, 0, st>>>(ri, ci, V, v_en, out, accumulate, mask, values, u); break;
        case 2: k_scatter_sum<2><<<g, 128, 0, st>>>(ri, ci, V, v_en, out, accumulate, mask, values, u); break;
        case 3: k_scatter_sum<3><<<g, 128, 0, st>>>(ri, ci, V, v_en, out, accumulate, mask, values, u); break;
        default: FEB_REQUIRE(false, FEB200_ERR_UNSUPPORTED, "scatter_sum: U=%d not supported", U);
    }
    FEB_CHECK_LAUNCH("scatter_sum");
    return FEB200_OK;
}

int feb200_dirichlet_blend(feb200_stream s, int64_t F, const double* u, const double* values, const double* mask, double* u_w) {
    FEB_REQUIRE(u && values && mask && u_w, FEB200_ERR_INVALID, "dirichlet_blend: null pointer");
    if (F == 0) return FEB200_OK;
    k_dirichlet_blend<<<grid_for(F, 256), 256, 0, as_stream(s)>>>(F, u, values, mask, u_w);
    FEB_CHECK_LAUNCH("dirichlet_blend");
    return FEB200_OK;
}

int feb200_dirichlet_residual(feb200_stream s, int64_t F, const double* u, const double* values, const double* mask, double* R) {
    FEB_REQUIRE(u && values && mask && R, FEB200_ERR_INVALID, "dirichlet_residual: null pointer");
    if (F == 0) return FEB200_OK;
    k_dirichlet_residual<<<grid_for(F, 256), 256, 0, as_stream(s)>>>(F, u, values, mask, R);
    FEB_CHECK_LAUNCH("dirichlet_residual");
    return FEB200_OK;
}

int feb200_axpby(feb200_stream s, int64_t n, double a, const double* x, double b, const double* y, double* out) {
    FEB_REQUIRE(x && out, FEB200_ERR_INVALID, "axpby: null pointer");
    if (n == 0) return FEB200_OK;
    k_axpby<<<grid_for(n, 256), 256, 0, as_stream(s)>>>(n, a, x, b, y, out);
    FEB_CHECK_LAUNCH("axpby");
    return FEB200_OK;
}

int feb200_scatter_set(feb200_stream s, int64_t n_idx, const int64_t* idx, const double* vals, double scale_vals,
                       const double* minus_src, double* dst) {
    FEB_REQUIRE(idx && vals && dst, FEB200_ERR_INVALID, "scatter_set: null pointer");
    if (n_idx == 0) return FEB200_OK;
    k_scatter_set<<<grid_for(n_idx, 256), 256, 0, as_stream(s)>>>(n_idx, idx, vals, scale_vals, minus_src, dst);
    FEB_CHECK_LAUNCH("scatter_set");
    return FEB200_OK;
}

}  // extern "C"
