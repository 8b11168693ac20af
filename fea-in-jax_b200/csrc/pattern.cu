// Jacobian sparsity at node level, element -> CSR slot map, and the DoF-level index arrays.
//
// The reference finds the pattern by lexsort + diff over all E*(N*U)^2 (row, col) pairs
// (/root/reference/src/fe_jax/fea.py:490-516) and sums duplicates after another sort
// (src/fe_jax/sparse_matrix.py:44-110).  Because DoFs are node-blocked (fea.py:341-345) the
// sorted-unique DoF pattern is the U x U block expansion of the sorted-unique NODE adjacency, so it
// is built here from the assembly maps without any sort over element pairs: each vertex walks its
// incident elements (ascending slots) and keeps a sorted unique neighbour list.
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace feb {

struct BatchMaps {
    feb200_batch_map b[FEB200_MAX_BATCHES];
    int nb;
};

// sorted-unique insertion; returns false on overflow
__device__ __forceinline__ bool insert_sorted(int32_t* list, int& cnt, int32_t key) {
    int lo = 0, hi = cnt;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (list[mid] < key) lo = mid + 1; else hi = mid;
    }
    if (lo < cnt && list[lo] == key) return true;
    if (cnt >= FEB200_MAX_NODE_DEGREE) return false;
    for (int i = cnt; i > lo; --i) list[i] = list[i - 1];
    list[lo] = key;
    ++cnt;
    return true;
}

__device__ __forceinline__ int collect_neighbours(const BatchMaps& m, int64_t a, int32_t* list, bool& ok) {
    int cnt = 0;
    ok = true;
    for (int ib = 0; ib < m.nb; ++ib) {
        const feb200_batch_map& bm = m.b[ib];
        const unsigned long long s0 = bm.row_ind[a], s1 = bm.row_ind[a + 1];
        for (unsigned long long k = s0; k < s1; ++k) {
            const int64_t e = (int64_t)(bm.col_ind[k] / (unsigned long long)bm.N);
            const int32_t* c = bm.conn + e * bm.N;
            for (int n = 0; n < bm.N; ++n) ok = insert_sorted(list, cnt, c[n]) && ok;
        }
    }
    return cnt;
}

__global__ void k_pattern_count(BatchMaps m, int64_t V, int32_t* deg, int32_t* err_flag) {
    int32_t list[FEB200_MAX_NODE_DEGREE];
    for (int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; a < V; a += (int64_t)gridDim.x * blockDim.x) {
        bool ok;
        deg[a] = collect_neighbours(m, a, list, ok);
        if (!ok) atomicExch(err_flag, 1);
    }
}

__global__ void k_pattern_fill(BatchMaps m, int64_t V, const int64_t* __restrict__ nindptr, int32_t* nindices,
                               int32_t* diag_rank) {
    int32_t list[FEB200_MAX_NODE_DEGREE];
    for (int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; a < V; a += (int64_t)gridDim.x * blockDim.x) {
        bool ok;
        const int cnt = collect_neighbours(m, a, list, ok);
        const int64_t p0 = nindptr[a];
        int dr = -1;
        for (int r = 0; r < cnt; ++r) {
            nindices[p0 + r] = list[r];
            if (list[r] == (int32_t)a) dr = r;
        }
        if (diag_rank) diag_rank[a] = dr;
    }
}

__global__ void k_slot_rank(const int32_t* __restrict__ conn, int64_t n_slots, int N, const int64_t* __restrict__ nindptr,
                            const int32_t* __restrict__ nindices, uint16_t* __restrict__ slot_rank) {
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < n_slots; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = s / N;
        const int32_t a = conn[s];
        const int64_t p0 = nindptr[a];
        const int deg = (int)(nindptr[a + 1] - p0);
        const int32_t* nb = nindices + p0;
        for (int n = 0; n < N; ++n) {
            const int32_t key = conn[e * N + n];
            int lo = 0, hi = deg;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (nb[mid] < key) lo = mid + 1; else hi = mid;
            }
            slot_rank[s * N + n] = (uint16_t)lo;
        }
    }
}

// one thread per (vertex, neighbour rank): writes the U*U COO entries of that block
__global__ void k_expand_coo(int64_t V, int U, const int64_t* __restrict__ nindptr, const int32_t* __restrict__ nindices,
                             int64_t* __restrict__ coo_row, int64_t* __restrict__ coo_col) {
    for (int64_t a = blockIdx.x; a < V; a += gridDim.x) {
        const int64_t p0 = nindptr[a];
        const int deg = (int)(nindptr[a + 1] - p0);
        const int64_t base = (int64_t)U * U * p0;
        const int row_len = U * deg;
        for (int t = threadIdx.x; t < U * row_len; t += blockDim.x) {
            const int i = t / row_len;
            const int c = t - i * row_len;
            const int r = c / U, j = c - r * U;
            if (coo_row) coo_row[base + t] = a * U + i;
            if (coo_col) coo_col[base + t] = (int64_t)nindices[p0 + r] * U + j;
        }
    }
}
__global__ void k_expand_indptr(int64_t V, int U, const int64_t* __restrict__ nindptr, int64_t* __restrict__ indptr) {
    for (int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; a <= V; a += (int64_t)gridDim.x * blockDim.x) {
        if (a == V) { indptr[(int64_t)U * V] = (int64_t)U * U * nindptr[V]; continue; }
        const int64_t p0 = nindptr[a];
        const int64_t deg = nindptr[a + 1] - p0;
        for (int i = 0; i < U; ++i) indptr[a * U + i] = (int64_t)U * U * p0 + (int64_t)i * U * deg;
    }
}

__global__ void k_set_first_zero(int64_t* p) { p[0] = 0; }

static int make_maps(const feb200_batch_map* maps, int n, BatchMaps& out) {
    FEB_REQUIRE(maps && n >= 1 && n <= FEB200_MAX_BATCHES, FEB200_ERR_INVALID, "pattern: 1..%d batches supported", FEB200_MAX_BATCHES);
    out.nb = n;
    for (int i = 0; i < n; ++i) {
        FEB_REQUIRE(maps[i].conn && maps[i].row_ind && maps[i].col_ind && maps[i].N > 0, FEB200_ERR_INVALID, "pattern: bad batch map %d", i);
        out.b[i] = maps[i];
    }
    return FEB200_OK;
}

}  // namespace feb

using namespace feb;

extern "C" {

int feb200_pattern_count(feb200_stream s, const feb200_batch_map* maps, int32_t n_batches, int64_t V, int32_t* deg,
                         int32_t* err_flag) {
    BatchMaps m;
    int rc = make_maps(maps, n_batches, m);
    if (rc) return rc;
    FEB_REQUIRE(deg && err_flag, FEB200_ERR_INVALID, "pattern_count: null output");
    if (V == 0) return FEB200_OK;
    k_pattern_count<<<grid_for(V, 64), 64, 0, as_stream(s)>>>(m, V, deg, err_flag);
    FEB_CHECK_LAUNCH("pattern_count");
    return FEB200_OK;
}

int feb200_pattern_scan(feb200_stream s, const int32_t* deg, int64_t V, int64_t* nindptr, void* ws, size_t ws_bytes) {
    FEB_REQUIRE(deg && nindptr, FEB200_ERR_INVALID, "pattern_scan: null pointer");
    FEB_REQUIRE(ws_bytes >= feb200_scan_workspace_bytes(V), FEB200_ERR_WORKSPACE, "pattern_scan: workspace too small");
    cudaStream_t st = as_stream(s);
    k_set_first_zero<<<1, 1, 0, st>>>(nindptr);
    if (V > 0) {
        size_t bytes = ws_bytes;
        FEB_CUDA(cub::DeviceScan::InclusiveSum(ws, bytes, deg, nindptr + 1, V, st));
    }
    FEB_CHECK_LAUNCH("pattern_scan");
    return FEB200_OK;
}

int feb200_pattern_fill(feb200_stream s, const feb200_batch_map* maps, int32_t n_batches, int64_t V,
                        const int64_t* nindptr, int32_t* nindices, int32_t* diag_rank) {
    BatchMaps m;
    int rc = make_maps(maps, n_batches, m);
    if (rc) return rc;
    FEB_REQUIRE(nindptr && nindices, FEB200_ERR_INVALID, "pattern_fill: null pointer");
    if (V == 0) return FEB200_OK;
    k_pattern_fill<<<grid_for(V, 64), 64, 0, as_stream(s)>>>(m, V, nindptr, nindices, diag_rank);
    FEB_CHECK_LAUNCH("pattern_fill");
    return FEB200_OK;
}

int feb200_slot_rank(feb200_stream s, const int32_t* conn, int64_t E, int32_t N, const int64_t* nindptr,
                     const int32_t* nindices, uint16_t* slot_rank) {
    FEB_REQUIRE(conn && nindptr && nindices && slot_rank && N > 0, FEB200_ERR_INVALID, "slot_rank: bad arguments");
    if (E == 0) return FEB200_OK;
    k_slot_rank<<<grid_for(E * N, 128), 128, 0, as_stream(s)>>>(conn, E * (int64_t)N, N, nindptr, nindices, slot_rank);
    FEB_CHECK_LAUNCH("slot_rank");
    return FEB200_OK;
}

int feb200_pattern_expand(feb200_stream s, int64_t V, int32_t U, const int64_t* nindptr, const int32_t* nindices,
                          int64_t* coo_row, int64_t* coo_col, int64_t* csr_indptr) {
    FEB_REQUIRE(nindptr && nindices && U > 0, FEB200_ERR_INVALID, "pattern_expand: bad arguments");
    cudaStream_t st = as_stream(s);
    if (V > 0 && (coo_row || coo_col)) {
        int g = (int)(V < (int64_t)sm_count() * 32 ? V : (int64_t)sm_count() * 32);
        k_expand_coo<<<g, 128, 0, st>>>(V, U, nindptr, nindices, coo_row, coo_col);
    }
    if (csr_indptr) k_expand_indptr<<<grid_for(V + 1, 256), 256, 0, st>>>(V, U, nindptr, csr_indptr);
    FEB_CHECK_LAUNCH("pattern_expand");
    return FEB200_OK;
}

}  // extern "C"
