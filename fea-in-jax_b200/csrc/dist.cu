// Multi-GPU data plane of the element-partitioned mode: interface halo-sum and the CG loop with its scalar
// all-reduces, enqueued from C on the caller's stream -- no host code between two collectives.
//
// Message pattern after the reference's only distributed code, the mpi4py demo
// tests/test_mpi4py_cuda/test_mpi4py_cuda_jax_async/test_mpi4py_cuda_jax_async.py:132,229-244 (neighbour sends / receives
// of boundary values, sum on arrival), and the remark fea.py:331-332 that rank-local numbering is future work.
//   halo_sum   pack the interface DoFs shared with every neighbour -> ncclSend / ncclRecv in one group -> add the received
//              partial sums.  Every interface vertex is shared by exactly two ranks (slab / strip partitions), so the sum
//              own + received is the same bit pattern on both sharers.
//   cg_*_dist  the staged CG of krylov.cu (feb200_cg_stage) with the halo-sum after the local SpMV and ncclAllReduce of the
//              dot products between the stages; kernels exit at once when `done` is set, so the caller enqueues iterations in
//              chunks and polls the flag in between (one host sync per chunk).
// NCCL is bound at run time with dlopen (the library torch already loaded; no link-time dependency): a build box without
// NCCL still produces libfeb200.so, and the entry points report FEB200_ERR_UNSUPPORTED there.
#include <dlfcn.h>

#include "common.cuh"

namespace feb {

typedef struct ncclComm* nccl_comm_t;
struct nccl_uid { char internal[128]; };
enum { NCCL_SUM = 0, NCCL_FLOAT64 = 8 };     // nccl.h: ncclSum = 0, ncclFloat64 = 8

struct NcclApi {
    bool tried = false, ok = false;
    int (*GetUniqueId)(nccl_uid*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, nccl_uid, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static bool nccl_load() {
    if (g_nccl.tried) return g_nccl.ok;
    g_nccl.tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);       // the copy the host framework already mapped
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return false;
#define FEB_NCCL_SYM(field, name)                                          \
    *reinterpret_cast<void**>(&g_nccl.field) = dlsym(h, name);            \
    if (!g_nccl.field) return false;
    FEB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    FEB_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    FEB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    FEB_NCCL_SYM(GroupStart, "ncclGroupStart")
    FEB_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    FEB_NCCL_SYM(Send, "ncclSend")
    FEB_NCCL_SYM(Recv, "ncclRecv")
    FEB_NCCL_SYM(AllReduce, "ncclAllReduce")
    FEB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef FEB_NCCL_SYM
    g_nccl.ok = true;
    return true;
}

#define FEB_NCCL(call)                                                                          \
    do {                                                                                        \
        int r__ = (call);                                                                       \
        if (r__ != 0) {                                                                         \
            feb::set_error("%s failed: %s", #call, feb::g_nccl.GetErrorString(r__));            \
            return FEB200_ERR_CUDA;                                                             \
        }                                                                                       \
    } while (0)

// buf[i*U + j] = v[idx[i]*U + j]
__global__ void k_halo_pack(int64_t n, int U, const int64_t* __restrict__ idx, const double* __restrict__ v,
                            double* __restrict__ buf) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n * U; t += stride) {
        const int64_t i = t / U;
        buf[t] = v[idx[i] * U + (t - i * U)];
    }
}
// v[idx[i]*U + j] += buf[i*U + j]   (idx has no repeated entries: plain read-modify-write, no atomics)
__global__ void k_halo_unpack_add(int64_t n, int U, const int64_t* __restrict__ idx, const double* __restrict__ buf,
                                  double* __restrict__ v) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n * U; t += stride) {
        const int64_t i = t / U;
        const int64_t d = idx[i] * U + (t - i * U);
        v[d] = v[d] + buf[t];
    }
}

static int check_halo(const feb200_halo* h) {
    FEB_REQUIRE(h && h->n_neighbours >= 0 && h->n_neighbours <= FEB200_MAX_NEIGHBOURS && h->U >= 1, FEB200_ERR_INVALID,
                "halo: bad descriptor");
    for (int k = 0; k < h->n_neighbours; ++k) {
        FEB_REQUIRE(h->count[k] >= 0 && (h->count[k] == 0 || (h->idx[k] && h->send[k] && h->recv[k])), FEB200_ERR_INVALID,
                    "halo: neighbour %d has null buffers", k);
        FEB_REQUIRE(k == 0 || h->peer[k] > h->peer[k - 1], FEB200_ERR_INVALID, "halo: neighbours must be listed in ascending rank order");
    }
    return FEB200_OK;
}

static int halo_pack(cudaStream_t cs, const feb200_halo* h, const double* v) {
    for (int k = 0; k < h->n_neighbours; ++k)
        if (h->count[k] > 0)
            k_halo_pack<<<grid_for(h->count[k] * h->U, 256), 256, 0, cs>>>(h->count[k], h->U, h->idx[k], v, h->send[k]);
    return FEB200_OK;
}
static int halo_unpack_add(cudaStream_t cs, const feb200_halo* h, double* v) {
    for (int k = 0; k < h->n_neighbours; ++k)      // ascending rank order
        if (h->count[k] > 0)
            k_halo_unpack_add<<<grid_for(h->count[k] * h->U, 256), 256, 0, cs>>>(h->count[k], h->U, h->idx[k], h->recv[k], v);
    return FEB200_OK;
}
static int halo_exchange(cudaStream_t cs, nccl_comm_t comm, const feb200_halo* h) {
    if (h->n_neighbours == 0) return FEB200_OK;
    FEB_NCCL(g_nccl.GroupStart());
    for (int k = 0; k < h->n_neighbours; ++k) {
        if (h->count[k] == 0) continue;
        FEB_NCCL(g_nccl.Send(h->send[k], (size_t)(h->count[k] * h->U), NCCL_FLOAT64, h->peer[k], comm, cs));
        FEB_NCCL(g_nccl.Recv(h->recv[k], (size_t)(h->count[k] * h->U), NCCL_FLOAT64, h->peer[k], comm, cs));
    }
    FEB_NCCL(g_nccl.GroupEnd());
    return FEB200_OK;
}
static int halo_sum(cudaStream_t cs, nccl_comm_t comm, const feb200_halo* h, double* v) {
    int rc;
    if ((rc = halo_pack(cs, h, v))) return rc;
    if ((rc = halo_exchange(cs, comm, h))) return rc;
    return halo_unpack_add(cs, h, v);
}

}  // namespace feb

using namespace feb;

extern "C" {

int feb200_nccl_available(void) { return nccl_load() ? 1 : 0; }

int feb200_nccl_unique_id(uint8_t* id128) {
    FEB_REQUIRE(id128, FEB200_ERR_INVALID, "nccl_unique_id: null pointer");
    FEB_REQUIRE(nccl_load(), FEB200_ERR_UNSUPPORTED, "nccl_unique_id: libnccl.so.2 not found");
    nccl_uid id;
    FEB_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, id.internal, 128);
    return FEB200_OK;
}

int feb200_nccl_comm_init(int32_t rank, int32_t world, const uint8_t* id128, void** comm_out) {
    FEB_REQUIRE(id128 && comm_out && world >= 1 && rank >= 0 && rank < world, FEB200_ERR_INVALID, "nccl_comm_init: bad argument");
    FEB_REQUIRE(nccl_load(), FEB200_ERR_UNSUPPORTED, "nccl_comm_init: libnccl.so.2 not found");
    nccl_uid id;
    memcpy(id.internal, id128, 128);
    nccl_comm_t c = nullptr;
    FEB_NCCL(g_nccl.CommInitRank(&c, world, id, rank));
    *comm_out = c;
    return FEB200_OK;
}

int feb200_nccl_comm_destroy(void* comm) {
    if (!comm) return FEB200_OK;
    FEB_REQUIRE(nccl_load(), FEB200_ERR_UNSUPPORTED, "nccl_comm_destroy: libnccl.so.2 not found");
    FEB_NCCL(g_nccl.CommDestroy(reinterpret_cast<nccl_comm_t>(comm)));
    return FEB200_OK;
}

int feb200_halo_pack(feb200_stream s, const feb200_halo* h, const double* v) {
    int rc = check_halo(h);
    if (rc) return rc;
    FEB_REQUIRE(v, FEB200_ERR_INVALID, "halo_pack: null vector");
    rc = halo_pack(as_stream(s), h, v);
    FEB_CHECK_LAUNCH("halo_pack");
    return rc;
}

int feb200_halo_unpack_add(feb200_stream s, const feb200_halo* h, double* v) {
    int rc = check_halo(h);
    if (rc) return rc;
    FEB_REQUIRE(v, FEB200_ERR_INVALID, "halo_unpack_add: null vector");
    rc = halo_unpack_add(as_stream(s), h, v);
    FEB_CHECK_LAUNCH("halo_unpack_add");
    return rc;
}

int feb200_halo_sum(feb200_stream s, void* comm, const feb200_halo* h, double* v) {
    int rc = check_halo(h);
    if (rc) return rc;
    FEB_REQUIRE(comm && v, FEB200_ERR_INVALID, "halo_sum: null communicator / vector");
    FEB_REQUIRE(nccl_load(), FEB200_ERR_UNSUPPORTED, "halo_sum: libnccl.so.2 not found");
    rc = halo_sum(as_stream(s), reinterpret_cast<nccl_comm_t>(comm), h, v);
    if (rc) return rc;
    FEB_CHECK_LAUNCH("halo_sum");
    return FEB200_OK;
}

int feb200_dot_allreduce(feb200_stream s, void* comm, double* buf, int32_t n) {
    FEB_REQUIRE(comm && buf && n >= 0, FEB200_ERR_INVALID, "dot_allreduce: bad argument");
    FEB_REQUIRE(nccl_load(), FEB200_ERR_UNSUPPORTED, "dot_allreduce: libnccl.so.2 not found");
    if (n == 0) return FEB200_OK;
    FEB_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, NCCL_FLOAT64, NCCL_SUM, reinterpret_cast<nccl_comm_t>(comm), as_stream(s)));
    return FEB200_OK;
}

// scalars: {gamma, pAp, rr, gamma_new, atol2, bb, tol, atol} (krylov.cu)
int feb200_cg_init_dist(feb200_stream s, void* comm, const feb200_cg_state* st, const feb200_halo* h, double tol, double atol) {
    int rc = check_halo(h);
    if (rc) return rc;
    FEB_REQUIRE(comm && st, FEB200_ERR_INVALID, "cg_init_dist: null communicator / state");
    FEB_REQUIRE(nccl_load(), FEB200_ERR_UNSUPPORTED, "cg_init_dist: libnccl.so.2 not found");
    nccl_comm_t c = reinterpret_cast<nccl_comm_t>(comm);
    cudaStream_t cs = as_stream(s);
    if ((rc = feb200_cg_stage(s, st, FEB200_CG_INIT_SPMV, tol, atol))) return rc;
    if ((rc = halo_sum(cs, c, h, st->Ap))) return rc;
    if ((rc = feb200_cg_stage(s, st, FEB200_CG_INIT_VECTORS, tol, atol))) return rc;
    FEB_NCCL(g_nccl.AllReduce(st->scalars + 2, st->scalars + 2, 2, NCCL_FLOAT64, NCCL_SUM, c, cs));      // r.r, r.z
    FEB_NCCL(g_nccl.AllReduce(st->scalars + 5, st->scalars + 5, 1, NCCL_FLOAT64, NCCL_SUM, c, cs));      // b.b
    return feb200_cg_stage(s, st, FEB200_CG_INIT_FINALIZE, tol, atol);
}

int feb200_cg_iterate_dist(feb200_stream s, void* comm, const feb200_cg_state* st, const feb200_halo* h, int32_t n_iters) {
    int rc = check_halo(h);
    if (rc) return rc;
    FEB_REQUIRE(comm && st && n_iters >= 0, FEB200_ERR_INVALID, "cg_iterate_dist: bad argument");
    FEB_REQUIRE(nccl_load(), FEB200_ERR_UNSUPPORTED, "cg_iterate_dist: libnccl.so.2 not found");
    nccl_comm_t c = reinterpret_cast<nccl_comm_t>(comm);
    cudaStream_t cs = as_stream(s);
    for (int it = 0; it < n_iters; ++it) {
        if ((rc = feb200_cg_stage(s, st, FEB200_CG_SPMV, 0, 0))) return rc;                              // Ap = A_loc p, local p.(A_loc p)
        if ((rc = halo_sum(cs, c, h, st->Ap))) return rc;
        FEB_NCCL(g_nccl.AllReduce(st->scalars + 1, st->scalars + 1, 1, NCCL_FLOAT64, NCCL_SUM, c, cs));  // p.Ap
        if ((rc = feb200_cg_stage(s, st, FEB200_CG_UPDATE_XR, 0, 0))) return rc;
        FEB_NCCL(g_nccl.AllReduce(st->scalars + 2, st->scalars + 2, 2, NCCL_FLOAT64, NCCL_SUM, c, cs));  // r.r, r.z
        if ((rc = feb200_cg_stage(s, st, FEB200_CG_UPDATE_P, 0, 0))) return rc;
    }
    return FEB200_OK;
}

}  // extern "C"
