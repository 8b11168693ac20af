/*
 * feb200 -- C ABI of the B200-native hot path for fea-in-jax.
 *
 * Every entry point
 *   - takes the CUDA stream to enqueue on, raw DEVICE pointers, int64 sizes and enum codes;
 *   - returns 0 on success or a negative feb200_status (message: feb200_last_error(), thread local);
 *   - never throws, never synchronises the stream and never allocates: scratch comes from the
 *     caller (size via the matching *_workspace_bytes function);
 *   - is re-entrant and CUDA-graph capturable.
 * This is what an XLA-FFI handler (jax.ffi) or a ctypes stub binds; see INTEGRATION.md.
 *
 * The reference (KeithBallard/fea-in-jax) has no FFI of its own for this path: each function
 * below cites the Python function whose device work it replaces (paths relative to the
 * reference root).  FP64 everywhere (src/fe_jax/__init__.py:3 enables x64).
 *
 * Index conventions: dof(v, j) = U*v + j (src/fe_jax/fea.py:341-345); slot(e, n) = e*N + n.
 */
#ifndef FEB200_H
#define FEB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* feb200_stream;   /* cudaStream_t */

typedef enum {
    FEB200_OK = 0,
    FEB200_ERR_INVALID = -1,      /* bad argument (shape / enum / null pointer) */
    FEB200_ERR_UNSUPPORTED = -2,  /* no kernel variant for this element / material */
    FEB200_ERR_WORKSPACE = -3,    /* workspace too small */
    FEB200_ERR_CUDA = -4,         /* a CUDA runtime call failed (launch error) */
    FEB200_ERR_CAPACITY = -5      /* a fixed device-side capacity was exceeded */
} feb200_status;

/* element_residual_func / constitutive_model pairs of the reference, identified by code */
typedef enum { FEB200_KIND_LINEAR = 0,   /* linear_elasticity_residual  linear_elasticity.py:145-206 */
               FEB200_KIND_HYPER = 1     /* hyperelasticity_residual    hyperelasticity.py:93-139   */
} feb200_kind;
typedef enum { FEB200_MAT_ISOTROPIC = 0,     /* elastic_isotropic    linear_elasticity.py:9-59   */
               FEB200_MAT_ORTHOTROPIC = 1,   /* elastic_orthotropic  linear_elasticity.py:62-142 */
               FEB200_MAT_SVK = 2,           /* st_venant_kirchhoff  hyperelasticity.py:9-53     */
               FEB200_MAT_MOONEY_RIVLIN = 3  /* mooney_rivlin        hyperelasticity.py:56-91    */
} feb200_material;
/* material_params array type, fea.py:77-80 */
typedef enum { FEB200_LAYOUT_M = 0, FEB200_LAYOUT_EM = 1, FEB200_LAYOUT_EQM = 2 } feb200_layout;

const char* feb200_last_error(void);
int feb200_version(void);

/* One element batch as the kernels see it (device pointers).  Mirrors the per-batch slices of
 * ElementBatchCollection, fea.py:88-345. */
typedef struct {
    int64_t E;                 /* elements */
    int32_t N;                 /* nodes per element */
    int32_t Q;                 /* quadrature points */
    int32_t D;                 /* spatial = parametric dimension (= U dofs per node) */
    int32_t kind;              /* feb200_kind */
    int32_t material;          /* feb200_material */
    int32_t layout;            /* feb200_layout */
    int32_t M;                 /* material parameters per point */
    int32_t reserved;
    const int32_t* conn;       /* (E, N) vertex ids */
    const double* dphi;        /* (Q, N, D) d phi / d xi */
    const double* W;           /* (Q,) weights as applied (the shim passes W[0] broadcast, fea.py:285-297) */
    const double* mp;          /* material params in `layout` */
} feb200_batch;

/* ------------------------------------------------------------------ index maps (setup.py:212-246)
 * build_row_ind: row_ind (V+1) uint64 = exclusive scan of vertex valence.        setup.py:212-224
 * build_col_ind: col_ind (E*N) uint64 = per vertex its slots ascending; data = 1. setup.py:227-246 */
size_t feb200_build_row_ind_workspace_bytes(int64_t V);
int feb200_build_row_ind(feb200_stream s, const int64_t* cells, int64_t E, int32_t N, int64_t V,
                         uint64_t* row_ind, void* ws, size_t ws_bytes);
size_t feb200_build_col_ind_workspace_bytes(int64_t V);
int feb200_build_col_ind(feb200_stream s, const int64_t* cells, int64_t E, int32_t N, int64_t V,
                         const uint64_t* row_ind, uint64_t* col_ind, double* data, void* ws, size_t ws_bytes);

/* ------------------------------------------------------------------ gather / scatter (setup.py:304-356)
 * gather:      out[e,n,:] = v[conn[e,n],:]                               setup.py:304-322
 * scatter_sum: out[v,:] (+)= sum over the slots of v, ascending, of v_en  setup.py:344-356
 *              accumulate != 0 adds to `out` (batch loop of fea.py:885-889).
 *              If mask != NULL the Dirichlet wrapper of fea.py:961-963 is fused:
 *              out = (1-m)*out + m*(u - g). */
int feb200_gather(feb200_stream s, const int32_t* conn, int64_t E, int32_t N, int32_t U, const double* v,
                  double* out);
int feb200_scatter_sum(feb200_stream s, const uint64_t* row_ind, const uint64_t* col_ind, int64_t V, int32_t U,
                       const double* v_en, double* out, int32_t accumulate, const double* mask,
                       const double* values, const double* u);
/* u_w = (1-m)*u + m*g   fea.py:948-950 ;  R = (1-m)*R + m*(u-g)   fea.py:961-963 */
int feb200_dirichlet_blend(feb200_stream s, int64_t F, const double* u, const double* values, const double* mask,
                           double* u_w);
int feb200_dirichlet_residual(feb200_stream s, int64_t F, const double* u, const double* values,
                              const double* mask, double* R);

/* ------------------------------------------------------------------ element kernels
 * residual: R_e (E,N,D) for one batch; gathers x from the vertex array x_vd (V,D) and u from the
 *           global vector u (V*D).  If mask != NULL, u is blended on the fly (fea.py:948-950).
 *           Replaces _calculate_residual_wo_dirichlet_batch, fea.py:800-847.
 * tangent:  K_e stored node-block major: Ke[(e*N+a)][b][D][D] = d R[a,:] / d u[b,:]
 *           Replaces _calculate_jacobian_batch_element_kernel (vmap(jacfwd)), fea.py:519-572. */
int feb200_element_residual(feb200_stream s, const feb200_batch* b, const double* x_vd, const double* u,
                            const double* mask, const double* values, double* R_e);
int feb200_element_tangent(feb200_stream s, const feb200_batch* b, const double* x_vd, const double* u,
                           const double* mask, const double* values, double* K_e);

/* diag_e (E, N, D): diagonal of the element tangent of the linear isotropic law, K_e never formed; scatter-summed by the
 * caller it is calculate_jacobian_diag_wo_dirichlet (fea.py:759-797) without assembling the matrix.  Other laws:
 * FEB200_ERR_UNSUPPORTED (the caller reads the diagonal off the assembled CSR instead). */
int feb200_element_tangent_diag(feb200_stream s, const feb200_batch* b, const double* x_vd, double* diag_e);

/* ------------------------------------------------------------------ Jacobian sparsity (fea.py:490-516)
 * The pattern is held at NODE level: nindptr (V+1) int64, nindices (nnz_nodes) int32 = sorted unique
 * neighbour list of each vertex over all batches.  The DoF-level CSR the reference produces
 * (sorted-unique (row, col), sparse_matrix.py:44-153) follows without searching:
 *   indptr[U*a+i] = U*U*nindptr[a] + i*U*deg(a),  indices[indptr[U*a+i] + U*r + j] = U*nindices[nindptr[a]+r] + j
 * so CSR `data` is stored in exactly the reference order and nnz = U*U*nnz_nodes. */
typedef struct {
    int64_t E;
    int32_t N;
    int32_t reserved;
    const int32_t* conn;       /* (E, N) */
    const uint64_t* row_ind;   /* (V+1) assembly map of this batch */
    const uint64_t* col_ind;   /* (E*N) */
} feb200_batch_map;
#define FEB200_MAX_BATCHES 16
#define FEB200_MAX_NODE_DEGREE 512

/* pass 1: deg[a] (int32) ; pass 2 (after the caller scanned deg into nindptr): nindices.
 * err_flag (device int32) is set non-zero if a vertex has more than FEB200_MAX_NODE_DEGREE neighbours. */
int feb200_pattern_count(feb200_stream s, const feb200_batch_map* maps, int32_t n_batches, int64_t V,
                         int32_t* deg, int32_t* err_flag);
size_t feb200_scan_workspace_bytes(int64_t n);
/* nindptr[0] = 0, nindptr[a+1] = sum_{a' <= a} deg[a'] */
int feb200_pattern_scan(feb200_stream s, const int32_t* deg, int64_t V, int64_t* nindptr, void* ws, size_t ws_bytes);
int feb200_pattern_fill(feb200_stream s, const feb200_batch_map* maps, int32_t n_batches, int64_t V,
                        const int64_t* nindptr, int32_t* nindices, int32_t* diag_rank);
/* slot_rank[(e*N+a)*N + b] = position of conn[e,b] in the neighbour list of conn[e,a] (uint16). */
int feb200_slot_rank(feb200_stream s, const int32_t* conn, int64_t E, int32_t N, const int64_t* nindptr,
                     const int32_t* nindices, uint16_t* slot_rank);
/* DoF-level COO / CSR index arrays exactly as the reference returns them (int64). Any output may be NULL. */
int feb200_pattern_expand(feb200_stream s, int64_t V, int32_t U, const int64_t* nindptr, const int32_t* nindices,
                          int64_t* coo_row, int64_t* coo_col, int64_t* csr_indptr);

/* ------------------------------------------------------------------ deterministic CSR assembly
 * values (U*U*nnz_nodes) (+)= segmented sum of the element blocks, ascending element order per entry
 * (the order a stable (row,col) sort + sequential bincount gives, fea.py:651-658).
 * Replaces COO-concatenate + _sort_indices + coo_sum_duplicates. max_deg = max vertex degree.
 * If mask != NULL (with nindices) the Dirichlet elimination of feb200_dirichlet_lhs is applied to the
 * finished sums while they are stored (pass it with the LAST batch only). */
int feb200_assemble_csr(feb200_stream s, const feb200_batch_map* map, int64_t V, int32_t U, const double* K_e,
                        const uint16_t* slot_rank, const int64_t* nindptr, int32_t max_deg, double* values,
                        int32_t accumulate, const int32_t* nindices, const double* mask);
/* The same with the constraints as one byte per vertex (bit j set when DoF U*v+j is constrained; NULL = no
 * elimination) instead of the (U*V) mask of doubles: the fused elimination then reads one byte per CSR block. */
int feb200_assemble_csr_flags(feb200_stream s, const feb200_batch_map* map, int64_t V, int32_t U, const double* K_e,
                              const uint16_t* slot_rank, const int64_t* nindptr, int32_t max_deg, double* values,
                              int32_t accumulate, const int32_t* nindices, const uint8_t* vertex_flags);
/* Fused assembly for linear isotropic hex8 batches (material constant per element): element tangents are
 * computed and summed into the CSR row panels in ONE kernel (K_e never reaches HBM).  The tile plan groups
 * vertices that are close in space; per tile: its vertices, the elements touching them and one packed item per
 * (vertex, incident slot): le | n_a << 12 | a_local << 16 | k << 24, sorted by (le, n_a).  Built by
 * fe_jax/fused_plan.py for any connectivity with vertex valence <= 8.  values are overwritten (single batch);
 * vertex_flags != NULL fuses apply_dirichlet_bcs_lhs: one byte per vertex, bit j set when DoF 3*v+j is
 * constrained.  Vertices may have at most 8 incident slots and 32 neighbours; te_cap <= 96.
 * Same summation order as feb200_assemble_csr. */
typedef struct {
    int32_t n_tiles;
    int32_t tn_cap;               /* max vertices per tile (multiple of 4, <= 32) */
    int32_t te_cap;               /* max elements per tile */
    int32_t reserved;
    const int32_t* tile_node_ptr; /* (n_tiles+1) */
    const int32_t* tile_nodes;    /* vertex ids, tile-major */
    const int32_t* tile_elem_ptr; /* (n_tiles+1) */
    const int32_t* tile_elems;    /* element ids, ascending within a tile */
    const int32_t* tile_item_ptr; /* (n_tiles+1) */
    const uint32_t* tile_items;
    const int64_t* tile_node_p0;  /* nindptr[vertex], parallel to tile_nodes */
    const int32_t* tile_node_meta;/* degree | valence << 8, parallel to tile_nodes */
} feb200_fused_plan;
size_t feb200_fused_smem_bytes(int32_t tn_cap, int32_t te_cap);
int feb200_assemble_fused_hex8(feb200_stream s, const feb200_batch* b, const feb200_fused_plan* plan, const double* x_vd,
                               const uint64_t* row_ind, const uint16_t* slot_rank, const int64_t* nindptr,
                               const int32_t* nindices, const uint8_t* vertex_flags, double* values);
/* The same contract (same plan, same outputs) with the block products of an element evaluated as one 24 x 24 x 8
 * FP64 tensor-core product (mma.sync m8n8k4) instead of 3x3 blocks one by one; material layouts (M,) and (E,M), and
 * the per-point layout (E,Q,M) when the Poisson ratio is constant within each element (the caller checks; fields
 * whose Poisson ratio varies inside an element go to feb200_assemble_fused_hex8).
 * material_hints: FEB200_HINT_UNIFORM_POISSON = every entry of the material array has the same Poisson ratio (heterogeneous
 * stiffness, one Poisson ratio): lambda / mu is then one number and the blocks are combined once per CSR block instead
 * of once per contribution.  The caller guarantees what it hints. */
#define FEB200_HINT_UNIFORM_POISSON 1
size_t feb200_fused_mma_smem_bytes(int32_t tn_cap, int32_t te_cap, int32_t material_layout);
int feb200_assemble_fused_hex8_mma(feb200_stream s, const feb200_batch* b, const feb200_fused_plan* plan, const double* x_vd,
                                   const uint16_t* slot_rank, const int32_t* nindices, const uint8_t* vertex_flags,
                                   int32_t material_hints, double* values);

/* Streaming variant of the fused hex8 assembly (same outputs as feb200_assemble_fused_hex8_mma up to the summation order,
 * which here is the plan's visit order; results are bit-reproducible from run to run).  Every warp walks "segments" of
 * vertex strips: a segment is a list of chunks (4 vertices each; the two chunks c, c+1 are live at step c and own the
 * 8 panel slots of the warp, chunk c in slots 4*(c&1)..) and a list of visits sorted by step.  A visit is one 48-byte
 * record of 12 int32 {element, slotw, perm, flags, conn[8]}: slotw nibble n = panel slot of local node n (0xF: the node is
 * not in the window), perm nibble i = i-th in-window node (ascending), flags = count of in-window nodes | (chunks
 * completed by this visit) << 4 | (runs interleaved with the next visit: disjoint panel slots, positions 2m / 2m+1 of the
 * segment, both with <= 4 in-window nodes) << 8, conn = the element's vertex ids.  slot_rank8 = the slot-rank map of
 * feb200_slot_rank narrowed to one byte per node pair (degree <= 32).  A chunk record is 64 bytes: int32 v[4] (vertex id,
 * -1: empty slot), int32 deg[4] (its degree), int64 p0[4] (nindptr[vertex]).
 * Built by fe_jax/sweep_plan.py for any hex8 connectivity with vertex valence <= 8 and degree <= 32.  Needs one Poisson
 * ratio for the batch: layout (M,) or FEB200_HINT_UNIFORM_POISSON.  block_flags != NULL fuses apply_dirichlet_bcs_lhs:
 * 32 bytes per vertex from feb200_sweep_block_flags (byte r = flags of CSR block r of the vertex row).  accumulate != 0 adds
 * to `values` (later batches of a multi-batch mesh, fea.py:885-889; rows of vertices the batch does not touch stay as they
 * are; pass block_flags with the LAST batch only: the elimination acts on the finished sums).  Replaces
 * fea.py:519-572 + :613-660 + sparse_matrix.py:157-184 like the kernels above. */
typedef struct {
    int32_t n_segments;
    int32_t max_deg;               /* largest vertex degree of the mesh (panel slot = 9 * max_deg doubles) */
    int32_t reserved[2];
    const int32_t* seg_visit_ptr;  /* (n_segments+1) */
    const int32_t* seg_chunk_ptr;  /* (n_segments+1) */
    const int32_t* visits;         /* (n_visits, 12), 16-byte aligned */
    const int32_t* chunk_rec;      /* (n_chunks, 16), 16-byte aligned */
} feb200_sweep_plan;
int32_t feb200_fused_sweep_warps(int32_t max_deg);   /* workers (warps) per SM the kernel will run with; 0 = does not fit */
/* block_flags[32*v + r] = (constrained components of neighbour r of v) | (neighbour == v) << 3 | (constrained
 * components of v) << 4, from the per-vertex flags (bit j: DoF 3*v+j constrained) and the node pattern. */
int feb200_sweep_block_flags(feb200_stream s, int64_t V, const int64_t* nindptr, const int32_t* nindices,
                             const uint8_t* vertex_flags, uint8_t* block_flags);
int feb200_assemble_fused_hex8_sweep(feb200_stream s, const feb200_batch* b, const feb200_sweep_plan* plan, const double* x_vd,
                                     const uint8_t* slot_rank8, const uint8_t* block_flags, int32_t material_hints,
                                     double* values, int32_t accumulate);

/* apply_dirichlet_bcs_lhs on the node-pattern CSR (sparse_matrix.py:157-184). mask: (U*V) 0/1 doubles. */
int feb200_dirichlet_lhs(feb200_stream s, int64_t V, int32_t U, const int64_t* nindptr, const int32_t* nindices,
                         const double* mask, double* values);
/* the same on arbitrary COO triples */
int feb200_dirichlet_lhs_coo(feb200_stream s, int64_t nnz, const int64_t* row, const int64_t* col,
                             const double* mask, const double* data_in, double* data_out);
/* diag[U*a+i] = values entry (U*a+i, U*a+i)  (calculate_jacobian_diag_wo_dirichlet, fea.py:759-797) */
int feb200_csr_diagonal(feb200_stream s, int64_t V, int32_t U, const int64_t* nindptr, const int32_t* diag_rank,
                        const double* values, double* diag);

/* ------------------------------------------------------------------ generic COO helpers (sparse_matrix.py)
 * sorted-unique of pre-sorted (row, col) with summed data (coo_arrays_sum_duplicates_jit :44-92).
 * head[i] (int32) = 1 where entry i starts a new (row,col); seg[i] (int64) = inclusive scan of head - 1. */
int feb200_coo_mark_heads(feb200_stream s, int64_t n, const int64_t* row, const int64_t* col, int32_t* head);
int feb200_coo_scan_heads(feb200_stream s, int64_t n, const int32_t* head, int64_t* seg, void* ws, size_t ws_bytes);
int feb200_coo_segment_sum(feb200_stream s, int64_t n, const int64_t* row, const int64_t* col, const double* data,
                           const int32_t* head, const int64_t* seg, int64_t n_out, double* out_data,
                           int64_t* out_row, int64_t* out_col);
/* indptr[r+1] = #entries with row <= r  for sorted rows (coo_to_csr :113-153) */
int feb200_coo_rows_to_indptr(feb200_stream s, int64_t nnz, const int64_t* row, int64_t n_rows, int64_t* indptr);

/* ------------------------------------------------------------------ Krylov (solve_cg.py:20-63)
 * y = A x on the node-pattern CSR.  If dot_out != NULL also dot_out[0] = x.y (deterministic). */
size_t feb200_reduce_workspace_bytes(void);
int feb200_spmv(feb200_stream s, int64_t V, int32_t U, const int64_t* nindptr, const int32_t* nindices,
                const double* values, const double* x, double* y, double* dot_out, void* ws, size_t ws_bytes);
/* generic CSR (int64 indptr / indices), one warp per row */
int feb200_spmv_csr(feb200_stream s, int64_t n_rows, const int64_t* indptr, const int64_t* indices,
                    const double* values, const double* x, double* y);
/* out[0] = a.b */
int feb200_dot(feb200_stream s, int64_t n, const double* a, const double* b, double* out, void* ws, size_t ws_bytes);

/* Device-resident CG state: scalars (double[8]) = {gamma, pAp, rr, gamma_new, atol2, bb, tol, atol},
 * counters (int32[4]) = {k, done, maxiter, -}.  cg_init computes r0 = b - A x0, z0 = M r0, p0 = z0,
 * gamma0, atol2 = max(tol^2 b.b, atol^2), hist[0] = |r0| and the first loop test; cg_iterate enqueues
 * `n_iters` iterations whose kernels exit at once when `done` is set, so the caller may enqueue maxiter
 * of them without ever synchronising.  inv_diag == NULL means M = identity (rs = gamma), otherwise the
 * Jacobi preconditioner z = r * inv_diag (rs = r.r), sparse_linear_solve.py:301-306. */
typedef struct {
    int64_t V;
    int32_t U;
    int32_t maxiter;
    const int64_t* nindptr;
    const int32_t* nindices;
    const double* values;
    const double* inv_diag;   /* may be NULL */
    const double* weights;    /* may be NULL; multi-GPU ownership weights (1 = DoF owned by this rank, else 0) */
    const double* b;
    double* x;                /* in: x0, out: solution */
    double* r;
    double* p;
    double* Ap;
    double* hist;             /* (maxiter,) residual-norm history */
    double* scalars;          /* 8 doubles */
    int32_t* counters;        /* 4 int32 */
    void* ws;                 /* feb200_reduce_workspace_bytes() */
    size_t ws_bytes;
} feb200_cg_state;
int feb200_cg_init(feb200_stream s, const feb200_cg_state* st, double tol, double atol);
int feb200_cg_iterate(feb200_stream s, const feb200_cg_state* st, int32_t n_iters);
/* The same loop cut into stages, so that an element-partitioned multi-GPU driver can put the interface
 * halo-sum (after the SpMV stages) and the scalar all-reduces (after INIT_VECTORS: scalars[2,3,5]; after SPMV:
 * scalars[1] = p.(A_local p), whose sum over the ranks is the global p.Ap; after UPDATE_XR: scalars[2,3]) between them.
 * PAP (weighted p.Ap from the halo-summed product) gives the same number and is kept for drivers that want it.  With `weights` the dot products count every DoF
 * once (owned DoFs only). */
typedef enum { FEB200_CG_INIT_SPMV = 0, FEB200_CG_INIT_VECTORS = 1, FEB200_CG_INIT_FINALIZE = 2, FEB200_CG_SPMV = 3,
               FEB200_CG_PAP = 4, FEB200_CG_UPDATE_XR = 5, FEB200_CG_UPDATE_P = 6 } feb200_cg_stage_id;
int feb200_cg_stage(feb200_stream s, const feb200_cg_state* st, int32_t stage, double tol, double atol);

/* ------------------------------------------------------------------ element-partitioned multi-GPU data plane
 * The reference has no distributed solver (fea.py:331-332 anticipates rank-local numbering; the only message-passing code
 * is the mpi4py demo tests/test_mpi4py_cuda/test_mpi4py_cuda_jax_async/test_mpi4py_cuda_jax_async.py:132,229-244).  Every
 * rank holds the sub-assembled matrix of its elements; interface rows hold partial sums.
 * NCCL is bound at run time (dlopen of the libnccl.so.2 the host framework loaded); feb200_nccl_available() == 0 means the
 * entry points below that take a communicator return FEB200_ERR_UNSUPPORTED. */
#define FEB200_MAX_NEIGHBOURS 8
typedef struct {
    int32_t n_neighbours;                       /* ranks this rank shares vertices with, ascending rank order */
    int32_t U;                                  /* DoFs per vertex */
    int32_t peer[FEB200_MAX_NEIGHBOURS];
    int32_t reserved[2];
    int64_t count[FEB200_MAX_NEIGHBOURS];       /* shared vertices per neighbour */
    const int64_t* idx[FEB200_MAX_NEIGHBOURS];  /* device: their local vertex ids, listed in the same order on both ranks */
    double* send[FEB200_MAX_NEIGHBOURS];        /* device scratch, count * U doubles each */
    double* recv[FEB200_MAX_NEIGHBOURS];
} feb200_halo;
int feb200_nccl_available(void);
int feb200_nccl_unique_id(uint8_t* id128);                                   /* host: 128-byte ncclUniqueId (rank 0 creates, all ranks share it) */
int feb200_nccl_comm_init(int32_t rank, int32_t world, const uint8_t* id128, void** comm_out);
int feb200_nccl_comm_destroy(void* comm);
/* send[k][i*U+j] = v[idx[k][i]*U+j]  /  v[idx[k][i]*U+j] += recv[k][i*U+j] in ascending neighbour order */
int feb200_halo_pack(feb200_stream s, const feb200_halo* h, const double* v);
int feb200_halo_unpack_add(feb200_stream s, const feb200_halo* h, double* v);
/* In place: every interface DoF becomes own + the neighbours' partial values (pack, one ncclSend/ncclRecv group, add).
 * Every interface vertex must be shared by exactly two ranks for the copies to stay bit-identical. */
int feb200_halo_sum(feb200_stream s, void* comm, const feb200_halo* h, double* v);
/* buf[0..n) <- sum over ranks (ncclAllReduce, in place): the CG dot products and residual norms */
int feb200_dot_allreduce(feb200_stream s, void* comm, double* buf, int32_t n);
/* The staged CG with the halo-sum after every local SpMV and the all-reduces of the dot products enqueued between the
 * stages (st->weights = ownership weights); same loop and stopping test as feb200_cg_init / feb200_cg_iterate. */
int feb200_cg_init_dist(feb200_stream s, void* comm, const feb200_cg_state* st, const feb200_halo* h, double tol, double atol);
int feb200_cg_iterate_dist(feb200_stream s, void* comm, const feb200_cg_state* st, const feb200_halo* h, int32_t n_iters);

/* Batched CG for many small systems on one mesh (vmap(cg) of the reference, solve_cg.py:20-137 under
 * tests/test_microscale_bvp.py): n_copies independent systems laid side by side, copy c owning the DoFs
 * [c*len, (c+1)*len) of every vector, the block-diagonal matrix applied with feb200_spmv.  Every copy keeps its own
 * alpha, beta, iteration count and stopping test.  scal: 4 doubles per copy {gamma, r.r, b.b, atol2}; cnt: 2 int32 per
 * copy {k, done}; hist: maxiter doubles per copy.  bcg_init: r = b - Ax0, z = M r, p = z and the first loop test;
 * bcg_step: one CG iteration after the caller's  Ap = A p  (alpha, x, r, z, the dot products, beta, p, loop test in one
 * launch; finished copies are skipped).  inv_diag == NULL: no preconditioner. */
int feb200_bcg_init(feb200_stream s, int64_t n_copies, int64_t len, const double* b, const double* Ax0, const double* inv_diag,
                    double* r, double* p, double* scal, int32_t* cnt, double* hist, double tol, double atol,
                    int32_t maxiter);
int feb200_bcg_step(feb200_stream s, int64_t n_copies, int64_t len, const double* Ap, const double* inv_diag, double* x,
                    double* r, double* p, double* scal, int32_t* cnt, double* hist, int32_t maxiter);

/* ------------------------------------------------------------------ geometric multigrid transfers (structured hex8 grids)
 * Stronger preconditioner than the reference's Jacobi (sparse_linear_solve.py:301-306; the reference reaches for AMGX,
 * :704-729).  Vertex id = (i*(ny+1) + j)*(nz+1) + k, U DoFs per vertex; the fine grid has (2 nxc, 2 nyc, 2 nzc) elements.
 * prolong_add: fine += P coarse (trilinear), skipping entries with mask_fine != 0;  restrict: coarse = P^T fine with
 * entries with mask_coarse != 0 set to zero.  Masks may be NULL. */
int feb200_mg_prolong_add(feb200_stream s, int32_t nxc, int32_t nyc, int32_t nzc, int32_t U, const double* coarse,
                          const double* mask_fine, double* fine);
int feb200_mg_restrict(feb200_stream s, int32_t nxc, int32_t nyc, int32_t nzc, int32_t U, const double* fine,
                       const double* mask_coarse, double* coarse);

/* fused vector helpers used by linear_solve / Newton (sparse_linear_solve.py:287-292, 527; fea.py:1103) */
int feb200_axpby(feb200_stream s, int64_t n, double a, const double* x, double b, const double* y, double* out);
int feb200_scatter_set(feb200_stream s, int64_t n_idx, const int64_t* idx, const double* vals, double scale_vals,
                       const double* minus_src, double* dst);

#ifdef __cplusplus
}
#endif
#endif /* FEB200_H */
