// XLA-FFI handlers over the feb200 C ABI (include/feb200.h).
//
// NOT compiled in this image: jax / jaxlib (and therefore xla/ffi/api/ffi.h, found through
// jax.ffi.include_dir()) are not installable here.  Build where JAX exists with
//   g++ -std=c++17 -shared -fPIC -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") -I/usr/local/cuda/include \
//       -Iinclude fea-in-jax_b200/csrc/xla_ffi_shim.cc -L fea-in-jax_b200/lib -lfeb200 -L/usr/local/cuda/lib64 -lcudart \
//       -o libfeb200_xla.so
// Each handler only forwards XLA's stream and buffers: XLA owns all operand / result buffers, the
// handler enqueues on the given stream and returns (no sync, no allocation => usable inside
// lax.while_loop and CUDA-graph capture).  See INTEGRATION.md for the Python registration.
#include <cuda_runtime.h>

#include "xla/ffi/api/ffi.h"

#include "feb200.h"

namespace ffi = xla::ffi;

static ffi::Error status(int rc) {
    return rc == 0 ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, feb200_last_error());
}
// optional operands (Dirichlet mask / values): a zero-size buffer stands for "not given"
template <typename B> static auto opt_data(const B& b) -> decltype(b.typed_data()) {
    return b.element_count() == 0 ? nullptr : b.typed_data();
}

// y = A x on the node-pattern CSR (replaces the JVP operator of sparse_linear_solve.py:251-258)
static ffi::Error SpmvImpl(cudaStream_t stream, ffi::Buffer<ffi::S64> nindptr, ffi::Buffer<ffi::S32> nindices,
                           ffi::Buffer<ffi::F64> values, ffi::Buffer<ffi::F64> x, ffi::ResultBuffer<ffi::F64> y,
                           int64_t U) {
    const int64_t V = nindptr.element_count() - 1;
    return status(feb200_spmv(stream, V, (int32_t)U, nindptr.typed_data(), nindices.typed_data(), values.typed_data(),
                              x.typed_data(), y->typed_data(), nullptr, nullptr, 0));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(feb200_xla_spmv, SpmvImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::S64>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Attr<int64_t>("U"));

static feb200_batch make_batch(const ffi::Buffer<ffi::S32>& conn, const ffi::Buffer<ffi::F64>& dphi,
                               const ffi::Buffer<ffi::F64>& W, const ffi::Buffer<ffi::F64>& mp, int64_t kind, int64_t material,
                               int64_t layout) {
    feb200_batch b{};
    b.E = conn.dimensions()[0];
    b.N = (int32_t)conn.dimensions()[1];
    b.Q = (int32_t)dphi.dimensions()[0];
    b.D = (int32_t)dphi.dimensions()[2];
    b.kind = (int32_t)kind;
    b.material = (int32_t)material;
    b.layout = (int32_t)layout;
    b.M = (int32_t)mp.dimensions().back();
    b.conn = conn.typed_data();
    b.dphi = dphi.typed_data();
    b.W = W.typed_data();
    b.mp = mp.typed_data();
    return b;
}

// element residual of one batch: R_e (E, N, D)   (fea.py:800-847)
static ffi::Error ResidualImpl(cudaStream_t stream, ffi::Buffer<ffi::S32> conn, ffi::Buffer<ffi::F64> dphi,
                               ffi::Buffer<ffi::F64> W, ffi::Buffer<ffi::F64> mp, ffi::Buffer<ffi::F64> x_vd,
                               ffi::Buffer<ffi::F64> u, ffi::Buffer<ffi::F64> mask, ffi::Buffer<ffi::F64> values,
                               ffi::ResultBuffer<ffi::F64> R_e, int64_t kind, int64_t material, int64_t layout) {
    const feb200_batch b = make_batch(conn, dphi, W, mp, kind, material, layout);
    return status(feb200_element_residual(stream, &b, x_vd.typed_data(), u.typed_data(), opt_data(mask),
                                          opt_data(values), R_e->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(feb200_xla_element_residual, ResidualImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Attr<int64_t>("kind")
                                  .Attr<int64_t>("material")
                                  .Attr<int64_t>("layout"));
// R (U V) = segmented sum of R_e over the assembly map, Dirichlet wrapper fused   (setup.py:344-356, fea.py:961-963)
static ffi::Error ScatterSumImpl(cudaStream_t stream, ffi::Buffer<ffi::U64> row_ind, ffi::Buffer<ffi::U64> col_ind,
                                 ffi::Buffer<ffi::F64> R_e, ffi::Buffer<ffi::F64> mask, ffi::Buffer<ffi::F64> values,
                                 ffi::Buffer<ffi::F64> u, ffi::ResultBuffer<ffi::F64> R, int64_t U) {
    const int64_t V = row_ind.element_count() - 1;
    return status(feb200_scatter_sum(stream, row_ind.typed_data(), col_ind.typed_data(), V, (int32_t)U, R_e.typed_data(),
                                     R->typed_data(), 0, opt_data(mask), opt_data(values), opt_data(u)));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(feb200_xla_scatter_sum, ScatterSumImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::U64>>()
                                  .Arg<ffi::Buffer<ffi::U64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Attr<int64_t>("U"));

// Jacobian CSR values of a linear isotropic hex8 batch, Dirichlet elimination fused   (fea.py:519-660, sparse_matrix.py:157-184).
// The tile plan is an ordinary pytree of device arrays built once per mesh (fe_jax/fused_plan.py) and passed as operands.
static ffi::Error AssembleFusedImpl(cudaStream_t stream, ffi::Buffer<ffi::S32> conn, ffi::Buffer<ffi::F64> dphi,
                                    ffi::Buffer<ffi::F64> W, ffi::Buffer<ffi::F64> mp, ffi::Buffer<ffi::F64> x_vd,
                                    ffi::Buffer<ffi::S32> tile_node_ptr, ffi::Buffer<ffi::S32> tile_nodes,
                                    ffi::Buffer<ffi::S32> tile_elem_ptr, ffi::Buffer<ffi::S32> tile_elems,
                                    ffi::Buffer<ffi::S32> tile_item_ptr, ffi::Buffer<ffi::U32> tile_items,
                                    ffi::Buffer<ffi::S64> tile_node_p0, ffi::Buffer<ffi::S32> tile_node_meta,
                                    ffi::Buffer<ffi::U16> slot_rank, ffi::Buffer<ffi::S32> nindices,
                                    ffi::Buffer<ffi::U8> vertex_flags, ffi::ResultBuffer<ffi::F64> values, int64_t layout,
                                    int64_t tn_cap, int64_t te_cap, int64_t material_hints) {
    const feb200_batch b = make_batch(conn, dphi, W, mp, FEB200_KIND_LINEAR, FEB200_MAT_ISOTROPIC, layout);
    feb200_fused_plan plan{};
    plan.n_tiles = (int32_t)(tile_node_ptr.element_count() - 1);
    plan.tn_cap = (int32_t)tn_cap;
    plan.te_cap = (int32_t)te_cap;
    plan.tile_node_ptr = tile_node_ptr.typed_data();
    plan.tile_nodes = tile_nodes.typed_data();
    plan.tile_elem_ptr = tile_elem_ptr.typed_data();
    plan.tile_elems = tile_elems.typed_data();
    plan.tile_item_ptr = tile_item_ptr.typed_data();
    plan.tile_items = tile_items.typed_data();
    plan.tile_node_p0 = tile_node_p0.typed_data();
    plan.tile_node_meta = tile_node_meta.typed_data();
    return status(feb200_assemble_fused_hex8_mma(stream, &b, &plan, x_vd.typed_data(), slot_rank.typed_data(),
                                                 nindices.typed_data(), vertex_flags.typed_data(), (int32_t)material_hints,
                                                 values->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(feb200_xla_assemble_fused_hex8, AssembleFusedImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Arg<ffi::Buffer<ffi::S64>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::U16>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Attr<int64_t>("layout")
                                  .Attr<int64_t>("tn_cap")
                                  .Attr<int64_t>("te_cap")
                                  .Attr<int64_t>("material_hints"));

// generic path, pass 1: K_e (E, N, N, U, U)   (vmap(jacfwd(residual_kernel)), fea.py:519-572)
static ffi::Error TangentImpl(cudaStream_t stream, ffi::Buffer<ffi::S32> conn, ffi::Buffer<ffi::F64> dphi,
                              ffi::Buffer<ffi::F64> W, ffi::Buffer<ffi::F64> mp, ffi::Buffer<ffi::F64> x_vd,
                              ffi::Buffer<ffi::F64> u, ffi::Buffer<ffi::F64> mask, ffi::Buffer<ffi::F64> values,
                              ffi::ResultBuffer<ffi::F64> K_e, int64_t kind, int64_t material, int64_t layout) {
    const feb200_batch b = make_batch(conn, dphi, W, mp, kind, material, layout);
    return status(feb200_element_tangent(stream, &b, x_vd.typed_data(), u.typed_data(), opt_data(mask),
                                         opt_data(values), K_e->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(feb200_xla_element_tangent, TangentImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Attr<int64_t>("kind")
                                  .Attr<int64_t>("material")
                                  .Attr<int64_t>("layout"));

// generic path, pass 2: CSR values from K_e through the slot-rank map   (fea.py:613-660, sparse_matrix.py:44-110)
static ffi::Error AssembleCsrImpl(cudaStream_t stream, ffi::Buffer<ffi::S32> conn, ffi::Buffer<ffi::U64> row_ind,
                                  ffi::Buffer<ffi::U64> col_ind, ffi::Buffer<ffi::F64> K_e, ffi::Buffer<ffi::U16> slot_rank,
                                  ffi::Buffer<ffi::S64> nindptr, ffi::Buffer<ffi::S32> nindices, ffi::Buffer<ffi::F64> mask,
                                  ffi::ResultBuffer<ffi::F64> values, int64_t U, int64_t max_deg) {
    feb200_batch_map m{};
    m.E = conn.dimensions()[0];
    m.N = (int32_t)conn.dimensions()[1];
    m.conn = conn.typed_data();
    m.row_ind = row_ind.typed_data();
    m.col_ind = col_ind.typed_data();
    const int64_t V = nindptr.element_count() - 1;
    return status(feb200_assemble_csr(stream, &m, V, (int32_t)U, K_e.typed_data(), slot_rank.typed_data(),
                                      nindptr.typed_data(), (int32_t)max_deg, values->typed_data(), 0,
                                      mask.element_count() ? nindices.typed_data() : nullptr, opt_data(mask)));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(feb200_xla_assemble_csr, AssembleCsrImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::U64>>()
                                  .Arg<ffi::Buffer<ffi::U64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::U16>>()
                                  .Arg<ffi::Buffer<ffi::S64>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Attr<int64_t>("U")
                                  .Attr<int64_t>("max_deg"));

// Whole CG solve on the node-pattern CSR, device resident   (solve_cg.py:20-137): the handler enqueues cg_init and
// maxiter iterations whose kernels exit at once after convergence; scalars / counters / hist are results.
static ffi::Error CgImpl(cudaStream_t stream, ffi::Buffer<ffi::S64> nindptr, ffi::Buffer<ffi::S32> nindices,
                         ffi::Buffer<ffi::F64> values, ffi::Buffer<ffi::F64> inv_diag, ffi::Buffer<ffi::F64> rhs,
                         ffi::Buffer<ffi::F64> x0, ffi::ResultBuffer<ffi::F64> x, ffi::ResultBuffer<ffi::F64> r,
                         ffi::ResultBuffer<ffi::F64> p, ffi::ResultBuffer<ffi::F64> Ap, ffi::ResultBuffer<ffi::F64> hist,
                         ffi::ResultBuffer<ffi::F64> scalars, ffi::ResultBuffer<ffi::S32> counters,
                         ffi::ResultBuffer<ffi::U8> ws, int64_t U, int64_t maxiter, double tol, double atol,
                         bool preconditioned) {
    feb200_cg_state st{};
    st.V = nindptr.element_count() - 1;
    st.U = (int32_t)U;
    st.maxiter = (int32_t)maxiter;
    st.nindptr = nindptr.typed_data();
    st.nindices = nindices.typed_data();
    st.values = values.typed_data();
    st.inv_diag = preconditioned ? inv_diag.typed_data() : nullptr;
    st.weights = nullptr;
    st.b = rhs.typed_data();
    st.x = x->typed_data();
    st.r = r->typed_data();
    st.p = p->typed_data();
    st.Ap = Ap->typed_data();
    st.hist = hist->typed_data();
    st.scalars = scalars->typed_data();
    st.counters = counters->typed_data();
    st.ws = ws->typed_data();
    st.ws_bytes = ws->element_count();
    if (cudaMemcpyAsync(st.x, x0.typed_data(), sizeof(double) * x0.element_count(), cudaMemcpyDeviceToDevice, stream) !=
        cudaSuccess)
        return ffi::Error(ffi::ErrorCode::kInternal, "feb200_xla_cg: copy of x0 failed");
    // `maxiter` is the number of iterations ENQUEUED (kernels exit at once when the device-side `done` flag is set, 3-4
    // launches each): the caller passes an explicit bound -- sparse_linear_solve.py:106 caps CG at 1000 -- and, for longer
    // solves, calls the handler again from a lax.while_loop on the returned `counters` (k, done) instead of enqueueing
    // the 10*n iterations of the CG_JAX_SCIPY default in one go.
    if (maxiter < 0 || maxiter > 4096)
        return ffi::Error(ffi::ErrorCode::kInvalidArgument, "feb200_xla_cg: maxiter must be in [0, 4096] per call");
    int rc = feb200_cg_init(stream, &st, tol, atol);
    if (rc == 0) rc = feb200_cg_iterate(stream, &st, (int32_t)maxiter);
    return status(rc);
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(feb200_xla_cg, CgImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::S64>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int64_t>("U")
                                  .Attr<int64_t>("maxiter")
                                  .Attr<double>("tol")
                                  .Attr<double>("atol")
                                  .Attr<bool>("preconditioned"));
// The map and pattern builders (build_row_ind / build_col_ind, pattern_count / scan / fill, slot_rank) run once per mesh
// and follow the same three lines: unpack buffers, call, status().
