// XLA-FFI handlers over the feb200 C ABI (include/feb200.h).
//
// NOT compiled in this image: jax / jaxlib (and therefore xla/ffi/api/ffi.h, found through
// jax.ffi.include_dir()) are not installable here.  Build where JAX exists with
//   g++ -std=c++17 -shared -fPIC -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \
//       -Iinclude fea-in-jax_b200/csrc/xla_ffi_shim.cc -L fea-in-jax_b200/lib -lfeb200 -o libfeb200_xla.so
// Each handler only forwards XLA's stream and buffers: XLA owns all operand / result buffers, the
// handler enqueues on the given stream and returns (no sync, no allocation => usable inside
// lax.while_loop and CUDA-graph capture).  See INTEGRATION.md for the Python registration.
#include "xla/ffi/api/ffi.h"

#include "feb200.h"

namespace ffi = xla::ffi;

static ffi::Error status(int rc) {
    return rc == 0 ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, feb200_last_error());
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

// element residual of one batch: R_e (E, N, D)   (fea.py:800-847)
static ffi::Error ResidualImpl(cudaStream_t stream, ffi::Buffer<ffi::S32> conn, ffi::Buffer<ffi::F64> dphi,
                               ffi::Buffer<ffi::F64> W, ffi::Buffer<ffi::F64> mp, ffi::Buffer<ffi::F64> x_vd,
                               ffi::Buffer<ffi::F64> u, ffi::Buffer<ffi::F64> mask, ffi::Buffer<ffi::F64> values,
                               ffi::ResultBuffer<ffi::F64> R_e, int64_t kind, int64_t material, int64_t layout) {
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
    return status(feb200_element_residual(stream, &b, x_vd.typed_data(), u.typed_data(), mask.typed_data(),
                                          values.typed_data(), R_e->typed_data()));
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
// The remaining entry points (scatter_sum, element_tangent, assemble_csr, dirichlet_lhs, cg_init /
// cg_iterate, the map and pattern builders) follow the same three lines: unpack buffers, call, status().
