// TEST-ONLY host build of the __host__ __device__ element math (csrc/element_math.cuh), so the
// formulas the CUDA kernels use can be compared with the oracle on the CPU-only build box.
// Never linked into libfeb200 and never used by the product path.
#include "../../fea-in-jax_b200/csrc/element_math.cuh"

#include <vector>

extern "C" {

int hc_residual(int D, int kind, int material, int N, int Q, const double* x, const double* u, const double* dphi,
                const double* W, const double* mp, int mp_q_stride, double* R) {
    std::vector<double> g(N * D);
    if (D == 2) feb::element_residual_rt<2>(kind, material, N, Q, x, u, dphi, W, mp, mp_q_stride, g.data(), R);
    else if (D == 3) feb::element_residual_rt<3>(kind, material, N, Q, x, u, dphi, W, mp, mp_q_stride, g.data(), R);
    else return -1;
    return 0;
}

int hc_tangent(int D, int kind, int material, int N, int Q, const double* x, const double* u, const double* dphi,
               const double* W, const double* mp, int mp_q_stride, double* K) {
    std::vector<double> g(N * D);
    if (D == 2) feb::element_tangent_rt<2>(kind, material, N, Q, x, u, dphi, W, mp, mp_q_stride, g.data(), K);
    else if (D == 3) feb::element_tangent_rt<3>(kind, material, N, Q, x, u, dphi, W, mp, mp_q_stride, g.data(), K);
    else return -1;
    return 0;
}

// compile-time-N geometry used by the fast kernels
int hc_shape_gradients_hex8(const double* x, const double* dphi_q, double* g, double* det) {
    *det = feb::shape_gradients<8, 3>(x, dphi_q, g);
    return 0;
}
}
