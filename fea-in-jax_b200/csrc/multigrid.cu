// Grid-transfer kernels of the geometric multigrid preconditioner (fe_jax/multigrid.py) on structured hex8 vertex grids:
// vertex id = (i * (ny + 1) + j) * (nz + 1) + k, U DoFs per vertex (the numbering of uniform_hex_grid, the 3-D extension
// of setup.py:41-53).  The reference has no multilevel solver of its own -- it hands the system to AMGX when asked for one
// (sparse_linear_solve.py:704-729); this is the B200-side answer to SURVEY 8(f) rank 4 ("a stronger preconditioner than
// Jacobi", sparse_linear_solve.py:301-306).  Both kernels gather (thread per output entry): no atomics, fixed order.
#include "common.cuh"

namespace feb {

// fine += P coarse : trilinear interpolation from the (nxc, nyc, nzc)-element grid to the (2 nxc, 2 nyc, 2 nzc) one;
// entries with mask != 0 (constrained DoFs) receive nothing.
__global__ void k_mg_prolong_add(int nxc, int nyc, int nzc, int U, const double* __restrict__ coarse,
                                 const double* __restrict__ mask_f, double* __restrict__ fine) {
    const int nyf = 2 * nyc + 1, nzf = 2 * nzc + 1;
    const int64_t nf = (int64_t)(2 * nxc + 1) * nyf * nzf * U;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nf; t += stride) {
        if (mask_f && mask_f[t] != 0.0) continue;
        const int c = (int)(t % U);
        const int64_t v = t / U;
        const int k = (int)(v % nzf), j = (int)((v / nzf) % nyf), i = (int)(v / ((int64_t)nzf * nyf));
        const int i0 = i >> 1, j0 = j >> 1, k0 = k >> 1;
        const int ni = (i & 1) ? 2 : 1, nj = (j & 1) ? 2 : 1, nk = (k & 1) ? 2 : 1;
        double acc = 0.0;
        for (int a = 0; a < ni; ++a)
            for (int b = 0; b < nj; ++b)
                for (int d = 0; d < nk; ++d)
                    acc += coarse[(((int64_t)(i0 + a) * (nyc + 1) + (j0 + b)) * (nzc + 1) + (k0 + d)) * U + c];
        fine[t] += acc / (double)(ni * nj * nk);
    }
}

// coarse = P^T fine (full weighting, the transpose of the interpolation); entries with mask != 0 are set to zero.
__global__ void k_mg_restrict(int nxc, int nyc, int nzc, int U, const double* __restrict__ fine,
                              const double* __restrict__ mask_c, double* __restrict__ coarse) {
    const int nxf = 2 * nxc + 1, nyf = 2 * nyc + 1, nzf = 2 * nzc + 1;
    const int64_t nc = (int64_t)(nxc + 1) * (nyc + 1) * (nzc + 1) * U;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nc; t += stride) {
        if (mask_c && mask_c[t] != 0.0) { coarse[t] = 0.0; continue; }
        const int c = (int)(t % U);
        const int64_t v = t / U;
        const int K = (int)(v % (nzc + 1)), J = (int)((v / (nzc + 1)) % (nyc + 1)), I = (int)(v / ((int64_t)(nzc + 1) * (nyc + 1)));
        double acc = 0.0;
        for (int a = -1; a <= 1; ++a) {
            const int i = 2 * I + a;
            if (i < 0 || i >= nxf) continue;
            for (int b = -1; b <= 1; ++b) {
                const int j = 2 * J + b;
                if (j < 0 || j >= nyf) continue;
                for (int d = -1; d <= 1; ++d) {
                    const int k = 2 * K + d;
                    if (k < 0 || k >= nzf) continue;
                    const double w = (a ? 0.5 : 1.0) * (b ? 0.5 : 1.0) * (d ? 0.5 : 1.0);
                    acc += w * fine[(((int64_t)i * nyf + j) * nzf + k) * U + c];
                }
            }
        }
        coarse[t] = acc;
    }
}

}  // namespace feb

using namespace feb;

extern "C" {

int feb200_mg_prolong_add(feb200_stream s, int32_t nxc, int32_t nyc, int32_t nzc, int32_t U, const double* coarse,
                          const double* mask_fine, double* fine) {
    FEB_REQUIRE(coarse && fine && nxc >= 1 && nyc >= 1 && nzc >= 1 && U >= 1, FEB200_ERR_INVALID, "mg_prolong_add: bad argument");
    const int64_t nf = (int64_t)(2 * nxc + 1) * (2 * nyc + 1) * (2 * nzc + 1) * U;
    k_mg_prolong_add<<<grid_for(nf, 256), 256, 0, as_stream(s)>>>(nxc, nyc, nzc, U, coarse, mask_fine, fine);
    FEB_CHECK_LAUNCH("mg_prolong_add");
    return FEB200_OK;
}

int feb200_mg_restrict(feb200_stream s, int32_t nxc, int32_t nyc, int32_t nzc, int32_t U, const double* fine,
                       const double* mask_coarse, double* coarse) {
    FEB_REQUIRE(coarse && fine && nxc >= 1 && nyc >= 1 && nzc >= 1 && U >= 1, FEB200_ERR_INVALID, "mg_restrict: bad argument");
    const int64_t nc = (int64_t)(nxc + 1) * (nyc + 1) * (nzc + 1) * U;
    k_mg_restrict<<<grid_for(nc, 256), 256, 0, as_stream(s)>>>(nxc, nyc, nzc, U, fine, mask_coarse, coarse);
    FEB_CHECK_LAUNCH("mg_restrict");
    return FEB200_OK;
}

}  // extern "C"
