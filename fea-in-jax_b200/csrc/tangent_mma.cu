// Element tangent of the linear isotropic law for elements with many nodes (hex27: 81 x 81 matrices, 27 quadrature points)
// as FP64 tensor-core products -- the two-pass path's first pass for BASELINE configs[3] (reference:
// _calculate_jacobian_batch_element_kernel = vmap(jacfwd(residual)), fea.py:519-572, on the basix hex27 / 27-point tables of
// basis_quadrature.py:31-87).  The runtime-N scalar kernel (k_tangent_general_rt) needed 14 ms for 32 768 hex27 elements,
// about 15 x the FP64 floor of the 2 * 81 * 81 * 27 multiply-adds per element.
//
// One warp per element.  lane = quadrature point computes J, its inverse and the 3 N physical gradients into a
// shared-memory buffer laid out as DMMA fragments, component-major: G[d][k-step][node][4 points].  Then for every pair of
// node tiles (ta <= tb, 8 nodes each) the nine products  S^{dk}[a][b] = sum_q w_q g_q[a,d] g_q[b,k]  (mma.sync.m8n8k4.f64,
// ceil(Q/4) k-steps) leave a lane with the complete 3 x 3 blocks of its two node pairs, so
//   K_ab[d][k] = S_lam[d][k] + S_mu[k][d] + delta_dk tr(S_mu)       (weights w lam_q / w mu_q: per-point material)
//   K_ab[d][k] = lam S[d][k] + mu S[k][d] + delta_dk mu tr(S)       (material constant over the element: one product)
// is formed in registers and written as 72-byte blocks of K_e[(e*N + a)][b][3][3]; tiles below the diagonal are the
// transposes (K_ba = K_ab^T) and are written from the same registers.
#include "common.cuh"
#include "element_math.cuh"
#include "hex8_mma.cuh"

namespace feb {

constexpr int TM_WARPS = 4;

struct TmSmem {
    int TA, KS, NA, ks_stride, d_stride, dphi_stride;
    size_t off_dphi, off_warps, per_warp, off_w, off_x, total;
};
static inline TmSmem tm_layout(int N, int Q) {
    TmSmem L;
    L.TA = (N + 7) / 8;
    L.KS = (Q + 3) / 4;
    L.NA = 8 * L.TA;
    L.ks_stride = L.NA * 4 + 4;                   // +4: the geometry stores of the lanes of a half-warp hit distinct banks
    L.d_stride = L.KS * L.ks_stride;
    L.dphi_stride = (N * 3) | 1;                  // odd: lane = point reads its own table row without conflicts
    size_t o = 0;
    L.off_dphi = o; o += (size_t)Q * L.dphi_stride * 8;
    o = (o + 15) & ~(size_t)15;
    L.off_warps = o;
    size_t w = (size_t)3 * L.d_stride * 8;        // G
    L.off_w = w;  w += (size_t)2 * L.KS * 4 * 8;  // w lam, w mu per point (padded to the k-steps)
    L.off_x = w;  w += (size_t)N * 3 * 8;
    L.per_warp = (w + 15) & ~(size_t)15;
    L.total = L.off_warps + TM_WARPS * L.per_warp;
    return L;
}

template <bool CONST_MAT>
__global__ void __launch_bounds__(TM_WARPS * 32) k_tangent_iso_mma(feb200_batch b, TmSmem L, const double* __restrict__ x_vd,
                                                                  double* __restrict__ K_e) {
    extern __shared__ __align__(16) unsigned char tm_raw[];
    double* s_dphi = reinterpret_cast<double*>(tm_raw + L.off_dphi);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
    unsigned char* mine = tm_raw + L.off_warps + (size_t)warp * L.per_warp;
    double* G = reinterpret_cast<double*>(mine);
    double* s_wl = reinterpret_cast<double*>(mine + L.off_w);
    double* s_wm = s_wl + L.KS * 4;
    double* xs = reinterpret_cast<double*>(mine + L.off_x);
    const int N = b.N, Q = b.Q, NU = N * 3;
    for (int i = threadIdx.x; i < Q * NU; i += blockDim.x) s_dphi[(i / NU) * L.dphi_stride + i % NU] = b.dphi[i];
    for (int i = lane; i < 3 * L.d_stride; i += 32) G[i] = 0.0;            // rows of the padding nodes stay zero
    __syncthreads();
    const int64_t n_warps = (int64_t)gridDim.x * TM_WARPS;
    for (int64_t e = (int64_t)blockIdx.x * TM_WARPS + warp; e < b.E; e += n_warps) {
        __syncwarp();
        for (int i = lane; i < NU; i += 32) xs[i] = x_vd[(int64_t)b.conn[e * N + i / 3] * 3 + i % 3];
        __syncwarp();
        // ---- geometry: lane = quadrature point (lanes past Q fill the k-step padding with zeros)
        if (lane < 4 * L.KS) {
            const int q = lane;
            double* dst = G + (q >> 2) * L.ks_stride + (q & 3);
            if (q < Q) {
                const double* dq = s_dphi + q * L.dphi_stride;
                double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, Ji[3][3];
                for (int n = 0; n < N; ++n)
#pragma unroll
                    for (int pp = 0; pp < 3; ++pp)
#pragma unroll
                        for (int dd = 0; dd < 3; ++dd) J[pp][dd] += xs[n * 3 + dd] * dq[n * 3 + pp];
                const double dt = inverse(J, Ji);
                for (int n = 0; n < N; ++n)
#pragma unroll
                    for (int dd = 0; dd < 3; ++dd) {
                        double sacc = 0.0;
#pragma unroll
                        for (int pp = 0; pp < 3; ++pp) sacc += Ji[dd][pp] * dq[n * 3 + pp];
                        dst[dd * L.d_stride + n * 4] = sacc;
                    }
                int qs;
                const double* mp = mp_of(b, e, qs);
                double lam, mu;
                isotropic_lame<3>(mp + (int64_t)q * qs, lam, mu);
                const double w = dt * b.W[q];
                s_wl[q] = CONST_MAT ? w : w * lam;
                s_wm[q] = CONST_MAT ? w : w * mu;
            } else {
                for (int n = 0; n < N; ++n)
#pragma unroll
                    for (int dd = 0; dd < 3; ++dd) dst[dd * L.d_stride + n * 4] = 0.0;
                s_wl[q] = 0.0;
                s_wm[q] = 0.0;
            }
        }
        __syncwarp();
        double c_lam = 0.0, c_mu = 0.0;
        if (CONST_MAT) {
            int qs;
            isotropic_lame<3>(mp_of(b, e, qs), c_lam, c_mu);
        }
        double* Ke = K_e + e * (int64_t)N * N * 9;
        // ---- block products, node tiles ta <= tb
        for (int ta = 0; ta < L.TA; ++ta)
            for (int tb = ta; tb < L.TA; ++tb) {
                double Cm[3][3][2], Cl[3][3][2];
#pragma unroll
                for (int d = 0; d < 3; ++d)
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        Cm[d][k][0] = Cm[d][k][1] = 0.0;
                        Cl[d][k][0] = Cl[d][k][1] = 0.0;
                    }
                const double* Ga = G + (8 * ta + gid) * 4 + tig;
                const double* Gb = G + (8 * tb + gid) * 4 + tig;
                for (int ks = 0; ks < L.KS; ++ks) {
                    const double wm = s_wm[4 * ks + tig];
                    const double wl = CONST_MAT ? 0.0 : s_wl[4 * ks + tig];
                    double Am[3], Al[3], Bk[3];
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const double ga = Ga[d * L.d_stride + ks * L.ks_stride];
                        Am[d] = wm * ga;
                        Al[d] = wl * ga;
                        Bk[d] = Gb[d * L.d_stride + ks * L.ks_stride];
                    }
#pragma unroll
                    for (int d = 0; d < 3; ++d)
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            dmma884(Cm[d][k][0], Cm[d][k][1], Am[d], Bk[k]);
                            if (!CONST_MAT) dmma884(Cl[d][k][0], Cl[d][k][1], Al[d], Bk[k]);
                        }
                }
                const int a = 8 * ta + gid;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int bb = 8 * tb + 2 * tig + j;
                    if (a >= N || bb >= N) continue;
                    double K[9];
                    const double trm = Cm[0][0][j] + Cm[1][1][j] + Cm[2][2][j];
#pragma unroll
                    for (int d = 0; d < 3; ++d)
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            if (CONST_MAT) K[d * 3 + k] = c_lam * Cm[d][k][j] + c_mu * Cm[k][d][j] + (d == k ? c_mu * trm : 0.0);
                            else K[d * 3 + k] = Cl[d][k][j] + Cm[k][d][j] + (d == k ? trm : 0.0);
                        }
                    double* dst = Ke + ((int64_t)a * N + bb) * 9;
#pragma unroll
                    for (int c = 0; c < 9; ++c) dst[c] = K[c];
                    if (tb > ta) {                                        // the tile below the diagonal: K_ba = K_ab^T
                        double* dt2 = Ke + ((int64_t)bb * N + a) * 9;
#pragma unroll
                        for (int d = 0; d < 3; ++d)
#pragma unroll
                            for (int k = 0; k < 3; ++k) dt2[k * 3 + d] = K[d * 3 + k];
                    }
                }
            }
    }
}

// K_e of the linear isotropic law through the tensor-core kernel (3-D, 8 < N <= 32 nodes, Q <= 32 points); returns
// FEB200_ERR_UNSUPPORTED for anything else (the caller then uses the scalar kernels).
int tangent_iso_mma(cudaStream_t st, const feb200_batch& b, const double* x_vd, double* K_e) {
    if (!(b.D == 3 && b.N > 8 && b.N <= 32 && b.Q <= 32 && b.kind == KIND_LINEAR && b.material == MAT_ISOTROPIC))
        return FEB200_ERR_UNSUPPORTED;
    const TmSmem L = tm_layout(b.N, b.Q);
    if (L.total > 200 * 1024) return FEB200_ERR_UNSUPPORTED;
    const bool cm = (b.layout != LAYOUT_EQM);
    static bool attr_set[2] = {false, false};
    int64_t g = ceil_div(b.E, TM_WARPS);
    const int64_t cap = (int64_t)sm_count() * (L.total <= 110 * 1024 ? 2 : 1);
    if (g > cap) g = cap;
    if (cm) {
        if (!attr_set[1]) {
            FEB_CUDA(cudaFuncSetAttribute(k_tangent_iso_mma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set[1] = true;
        }
        k_tangent_iso_mma<true><<<(int)g, TM_WARPS * 32, L.total, st>>>(b, L, x_vd, K_e);
    } else {
        if (!attr_set[0]) {
            FEB_CUDA(cudaFuncSetAttribute(k_tangent_iso_mma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set[0] = true;
        }
        k_tangent_iso_mma<false><<<(int)g, TM_WARPS * 32, L.total, st>>>(b, L, x_vd, K_e);
    }
    return FEB200_OK;
}

}  // namespace feb
