// Element tangent of a state-dependent law (hyperelasticity: st_venant_kirchhoff, mooney_rivlin; also the orthotropic
// linear law) on hex8 with 8 quadrature points as FP64 tensor-core products -- the first pass of the two-pass assembly of
// BASELINE configs[2] (reference: _calculate_jacobian_batch_element_kernel = vmap(jacfwd(residual_kernel)), fea.py:519-572,
// with hyperelasticity.py:93-139 inside).
//
//   K[(a,d),(b,k)] = sum_q w_q sum_{i,l} g_q[a,i] A_q[(i,d),(k,l)] g_q[b,l],      A_q = d S[i][d] / d H[k][l] at point q
//
// (19 008 multiply-adds per element in the factorised form below, plus A_q).  The scalar kernel k_tangent_general_ct ran
// this at a third of the FP64 peak: one thread per block row, 255 registers, the 81 entries of A_q re-read by every row.
//
// Here every warp is an independent worker (no block barrier after start-up) that takes groups of 4 elements:
//   phase 1  lane = (element, quadrature point): coordinates / solution from the vertex arrays (Dirichlet blend fused),
//            J, J^-1, the 24 physical gradients, H, then w A_q (closed form for St. Venant-Kirchhoff, forward-mode duals
//            on the constitutive law otherwise) into shared memory;
//   phase 2  per element, as nine 8 x 8 products over the 24 (point, i) pairs on mma.sync.m8n8k4.f64:
//              K^{dk}[a][b] = sum_{(q,i)} g_q[a,i] * Y^{dk}[(q,i)][b],     Y^{dk}[(q,i)][b] = sum_l w A_q[(i,d),(k,l)] g_q[b,l]
//            the lane that owns B-fragment entry ((q,i), b) forms Y in registers (3 multiply-adds per entry, A_q rows
//            read as shared-memory broadcasts) and feeds it straight to the tensor core; the A fragment is one of the three
//            gradient values it already holds.  A lane ends up with the two complete 3 x 3 blocks K[a][2t], K[a][2t+1] =
//            144 contiguous bytes of K_e[(e*8 + a)][b][3][3];
//   store    the element's 4608 bytes are staged in the shared memory its A_q occupied and leave as ONE bulk asynchronous
//            copy (cp.async.bulk shared -> global), overlapping the next element's products.
#include "common.cuh"
#include "element_math.cuh"
#include "hex8_mma.cuh"

#ifndef FEB_TG_WARPS
#define FEB_TG_WARPS 8
#endif

namespace feb {

constexpr int TG_GQ = 27;                 // doubles per (element, point) row of the gradient buffer: [node][3] + 3 (stores and
                                          // fragment loads of a half-warp then fall into distinct banks)
constexpr int TG_GE = 8 * TG_GQ;          // per element
constexpr int TG_AQ = 81;                 // w A_q: [(i,d)][(k,l)]
constexpr int TG_AE = 8 * TG_AQ;          // per element: 5184 bytes, reused as the 4608-byte staging of K_e
constexpr int TG_DPHI = 26;               // doubles per point of the basis table (24 used; 16-byte rows, conflict-free)
constexpr size_t TG_TABLE = (size_t)(8 * TG_DPHI + 8) * 8;
constexpr size_t TG_PER_WARP = (size_t)(4 * TG_GE + 4 * TG_AE) * 8;

__device__ __forceinline__ double tg_load_u(const double* __restrict__ u, const double* __restrict__ mask,
                                            const double* __restrict__ values, int64_t f) {
    const double v = u[f];
    if (!mask) return v;
    const double m = mask[f];
    return (1.0 - m) * v + m * values[f];
}

// w A_q of any law through forward-mode duals on the constitutive function (element_math.cuh); kept out of line so that
// its local arrays do not weigh on the register allocation of the kernel
__device__ __noinline__ void tg_generic_tangent(int material, const double* mp, const double (&H)[3][3], double w, double* ad) {
    double A[81];
    point_tangent<3>(material, mp, H, A);
    for (int i = 0; i < 81; ++i) ad[i] = w * A[i];
}

__global__ void __launch_bounds__(FEB_TG_WARPS * 32, 1)
k_tangent_general_mma(feb200_batch b, const double* __restrict__ x_vd, const double* __restrict__ u,
                      const double* __restrict__ mask, const double* __restrict__ values, double* __restrict__ K_e) {
    extern __shared__ __align__(16) unsigned char tg_raw[];
    double* s_dphi = reinterpret_cast<double*>(tg_raw);                     // [q] DPHI : [n][3]
    double* s_w = s_dphi + 8 * TG_DPHI;                                     // [q]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    double* Gw = reinterpret_cast<double*>(tg_raw + TG_TABLE + (size_t)warp * TG_PER_WARP);   // [4] GE : [q] GQ : [n][3]
    double* Aw = Gw + 4 * TG_GE;                                            // [4] AE : [q] AQ
    for (int i = tid; i < 8 * 24; i += blockDim.x) s_dphi[(i / 24) * TG_DPHI + (i % 24)] = b.dphi[i];
    if (tid < 8) s_w[tid] = b.W[tid];
    __syncthreads();                                                        // the only block-wide barrier

    const int jj = lane >> 3, qq = lane & 7;                                // phase 1: element of the group, quadrature point
    // phase 2, k-step s: this lane's contraction index m = 4 s + tig = 3 q + i
    int offG[6], offA[6], isel[6];
#pragma unroll
    for (int s = 0; s < 6; ++s) {
        const int m = 4 * s + tig, q = m / 3, i = m - 3 * q;
        offG[s] = q * TG_GQ + gid * 3;
        offA[s] = q * TG_AQ + i * 27;
        isel[s] = i;
    }
    const bool hyper = (b.kind == KIND_HYPER);
    const int64_t n_groups = (b.E + 3) >> 2;
    const int64_t n_workers = (int64_t)gridDim.x * (blockDim.x >> 5);

    for (int64_t grp = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; grp < n_groups; grp += n_workers) {
        const int64_t e = 4 * grp + jj;
        const bool act = e < b.E;
        // the bulk stores of the previous group have finished reading their staging areas (the A regions)
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        // ---- coordinates and solution of the group -> the (still unused) gradient area: [jj] 48 : x[8][3], u[8][3]
        if (act) {
            const int64_t v = b.conn[e * 8 + qq];
            double* xd = Gw + jj * 48 + qq * 3;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const double xv = x_vd[v * 3 + d];
                const double uv = tg_load_u(u, mask, values, v * 3 + d);
                xd[d] = xv;
                xd[24 + d] = hyper ? uv + xv : uv;
            }
        }
        __syncwarp();
        // ---- phase 1: geometry and H at (element jj, point qq)
        double g[24], H[3][3], w = 0.0;
        if (act) {
            const double dt = hex8_gradients(Gw + jj * 48, s_dphi + qq * TG_DPHI, g);
            w = dt * s_w[qq];
            const double* ue = Gw + jj * 48 + 24;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int d = 0; d < 3; ++d) H[i][d] = 0.0;
#pragma unroll
            for (int n = 0; n < 8; ++n)
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const double ui = ue[n * 3 + i];
#pragma unroll
                    for (int d = 0; d < 3; ++d) H[i][d] += g[n * 3 + d] * ui;
                }
        }
        __syncwarp();                                                       // every lane has read x / u: the area becomes G
        if (act) {
            double* gd = Gw + jj * TG_GE + qq * TG_GQ;
#pragma unroll
            for (int i = 0; i < 24; ++i) gd[i] = g[i];
            int qs;
            const double* mp = mp_of(b, e, qs);
            double* ad = Aw + jj * TG_AE + qq * TG_AQ;
            if (b.material == MAT_SVK) {
                // point_tangent's closed form with the isotropic C written out:
                // A[(i,d),(k,l)] = lam delta_id F[k][l] + mu (delta_il F[k][d] + delta_dl F[k][i]),  F = H
                double lam, mu;
                isotropic_lame<3>(mp + (int64_t)qq * qs, lam, mu);
                double lF[3][3], mF[3][3];
#pragma unroll
                for (int k = 0; k < 3; ++k)
#pragma unroll
                    for (int l = 0; l < 3; ++l) { lF[k][l] = (w * lam) * H[k][l]; mF[k][l] = (w * mu) * H[k][l]; }
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int d = 0; d < 3; ++d)
#pragma unroll
                        for (int k = 0; k < 3; ++k)
#pragma unroll
                            for (int l = 0; l < 3; ++l) {
                                double a = 0.0;
                                if (i == d) a += lF[k][l];
                                if (i == l) a += mF[k][d];
                                if (d == l) a += mF[k][i];
                                ad[(i * 3 + d) * 9 + k * 3 + l] = a;
                            }
            } else {
                tg_generic_tangent(b.material, mp + (int64_t)qq * qs, H, w, ad);
            }
        }
        __syncwarp();
        // ---- phase 2: the nine 8 x 8 x 24 products of every element of the group
#pragma unroll 1
        for (int jv = 0; jv < 4; ++jv) {
            const int64_t ev = 4 * grp + jv;
            if (ev >= b.E) break;
            const double* Ge = Gw + jv * TG_GE;
            double* Ae = Aw + jv * TG_AE;
            double C[3][3][2];
#pragma unroll
            for (int d = 0; d < 3; ++d)
#pragma unroll
                for (int k = 0; k < 3; ++k) C[d][k][0] = C[d][k][1] = 0.0;
#pragma unroll
            for (int s = 0; s < 6; ++s) {
                const double g0 = Ge[offG[s]], g1 = Ge[offG[s] + 1], g2 = Ge[offG[s] + 2];
                const double aop = isel[s] == 0 ? g0 : (isel[s] == 1 ? g1 : g2);
                const double* row = Ae + offA[s];
#pragma unroll
                for (int d = 0; d < 3; ++d)
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const double* r = row + d * 9 + k * 3;
                        const double y = r[0] * g0 + r[1] * g1 + r[2] * g2;
                        dmma884(C[d][k][0], C[d][k][1], aop, y);
                    }
            }
            __syncwarp();                                                   // all lanes are done with A of this element
            // ---- stage K_e[(a)][b][3][3] (a = gid, b = 2 tig, 2 tig + 1) where A lived, then one bulk store
            double2* st = reinterpret_cast<double2*>(Ae + (gid * 8 + 2 * tig) * 9);
            st[0] = make_double2(C[0][0][0], C[0][1][0]);
            st[1] = make_double2(C[0][2][0], C[1][0][0]);
            st[2] = make_double2(C[1][1][0], C[1][2][0]);
            st[3] = make_double2(C[2][0][0], C[2][1][0]);
            st[4] = make_double2(C[2][2][0], C[0][0][1]);
            st[5] = make_double2(C[0][1][1], C[0][2][1]);
            st[6] = make_double2(C[1][0][1], C[1][1][1]);
            st[7] = make_double2(C[1][2][1], C[2][0][1]);
            st[8] = make_double2(C[2][1][1], C[2][2][1]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                const unsigned src = (unsigned)__cvta_generic_to_shared(Ae);
                double* dst = K_e + ev * 576;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 4608;" ::"l"(dst), "r"(src) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
}

// K_e of any constitutive law on hex8 / 8 points through the tensor-core kernel; FEB200_ERR_UNSUPPORTED for anything else
// (the caller then uses the scalar kernels).
int tangent_general_mma(cudaStream_t st, const feb200_batch& b, const double* x_vd, const double* u, const double* mask,
                        const double* values, double* K_e) {
    if (!(b.D == 3 && b.N == 8 && b.Q == 8)) return FEB200_ERR_UNSUPPORTED;
    if (((uintptr_t)K_e & 15) != 0) return FEB200_ERR_UNSUPPORTED;          // bulk stores need 16-byte aligned rows
    const size_t smem = TG_TABLE + (size_t)FEB_TG_WARPS * TG_PER_WARP;
    static bool attr_set = false;
    if (!attr_set) {
        FEB_CUDA(cudaFuncSetAttribute(k_tangent_general_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int64_t want = ceil_div(ceil_div(b.E, 4), FEB_TG_WARPS);
    const int grid = (int)(want < sm_count() ? want : sm_count());
    k_tangent_general_mma<<<grid, FEB_TG_WARPS * 32, smem, st>>>(b, x_vd, u, mask, values, K_e);
    return FEB200_OK;
}

}  // namespace feb
