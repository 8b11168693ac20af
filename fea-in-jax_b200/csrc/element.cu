// Element kernels: batched residual R_e and tangent K_e.
//   residual: one thread per element (replaces vmap(element_residual_func),
//             /root/reference/src/fe_jax/fea.py:800-847 with linear_elasticity.py:145-206 /
//             hyperelasticity.py:93-139 inside)
//   tangent : N threads per element, thread (e, a) owns the block row K[a, :] (replaces
//             vmap(jacfwd(residual_kernel)), fea.py:519-572, by the hand-derived
//             K[(a,d),(b,k)] = sum_q detJ W sum_{i,l} g[a,i] dS[i,d]/dH[k,l] g[b,l]).
// Coordinates and the solution are gathered from vertex arrays through `conn`; the Dirichlet
// blend u_w = (1-m) u + m g (fea.py:948-950) is applied while gathering when a mask is given.
#include <stdlib.h>

#include "common.cuh"
#include "element_math.cuh"

namespace feb {

__device__ __forceinline__ double load_u(const double* __restrict__ u, const double* __restrict__ mask,
                                         const double* __restrict__ values, int64_t f) {
    const double v = u[f];
    if (!mask) return v;
    const double m = mask[f];
    return (1.0 - m) * v + m * values[f];
}

// ------------------------------------------------------------------------------------------------
// residual, compile-time N: registers for x, u, R; dphi / W staged in shared memory; the element-major
// result is transposed through shared memory so that the global stores are coalesced.
// ------------------------------------------------------------------------------------------------
#ifndef FEB_RES_MINBLOCKS
#define FEB_RES_MINBLOCKS 1
#endif
template <int N, int D, int BLOCK>
__global__ void __launch_bounds__(BLOCK, FEB_RES_MINBLOCKS) k_residual_ct(feb200_batch b, const double* __restrict__ x_vd,
                                                       const double* __restrict__ u, const double* __restrict__ mask,
                                                       const double* __restrict__ values, double* __restrict__ R_e) {
    constexpr int NU = N * D;
    extern __shared__ double smem[];
    double* s_dphi = smem;                         // Q*N*D
    double* s_W = s_dphi + b.Q * NU;               // Q
    double* s_out = s_W + b.Q;                     // BLOCK * (NU+1)
    for (int i = threadIdx.x; i < b.Q * NU; i += BLOCK) s_dphi[i] = b.dphi[i];
    for (int i = threadIdx.x; i < b.Q; i += BLOCK) s_W[i] = b.W[i];
    __syncthreads();

    for (int64_t e0 = (int64_t)blockIdx.x * BLOCK; e0 < b.E; e0 += (int64_t)gridDim.x * BLOCK) {
        const int64_t e = e0 + threadIdx.x;
        if (e < b.E) {
            double x[NU], ue[NU], R[NU];
#pragma unroll
            for (int n = 0; n < N; ++n) {
                const int64_t v = b.conn[e * N + n];
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    x[n * D + d] = x_vd[v * D + d];
                    ue[n * D + d] = load_u(u, mask, values, v * D + d);
                    R[n * D + d] = 0.0;
                }
            }
            if (b.kind == KIND_HYPER)
#pragma unroll
                for (int i = 0; i < NU; ++i) ue[i] += x[i];
            int qs;
            const double* mp = mp_of(b, e, qs);
            // Per point, in reference coordinates: the physical gradients g[n][d] = sum_p Ji[d][p] dphi[n][p] are never
            // formed (24 fewer live doubles for hex8).  H[i][d] = sum_n u[n][i] g[n][d] = sum_p Hr[i][p] Ji[d][p] with
            // Hr = sum_n u[n] (x) dphi[n];  R[n][d] += w sum_i g[n][i] S[i][d] = sum_p dphi[n][p] T[p][d] with
            // T = w Ji^T S.  Same quantities as linear_elasticity.py:145-206 / hyperelasticity.py:93-139.
            for (int q = 0; q < b.Q; ++q) {
                const volatile double* dq = s_dphi + q * NU;            // in program order: hoisting all 2 NU loads spills
                double J[D][D], Ji[D][D], Hr[D][D], H[D][D], S[D][D], T[D][D];
#pragma unroll
                for (int p = 0; p < D; ++p)
#pragma unroll
                    for (int d = 0; d < D; ++d) { J[p][d] = 0.0; Hr[p][d] = 0.0; }
#pragma unroll
                for (int n = 0; n < N; ++n)
#pragma unroll
                    for (int p = 0; p < D; ++p) {
                        const double dp = dq[n * D + p];
#pragma unroll
                        for (int d = 0; d < D; ++d) {
                            J[p][d] += x[n * D + d] * dp;
                            Hr[d][p] += ue[n * D + d] * dp;
                        }
                    }
                const double dt = inverse(J, Ji);                           // Ji[d][p] = d xi_p / d x_d
#pragma unroll
                for (int i = 0; i < D; ++i)
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        double sacc = 0.0;
#pragma unroll
                        for (int p = 0; p < D; ++p) sacc += Hr[i][p] * Ji[d][p];
                        H[i][d] = sacc;
                    }
                point_stress<D, double>(b.material, mp + (int64_t)q * qs, H, S);
                const double w = dt * s_W[q];
#pragma unroll
                for (int p = 0; p < D; ++p)
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        double sacc = 0.0;
#pragma unroll
                        for (int i = 0; i < D; ++i) sacc += Ji[i][p] * S[i][d];
                        T[p][d] = w * sacc;
                    }
#pragma unroll
                for (int n = 0; n < N; ++n)
#pragma unroll
                    for (int p = 0; p < D; ++p) {
                        const double dp = dq[n * D + p];
#pragma unroll
                        for (int d = 0; d < D; ++d) R[n * D + d] += dp * T[p][d];
                    }
            }
#pragma unroll
            for (int i = 0; i < NU; ++i) s_out[threadIdx.x * (NU + 1) + i] = R[i];
        }
        __syncthreads();
        const int n_here = (int)((b.E - e0 < BLOCK) ? (b.E - e0) : BLOCK);
        for (int i = threadIdx.x; i < n_here * NU; i += BLOCK) {
            const int t = i / NU, c = i - t * NU;                          // constant divisor, 32-bit
            R_e[e0 * NU + i] = s_out[t * (NU + 1) + c];
        }
        __syncthreads();
    }
}

// residual, runtime N (quadratic elements): local-memory arrays
constexpr int MAX_RT_N = 27;
template <int D>
__global__ void __launch_bounds__(64) k_residual_rt(feb200_batch b, const double* __restrict__ x_vd,
                                                    const double* __restrict__ u, const double* __restrict__ mask,
                                                    const double* __restrict__ values, double* __restrict__ R_e) {
    const int N = b.N, NU = N * D;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < b.E; e += (int64_t)gridDim.x * blockDim.x) {
        double x[MAX_RT_N * D], ue[MAX_RT_N * D], R[MAX_RT_N * D], g[MAX_RT_N * D];
        for (int n = 0; n < N; ++n) {
            const int64_t v = b.conn[e * N + n];
            for (int d = 0; d < D; ++d) {
                x[n * D + d] = x_vd[v * D + d];
                ue[n * D + d] = load_u(u, mask, values, v * D + d);
            }
        }
        int qs;
        const double* mp = mp_of(b, e, qs);
        element_residual_rt<D>(b.kind, b.material, N, b.Q, x, ue, b.dphi, b.W, mp, qs, g, R);
        for (int i = 0; i < NU; ++i) R_e[e * NU + i] = R[i];
    }
}

// ------------------------------------------------------------------------------------------------
// tangent: shared-memory staging common to all variants.  Per element in the block:
//   xe[N*D], ue[N*D], g[Q][N*D], wdet[Q], aux[Q][AUX]   (AUX = D^4 general, 2 isotropic fast path)
// Thread t of the block is (element e_l = t / N, node a = t % N).
// ------------------------------------------------------------------------------------------------
template <int D, int NCT = 0>
__device__ __forceinline__ void stage_geometry(const feb200_batch& b, int N, int e_l, int a, int64_t e, bool active,
                                               const double* __restrict__ x_vd, const double* __restrict__ u,
                                               const double* __restrict__ mask, const double* __restrict__ values,
                                               const double* s_dphi, double* s_el, int per_el, int AUX, bool iso_fast) {
    const int NU = N * D, Q = b.Q;
    double* xe = s_el + (size_t)e_l * per_el;
    double* ue = xe + NU;
    double* g = ue + NU;
    double* wdet = g + (size_t)Q * (NU + 1);
    double* aux = wdet + Q;
    if (active) {
        const int64_t v = b.conn[e * N + a];
        for (int d = 0; d < D; ++d) {
            xe[a * D + d] = x_vd[v * D + d];
            double uu = load_u(u, mask, values, v * D + d);
            ue[a * D + d] = uu;
        }
    }
    __syncthreads();
    if (active) {
        int qs;
        const double* mp = mp_of(b, e, qs);
        for (int q = a; q < Q; q += N) {
            double* gq = g + (size_t)q * (NU + 1);
            double dt;
            if (NCT > 0) {          // compile-time N: coordinates / basis table / gradients in registers
                constexpr int NL = (NCT > 0 ? NCT : 1) * D;
                double xl[NL], dl[NL], gl[NL];
#pragma unroll
                for (int i = 0; i < NL; ++i) { xl[i] = xe[i]; dl[i] = s_dphi[(size_t)q * (NU + 1) + i]; }
                dt = shape_gradients<(NCT > 0 ? NCT : 1), D>(xl, dl, gl);
#pragma unroll
                for (int i = 0; i < NL; ++i) gq[i] = gl[i];
            } else {
                dt = shape_gradients_rt<D>(N, xe, s_dphi + (size_t)q * (NU + 1), gq);
            }
            wdet[q] = dt * b.W[q];
            const double* mpq = mp + (int64_t)q * qs;
            if (iso_fast) {
                double lam, mu;
                isotropic_lame<D>(mpq, lam, mu);
                aux[q * 2 + 0] = lam;
                aux[q * 2 + 1] = mu;
            } else {
                double H[D][D];
                for (int i = 0; i < D; ++i)
                    for (int d = 0; d < D; ++d) H[i][d] = 0.0;
                for (int n = 0; n < N; ++n)
                    for (int i = 0; i < D; ++i) {
                        const double ui = (b.kind == KIND_HYPER) ? ue[n * D + i] + xe[n * D + i] : ue[n * D + i];
                        for (int d = 0; d < D; ++d) H[i][d] += gq[n * D + d] * ui;
                    }
                double A[D * D * D * D];
                point_tangent<D>(b.material, mpq, H, A);
                for (int i = 0; i < AUX; ++i) aux[(size_t)q * AUX + i] = A[i];
            }
        }
    }
    __syncthreads();
}

// isotropic linear elasticity, closed form, q outer with the whole block row K[a, 0..N) in registers:
//   K_ab[d][k] = sum_q w (lam g_a[d] g_b[k] + mu g_a[k] g_b[d]) + delta_dk sum_q w mu g_a.g_b
// CONST_MAT (material constant over the element: layouts (M,) and (E,M)): accumulate
//   S_ab[i][l] = sum_q w g_a[i] g_b[l]   (9 FMA per pair and point instead of 21)
// and apply lam, mu once: K_ab[d][k] = lam S[d][k] + mu S[k][d] + delta_dk mu tr S.
// The block row (N*D*D doubles, contiguous in K_e) is transposed through shared memory in PASSES
// pieces so that the global stores are full 32-byte sectors written by consecutive lanes.
template <int N, int D, bool CONST_MAT, int PASSES>
__global__ void __launch_bounds__(128) k_tangent_iso(feb200_batch b, int epb, int per_el, const double* __restrict__ x_vd,
                                                     const double* __restrict__ u, const double* __restrict__ mask,
                                                     const double* __restrict__ values, double* __restrict__ K_e) {
    constexpr int NU = N * D, D2 = D * D;
    constexpr int NB = N / PASSES;                  // blocks per pass
    constexpr int PIECE = NB * D2;                  // doubles per thread and pass
    constexpr int PSTRIDE = PIECE | 1;              // odd shared-memory stride: conflict-free staging writes
    static_assert(N % PASSES == 0, "PASSES must divide N");
    extern __shared__ double smem[];
    double* s_dphi = smem;
    double* s_el = smem + (size_t)b.Q * (NU + 1);         // geometry staging, later reused as store staging
    for (int i = threadIdx.x; i < b.Q * NU; i += blockDim.x) s_dphi[(i / NU) * (NU + 1) + i % NU] = b.dphi[i];   // padded rows
    const int e_l = threadIdx.x / N, a = threadIdx.x - e_l * N;
    const int Q = b.Q;
    const int n_slots_blk = epb * N;                // (element, node) slots handled per tile
    for (int64_t e0 = (int64_t)blockIdx.x * epb; e0 < b.E; e0 += (int64_t)gridDim.x * epb) {
        const int64_t e = e0 + e_l;
        const bool active = (e_l < epb) && (e < b.E);
        __syncthreads();
        stage_geometry<D, N>(b, N, e_l, a, e, active, x_vd, u, mask, values, s_dphi, s_el, per_el, 2, true);
        double acc[N][D2];
#pragma unroll
        for (int bn = 0; bn < N; ++bn)
#pragma unroll
            for (int c = 0; c < D2; ++c) acc[bn][c] = 0.0;
        if (active) {
            const double* g = s_el + (size_t)e_l * per_el + 2 * NU;
            const double* wdet = g + (size_t)Q * (NU + 1);
            const double* aux = wdet + Q;
            for (int q = 0; q < Q; ++q) {
                const double* gq = g + (size_t)q * (NU + 1);
                const double w = wdet[q];
                double ga[D];
#pragma unroll
                for (int d = 0; d < D; ++d) ga[d] = gq[a * D + d];
                if (CONST_MAT) {
                    double wga[D];
#pragma unroll
                    for (int d = 0; d < D; ++d) wga[d] = w * ga[d];
#pragma unroll
                    for (int bn = 0; bn < N; ++bn) {
                        double gb[D];
#pragma unroll
                        for (int d = 0; d < D; ++d) gb[d] = gq[bn * D + d];
#pragma unroll
                        for (int i = 0; i < D; ++i)
#pragma unroll
                            for (int l = 0; l < D; ++l) acc[bn][i * D + l] += wga[i] * gb[l];
                    }
                } else {
                    const double wl = w * aux[q * 2], wm = w * aux[q * 2 + 1];
                    double la[D], ma[D];
#pragma unroll
                    for (int d = 0; d < D; ++d) { la[d] = wl * ga[d]; ma[d] = wm * ga[d]; }
#pragma unroll
                    for (int bn = 0; bn < N; ++bn) {
                        double gb[D];
#pragma unroll
                        for (int d = 0; d < D; ++d) gb[d] = gq[bn * D + d];
                        double dot = 0.0;
#pragma unroll
                        for (int d = 0; d < D; ++d) dot += ma[d] * gb[d];
#pragma unroll
                        for (int d = 0; d < D; ++d)
#pragma unroll
                            for (int k = 0; k < D; ++k) acc[bn][d * D + k] += la[d] * gb[k] + ma[k] * gb[d];
#pragma unroll
                        for (int d = 0; d < D; ++d) acc[bn][d * D + d] += dot;
                    }
                }
            }
            if (CONST_MAT) {
                const double lam = aux[0], mu = aux[1];
#pragma unroll
                for (int bn = 0; bn < N; ++bn) {
                    double S[D2];
                    double tr = 0.0;
#pragma unroll
                    for (int c = 0; c < D2; ++c) S[c] = acc[bn][c];
#pragma unroll
                    for (int d = 0; d < D; ++d) tr += S[d * D + d];
#pragma unroll
                    for (int d = 0; d < D; ++d)
#pragma unroll
                        for (int k = 0; k < D; ++k)
                            acc[bn][d * D + k] = lam * S[d * D + k] + mu * S[k * D + d] + (d == k ? mu * tr : 0.0);
                }
            }
        }
        // ---- coalesced store of the block rows through shared memory
        const int64_t slot0 = e0 * N;                                            // first slot of this tile
        const int64_t n_slots = ((b.E - e0) < epb ? (b.E - e0) : epb) * (int64_t)N;
#pragma unroll
        for (int pass = 0; pass < PASSES; ++pass) {
            __syncthreads();                                                     // geometry / previous pass consumed
            if (active) {
                double* dst = s_el + (size_t)threadIdx.x * PSTRIDE;
#pragma unroll
                for (int bn = 0; bn < NB; ++bn)
#pragma unroll
                    for (int c = 0; c < D2; ++c) dst[bn * D2 + c] = acc[pass * NB + bn][c];
            }
            __syncthreads();
            const int n_items = (int)n_slots * PIECE;
            for (int i = threadIdx.x; i < n_items; i += blockDim.x) {
                const int t = i / PIECE;                                   // constant divisor, 32-bit
                const int c = i - t * PIECE;
                K_e[(slot0 + t) * (int64_t)(N * D2) + pass * PIECE + c] = s_el[t * PSTRIDE + c];
            }
        }
        (void)n_slots_blk;
    }
}

// general material, compile-time N: q outer, block row kept in registers
template <int N, int D>
__global__ void k_tangent_general_ct(feb200_batch b, int epb, int per_el, const double* __restrict__ x_vd,
                                     const double* __restrict__ u, const double* __restrict__ mask,
                                     const double* __restrict__ values, double* __restrict__ K_e) {
    constexpr int NU = N * D, D2 = D * D, D4 = D2 * D2;
    extern __shared__ double smem[];
    double* s_dphi = smem;
    double* s_el = smem + (size_t)b.Q * (NU + 1);
    for (int i = threadIdx.x; i < b.Q * NU; i += blockDim.x) s_dphi[(i / NU) * (NU + 1) + i % NU] = b.dphi[i];   // padded rows
    const int e_l = threadIdx.x / N, a = threadIdx.x - e_l * N;
    const int Q = b.Q;
    for (int64_t e0 = (int64_t)blockIdx.x * epb; e0 < b.E; e0 += (int64_t)gridDim.x * epb) {
        const int64_t e = e0 + e_l;
        const bool active = (e_l < epb) && (e < b.E);
        __syncthreads();
        stage_geometry<D, N>(b, N, e_l, a, e, active, x_vd, u, mask, values, s_dphi, s_el, per_el, D4, false);
        if (!active) continue;
        const double* g = s_el + (size_t)e_l * per_el + 2 * NU;
        const double* wdet = g + (size_t)Q * (NU + 1);
        const double* aux = wdet + Q;
        double acc[N][D2];
#pragma unroll
        for (int bn = 0; bn < N; ++bn)
#pragma unroll
            for (int c = 0; c < D2; ++c) acc[bn][c] = 0.0;
        for (int q = 0; q < Q; ++q) {
            const double* A = aux + (size_t)q * D4;
            const double* gq = g + (size_t)q * (NU + 1);
            double t[D2][D];   // t[d*D+k][l] = w * sum_i g_a[i] A[i,d,k,l]
            const double w = wdet[q];
#pragma unroll
            for (int dk = 0; dk < D2; ++dk)
#pragma unroll
                for (int l = 0; l < D; ++l) {
                    const int d = dk / D, k = dk - d * D;
                    double sacc = 0.0;
#pragma unroll
                    for (int i = 0; i < D; ++i) sacc += gq[a * D + i] * A[(i * D + d) * D2 + k * D + l];
                    t[dk][l] = w * sacc;
                }
#pragma unroll
            for (int bn = 0; bn < N; ++bn)
#pragma unroll
                for (int dk = 0; dk < D2; ++dk) {
                    double sacc = acc[bn][dk];
#pragma unroll
                    for (int l = 0; l < D; ++l) sacc += t[dk][l] * gq[bn * D + l];
                    acc[bn][dk] = sacc;
                }
        }
        double* out = K_e + ((e * N + a) * (int64_t)N) * D2;
        if ((N * D2) % 2 == 0) {       // the row block of a node is N * D2 doubles: 16-byte aligned, stored as pairs
            double2* out2 = reinterpret_cast<double2*>(out);
#pragma unroll
            for (int i = 0; i < N * D2 / 2; ++i)
                out2[i] = make_double2(acc[(2 * i) / D2][(2 * i) % D2], acc[(2 * i + 1) / D2][(2 * i + 1) % D2]);
        } else {
#pragma unroll
            for (int bn = 0; bn < N; ++bn)
#pragma unroll
                for (int c = 0; c < D2; ++c) out[bn * D2 + c] = acc[bn][c];
        }
    }
}

// general material, runtime N: b outer, q inner, t recomputed (quadratic elements; coverage path)
template <int D>
__global__ void k_tangent_general_rt(feb200_batch b, int epb, int per_el, const double* __restrict__ x_vd,
                                     const double* __restrict__ u, const double* __restrict__ mask,
                                     const double* __restrict__ values, double* __restrict__ K_e) {
    constexpr int D2 = D * D, D4 = D2 * D2;
    const int N = b.N, NU = N * D, Q = b.Q;
    extern __shared__ double smem[];
    double* s_dphi = smem;
    double* s_el = smem + (size_t)Q * (NU + 1);
    for (int i = threadIdx.x; i < Q * NU; i += blockDim.x) s_dphi[(i / NU) * (NU + 1) + i % NU] = b.dphi[i];   // padded rows
    const int e_l = threadIdx.x / N, a = threadIdx.x - e_l * N;
    for (int64_t e0 = (int64_t)blockIdx.x * epb; e0 < b.E; e0 += (int64_t)gridDim.x * epb) {
        const int64_t e = e0 + e_l;
        const bool active = (e_l < epb) && (e < b.E);
        __syncthreads();
        stage_geometry<D>(b, N, e_l, a, e, active, x_vd, u, mask, values, s_dphi, s_el, per_el, D4, false);
        if (!active) continue;
        const double* g = s_el + (size_t)e_l * per_el + 2 * NU;
        const double* wdet = g + (size_t)Q * (NU + 1);
        const double* aux = wdet + Q;
        double* out = K_e + ((e * N + a) * (int64_t)N) * D2;
        for (int bn = 0; bn < N; ++bn) {
            double acc[D2];
#pragma unroll
            for (int c = 0; c < D2; ++c) acc[c] = 0.0;
            for (int q = 0; q < Q; ++q) {
                const double* A = aux + (size_t)q * D4;
                const double* gq = g + (size_t)q * (NU + 1);
                const double w = wdet[q];
#pragma unroll
                for (int dk = 0; dk < D2; ++dk) {
                    const int d = dk / D, k = dk - d * D;
                    double sacc = 0.0;
#pragma unroll
                    for (int l = 0; l < D; ++l) {
                        double tl = 0.0;
#pragma unroll
                        for (int i = 0; i < D; ++i) tl += gq[a * D + i] * A[(i * D + d) * D2 + k * D + l];
                        sacc += tl * gq[bn * D + l];
                    }
                    acc[dk] += w * sacc;
                }
            }
#pragma unroll
            for (int c = 0; c < D2; ++c) out[bn * D2 + c] = acc[c];
        }
    }
}

// diag(K_e) of the linear isotropic law without forming K_e: K[(a,d),(a,d)] = sum_q w (lam* g_ad^2 + mu g_ad^2 + mu |g_a|^2)
// (SURVEY appendix A.2 with a = b, d = k; calculate_jacobian_diag_wo_dirichlet, fea.py:759-797).  Thread per element.
template <int D>
__global__ void __launch_bounds__(128) k_tangent_diag_iso(feb200_batch b, const double* __restrict__ x_vd, double* __restrict__ diag_e) {
    const int N = b.N;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < b.E; e += (int64_t)gridDim.x * blockDim.x) {
        double x[MAX_RT_N * D], g[MAX_RT_N * D], acc[MAX_RT_N * D];
        for (int n = 0; n < N; ++n) {
            const int64_t v = b.conn[e * N + n];
            for (int d = 0; d < D; ++d) { x[n * D + d] = x_vd[v * D + d]; acc[n * D + d] = 0.0; }
        }
        int qs;
        const double* mp = mp_of(b, e, qs);
        for (int q = 0; q < b.Q; ++q) {
            const double dt = shape_gradients_rt<D>(N, x, b.dphi + (int64_t)q * N * D, g);
            double lam, mu;
            isotropic_lame<D>(mp + (int64_t)q * qs, lam, mu);
            const double w = dt * b.W[q];
            for (int n = 0; n < N; ++n) {
                double g2 = 0.0;
                for (int d = 0; d < D; ++d) g2 += g[n * D + d] * g[n * D + d];
                for (int d = 0; d < D; ++d) acc[n * D + d] += w * ((lam + mu) * g[n * D + d] * g[n * D + d] + mu * g2);
            }
        }
        for (int i = 0; i < N * D; ++i) diag_e[e * N * D + i] = acc[i];
    }
}

static int check_batch(const feb200_batch* b, const char* who) {
    FEB_REQUIRE(b, FEB200_ERR_INVALID, "%s: null batch", who);
    FEB_REQUIRE(b->E >= 0 && b->N > 0 && b->Q > 0 && (b->D == 2 || b->D == 3), FEB200_ERR_INVALID,
                "%s: bad sizes E=%lld N=%d Q=%d D=%d", who, (long long)b->E, b->N, b->Q, b->D);
    FEB_REQUIRE(b->N <= MAX_RT_N, FEB200_ERR_UNSUPPORTED, "%s: N=%d exceeds %d", who, b->N, MAX_RT_N);
    FEB_REQUIRE(b->conn && b->dphi && b->W && b->mp, FEB200_ERR_INVALID, "%s: null table pointer", who);
    FEB_REQUIRE(b->kind == KIND_LINEAR || b->kind == KIND_HYPER, FEB200_ERR_INVALID, "%s: bad kind %d", who, b->kind);
    FEB_REQUIRE(b->material >= 0 && b->material <= 3, FEB200_ERR_INVALID, "%s: bad material %d", who, b->material);
    FEB_REQUIRE(b->layout >= 0 && b->layout <= 2, FEB200_ERR_INVALID, "%s: bad layout %d", who, b->layout);
    const bool lin = (b->material == MAT_ISOTROPIC || b->material == MAT_ORTHOTROPIC);
    FEB_REQUIRE(lin == (b->kind == KIND_LINEAR), FEB200_ERR_UNSUPPORTED,
                "%s: material %d cannot be used with residual kind %d", who, b->material, b->kind);
    int need = 2;
    if (b->material == MAT_ORTHOTROPIC) need = (b->D == 2) ? 4 : 9;
    if (b->material == MAT_MOONEY_RIVLIN) need = 3;
    FEB_REQUIRE(b->M >= need, FEB200_ERR_INVALID, "%s: material needs %d parameters, got %d", who, need, b->M);
    return FEB200_OK;
}

template <int N, int D>
static int launch_residual_ct(cudaStream_t st, const feb200_batch& b, const double* x_vd, const double* u,
                              const double* mask, const double* values, double* R_e) {
    constexpr int BLOCK = 128;
    const size_t smem = ((size_t)b.Q * N * D + b.Q + (size_t)BLOCK * (N * D + 1)) * sizeof(double);
    auto kern = k_residual_ct<N, D, BLOCK>;
    if (smem > 48 * 1024) FEB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t g = ceil_div(b.E, BLOCK);
    const int64_t cap = (int64_t)sm_count() * 8;
    if (g > cap) g = cap;
    kern<<<(int)g, BLOCK, smem, st>>>(b, x_vd, u, mask, values, R_e);
    return FEB200_OK;
}

struct TangentCfg { int epb, per_el, threads; size_t smem; };
static int tangent_cfg(const feb200_batch& b, int aux, TangentCfg& c, int max_threads = 256, int stage_doubles_per_thread = 0) {
    const int NU = b.N * b.D;
    c.per_el = 2 * NU + b.Q * (NU + 1) + b.Q + b.Q * aux;
    if (c.per_el < stage_doubles_per_thread * b.N) c.per_el = stage_doubles_per_thread * b.N;   // store staging aliases it
    while (c.per_el % 16 != 4) ++c.per_el;   // elements of one warp hit distinct shared-memory banks on broadcast reads
    const size_t table = (size_t)b.Q * (NU + 1) * sizeof(double);
    const size_t budget = 96 * 1024;
    int epb = max_threads / b.N;
    while (epb > 1 && table + (size_t)epb * c.per_el * sizeof(double) > budget) --epb;
    if (epb < 1) epb = 1;
    c.epb = epb;
    c.smem = table + (size_t)epb * c.per_el * sizeof(double);
    FEB_REQUIRE(c.smem <= 200 * 1024, FEB200_ERR_CAPACITY, "tangent: element needs %zu bytes of shared memory", c.smem);
    c.threads = ((epb * b.N + 31) / 32) * 32;
    return FEB200_OK;
}

template <typename Kern>
static int launch_tangent(Kern kern, cudaStream_t st, const feb200_batch& b, const TangentCfg& c, const double* x_vd,
                          const double* u, const double* mask, const double* values, double* K_e) {
    if (c.smem > 48 * 1024) FEB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
    int64_t g = ceil_div(b.E, c.epb);
    const int64_t cap = (int64_t)sm_count() * 16;
    if (g > cap) g = cap;
    kern<<<(int)g, c.threads, c.smem, st>>>(b, c.epb, c.per_el, x_vd, u, mask, values, K_e);
    return FEB200_OK;
}

int tangent_iso_mma(cudaStream_t st, const feb200_batch& b, const double* x_vd, double* K_e);   // tangent_mma.cu
int tangent_general_mma(cudaStream_t st, const feb200_batch& b, const double* x_vd, const double* u, const double* mask,
                        const double* values, double* K_e);                                       // tangent_general_mma.cu

}  // namespace feb

using namespace feb;

extern "C" {

int feb200_element_residual(feb200_stream s, const feb200_batch* bp, const double* x_vd, const double* u,
                            const double* mask, const double* values, double* R_e) {
    int rc = check_batch(bp, "element_residual");
    if (rc) return rc;
    FEB_REQUIRE(x_vd && u && R_e, FEB200_ERR_INVALID, "element_residual: null pointer");
    FEB_REQUIRE(!mask || values, FEB200_ERR_INVALID, "element_residual: mask without values");
    const feb200_batch& b = *bp;
    if (b.E == 0) return FEB200_OK;
    cudaStream_t st = as_stream(s);
    rc = FEB200_OK;
    if (b.D == 2 && b.N == 3) rc = launch_residual_ct<3, 2>(st, b, x_vd, u, mask, values, R_e);
    else if (b.D == 2 && b.N == 4) rc = launch_residual_ct<4, 2>(st, b, x_vd, u, mask, values, R_e);
    else if (b.D == 3 && b.N == 4) rc = launch_residual_ct<4, 3>(st, b, x_vd, u, mask, values, R_e);
    else if (b.D == 3 && b.N == 8) rc = launch_residual_ct<8, 3>(st, b, x_vd, u, mask, values, R_e);
    else if (b.D == 2 && b.N == 6) rc = launch_residual_ct<6, 2>(st, b, x_vd, u, mask, values, R_e);      // tri6
    else if (b.D == 3 && b.N == 10) rc = launch_residual_ct<10, 3>(st, b, x_vd, u, mask, values, R_e);    // tet10
    else if (b.D == 2 && b.N == 9) rc = launch_residual_ct<9, 2>(st, b, x_vd, u, mask, values, R_e);      // quad9
    else if (b.D == 2) k_residual_rt<2><<<grid_for(b.E, 64), 64, 0, st>>>(b, x_vd, u, mask, values, R_e);
    else k_residual_rt<3><<<grid_for(b.E, 64), 64, 0, st>>>(b, x_vd, u, mask, values, R_e);
    if (rc) return rc;
    FEB_CHECK_LAUNCH("element_residual");
    return FEB200_OK;
}

int feb200_element_tangent(feb200_stream s, const feb200_batch* bp, const double* x_vd, const double* u,
                           const double* mask, const double* values, double* K_e) {
    int rc = check_batch(bp, "element_tangent");
    if (rc) return rc;
    FEB_REQUIRE(x_vd && u && K_e, FEB200_ERR_INVALID, "element_tangent: null pointer");
    FEB_REQUIRE(!mask || values, FEB200_ERR_INVALID, "element_tangent: mask without values");
    const feb200_batch& b = *bp;
    if (b.E == 0) return FEB200_OK;
    cudaStream_t st = as_stream(s);
    const bool iso = (b.material == MAT_ISOTROPIC && b.kind == KIND_LINEAR);
    const int D4 = b.D * b.D * b.D * b.D;
    TangentCfg c;
    const bool cm = (b.layout != LAYOUT_EQM);
    constexpr int PASSES_HEX = 2;
    if (iso && b.D == 3 && b.N > 10) {       // many-node elements (hex27): block products on the FP64 tensor cores
        rc = tangent_iso_mma(st, b, x_vd, K_e);
        if (rc == FEB200_OK) {
            FEB_CHECK_LAUNCH("element_tangent");
            return FEB200_OK;
        }
        if (rc != FEB200_ERR_UNSUPPORTED) return rc;
    }
    if (!iso && b.D == 3 && b.N == 8 && b.Q == 8) {      // hex8, state-dependent law: nine 8 x 8 x 24 tensor-core products
        static const bool scalar_only = getenv("FEB200_TANGENT_SCALAR") != nullptr;     // A/B switch for the profile drivers
        if (!scalar_only) {
            rc = tangent_general_mma(st, b, x_vd, u, mask, values, K_e);
            if (rc == FEB200_OK) {
                FEB_CHECK_LAUNCH("element_tangent");
                return FEB200_OK;
            }
            if (rc != FEB200_ERR_UNSUPPORTED) return rc;
        }
    }
#ifndef FEB_GENERAL_CT_MAX_NU
#define FEB_GENERAL_CT_MAX_NU 30      /* general (state-dependent) tangent: longer block rows (hex27, quad9 ...) go to the runtime-N kernel; tet10 (30) spills but is 3x faster compiled */
#endif
#define FEB_TANGENT_CASE(NN, DD, PP)                                                                               \
    if (b.N == NN && b.D == DD && (iso || NN * DD <= FEB_GENERAL_CT_MAX_NU)) {                                     \
        if (iso) {                                                                                                 \
            rc = tangent_cfg(b, 2, c, 128, ((NN / PP) * DD * DD) | 1);                                                   \
            if (rc) return rc;                                                                                     \
            rc = cm ? launch_tangent(k_tangent_iso<NN, DD, true, PP>, st, b, c, x_vd, u, mask, values, K_e)        \
                    : launch_tangent(k_tangent_iso<NN, DD, false, PP>, st, b, c, x_vd, u, mask, values, K_e);      \
        } else {                                                                                                   \
            rc = tangent_cfg(b, D4, c);                                                                            \
            if (rc) return rc;                                                                                     \
            rc = launch_tangent(k_tangent_general_ct<NN, DD>, st, b, c, x_vd, u, mask, values, K_e);               \
        }                                                                                                          \
    } else
    FEB_TANGENT_CASE(3, 2, 1)
    FEB_TANGENT_CASE(4, 2, 1)
    FEB_TANGENT_CASE(4, 3, 1)
    FEB_TANGENT_CASE(6, 2, 1)      /* tri6 */
    FEB_TANGENT_CASE(10, 3, 2)     /* tet10 */
    FEB_TANGENT_CASE(9, 2, 1)      /* quad9 */
    FEB_TANGENT_CASE(8, 3, PASSES_HEX) {
        rc = tangent_cfg(b, D4, c);     // the generic kernel stages the full D^4 tangent
        if (rc) return rc;
        rc = (b.D == 2) ? launch_tangent(k_tangent_general_rt<2>, st, b, c, x_vd, u, mask, values, K_e)
                        : launch_tangent(k_tangent_general_rt<3>, st, b, c, x_vd, u, mask, values, K_e);
    }
#undef FEB_TANGENT_CASE
    if (rc) return rc;
    FEB_CHECK_LAUNCH("element_tangent");
    return FEB200_OK;
}

int feb200_element_tangent_diag(feb200_stream s, const feb200_batch* bp, const double* x_vd, double* diag_e) {
    int rc = check_batch(bp, "element_tangent_diag");
    if (rc) return rc;
    FEB_REQUIRE(x_vd && diag_e, FEB200_ERR_INVALID, "element_tangent_diag: null pointer");
    FEB_REQUIRE(bp->kind == KIND_LINEAR && bp->material == MAT_ISOTROPIC, FEB200_ERR_UNSUPPORTED,
                "element_tangent_diag: only the linear isotropic law has a diagonal-only kernel");
    if (bp->E == 0) return FEB200_OK;
    const int grid = grid_for(bp->E, 128);
    if (bp->D == 3) k_tangent_diag_iso<3><<<grid, 128, 0, as_stream(s)>>>(*bp, x_vd, diag_e);
    else k_tangent_diag_iso<2><<<grid, 128, 0, as_stream(s)>>>(*bp, x_vd, diag_e);
    FEB_CHECK_LAUNCH("element_tangent_diag");
    return FEB200_OK;
}

}  // extern "C"
