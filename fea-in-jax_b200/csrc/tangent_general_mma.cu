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
// Here every warp is an independent worker (12 per SM, no block barrier after start-up) that takes groups of 4 elements:
//   phase 1  lane = (element, quadrature point): coordinates / solution from the vertex arrays (Dirichlet blend fused) into
//            shared memory, J, J^-1, the 24 physical gradients (stored as they are formed), H; then, for two elements at a time,
//            w A_q (closed form for St. Venant-Kirchhoff, forward-mode duals on the constitutive law otherwise) into shared
//            memory -- the lanes of the other two elements keep H and w in registers until their turn;
//   phase 2  per element (two elements run as one instruction stream), two chained tensor-core products
//            (mma.sync.m8n8k4.f64), nothing scalar in between:
//              product 1, per point:  Y_q[b][(i,d,k)] = sum_l g_q[b,l] * w A_q[(i,d),(k,l)]          (8 x 27 x 3: 5 DMMAs)
//              product 2, per (d,k):  K^{dk}[a][b]    = sum_{(q,i)} g_q[a,i] * Y_q[b][(i,d,k)]       (8 x 8 x 24: 6 DMMAs)
//            The accumulator fragment of product 1 IS the B fragment of product 2: the columns of product 1 are ordered so
//            that the lane holding Y_q[b][(i,d,k)] is the lane that owns contraction slot (q,i) / column b of product 2
//            (ownership of i rotates with the point through the four lanes of a quad; one register select per entry).
//            Operands reach the lanes as fragments (no replicated reads): 54 shared-memory loads per element instead of the
//            180 of the first version, whose scalar Y stage was bound by shared-memory wavefronts (profiles/
//            r02_tangent_general_mma_v1_ncu_summary.txt).  A lane ends up with the two complete 3 x 3 blocks K[a][2t],
//            K[a][2t+1] = 144 contiguous bytes of K_e[(e*8 + a)][b][3][3];
//   store    the element's 4608 bytes are staged in the shared memory its A_q occupied and leave as ONE bulk asynchronous
//            copy (cp.async.bulk shared -> global), overlapping the next element's products.
#include <stdlib.h>

#include "common.cuh"
#include "element_math.cuh"
#include "hex8_mma.cuh"

#ifndef FEB_TG_WARPS
#define FEB_TG_WARPS 12
#endif
#ifndef FEB_TG_GROUP
#define FEB_TG_GROUP 4                    /* elements per group (lanes 8 * GROUP .. 31 idle in phase 1 when < 4) */
#endif
#ifndef FEB_TG_PAIR
#define FEB_TG_PAIR 1                     /* 1 = phase 2 runs two elements as one instruction stream */
#endif
#ifndef FEB_TG_AGROUP
#define FEB_TG_AGROUP 2                   /* elements of a group whose point tangents w A_q sit in shared memory at a time: the
                                             lanes of the other elements keep H and w in registers and fill the same area
                                             when their turn comes (20.7 KB of A per warp would cap the SM at 8 warps) */
#endif

namespace feb {

constexpr int TG_GQ = 27;                 // doubles per (element, point) row of the gradient buffer: [node][3] + 3 (stores and
                                          // fragment loads of a half-warp then fall into distinct banks)
constexpr int TG_GE = 8 * TG_GQ;          // per element
constexpr int TG_AQ = 81;                 // w A_q: [(i,d)][(k,l)]
constexpr int TG_AE = 8 * TG_AQ;          // per element: 5184 bytes, reused as the 4608-byte staging of K_e
constexpr int TG_DPHI = 26;               // doubles per point of the basis table (24 used; 16-byte rows, conflict-free)
constexpr size_t TG_TABLE = (size_t)(8 * TG_DPHI + 8) * 8;
constexpr int TG_GRP = FEB_TG_GROUP;
constexpr int TG_AGRP = FEB_TG_AGROUP;    // elements whose w A_q is resident at a time
constexpr int TG_NP = FEB_TG_PAIR ? 2 : 1; // elements per instruction stream in phase 2
static_assert(TG_AGRP % TG_NP == 0, "FEB_TG_PAIR needs an even FEB_TG_AGROUP");
static_assert(TG_GRP % TG_AGRP == 0, "FEB_TG_AGROUP must divide FEB_TG_GROUP");
constexpr size_t TG_PER_WARP = (size_t)(TG_GRP * TG_GE + TG_AGRP * TG_AE + TG_GRP * 48) * 8;   // G, w A_q, x / u of the group

// w A_q of any law through forward-mode duals on the constitutive function (element_math.cuh), three directions at a time
// (the derivatives with respect to H[k][0..2] per pass: 36 live dual components instead of the 90 of point_tangent, which
// ptxas keeps in local memory); out of line so that its registers do not weigh on the kernel.
__device__ __noinline__ void tg_generic_tangent(int material, const double* mp, const double (&H)[3][3], double w, double* ad) {
    if (material == MAT_ISOTROPIC || material == MAT_ORTHOTROPIC) {         // constant tangent: no duals needed
        double A[81];
        point_tangent<3>(material, mp, H, A);
        for (int i = 0; i < 81; ++i) ad[i] = w * A[i];
        return;
    }
#pragma unroll 1
    for (int k = 0; k < 3; ++k) {
        Dual<3> Hd[3][3], Sd[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int l = 0; l < 3; ++l) {
                Hd[i][l] = dual_const<3>(H[i][l]);
                if (i == k) Hd[i][l].d[l] = 1.0;
            }
        point_stress<3, Dual<3>>(material, mp, Hd, Sd);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int d = 0; d < 3; ++d)
#pragma unroll
                for (int l = 0; l < 3; ++l) ad[(i * 3 + d) * 9 + k * 3 + l] = w * Sd[i][d].d[l];
    }
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
    double* Aw = Gw + TG_GRP * TG_GE;                                            // [4] AE : [q] AQ
    double* Xw = Aw + TG_AGRP * TG_AE;                                           // [4] 48 : x[8][3], u[8][3] of the group
    for (int i = tid; i < 8 * 24; i += blockDim.x) s_dphi[(i / 24) * TG_DPHI + (i % 24)] = b.dphi[i];
    if (tid < 8) s_w[tid] = b.W[tid];
    __syncthreads();                                                        // the only block-wide barrier

    const int jj = lane >> 3, qq = lane & 7;                                // phase 1: element of the group, quadrature point
    // phase 2, per-lane constants (tests/test_tangent_mma_host.py restates this index scheme on the CPU).
    // Product 1 (Y = g A^T per point q = 4 h + r): rows = node b (gid), contraction = l (tig; 3 is padding), columns
    // c = 8 nt + gid: column c lands in the accumulators of lane-in-quad tig_o = (c % 8) / 2, slot dk = 2 nt + c % 2, and
    // carries gradient component i = (tig_o - r) mod 4 (3 = idle): every point hands each (d,k) product three of the four
    // contraction slots of a quad, rotating with the point, so the 24 (q,i) pairs fill exactly 6 k-steps of product 2.
    const int tig_o = gid >> 1, jcol = gid & 1;
    const int lB = tig < 3 ? tig : 2;                                       // tig = 3 multiplies a zero A entry: any finite value
    const int offA1 = gid * 3 + lB;
    int cB[5], io27[4];
#pragma unroll
    for (int nt = 0; nt < 5; ++nt) {
        const int dk = 2 * nt + jcol;
        cB[nt] = (dk < 8 ? dk : 8) * 3 + lB;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = (tig_o - r) & 3;
        io27[r] = 27 * (i < 3 ? i : 2);
    }
    // Product 2, step (h, st): this lane's slot holds point 4 h + rsel[st], component (tig - rsel[st]) mod 4
    const bool sel0 = (tig == 3), sel1 = (tig == 0 || tig == 3), sel2 = (tig == 2);
    int offAop[6];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int r0 = sel0 ? 1 : 0, r1 = sel1 ? 2 : 1, r2 = sel2 ? 2 : 3;
        offAop[3 * h + 0] = (4 * h + r0) * TG_GQ + gid * 3 + ((tig - r0) & 3);
        offAop[3 * h + 1] = (4 * h + r1) * TG_GQ + gid * 3 + ((tig - r1) & 3);
        offAop[3 * h + 2] = (4 * h + r2) * TG_GQ + gid * 3 + ((tig - r2) & 3);
    }
    const bool hyper = (b.kind == KIND_HYPER);
    const int64_t n_groups = (b.E + TG_GRP - 1) / TG_GRP;
    const int64_t n_workers = (int64_t)gridDim.x * (blockDim.x >> 5);

    for (int64_t grp = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; grp < n_groups; grp += n_workers) {
        const int64_t e = TG_GRP * grp + jj;
        const bool act = jj < TG_GRP && e < b.E;
        // the bulk stores of the previous group have finished reading their staging areas (the A regions)
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        // ---- coordinates and solution of the group -> shared memory: [jj] 48 : x[8][3], u[8][3]
        // (a register software pipeline of these gathers was measured: the compiler spills the prefetched values, 19.2 -> 20.9 ms)
        if (act) {
            const int64_t v = b.conn[e * 8 + qq];
            double* xd = Xw + jj * 48 + qq * 3;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const double xv = x_vd[v * 3 + d];
                double uv = u[v * 3 + d];
                if (mask) {
                    const double m = mask[v * 3 + d];
                    uv = (1.0 - m) * uv + m * values[v * 3 + d];
                }
                xd[d] = xv;
                xd[24 + d] = hyper ? uv + xv : uv;
            }
        }
        __syncwarp();
        // ---- phase 1: geometry and H at (element jj, point qq); every gradient goes to shared memory as soon as it exists
        //      (holding the 24 of them in registers across the stores made ptxas spill them: 88 local-memory accesses per lane)
        double H[3][3], w = 0.0;
        if (act) {
            const double2* x2 = reinterpret_cast<const double2*>(Xw + jj * 48);
            const double2* d2 = reinterpret_cast<const double2*>(s_dphi + qq * TG_DPHI);
            double dq[24];
            double J[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
#pragma unroll
            for (int h = 0; h < 4; ++h) {                            // two nodes = three double2 per step
                const double2 xa = x2[3 * h], xb = x2[3 * h + 1], xc = x2[3 * h + 2];
                const double2 da = d2[3 * h], db = d2[3 * h + 1], dc = d2[3 * h + 2];
                const double xx[6] = {xa.x, xa.y, xb.x, xb.y, xc.x, xc.y};
                dq[6 * h] = da.x; dq[6 * h + 1] = da.y; dq[6 * h + 2] = db.x; dq[6 * h + 3] = db.y; dq[6 * h + 4] = dc.x; dq[6 * h + 5] = dc.y;
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int pp = 0; pp < 3; ++pp)
#pragma unroll
                        for (int dd = 0; dd < 3; ++dd) J[pp][dd] += xx[3 * m + dd] * dq[6 * h + 3 * m + pp];
            }
            double Ji[3][3];
            const double dt = inverse(J, Ji);
            w = dt * s_w[qq];
            const double* ue = Xw + jj * 48 + 24;
            double* gd = Gw + jj * TG_GE + qq * TG_GQ;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int d = 0; d < 3; ++d) H[i][d] = 0.0;
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                double gn[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    double sacc = 0.0;
#pragma unroll
                    for (int pp = 0; pp < 3; ++pp) sacc += Ji[d][pp] * dq[n * 3 + pp];
                    gn[d] = sacc;
                    gd[n * 3 + d] = sacc;
                }
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const double ui = ue[n * 3 + i];
#pragma unroll
                    for (int d = 0; d < 3; ++d) H[i][d] += gn[d] * ui;
                }
            }
        }
        // ---- per sub-group of AGRP elements: their w A_q -> shared memory, then phase 2 for them
#pragma unroll 1
        for (int sub = 0; sub < TG_GRP / TG_AGRP; ++sub) {
        if (sub > 0) {                  // the staged K_e of the previous sub-group has left the A area
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
        }
        if (act && jj / TG_AGRP == sub) {
            int qs;
            const double* mp = mp_of(b, e, qs);
            double* ad = Aw + (jj % TG_AGRP) * TG_AE + qq * TG_AQ;
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
        // ---- phase 2: the chained products of the sub-group's elements, NP elements per instruction stream (two independent
        //      load -> DMMA -> select -> DMMA chains keep the FP64 tensor pipe busy from one warp)
#pragma unroll 1
        for (int jv = sub * TG_AGRP; jv < (sub + 1) * TG_AGRP; jv += TG_NP) {
            const int64_t ev = TG_GRP * grp + jv;
            if (ev >= b.E) break;
            const double* Ge[TG_NP];
            double* Ae[TG_NP];
            double C[TG_NP][9][2];
#pragma unroll
            for (int p = 0; p < TG_NP; ++p) {
                Ge[p] = Gw + (jv + p) * TG_GE;
                Ae[p] = Aw + ((jv + p) % TG_AGRP) * TG_AE;
#pragma unroll
                for (int dk = 0; dk < 9; ++dk) C[p][dk][0] = C[p][dk][1] = 0.0;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                double Y[TG_NP][4][5][2];
                auto product1 = [&](int r) {                                // Y[r][nt] = g_q A_q^T, q = 4 h + r
                    const int q = 4 * h + r;
#pragma unroll
                    for (int p = 0; p < TG_NP; ++p) {
                        double a1 = Ge[p][q * TG_GQ + offA1];
                        if (tig == 3) a1 = 0.0;
                        const double* Aq = Ae[p] + q * TG_AQ + io27[r];
#pragma unroll
                        for (int nt = 0; nt < 5; ++nt) {
                            Y[p][r][nt][0] = Y[p][r][nt][1] = 0.0;
                            dmma884(Y[p][r][nt][0], Y[p][r][nt][1], a1, Aq[cB[nt]]);
                        }
                    }
                };
                auto product2 = [&](int st, bool sel, int ra, int rb) {     // slot value: sel ? Y[ra] : Y[rb]
#pragma unroll
                    for (int p = 0; p < TG_NP; ++p) {
                        const double aop = Ge[p][offAop[3 * h + st]];
#pragma unroll
                        for (int dk = 0; dk < 9; ++dk) {
                            const double y = sel ? Y[p][ra][dk >> 1][dk & 1] : Y[p][rb][dk >> 1][dk & 1];
                            dmma884(C[p][dk][0], C[p][dk][1], aop, y);
                        }
                    }
                };
                product1(0);
                product1(1);
                product2(0, sel0, 1, 0);
                product1(2);
                product2(1, sel1, 2, 1);
                product1(3);
                product2(2, sel2, 2, 3);
            }
            __syncwarp();                                                   // all lanes are done with A of these elements
            // ---- stage K_e[(a)][b][3][3] (a = gid, b = 2 tig, 2 tig + 1) where A lived, then one bulk store per element
#pragma unroll
            for (int p = 0; p < TG_NP; ++p) {
                double2* st = reinterpret_cast<double2*>(Ae[p] + (gid * 8 + 2 * tig) * 9);
                st[0] = make_double2(C[p][0][0], C[p][1][0]);
                st[1] = make_double2(C[p][2][0], C[p][3][0]);
                st[2] = make_double2(C[p][4][0], C[p][5][0]);
                st[3] = make_double2(C[p][6][0], C[p][7][0]);
                st[4] = make_double2(C[p][8][0], C[p][0][1]);
                st[5] = make_double2(C[p][1][1], C[p][2][1]);
                st[6] = make_double2(C[p][3][1], C[p][4][1]);
                st[7] = make_double2(C[p][5][1], C[p][6][1]);
                st[8] = make_double2(C[p][7][1], C[p][8][1]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int p = 0; p < TG_NP; ++p)
                    if (ev + p < b.E) {
                        const unsigned src = (unsigned)__cvta_generic_to_shared(Ae[p]);
                        double* dst = K_e + (ev + p) * 576;
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 4608;" ::"l"(dst), "r"(src) : "memory");
                    }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        }   // sub-group
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
    static int warps = 0;                                                   // FEB200_TG_WARPS: fewer workers per SM (profile drivers)
    if (warps == 0) {
        const char* env = getenv("FEB200_TG_WARPS");
        warps = env ? atoi(env) : FEB_TG_WARPS;
        if (warps < 1 || warps > FEB_TG_WARPS) warps = FEB_TG_WARPS;
    }
    const size_t smem = TG_TABLE + (size_t)warps * TG_PER_WARP;
    static bool attr_set = false;
    if (!attr_set) {
        FEB_CUDA(cudaFuncSetAttribute(k_tangent_general_mma, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(TG_TABLE + (size_t)FEB_TG_WARPS * TG_PER_WARP)));
        attr_set = true;
    }
    const int64_t want = ceil_div(ceil_div(b.E, TG_GRP), warps);
    const int grid = (int)(want < sm_count() ? want : sm_count());
    k_tangent_general_mma<<<grid, warps * 32, smem, st>>>(b, x_vd, u, mask, values, K_e);
    return FEB200_OK;
}

}  // namespace feb
