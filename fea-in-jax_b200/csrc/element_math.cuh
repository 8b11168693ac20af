// Per-quadrature-point math of the element kernels: isoparametric map, shape-function
// gradients, the four constitutive laws and their consistent tangents.
//
// Everything here is __host__ __device__ so that tests/hostcheck can compile the same
// templates with g++ and compare them with the oracle on the CPU box (the kernels that
// use them live in element.cu and only run on the GPU).
//
// Reference formulas (paths relative to /root/reference):
//   geometry              src/fe_jax/linear_elasticity.py:180-184
//   small strain residual src/fe_jax/linear_elasticity.py:186-204
//   elastic_isotropic     src/fe_jax/linear_elasticity.py:9-59   (plane stress in 2-D)
//   elastic_orthotropic   src/fe_jax/linear_elasticity.py:62-142
//   st_venant_kirchhoff   src/fe_jax/hyperelasticity.py:9-53
//   mooney_rivlin         src/fe_jax/hyperelasticity.py:56-91
//   Voigt order           src/fe_jax/utils.py:202-253  (xx,yy,zz,yz,xz,xy ; shear doubled)
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define FEB_HD __host__ __device__ __forceinline__
#else
#define FEB_HD inline
#endif

namespace feb {

enum Material : int { MAT_ISOTROPIC = 0, MAT_ORTHOTROPIC = 1, MAT_SVK = 2, MAT_MOONEY_RIVLIN = 3 };
enum ResidualKind : int { KIND_LINEAR = 0, KIND_HYPER = 1 };
enum MaterialLayout : int { LAYOUT_M = 0, LAYOUT_EM = 1, LAYOUT_EQM = 2 };

template <int D> struct VoigtSize { static constexpr int value = (D == 3) ? 6 : (D == 2 ? 3 : 1); };

// Voigt slot of tensor entry (i, j)
template <int D> FEB_HD int voigt(int i, int j) {
    if (D == 2) return (i == j) ? i : 2;
    if (i == j) return i;
    return 6 - i - j;   // (1,2)->3, (0,2)->4, (0,1)->5
}

// ---------------------------------------------------------------------------------------------
// forward-mode dual number with NV directional derivatives (used only for d stress / d F)
// ---------------------------------------------------------------------------------------------
template <int NV> struct Dual {
    double v;
    double d[NV];
};
template <int NV> FEB_HD Dual<NV> dual_const(double c) {
    Dual<NV> r; r.v = c;
    for (int k = 0; k < NV; ++k) r.d[k] = 0.0;
    return r;
}
template <int NV> FEB_HD Dual<NV> operator+(const Dual<NV>& a, const Dual<NV>& b) {
    Dual<NV> r; r.v = a.v + b.v;
    for (int k = 0; k < NV; ++k) r.d[k] = a.d[k] + b.d[k];
    return r;
}
template <int NV> FEB_HD Dual<NV> operator-(const Dual<NV>& a, const Dual<NV>& b) {
    Dual<NV> r; r.v = a.v - b.v;
    for (int k = 0; k < NV; ++k) r.d[k] = a.d[k] - b.d[k];
    return r;
}
template <int NV> FEB_HD Dual<NV> operator*(const Dual<NV>& a, const Dual<NV>& b) {
    Dual<NV> r; r.v = a.v * b.v;
    for (int k = 0; k < NV; ++k) r.d[k] = a.d[k] * b.v + a.v * b.d[k];
    return r;
}
template <int NV> FEB_HD Dual<NV> operator/(const Dual<NV>& a, const Dual<NV>& b) {
    Dual<NV> r; const double ib = 1.0 / b.v; r.v = a.v * ib;
    for (int k = 0; k < NV; ++k) r.d[k] = (a.d[k] - r.v * b.d[k]) * ib;
    return r;
}
template <int NV> FEB_HD Dual<NV> operator+(const Dual<NV>& a, double b) { Dual<NV> r = a; r.v += b; return r; }
template <int NV> FEB_HD Dual<NV> operator-(const Dual<NV>& a, double b) { Dual<NV> r = a; r.v -= b; return r; }
template <int NV> FEB_HD Dual<NV> operator*(const Dual<NV>& a, double b) {
    Dual<NV> r; r.v = a.v * b;
    for (int k = 0; k < NV; ++k) r.d[k] = a.d[k] * b;
    return r;
}
template <int NV> FEB_HD Dual<NV> operator*(double b, const Dual<NV>& a) { return a * b; }
template <int NV> FEB_HD Dual<NV> operator/(double a, const Dual<NV>& b) { return dual_const<NV>(a) / b; }
template <int NV> FEB_HD Dual<NV> operator/(const Dual<NV>& a, double b) { return a * (1.0 / b); }

FEB_HD double feb_pow(double x, double p) { return pow(x, p); }
template <int NV> FEB_HD Dual<NV> feb_pow(const Dual<NV>& x, double p) {
    Dual<NV> r; r.v = pow(x.v, p);
    const double f = p * r.v / x.v;
    for (int k = 0; k < NV; ++k) r.d[k] = f * x.d[k];
    return r;
}
FEB_HD double feb_value(double x) { return x; }
template <int NV> FEB_HD double feb_value(const Dual<NV>& x) { return x.v; }

// ---------------------------------------------------------------------------------------------
// small dense helpers (templated on scalar so that they also run on Dual)
// ---------------------------------------------------------------------------------------------
template <typename T> FEB_HD T det2(const T (&A)[2][2]) { return A[0][0] * A[1][1] - A[0][1] * A[1][0]; }
template <typename T> FEB_HD T det3(const T (&A)[3][3]) {
    return A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
           A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
}
template <typename T> FEB_HD T det(const T (&A)[2][2]) { return det2(A); }
template <typename T> FEB_HD T det(const T (&A)[3][3]) { return det3(A); }

// B = inverse(A); returns det(A)
template <typename T> FEB_HD T inverse(const T (&A)[2][2], T (&B)[2][2]) {
    const T dt = det2(A);
    const T id = 1.0 / dt;
    B[0][0] = A[1][1] * id;
    B[0][1] = (A[0][1] * id) * (-1.0);
    B[1][0] = (A[1][0] * id) * (-1.0);
    B[1][1] = A[0][0] * id;
    return dt;
}
template <typename T> FEB_HD T inverse(const T (&A)[3][3], T (&B)[3][3]) {
    const T c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1];
    const T c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2];
    const T c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
    const T dt = A[0][0] * c00 + A[0][1] * c01 + A[0][2] * c02;
    const T id = 1.0 / dt;
    B[0][0] = c00 * id;
    B[1][0] = c01 * id;
    B[2][0] = c02 * id;
    B[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) * id;
    B[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * id;
    B[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) * id;
    B[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) * id;
    B[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) * id;
    B[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * id;
    return dt;
}

// ---------------------------------------------------------------------------------------------
// geometry at one quadrature point
//   J[p][d] = sum_n x[n][d] dphi[n][p] ; G[p][d] = inverse(J)[d][p] ; g[n][d] = sum_p G[p][d] dphi[n][p]
// x: N x D row-major, dphi_q: N x D (N x P) row-major.  Returns det J (signed).
// ---------------------------------------------------------------------------------------------
template <int N, int D>
FEB_HD double shape_gradients(const double* x, const double* dphi_q, double* g /* N*D */) {
    double J[D][D], Ji[D][D];
    for (int p = 0; p < D; ++p)
        for (int d = 0; d < D; ++d) J[p][d] = 0.0;
    for (int n = 0; n < N; ++n)
        for (int p = 0; p < D; ++p)
            for (int d = 0; d < D; ++d) J[p][d] += x[n * D + d] * dphi_q[n * D + p];
    const double dt = inverse(J, Ji);   // Ji[d][p] = d xi_p / d x_d
    for (int n = 0; n < N; ++n)
        for (int d = 0; d < D; ++d) {
            double s = 0.0;
            for (int p = 0; p < D; ++p) s += Ji[d][p] * dphi_q[n * D + p];
            g[n * D + d] = s;
        }
    return dt;
}
// runtime-N variant (quadratic elements handled by the generic kernels)
template <int D>
FEB_HD double shape_gradients_rt(int N, const double* x, const double* dphi_q, double* g) {
    double J[D][D], Ji[D][D];
    for (int p = 0; p < D; ++p)
        for (int d = 0; d < D; ++d) J[p][d] = 0.0;
    for (int n = 0; n < N; ++n)
        for (int p = 0; p < D; ++p)
            for (int d = 0; d < D; ++d) J[p][d] += x[n * D + d] * dphi_q[n * D + p];
    const double dt = inverse(J, Ji);
    for (int n = 0; n < N; ++n)
        for (int d = 0; d < D; ++d) {
            double s = 0.0;
            for (int p = 0; p < D; ++p) s += Ji[d][p] * dphi_q[n * D + p];
            g[n * D + d] = s;
        }
    return dt;
}

// ---------------------------------------------------------------------------------------------
// elasticity matrices in Voigt form (closed forms of inverse(compliance))
// ---------------------------------------------------------------------------------------------
template <int D> FEB_HD void zero_C(double (&C)[VoigtSize<D>::value][VoigtSize<D>::value]) {
    for (int i = 0; i < VoigtSize<D>::value; ++i)
        for (int j = 0; j < VoigtSize<D>::value; ++j) C[i][j] = 0.0;
}

// lambda* and mu of the isotropic law: 3-D Lame, 2-D plane stress
template <int D> FEB_HD void isotropic_lame(const double* mp, double& lam, double& mu) {
    const double E = mp[0], nu = mp[1];
    mu = 0.5 * E / (1.0 + nu);
    if (D == 3) lam = E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu));
    else lam = E * nu / (1.0 - nu * nu);
}

template <int D> FEB_HD void isotropic_C(const double* mp, double (&C)[VoigtSize<D>::value][VoigtSize<D>::value]) {
    double lam, mu;
    isotropic_lame<D>(mp, lam, mu);
    zero_C<D>(C);
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) C[i][j] = (i == j) ? lam + 2.0 * mu : lam;
    for (int s = D; s < VoigtSize<D>::value; ++s) C[s][s] = mu;
}

template <int D> FEB_HD void orthotropic_C(const double* mp, double (&C)[VoigtSize<D>::value][VoigtSize<D>::value]) {
    zero_C<D>(C);
    if (D == 2) {
        const double Exx = mp[0], Eyy = mp[1], nxy = mp[2], Gxy = mp[3];
        const double s11 = 1.0 / Exx, s22 = 1.0 / Eyy, s12 = -nxy / Exx;
        const double dt = s11 * s22 - s12 * s12;
        C[0][0] = s22 / dt; C[1][1] = s11 / dt; C[0][1] = C[1][0] = -s12 / dt;
        C[2][2] = Gxy;
    } else {
        const double Exx = mp[0], Eyy = mp[1], Ezz = mp[2], nxy = mp[3], nyz = mp[4], nxz = mp[5];
        const double Gxy = mp[6], Gyz = mp[7], Gxz = mp[8];
        double S[3][3], Si[3][3];
        S[0][0] = 1.0 / Exx; S[1][1] = 1.0 / Eyy; S[2][2] = 1.0 / Ezz;
        S[0][1] = S[1][0] = -nxy / Exx;
        S[0][2] = S[2][0] = -nxz / Exx;
        S[1][2] = S[2][1] = -nyz / Eyy;
        inverse(S, Si);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) C[i][j] = Si[i][j];
        C[3][3] = Gyz; C[4][4] = Gxz; C[5][5] = Gxy;
    }
}

// sigma = C : sym(H) with engineering shear; works on double or Dual
template <int D, typename T>
FEB_HD void stress_from_C(const double (&C)[VoigtSize<D>::value][VoigtSize<D>::value], const T (&eps)[D][D],
                          T (&sig)[D][D]) {
    constexpr int S = VoigtSize<D>::value;
    T ev[S], sv[S];
    for (int i = 0; i < D; ++i) ev[i] = eps[i][i];
    if (D == 2) ev[2] = eps[0][1] * 2.0;
    if (D == 3) { ev[3] = eps[1][2] * 2.0; ev[4] = eps[0][2] * 2.0; ev[5] = eps[0][1] * 2.0; }
    for (int s = 0; s < S; ++s) {
        T acc = ev[0] * C[s][0];
        for (int t = 1; t < S; ++t) acc = acc + ev[t] * C[s][t];
        sv[s] = acc;
    }
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) sig[i][j] = sv[voigt<D>(i, j)];
}

// isotropic C applied in closed form: S = lam tr(e) I + 2 mu e  (the same numbers as isotropic_C + stress_from_C: C[i][i] =
// lam + 2 mu, C[i][j] = lam, shear mu on the doubled shear strains; plane stress in 2-D through isotropic_lame<2>)
template <int D, typename T>
FEB_HD void isotropic_stress(const double* mp, const T (&e)[D][D], T (&S)[D][D]) {
    double lam, mu;
    isotropic_lame<D>(mp, lam, mu);
    T tr = e[0][0];
    for (int i = 1; i < D; ++i) tr = tr + e[i][i];
    const T ltr = tr * lam;
    const double mu2 = 2.0 * mu;
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) {
            S[i][j] = e[i][j] * mu2;
            if (i == j) S[i][j] = S[i][j] + ltr;
        }
}

// ---------------------------------------------------------------------------------------------
// hyperelastic "stress" S(F) exactly as the reference returns it (used as is in the residual)
// ---------------------------------------------------------------------------------------------
template <int D, typename T>
FEB_HD void svk_stress(const double* mp, const T (&F)[D][D], T (&S)[D][D]) {
    T egl[D][D];
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) {
            T acc = F[0][i] * F[0][j];
            for (int d = 1; d < D; ++d) acc = acc + F[d][i] * F[d][j];
            egl[i][j] = (acc - (i == j ? 1.0 : 0.0)) * 0.5;
        }
    isotropic_stress<D, T>(mp, egl, S);
}

template <int D, typename T>
FEB_HD void mooney_rivlin_stress(const double* mp, const T (&F)[D][D], T (&S)[D][D]) {
    const double C1 = mp[0], C2 = mp[1], D1 = mp[2];
    const T J = det(F);
    const T a = feb_pow(J, -1.0 / 3.0);
    T Fb[D][D], Bb[D][D], BF[D][D], Fi[D][D];
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) Fb[i][j] = a * F[i][j];
    for (int j = 0; j < D; ++j)
        for (int k = 0; k < D; ++k) {
            T acc = Fb[0][j] * Fb[0][k];
            for (int d = 1; d < D; ++d) acc = acc + Fb[d][j] * Fb[d][k];
            Bb[j][k] = acc;
        }
    T I1 = Bb[0][0];
    for (int d = 1; d < D; ++d) I1 = I1 + Bb[d][d];
    T bb = Bb[0][0] * Bb[0][0];
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j)
            if (i + j > 0) bb = bb + Bb[i][j] * Bb[j][i];
    const T I2 = (I1 * I1 - bb) * 0.5;
    for (int j = 0; j < D; ++j)
        for (int k = 0; k < D; ++k) {
            T acc = Bb[j][0] * Fb[0][k];
            for (int d = 1; d < D; ++d) acc = acc + Bb[j][d] * Fb[d][k];
            BF[j][k] = acc;
        }
    inverse(Fb, Fi);
    const T c0 = (J - 1.0) / (a * a) * (2.0 / D1) - a * (I1 * C1 + I2 * (2.0 * C2)) * (2.0 / 3.0);
    const T c1 = a * (I1 * C2 + C1) * 2.0;
    const T c2 = a * (2.0 * C2);
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) S[i][j] = c0 * Fi[i][j] + c1 * Fb[i][j] + c2 * BF[i][j];
}

// ---------------------------------------------------------------------------------------------
// stress at a quadrature point from the displacement gradient H[i][d] = d u_i / d x_d
// (linear kinds) or from F[i][d] = d (u+x)_i / d X_d (hyper kinds)
// ---------------------------------------------------------------------------------------------
template <int D, typename T>
FEB_HD void point_stress(int material, const double* mp, const T (&H)[D][D], T (&S)[D][D]) {
    if (material == MAT_ISOTROPIC) {
        T eps[D][D];
        for (int i = 0; i < D; ++i)
            for (int j = 0; j < D; ++j) eps[i][j] = (H[i][j] + H[j][i]) * 0.5;
        isotropic_stress<D, T>(mp, eps, S);
        return;
    }
    if (material == MAT_ISOTROPIC || material == MAT_ORTHOTROPIC) {
        double C[VoigtSize<D>::value][VoigtSize<D>::value];
        if (material == MAT_ISOTROPIC) isotropic_C<D>(mp, C);
        else orthotropic_C<D>(mp, C);
        T eps[D][D];
        for (int i = 0; i < D; ++i)
            for (int j = 0; j < D; ++j) eps[i][j] = (H[i][j] + H[j][i]) * 0.5;
        stress_from_C<D, T>(C, eps, S);
    } else if (material == MAT_SVK) {
        svk_stress<D, T>(mp, H, S);
    } else {
        mooney_rivlin_stress<D, T>(mp, H, S);
    }
}

// A[(i*D+d)*D*D + k*D+l] = d S[i][d] / d H[k][l]
template <int D>
FEB_HD void point_tangent(int material, const double* mp, const double (&H)[D][D], double* A) {
    if (material == MAT_ISOTROPIC || material == MAT_ORTHOTROPIC) {
        double C[VoigtSize<D>::value][VoigtSize<D>::value];
        if (material == MAT_ISOTROPIC) isotropic_C<D>(mp, C);
        else orthotropic_C<D>(mp, C);
        for (int i = 0; i < D; ++i)
            for (int d = 0; d < D; ++d)
                for (int k = 0; k < D; ++k)
                    for (int l = 0; l < D; ++l)
                        A[(i * D + d) * D * D + k * D + l] = C[voigt<D>(i, d)][voigt<D>(k, l)];
        return;
    }
    if (material == MAT_SVK) {
        // closed form of jacfwd(st_venant_kirchhoff) (hyperelasticity.py:9-53): S = C : E, E = (F^T F - I) / 2 with
        // the engineering-shear Voigt convention of stress_from_C  =>  dS[i][d]/dF[k][l] = sum_n C[v(i,d)][v(l,n)] F[k][n]
        // (the dual-number route below gives the same numbers at ~6x the arithmetic)
        double C[VoigtSize<D>::value][VoigtSize<D>::value];
        isotropic_C<D>(mp, C);
        for (int i = 0; i < D; ++i)
            for (int d = 0; d < D; ++d) {
                const int r = voigt<D>(i, d);
                for (int k = 0; k < D; ++k)
                    for (int l = 0; l < D; ++l) {
                        double acc = C[r][voigt<D>(l, 0)] * H[k][0];
                        for (int n = 1; n < D; ++n) acc += C[r][voigt<D>(l, n)] * H[k][n];
                        A[(i * D + d) * D * D + k * D + l] = acc;
                    }
            }
        return;
    }
    constexpr int NV = D * D;
    Dual<NV> Hd[D][D], Sd[D][D];
    for (int k = 0; k < D; ++k)
        for (int l = 0; l < D; ++l) {
            Hd[k][l] = dual_const<NV>(H[k][l]);
            Hd[k][l].d[k * D + l] = 1.0;
        }
    point_stress<D, Dual<NV>>(material, mp, Hd, Sd);
    for (int i = 0; i < D; ++i)
        for (int d = 0; d < D; ++d)
            for (int c = 0; c < NV; ++c) A[(i * D + d) * NV + c] = Sd[i][d].d[c];
}

// ---------------------------------------------------------------------------------------------
// whole-element residual, one thread per element (runtime N, Q; D compile time)
//   R[n][d] = sum_q detJ_q W_q sum_i g[q][n][i] S_q[i][d]
// mp_q(q) returns the material parameters of quadrature point q.
// g_scratch: N*D doubles of scratch.
// ---------------------------------------------------------------------------------------------
template <int D>
FEB_HD void element_residual_rt(int kind, int material, int N, int Q, const double* x, const double* u,
                                const double* dphi, const double* W, const double* mp, int mp_q_stride,
                                double* g_scratch, double* R) {
    for (int i = 0; i < N * D; ++i) R[i] = 0.0;
    for (int q = 0; q < Q; ++q) {
        const double dt = shape_gradients_rt<D>(N, x, dphi + (long)q * N * D, g_scratch);
        double H[D][D], S[D][D];
        for (int i = 0; i < D; ++i)
            for (int d = 0; d < D; ++d) H[i][d] = 0.0;
        for (int n = 0; n < N; ++n)
            for (int i = 0; i < D; ++i) {
                const double ui = (kind == KIND_HYPER) ? u[n * D + i] + x[n * D + i] : u[n * D + i];
                for (int d = 0; d < D; ++d) H[i][d] += g_scratch[n * D + d] * ui;
            }
        point_stress<D, double>(material, mp + (long)q * mp_q_stride, H, S);
        const double w = dt * W[q];
        for (int n = 0; n < N; ++n)
            for (int d = 0; d < D; ++d) {
                double s = 0.0;
                for (int i = 0; i < D; ++i) s += g_scratch[n * D + i] * S[i][d];
                R[n * D + d] += w * s;
            }
    }
}

// whole-element tangent (host check / generic path): K[(a*D+d)*N*D + b*D+k]
template <int D>
FEB_HD void element_tangent_rt(int kind, int material, int N, int Q, const double* x, const double* u,
                               const double* dphi, const double* W, const double* mp, int mp_q_stride,
                               double* g_scratch, double* K) {
    const int NU = N * D;
    for (int i = 0; i < NU * NU; ++i) K[i] = 0.0;
    double A[D * D * D * D];
    for (int q = 0; q < Q; ++q) {
        const double dt = shape_gradients_rt<D>(N, x, dphi + (long)q * N * D, g_scratch);
        double H[D][D];
        for (int i = 0; i < D; ++i)
            for (int d = 0; d < D; ++d) H[i][d] = 0.0;
        for (int n = 0; n < N; ++n)
            for (int i = 0; i < D; ++i) {
                const double ui = (kind == KIND_HYPER) ? u[n * D + i] + x[n * D + i] : u[n * D + i];
                for (int d = 0; d < D; ++d) H[i][d] += g_scratch[n * D + d] * ui;
            }
        point_tangent<D>(material, mp + (long)q * mp_q_stride, H, A);
        const double w = dt * W[q];
        for (int a = 0; a < N; ++a)
            for (int d = 0; d < D; ++d)
                for (int k = 0; k < D; ++k)
                    for (int l = 0; l < D; ++l) {
                        double t = 0.0;
                        for (int i = 0; i < D; ++i) t += g_scratch[a * D + i] * A[(i * D + d) * D * D + k * D + l];
                        t *= w;
                        for (int b = 0; b < N; ++b) K[(a * D + d) * NU + b * D + k] += t * g_scratch[b * D + l];
                    }
    }
}

}  // namespace feb
