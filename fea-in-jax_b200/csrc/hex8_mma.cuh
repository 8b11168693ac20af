// Shared device helpers of the fused hex8 assembly kernels (fused_mma.cu, fused_sweep.cu): the FP64 tensor-core
// instruction and the shape-function gradients of one quadrature point with 16-byte shared-memory loads.
// Geometry follows linear_elasticity.py:180-184 of the reference (same arithmetic as shape_gradients<8,3>).
#pragma once

#include "element_math.cuh"

namespace feb {

// C (8x8, lane holds row lane/4, columns 2*(lane%4), +1) += A (8x4, lane holds [lane/4][lane%4]) * B (4x8, lane holds [lane%4][lane/4])
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double bb) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(bb));
}

// Gradients of the 8 shape functions at one quadrature point: xe = 24 coordinates [n][3] (16-byte aligned),
// dq = the point's basis table [n][3] (16-byte aligned); g[n*3+d] = d phi_n / d x_d; returns det J.
__device__ __forceinline__ double hex8_gradients(const double* __restrict__ xe, const double* __restrict__ dq, double (&g)[24]) {
    const double2* x2 = reinterpret_cast<const double2*>(xe);
    const double2* d2 = reinterpret_cast<const double2*>(dq);
    double d[24];
    double J[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
#pragma unroll
    for (int h = 0; h < 4; ++h) {                            // two nodes = three double2 per step
        const double2 xa = x2[3 * h], xb = x2[3 * h + 1], xc = x2[3 * h + 2];
        const double2 da = d2[3 * h], db = d2[3 * h + 1], dc = d2[3 * h + 2];
        const double x[6] = {xa.x, xa.y, xb.x, xb.y, xc.x, xc.y};
        d[6 * h] = da.x; d[6 * h + 1] = da.y; d[6 * h + 2] = db.x; d[6 * h + 3] = db.y; d[6 * h + 4] = dc.x; d[6 * h + 5] = dc.y;
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int pp = 0; pp < 3; ++pp)
#pragma unroll
                for (int dd = 0; dd < 3; ++dd) J[pp][dd] += x[3 * m + dd] * d[6 * h + 3 * m + pp];
    }
    double Ji[3][3];
    const double dt = inverse(J, Ji);
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int dd = 0; dd < 3; ++dd) {
            double sacc = 0.0;
#pragma unroll
            for (int pp = 0; pp < 3; ++pp) sacc += Ji[dd][pp] * d[n * 3 + pp];
            g[n * 3 + dd] = sacc;
        }
    return dt;
}

// The same with the point's basis table held in registers (dq[n*3+p]).
__device__ __forceinline__ double hex8_gradients_rd(const double* __restrict__ xe, const double (&dq)[24], double (&g)[24]) {
    const double2* x2 = reinterpret_cast<const double2*>(xe);
    double J[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        const double2 xa = x2[3 * h], xb = x2[3 * h + 1], xc = x2[3 * h + 2];
        const double x[6] = {xa.x, xa.y, xb.x, xb.y, xc.x, xc.y};
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int pp = 0; pp < 3; ++pp)
#pragma unroll
                for (int dd = 0; dd < 3; ++dd) J[pp][dd] += x[3 * m + dd] * dq[6 * h + 3 * m + pp];
    }
    double Ji[3][3];
    const double dt = inverse(J, Ji);
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int dd = 0; dd < 3; ++dd) {
            double sacc = 0.0;
#pragma unroll
            for (int pp = 0; pp < 3; ++pp) sacc += Ji[dd][pp] * dq[n * 3 + pp];
            g[n * 3 + dd] = sacc;
        }
    return dt;
}

}  // namespace feb
