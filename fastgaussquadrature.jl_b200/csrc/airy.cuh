// airy.cuh — Ai(z), Ai'(z) on the negative real axis for the Hermite and Laguerre kernels
// (the reference calls SpecialFunctions.airyai/airyaiprime -> AMOS zairy at
// src/gausshermite.jl:182-183 and src/gausslaguerre.jl:525).
//
//   -11.5 <= z <= 0.5 : Taylor expansion about the nearest integer c, coefficients from the ODE
//                       Ai'' = z Ai:  a_{j+2} = (c a_j + a_{j-1}) / ((j+2)(j+1)), anchored at
//                       (Ai(c), Ai'(c)) from special_tables.h.  |h| <= 0.5, 36 terms: error
//                       < 3e-16 of the local amplitude (AMOS itself: 4e-14 here).
//   z < -11.5         : DLMF 9.7.9-10, zeta = 2/3 |z|^{3/2} >= 26, 12 terms of each series; the
//                       error is the rounding of zeta (zeta * 1e-16), the same as AMOS's.
#ifndef FGQ_AIRY_CUH_
#define FGQ_AIRY_CUH_

#include "special_tables.h"

namespace fgq {

__constant__ double c_airy_anchor[FGQ_AIRY_NCENTERS][2] = FGQ_AIRY_ANCHORS;
__constant__ double c_airy_u[FGQ_AIRY_NUV] = FGQ_AIRY_U;
__constant__ double c_airy_v[FGQ_AIRY_NUV] = FGQ_AIRY_V;

__device__ __forceinline__ void airy_neg(double z, double& ai, double& aip) {
    if (z >= -11.5) {
        int i = (int)floor(0.5 - z);  // nearest integer to -z
        i = i < 0 ? 0 : (i > FGQ_AIRY_NCENTERS - 1 ? FGQ_AIRY_NCENTERS - 1 : i);
        const double c = -(double)i;
        const double h = z - c;
        constexpr int J = 36;
        double a[J + 1];
        a[0] = c_airy_anchor[i][0];
        a[1] = c_airy_anchor[i][1];
        a[2] = c * a[0] * 0.5;
        // Rolled on purpose: unrolled, the 37 coefficients live in 74 registers and set the register count (hence the
        // occupancy) of every kernel that inlines this function, for a branch that a few dozen nodes per rule take;
        // rolled, a[] is thread-local memory.  Same operations in the same order.
#pragma unroll 1
        for (int j = 1; j + 2 <= J; ++j) a[j + 2] = fma(c, a[j], a[j - 1]) / ((double)(j + 2) * (double)(j + 1));
        double val = 0.0, der = 0.0;
#pragma unroll 1
        for (int j = J; j >= 0; --j) {
            der = fma(der, h, val);
            val = fma(val, h, a[j]);
        }
        ai = val;
        aip = der;
        return;
    }
    const double x = -z;
    const double sx = sqrt(x);
    const double zeta = (2.0 / 3.0) * x * sx;
    const double iz = 1.0 / zeta;
    const double iz2 = iz * iz;
    constexpr int KK = 12;  // pairs
    double P = 0.0, Q = 0.0, R = 0.0, S = 0.0;
#pragma unroll
    for (int k = KK - 1; k >= 0; --k) {
        const double sg = (k & 1) ? -1.0 : 1.0;
        P = fma(P, iz2, sg * c_airy_u[2 * k]);
        Q = fma(Q, iz2, sg * c_airy_u[2 * k + 1]);
        R = fma(R, iz2, sg * c_airy_v[2 * k]);
        S = fma(S, iz2, sg * c_airy_v[2 * k + 1]);
    }
    Q *= iz;
    S *= iz;
    double s, c;
    sincos(zeta, &s, &c);
    const double r2 = 0.70710678118654752440;
    const double sp = (s + c) * r2;  // sin(zeta + pi/4)
    const double cp = (c - s) * r2;  // cos(zeta + pi/4)
    const double x14 = sqrt(sx);
    const double rsp = 0.56418958354775628695;  // 1/sqrt(pi)
    ai = (sp * P - cp * Q) * rsp / x14;
    aip = -(cp * R + sp * S) * rsp * x14;
}

}  // namespace fgq
#endif
