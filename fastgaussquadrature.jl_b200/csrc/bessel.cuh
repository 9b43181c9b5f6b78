// bessel.cuh — J_nu(x) of real order on the device, for the hard-edge (Bessel) region of the
// Gauss-Laguerre kernels (the reference calls SpecialFunctions.besselj -> AMOS zbesj at
// src/gausslaguerre.jl:484,501).
//
//   x < 25 : ascending series  (x/2)^nu / Gamma(nu+1) * sum_k (-x^2/4)^k / (k! (nu+1)_k), the
//            alternating sum in double-double (cancellation reaches e^x), prefactor in double;
//   x >= 25: Hankel's expansion (DLMF 10.17.3), |nu| < 5: the terms fall below 1e-17 before they
//            start to grow; cos/sin of the phase x - (nu/2 + 1/4) pi by the angle-difference
//            formulas so that the large argument x is never rounded against the constant.
// 1/Gamma(nu+1), cos and sin of (nu/2 + 1/4) pi are planned on the host (one order per launch).
#ifndef FGQ_BESSEL_CUH_
#define FGQ_BESSEL_CUH_

#include "jacobi_plan.h"

namespace fgq {

struct ddv {
    double hi, lo;
};
__host__ __device__ __forceinline__ ddv dd_two_sum(double a, double b) {
    const double s = a + b, bb = s - a;
    return {s, (a - (s - bb)) + (b - bb)};
}
__host__ __device__ __forceinline__ ddv dd_two_prod(double a, double b) {
    const double p = a * b;
    return {p, fma(a, b, -p)};
}
__host__ __device__ __forceinline__ ddv dd_add(ddv a, ddv b) {
    const ddv s = dd_two_sum(a.hi, b.hi);
    return dd_two_sum(s.hi, s.lo + a.lo + b.lo);
}
__host__ __device__ __forceinline__ ddv dd_mul(ddv a, ddv b) {
    const ddv p = dd_two_prod(a.hi, b.hi);
    return dd_two_sum(p.hi, p.lo + (a.hi * b.lo + a.lo * b.hi));
}
__host__ __device__ __forceinline__ ddv dd_div(ddv a, ddv b) {
    const double q1 = a.hi / b.hi;
    ddv r = dd_add(a, dd_mul({-q1, 0.0}, b));
    const double q2 = r.hi / b.hi;
    r = dd_add(r, dd_mul({-q2, 0.0}, b));
    const double q3 = r.hi / b.hi;
    return dd_add(dd_two_sum(q1, q2), {q3, 0.0});
}

// ascending series in double-double: any real order, x up to ~45 (cancellation e^x against 1e-32)
__device__ inline double besselj_series_dd(const BesselOrder& o, double x) {
    const double nu = o.nu;
    if (x == 0.0) return nu == 0.0 ? 1.0 : 0.0;
    const ddv q = dd_mul(dd_two_prod(x, x), {0.25, 0.0});
    ddv term = {1.0, 0.0}, sum = {1.0, 0.0};
    double maxterm = 1.0;
    for (int k = 0; k < 400; ++k) {
        const ddv den = dd_mul({(double)(k + 1), 0.0}, dd_two_sum(nu, (double)(k + 1)));
        term = dd_div(dd_mul(term, q), den);
        term.hi = -term.hi;
        term.lo = -term.lo;
        sum = dd_add(sum, term);
        maxterm = fmax(maxterm, fabs(term.hi));
        if (fabs(term.hi) < 1e-36 * maxterm && (double)k > 0.5 * x) break;
    }
    return o.sign * (pow(0.5 * x, nu) * o.rgamma) * sum.hi;
}

// The same series with the reciprocals 1 / ((k+1)(nu+k+1)) of its term ratios tabulated (they depend on the order
// only: BESSEL_INV_DEN entries planned on the host with these very double-double routines, besselj_inv_den_table).
// The three dependent divisions of dd_div leave the term recurrence, whose dependent chain is then two double-double
// products: ~7x shorter, and this chain is the latency of the first Bessel-region nodes of a Laguerre rule.
constexpr int BESSEL_INV_DEN = 96;

__host__ __device__ inline ddv besselj_term_den(double nu, int k) {
    return dd_mul({(double)(k + 1), 0.0}, dd_two_sum(nu, (double)(k + 1)));
}

inline void besselj_inv_den_table(double nu, ddv* inv_den) {
    for (int k = 0; k < BESSEL_INV_DEN; ++k) inv_den[k] = dd_div({1.0, 0.0}, besselj_term_den(nu, k));
}

// stride: distance between consecutive k of this order's reciprocals (1: a table of its own; 64: the interleaved
// per-order tables of the Jacobi boundary block, conflict-free across the lanes of a warp)
__device__ inline double besselj_series_dd_tab(const BesselOrder& o, double x, const ddv* __restrict__ inv_den,
                                               int stride = 1) {
    const double nu = o.nu;
    if (x == 0.0) return nu == 0.0 ? 1.0 : 0.0;
    const ddv q = dd_mul(dd_two_prod(x, x), {0.25, 0.0});
    ddv term = {1.0, 0.0}, sum = {1.0, 0.0};
    double maxterm = 1.0;
    for (int k = 0; k < 400; ++k) {
        if (k < BESSEL_INV_DEN) {
            term = dd_mul(dd_mul(term, q), inv_den[k * stride]);
        } else {
            term = dd_div(dd_mul(term, q), besselj_term_den(nu, k));
        }
        term.hi = -term.hi;
        term.lo = -term.lo;
        sum = dd_add(sum, term);
        maxterm = fmax(maxterm, fabs(term.hi));
        if (fabs(term.hi) < 1e-36 * maxterm && (double)k > 0.5 * x) break;
    }
    return o.sign * (pow(0.5 * x, nu) * o.rgamma) * sum.hi;
}

// inv_den: besselj_inv_den_table of the order, or null (plain series)
__device__ inline double besselj_dev(const BesselOrder& o, double x, const ddv* __restrict__ inv_den = nullptr) {
    const double nu = o.nu;
    if (x < 25.0) return inv_den ? besselj_series_dd_tab(o, x, inv_den) : besselj_series_dd(o, x);
    const double mu = 4.0 * nu * nu;
    const double i8x = 1.0 / (8.0 * x);
    double a = 1.0, P = 1.0, Q = 0.0, prev = 1.0;
    for (int k = 1; k <= 60; ++k) {
        const double odd = (double)(2 * k - 1);
        a = a * (mu - odd * odd) * i8x / (double)k;
        if (fabs(a) > prev) break;  // asymptotic series: stop at the smallest term
        prev = fabs(a);
        switch (k & 3) {
            case 1: Q += a; break;
            case 2: P -= a; break;
            case 3: Q -= a; break;
            default: P += a; break;
        }
        if (prev < 1e-18) break;
    }
    double sx, cx;
    sincos(x, &sx, &cx);
    const double cchi = cx * o.cosc + sx * o.sinc;  // cos(x - c)
    const double schi = sx * o.cosc - cx * o.sinc;  // sin(x - c)
    return o.sign * sqrt(0.63661977236758134308 / x) * (P * cchi - Q * schi);
}

}  // namespace fgq
#endif
