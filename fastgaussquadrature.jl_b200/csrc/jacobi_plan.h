// jacobi_plan.h — host-planned, parameter-only data of the Gauss-Jacobi path (shared between the
// host planner jacobi_boundary.cu and the kernels in jacobi.cu).
#ifndef FGQ_JACOBI_PLAN_H_
#define FGQ_JACOBI_PLAN_H_
#include <stdint.h>

namespace fgq {

// what the device series for J_nu needs, planned on the host (one order per entry)
struct BesselOrder {
    double nu;        // effective order (negative-integer reflection already applied)
    double sign;      // J_{-m} = (-1)^m J_m for integer m: multiply the result by this
    double rgamma;    // 1 / Gamma(nu + 1)
    double cosc, sinc;  // cos, sin of (nu/2 + 1/4) pi (Hankel branch only)
};

constexpr int JAC_M = 20;     // terms of the interior expansion (gaussjacobi.jl:256)
constexpr int JAC_NBDY = 10;  // boundary nodes per end (gaussjacobi.jl:174)

// One half of the interval: (a, b) = (alpha, beta) for x > 0, swapped for x < 0 (gaussjacobi.jl:225).
struct JacobiHalfPlan {
    double a, b;
    double C, C2;                 // gaussjacobi.jl:319-350
    double PHI[JAC_M][JAC_M];     // :274-283, [l][m]
    double PHI2[JAC_M][JAC_M];    // :285-294
};

// One boundary solve (asy2, gaussjacobi.jl:363-398) with parameters (al, be).
struct JacobiBoundaryPlan {
    int64_t n;
    double al, be;
    double rho, rho2;      // n + (al+be+1)/2, n + (al+be-1)/2
    double C, C2;          // :443-462
    double jk[JAC_NBDY];   // approx_besselroots(al, 10)
    BesselOrder ord[64];   // orders al + (o - 30), o = 0..61
    double wc;             // weightsConstant of the rule
    int side;              // 0: right end, 1: left end (mirrored)
    int64_t lo, hi;        // output slice
};
void jacobi_boundary_plan(int64_t n, double al, double be, JacobiBoundaryPlan* P);

void jacobi_interior_plan(int64_t n, double a, double b, JacobiHalfPlan* hp);
double jacobi_weights_constant(int64_t n, double a, double b);
void approx_besselroots(double nu, int n, double* x);
double besselj_real(double nu, double x);

}  // namespace fgq
#endif
