// jacobi_plan.h — host-planned, parameter-only data of the Gauss-Jacobi path (shared between the
// host planner jacobi_boundary.cu and the kernels in jacobi.cu).
#ifndef FGQ_JACOBI_PLAN_H_
#define FGQ_JACOBI_PLAN_H_
#include <stdint.h>

namespace fgq {

constexpr int JAC_M = 20;     // terms of the interior expansion (gaussjacobi.jl:256)
constexpr int JAC_NBDY = 10;  // boundary nodes per end (gaussjacobi.jl:174)

// One half of the interval: (a, b) = (alpha, beta) for x > 0, swapped for x < 0 (gaussjacobi.jl:225).
struct JacobiHalfPlan {
    double a, b;
    double C, C2;                 // gaussjacobi.jl:319-350
    double PHI[JAC_M][JAC_M];     // :274-283, [l][m]
    double PHI2[JAC_M][JAC_M];    // :285-294
};

void jacobi_interior_plan(int64_t n, double a, double b, JacobiHalfPlan* hp);
double jacobi_weights_constant(int64_t n, double a, double b);
void jacobi_asy2(int64_t n, double al, double be, int npts, double* x, double* w);
void approx_besselroots(double nu, int n, double* x);
double besselj_real(double nu, double x);

}  // namespace fgq
#endif
