/* jacobi_rec_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * C restatement of innerjacobi_rec! (src/gaussjacobi.jl:129-155) with muladd -> fma(), because the
 * recurrence is ill-conditioned enough (large alpha) that fused vs unfused evaluation moves P'_n by
 * hundreds of ulps; the rest of jacobi_rec is restated in oracle/jacobi_oracle.py. */
#include <math.h>
#include <stdint.h>

void oracle_innerjacobi_rec(int64_t n, int64_t N, const double* x, double a, double b, double* P, double* PP) {
    for (int64_t j = 0; j < N; ++j) {
        const double xj = x[j];
        double Pj = (a - b + (a + b + 2) * xj) / 2;
        double Pm1 = 1.0;
        double PPj = (a + b + 2) / 2;
        double PPm1 = 0.0;
        for (int64_t k = 1; k <= n - 1; ++k) {
            const double dk = (double)k;
            const double k0 = fma(2.0, dk, a + b);
            const double k1 = k0 + 1;
            const double k2 = k0 + 2;
            const double A = 2 * (dk + 1) * (dk + (a + b + 1)) * k0;
            const double B = k1 * (a * a - b * b);
            const double C = k0 * k1 * k2;
            const double D = 2 * (dk + a) * (dk + b) * k2;
            const double c1 = fma(C, xj, B);
            const double newP = fma(-D, Pm1, c1 * Pj) / A;
            Pm1 = Pj;
            Pj = newP;
            const double newPP = fma(c1, PPj, fma(-D, PPm1, C * Pm1)) / A;
            PPm1 = PPj;
            PPj = newPP;
        }
        P[j] = Pj;
        PP[j] = PPj;
    }
}
