/* laguerre_rec_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * Restatement in C of the forward-recurrence pieces of src/gausslaguerre.jl that the asymptotic
 * path uses as its `recompute` hook:
 *   evalLaguerreRec   src/gausslaguerre.jl:625-641
 *   gl_rec_newton     src/gausslaguerre.jl:547-579
 * (the rest of the Laguerre path is restated in oracle/laguerre_oracle.py).  The only shortcut:
 * the O(n) recurrence stops once pn AND pnprev are NaN — from there every later value is NaN
 * (NaN is absorbing in a*pn - b*pnprev), so the returned values are unchanged. */
#include <float.h>
#include <math.h>
#include <stdint.h>

static void eval_laguerre_rec(int64_t n, double alpha, double x, double* out_pn, double* out_pnd) {
    double pnprev = 0.0;
    double pn = 1.0 / sqrt(tgamma(alpha + 1.0));
    double pndprev = 0.0, pnd = 0.0;
    for (int64_t k = 1; k <= n; ++k) {
        const double dk = (double)k;
        const double pnold = pn;
        pn = (x - 2 * dk - alpha + 1) / sqrt(dk * (alpha + dk)) * pnold -
             sqrt((dk - 1 + alpha) * (dk - 1) / dk / (dk + alpha)) * pnprev;
        pnprev = pnold;
        const double pndold = pnd;
        pnd = (pnold + (x - 2 * dk - alpha + 1) * pndold) / sqrt(dk * (alpha + dk)) -
              sqrt((dk - 1 + alpha) * (dk - 1) / dk / (alpha + dk)) * pndprev;
        pndprev = pndold;
        if (pn != pn && pnprev != pnprev && pnd != pnd && pndprev != pndprev) break;
    }
    *out_pn = pn;
    *out_pnd = pnd;
}

void oracle_eval_laguerre_rec(int64_t n, double alpha, double x, double* pn, double* pnd) {
    eval_laguerre_rec(n, alpha, x, pn, pnd);
}

/* gl_rec_newton(x0, n, alpha; maxiter = 20, computeweight) -> (xk, wk) */
void oracle_gl_rec_newton2(double x0, int64_t n, double alpha, int computeweight, double* xk_out, double* wk_out);
void oracle_gl_rec_newton(double x0, int64_t n, double alpha, double* xk_out, double* wk_out) {
    oracle_gl_rec_newton2(x0, n, alpha, 1, xk_out, wk_out);
}
void oracle_gl_rec_newton2(double x0, int64_t n, double alpha, int computeweight, double* xk_out, double* wk_out) {
    const double eps = DBL_EPSILON;
    double step = x0;
    int iter = 0;
    const int maxiter = 20;
    double xk = x0, xk_prev = xk, pn_prev = DBL_MAX, pn_deriv = 0.0;
    while (fabs(step) > 40 * eps * xk && iter < maxiter) {
        iter += 1;
        double pn;
        eval_laguerre_rec(n, alpha, xk, &pn, &pn_deriv);
        if (fabs(pn) >= fabs(pn_prev) * (1 - 50 * eps)) {
            xk = xk_prev;
            break;
        }
        step = pn / pn_deriv;
        xk_prev = xk;
        xk -= step;
        pn_prev = pn;
    }
    *xk_out = xk;
    *wk_out = 0.0;
    if (computeweight) {
        double pn_min1, dummy;
        eval_laguerre_rec(n - 1, alpha, xk, &pn_min1, &dummy);
        const double nn = (double)n;
        *wk_out = pow(nn * nn + alpha * nn, -0.5) / pn_min1 / pn_deriv;
    }
}
