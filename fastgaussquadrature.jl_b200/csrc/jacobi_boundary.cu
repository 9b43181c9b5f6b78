// jacobi_boundary.cu — HOST planner pieces of the Gauss-Jacobi path: everything that is O(1) in n.
//
//   approx_besselroots / McMahon / piessens        src/besselroots.jl:23-165
//   constants of feval_asy2                        src/gaussjacobi.jl:443-462
//   weightsConstant                                src/gaussjacobi.jl:157-168
//   PHI/PHI2 tables and the constants C, C2 of feval_asy1   src/gaussjacobi.jl:274-294, 319-350
//
// Only parameter-only data is planned here (Bessel-zero guesses, Stirling-series constants, the
// PHI tables, 1/Gamma of the Bessel orders); the boundary solve itself runs on the device
// (jacobi_boundary.cuh), the O(n) interior in jacobi.cu.  besselj_real is the host twin of the device
// series, exported for tests (fgq_besselj).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "jacobi_plan.h"
#include "jacobi_tables.h"

namespace fgq {

static const double EPS = 2.220446049250313e-16;
static const double PI = 3.141592653589793;

// ---- double-double helpers (host only) ---------------------------------------------------------
struct dd {
    double hi, lo;
};
static inline dd two_sum(double a, double b) {
    double s = a + b, bb = s - a;
    return {s, (a - (s - bb)) + (b - bb)};
}
static inline dd two_prod(double a, double b) {
    double p = a * b;
    return {p, fma(a, b, -p)};
}
static inline dd dd_add(dd a, dd b) {
    dd s = two_sum(a.hi, b.hi);
    double lo = s.lo + a.lo + b.lo;
    dd r = two_sum(s.hi, lo);
    return r;
}
static inline dd dd_mul(dd a, dd b) {
    dd p = two_prod(a.hi, b.hi);
    double lo = p.lo + (a.hi * b.lo + a.lo * b.hi);
    return two_sum(p.hi, lo);
}
static inline dd dd_div(dd a, dd b) {
    double q1 = a.hi / b.hi;
    dd r = dd_add(a, dd_mul({-q1, 0.0}, b));
    double q2 = r.hi / b.hi;
    r = dd_add(r, dd_mul({-q2, 0.0}, b));
    double q3 = r.hi / b.hi;
    dd q = two_sum(q1, q2);
    return dd_add(q, {q3, 0.0});
}

// J_nu(x), real order (any sign), x > 0 moderate (<= ~60):
//   J_nu(x) = (x/2)^nu / Gamma(nu+1) * sum_k (-1)^k (x^2/4)^k / (k! (nu+1)_k)
// The alternating sum loses up to ~1e15 to cancellation at x ~ 35, hence double-double; the
// prefactor is a plain relative factor and stays in double.
double besselj_real(double nu, double x) {
    if (x == 0.0) return nu == 0.0 ? 1.0 : 0.0;
    if (nu < 0.0 && nu == floor(nu)) {  // negative integer order: J_{-m} = (-1)^m J_m
        const double m = -nu;
        const double v = besselj_real(m, x);
        return (fmod(m, 2.0) == 0.0) ? v : -v;
    }
    const dd q = dd_mul(two_prod(x, x), {0.25, 0.0});  // x^2/4 exactly
    dd term = {1.0, 0.0}, sum = {1.0, 0.0};
    double maxterm = 1.0;
    for (int k = 0; k < 600; ++k) {
        const dd den = dd_mul({(double)(k + 1), 0.0}, two_sum(nu, (double)(k + 1)));
        term = dd_div(dd_mul(term, q), den);
        term.hi = -term.hi;
        term.lo = -term.lo;
        sum = dd_add(sum, term);
        maxterm = std::max(maxterm, fabs(term.hi));
        if (fabs(term.hi) < 1e-36 * maxterm && (double)k > 0.5 * x) break;
    }
    const double pref = pow(0.5 * x, nu) / tgamma(nu + 1.0);
    return pref * sum.hi;
}

// ---- besselroots.jl:69-165 -----------------------------------------------------------------------
static double mcmahon(double nu, int k) {
    const double mu = 4 * nu * nu;
    const double a1 = 1.0 / 8;
    const double a3 = (7 * mu - 31) / 384;
    const double a5 = 4 * (3779 + mu * (-982 + 83 * mu)) / 61440;
    const double a7 = 6 * (-6277237 + mu * (1585743 + mu * (-153855 + 6949 * mu))) / 20643840;
    const double a9 = 144 * (2092163573.0 + mu * (-512062548 + mu * (48010494 + mu * (-2479316 + 70197 * mu)))) / 11890851840.0;
    const double a11 = 720 * (-8249725736393.0 + mu * (1982611456181.0 + mu * (-179289628602.0 + mu * (8903961290.0 + mu * (-287149133 + mu * 5592657))))) / 10463949619200.0;
    const double a13 = 576 * (423748443625564327.0 + mu * (-100847472093088506.0 + mu * (8929489333108377.0 + mu * (-426353946885548.0 + mu * (13172003634537.0 + mu * (-291245357370.0 + mu * 4148944183.0)))))) / 13059009124761600.0;
    const double b = 0.25 * (2 * nu + 4 * k - 1) * PI;
    const double b2 = b * b;
    return b - (mu - 1) * (((((((a13 / b2 + a11) / b2 + a9) / b2 + a7) / b2 + a5) / b2 + a3) / b2 + a1) / b);
}

static const double J0_ROOTS20[20] = {
    2.4048255576957728, 5.5200781102863106, 8.6537279129110122, 11.791534439014281,
    14.930917708487785, 18.071063967910922, 21.211636629879258, 24.352471530749302,
    27.493479132040254, 30.634606468431975, 33.775820213573568, 36.917098353664044,
    40.058425764628239, 43.199791713176730, 46.341188371661814, 49.482609897397817,
    52.624051841114996, 55.765510755019979, 58.906983926080942, 62.048469190227170};

void approx_besselroots(double nu, int n, double* x) {
    for (int k = 0; k < n; ++k) x[k] = 0.0;
    if (nu == 0) {
        for (int k = 1; k <= std::min(n, 20); ++k) x[k - 1] = J0_ROOTS20[k - 1];
        for (int k = std::min(n, 20) + 1; k <= n; ++k) x[k - 1] = mcmahon(nu, k);
    } else if (-1 <= nu && nu <= 5) {
        const double pt = (nu - 2) / 3;
        double T[30];
        T[0] = 1.0;
        T[1] = pt;
        for (int i = 2; i < 30; ++i) T[i] = 2 * pt * T[i - 1] - T[i - 2];
        for (int k = 1; k <= std::min(n, 6); ++k) {
            double y = 0.0;
            for (int i = 0; i < 30; ++i) y += PIESSENS_C[i][k - 1] * T[i];
            if (k == 1) y *= sqrt(nu + 1);
            x[k - 1] = y;
        }
        for (int k = std::min(n, 6) + 1; k <= n; ++k) x[k - 1] = mcmahon(nu, k);
    } else if (nu > 5) {
        for (int k = 1; k <= n; ++k) x[k - 1] = mcmahon(nu, k);
    }
}

// ---- gaussjacobi.jl:157-168 ------------------------------------------------------------------------
double jacobi_weights_constant(int64_t n, double a, double b) {
    const int M = (int)std::min<int64_t>(20, n - 1);
    double C = 1.0;
    double p = -a * b / (double)n;
    for (int m = 1; m <= M; ++m) {
        C += p;
        p *= -(m + a) * (m + b) / (m + 1) / (double)(n - m);
        if (fabs(p / C) < EPS / 100) break;
    }
    return pow(2.0, a + b + 1) * C;
}

static const double STIRLING_G[10] = {
    1.0, 1.0 / 12, 1.0 / 288, -139.0 / 51840, -571.0 / 2488320, 163879.0 / 209018880,
    5246819.0 / 75246796800.0, -534703531.0 / 902961561600.0,
    -4483131259.0 / 86684309913600.0, 432261921612371.0 / 514904800886784000.0};
static double stirling_f(double z) {
    double s = 0.0, p = 1.0;
    for (int i = 0; i < 10; ++i) {
        s += STIRLING_G[i] * p;
        p *= 1.0 / z;
    }
    return s;
}

// ---- parameter-only pieces of feval_asy1 (:274-294, :319-350) ----------------------------------------
void jacobi_interior_plan(int64_t n, double a, double b, JacobiHalfPlan* hp) {
    const int M = JAC_M;
    const double nn = (double)n;
    for (int which = 0; which < 2; ++which) {
        const double s1 = which == 0 ? 1.0 : -1.0, s2 = which == 0 ? 0.0 : -2.0;
        double P1[JAC_M];
        P1[0] = 1.0;
        for (int j = 1; j < M; ++j)
            P1[j] = P1[j - 1] * ((a + j - 0.5) * (-a + j - 0.5) / (2 * nn + a + b + j + s1) / j);
        for (int i = 2; i < M; i += 4) P1[i] = -P1[i];
        for (int i = 3; i < M; i += 4) P1[i] = -P1[i];
        double(*T)[JAC_M] = which == 0 ? hp->PHI : hp->PHI2;
        for (int l = 0; l < M; ++l)
            for (int m = 0; m < M; ++m) T[l][m] = 0.0;
        for (int l = 1; l <= M; ++l) {
            T[l - 1][l - 1] = 1.0;
            double cp = 1.0;
            for (int j = 1; j <= M - l - 2; ++j) {
                cp *= (b + j - 0.5) * (-b + j - 0.5) / (2 * nn + a + b + j + l + s2) / j;
                T[l - 1][l + j - 1] = cp;
            }
        }
        for (int l = 0; l < M; ++l)
            for (int m = 0; m < M; ++m) T[l][m] *= P1[l];
    }
    double dsa = a * a / (2 * nn), dsb = b * b / (2 * nn), dsab = (a + b) * (a + b) / (4 * nn);
    double ds = dsa + dsb - dsab, s = ds, dsold = ds;
    int i = 1;
    while (fabs(ds / s) + dsold > EPS / 10) {  // NaN (0/0 for a = b = 1) compares false, as in Julia
        dsold = fabs(ds / s);
        i += 1;
        const double tmp = -(double)(i - 1) / (i + 1) / nn;
        dsa = tmp * dsa * a;
        dsb = tmp * dsb * b;
        dsab = tmp * dsab * (a + b) / 2;
        ds = dsa + dsb - dsab;
        s = s + ds;
    }
    const double p2 = exp(s) * sqrt(2 * PI * (nn + a) * (nn + b) / (2 * nn + a + b)) / (2 * nn + a + b + 1);
    hp->C = 2 * p2 * (stirling_f(nn + a) * stirling_f(nn + b) / stirling_f(2 * nn + a + b)) / PI;
    hp->C2 = hp->C * (a + b + 2 * nn) * (a + b + 1 + 2 * nn) / (4 * (a + nn) * (b + nn));
    hp->a = a;
    hp->b = b;
}

// Parameter-only plan of one boundary solve (asy2 with (al, be)): Bessel-zero guesses, the constants
// C, C2 of feval_asy2 (:443-462) and, for every order alpha + nu (nu = -30..31), what the device series
// needs (reflection of negative integer orders, 1/Gamma(order + 1)).
void jacobi_boundary_plan(int64_t n, double al, double be, JacobiBoundaryPlan* P) {
    const double nn = (double)n;
    P->n = n;
    P->al = al;
    P->be = be;
    P->rho = nn + 0.5 * (al + be + 1);
    P->rho2 = nn + 0.5 * (al + be - 1);
    approx_besselroots(al, JAC_NBDY, P->jk);
    double ds = 0.5 * (al * al) / nn, s = ds;
    int jj = 1;
    while (fabs(ds / s) > EPS / 10) {  // NaN for al = 0: loop not entered, as in Julia
        jj += 1;
        ds = -(double)(jj - 1) / (jj + 1) / nn * (ds * al);
        s = s + ds;
    }
    const double p2 = exp(s) * sqrt((nn + al) / nn) * pow(nn / P->rho, al);
    P->C = p2 * (stirling_f(nn + al) / stirling_f(nn)) / sqrt(2.0);
    P->C2 = P->C * nn / (nn + al) * pow(P->rho / P->rho2, al);
    for (int o = 0; o < 64; ++o) {
        double nu = al + (double)(o - 30);
        BesselOrder& b = P->ord[o];
        b.sign = 1.0;
        if (nu < 0.0 && nu == floor(nu)) {
            const double m = -nu;
            b.sign = fmod(m, 2.0) == 0.0 ? 1.0 : -1.0;
            nu = m;
        }
        b.nu = nu;
        b.rgamma = 1.0 / tgamma(nu + 1.0);
        b.cosc = 0.0;
        b.sinc = 0.0;
    }
}

}  // namespace fgq
