// jacobi_boundary.cu — HOST planner pieces of the Gauss-Jacobi path: everything that is O(1) in n.
//
//   approx_besselroots / McMahon / piessens        src/besselroots.jl:23-165
//   asy2, feval_asy2                               src/gaussjacobi.jl:363-468   (10 nodes per end)
//   asy2_higherterms                               src/gaussjacobi.jl:470-568
//   bary, bessel_taylor (nck_mat)                  src/gaussjacobi.jl:601-667
//   weightsConstant                                src/gaussjacobi.jl:157-168
//   PHI/PHI2 tables and the constants C, C2 of feval_asy1   src/gaussjacobi.jl:274-294, 319-350
//
// The 2 x 10 boundary nodes cost a few thousand Bessel evaluations regardless of n, so they are
// planned on the host and handed to the patch kernel as launch parameters;
// the O(n) interior runs in jacobi.cu.  The reference's besselj (AMOS zbesj, outside its tree) is
// replaced by a real-order ascending series summed in double-double arithmetic.
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

// ---- boundary: :470-568 -------------------------------------------------------------------------------
struct HigherTerms {
    double t[10], v[10], tB1[10], A2[10], tB2[10], A3[10];
};

static void matvec10(const double M[10][10], double scale, const double* x, double* y) {
    for (int i = 0; i < 10; ++i) {
        double s = 0.0;
        for (int j = 0; j < 10; ++j) s += (scale * M[i][j]) * x[j];
        y[i] = s;
    }
}

static void asy2_higherterms(double al, double be, const double* theta, int npts, HigherTerms* H) {
    const int N = 10;
    const double A = 0.25 - al * al, B = 0.25 - be * be;
    double c = 0.5;
    for (int i = 0; i < npts; ++i) c = std::max(c, theta[i]);
    double* t = H->t;
    double* v = H->v;
    for (int i = 0; i < N; ++i) {
        t[i] = 0.5 * c * (sin(PI * (double)(-(N - 1) + 2 * i) / (2.0 * (N - 1))) + 1);
        v[i] = 1.0;
    }
    v[0] = 0.5;
    for (int i = 1; i < N; i += 2) v[i] = -1.0;
    v[N - 1] *= 0.5;
    double g[10], gp[10], gpp[10], B0[10], B0p[10], A1[10], A1p[10], A1p_t[10], f[10];
    for (int i = 0; i < N; ++i) {
        const double ti = t[i];
        const double cot = 1.0 / tan(ti / 2), csc = 1.0 / sin(ti / 2), sec = 1.0 / cos(ti / 2);
        g[i] = A * (cot - 2 / ti) - B * tan(ti / 2);
        gp[i] = A * (2 / (ti * ti) - 0.5 * csc * csc) - 0.5 * (0.25 - be * be) * sec * sec;
        const double s2 = sin(ti / 2), cs = 1.0 / sin(ti);
        gpp[i] = A * (-4 / (ti * ti * ti) + 0.25 * sin(ti) * (csc * csc) * (csc * csc)) - 4 * B * (s2 * s2) * (s2 * s2) * (cs * cs * cs);
    }
    g[0] = 0;
    gp[0] = -A / 6 - 0.5 * B;
    gpp[0] = 0;
    for (int i = 0; i < N; ++i) {
        B0[i] = 0.25 * g[i] / t[i];
        B0p[i] = 0.25 * (gp[i] / t[i] - g[i] / (t[i] * t[i]));
    }
    B0[0] = 0.25 * (-A / 6 - 0.5 * B);
    B0p[0] = 0;
    const double A10 = al * (A + 3 * B) / 24;
    for (int i = 0; i < N; ++i) {
        A1[i] = 0.125 * gp[i] - (1 + 2 * al) / 2 * B0[i] - g[i] * g[i] / 32 - A10;
        A1p[i] = 0.125 * gpp[i] - (1 + 2 * al) / 2 * B0p[i] - gp[i] * g[i] / 16;
        A1p_t[i] = A1p[i] / t[i];
    }
    A1p_t[0] = -A / 720 - A * A / 576 - A * B / 96 - B * B / 64 - B / 48 + al * (A / 720 + B / 48);
    for (int i = 0; i < N; ++i) {
        const double ti = t[i], t2 = ti * ti;
        const double ch = 2 * cos(ti / 2);
        const double fcos = B / (ch * ch);
        double fi;
        if (ti > 0.5) {
            const double sh = 2 * sin(ti / 2);
            fi = A * (1 / t2 - 1 / (sh * sh));
        } else {
            const double t4 = t2 * t2, t6 = t4 * t2, t8 = t4 * t4, t10 = t8 * t2, t12 = t6 * t6, t14 = t12 * t2, t16 = t8 * t8;
            fi = -A * (1.0 / 12 + t2 / 240 + t4 / 6048 + t6 / 172800 + t8 / 5322240 + 691 * t10 / 118879488000.0 +
                       t12 / 5748019200.0 + 3617 * t14 / 711374856192000.0 + 43867 * t16 / 300534953951232000.0);
        }
        f[i] = fi - fcos;
    }
    const double cs = 0.5 * c, dsca = 2 / c;
    double I[10], J[10], tmp[10], K[10], B1[10], A2p[10], A2p_t[10], B2[10], w[9], D1[10], C1[10], C2v[10];
    matvec10(CUMSUMMAT_10, cs, A1p_t, I);
    for (int i = 0; i < N; ++i) tmp[i] = f[i] * A1[i];
    matvec10(CUMSUMMAT_10, cs, tmp, J);
    double* tB1 = H->tB1;
    for (int i = 0; i < N; ++i) tB1[i] = -0.5 * A1p[i] - (0.5 + al) * I[i] + 0.5 * J[i];
    tB1[0] = 0;
    for (int i = 0; i < N; ++i) B1[i] = tB1[i] / t[i];
    B1[0] = A / 720 + A * A / 576 + A * B / 96 + B * B / 64 + B / 48 + al * (A * A / 576 + B * B / 64 + A * B / 96) -
            al * al * (A / 720 + B / 48);
    for (int i = 0; i < N; ++i) tmp[i] = f[i] * tB1[i];
    matvec10(CUMSUMMAT_10, cs, tmp, K);
    matvec10(DIFFMAT_10, dsca, tB1, D1);
    double* A2 = H->A2;
    for (int i = 0; i < N; ++i) A2[i] = 0.5 * D1[i] - (0.5 + al) * B1[i] - 0.5 * K[i];
    const double A20 = A2[0];
    for (int i = 0; i < N; ++i) A2[i] -= A20;
    matvec10(DIFFMAT_10, dsca, A2, A2p);
    const double A2p0 = A2p[0];
    for (int i = 0; i < N; ++i) A2p[i] -= A2p0;
    for (int i = 0; i < N; ++i) A2p_t[i] = A2p[i] / t[i];
    for (int i = 0; i < N - 1; ++i) w[i] = PI / 2 - t[i + 1];
    for (int i = 1; i < N - 1; i += 2) w[i] = -w[i];
    w[N - 2] = 0.5 * w[N - 2];
    double sw = 0.0, swa = 0.0;
    for (int i = 0; i < N - 1; ++i) {
        sw += w[i];
        swa += w[i] * A2p_t[i + 1];
    }
    A2p_t[0] = swa / sw;
    matvec10(CUMSUMMAT_10, cs, A2p_t, C1);
    for (int i = 0; i < N; ++i) tmp[i] = f[i] * A2[i];
    matvec10(CUMSUMMAT_10, cs, tmp, C2v);
    double* tB2 = H->tB2;
    for (int i = 0; i < N; ++i) tB2[i] = -0.5 * A2p[i] - (0.5 + al) * C1[i] + 0.5 * C2v[i];
    for (int i = 0; i < N; ++i) B2[i] = tB2[i] / t[i];
    double swb = 0.0;
    for (int i = 0; i < N - 1; ++i) swb += w[i] * B2[i + 1];
    B2[0] = swb / sw;
    for (int i = 0; i < N; ++i) tmp[i] = f[i] * tB2[i];
    matvec10(CUMSUMMAT_10, cs, tmp, K);
    matvec10(DIFFMAT_10, dsca, tB2, D1);
    double* A3 = H->A3;
    for (int i = 0; i < N; ++i) A3[i] = 0.5 * D1[i] - (0.5 + al) * B2[i] - 0.5 * K[i];
    const double A30 = A3[0];
    for (int i = 0; i < N; ++i) A3[i] -= A30;
}

// :601-622
static double bary(double x, const double* fvals, const double* xk, const double* vk) {
    double num = 0.0, den = 0.0;
    for (int k = 0; k < 10; ++k) {
        const double xx = vk[k] / (x - xk[k]);
        num += xx * fvals[k];
        den += xx;
    }
    double r = num / den;
    if (r != r) {
        for (int k = 0; k < 10; ++k)
            if (x == xk[k]) return fvals[k];
    }
    return r;
}

// :624-654 — J_a(z + t) by a Taylor expansion about z (binomials directly: see DESIGN.md on the
// reference's nck_mat for kmax <= 3)
static void bessel_taylor(const double* t, const double* z, int npts, double al, double* out) {
    double tinf = 0.0;
    for (int i = 0; i < npts; ++i) tinf = std::max(tinf, fabs(t[i]));
    int kmax = (int)ceil(fabs(log(EPS) / log(tinf)));
    if (kmax > 30) kmax = 30;
    if (kmax < 0) kmax = 0;
    std::vector<double> Bjk(2 * kmax + 1);
    for (int p = 0; p < npts; ++p) {
        for (int j = -kmax; j <= kmax; ++j) Bjk[j + kmax] = besselj_real(al + j, z[p]);
        double acc = Bjk[kmax];  // k = 0 term, H = t^0
        double fact = 1.0, tp = 1.0, two_k = 1.0;
        for (int k = 1; k <= kmax; ++k) {
            double aa = 0.0, sgn = 1.0, binom = 1.0;
            for (int l = 0; l <= k; ++l) {
                aa = aa + sgn * binom * Bjk[kmax + 2 * l - k];
                sgn = -sgn;
                binom = binom * (double)(k - l) / (double)(l + 1);
            }
            fact = k * fact;
            two_k *= 2.0;
            aa /= two_k * fact;
            tp *= t[p];
            acc += aa * tp;
        }
        out[p] = acc;
    }
}

// :405-468
static void feval_asy2(int64_t n, double al, double be, const double* t, int npts, const HigherTerms& H,
                       double* vals, double* ders) {
    const double nn = (double)n;
    const double rho = nn + 0.5 * (al + be + 1), rho2 = nn + 0.5 * (al + be - 1);
    const double A = 0.25 - al * al, B = 0.25 - be * be;
    double Ja[16], Jb[16], Jbb[16], Jab[16], mt[16], z[16];
    for (int i = 0; i < npts; ++i) {
        z[i] = rho * t[i];
        mt[i] = -t[i];
        Ja[i] = besselj_real(al, z[i]);
        Jb[i] = besselj_real(al + 1, z[i]);
        Jbb[i] = besselj_real(al + 1, rho2 * t[i]);
    }
    bessel_taylor(mt, z, npts, al, Jab);
    double ds = 0.5 * (al * al) / nn, s = ds;
    int jj = 1;
    while (fabs(ds / s) > EPS / 10) {
        jj += 1;
        ds = -(double)(jj - 1) / (jj + 1) / nn * (ds * al);
        s = s + ds;
    }
    const double p2 = exp(s) * sqrt((nn + al) / nn) * pow(nn / rho, al);
    const double C = p2 * (stirling_f(nn + al) / stirling_f(nn)) / sqrt(2.0);
    const double C2 = C * nn / (nn + al) * pow(rho / rho2, al);
    const double A10 = al * (A + 3 * B) / 24;
    for (int i = 0; i < npts; ++i) {
        const double ti = t[i];
        const double cot = 1.0 / tan(ti / 2), csc = 1.0 / sin(ti / 2), sec = 1.0 / cos(ti / 2);
        const double gt = A * (cot - (2 / ti)) - B * tan(ti / 2);
        const double gtdx = A * (2 / (ti * ti) - 0.5 * csc * csc) - 0.5 * B * sec * sec;
        const double tB0 = 0.25 * gt;
        const double A1 = gtdx / 8 - (1 + 2 * al) / 8 * gt / ti - gt * gt / 32 - A10;
        const double tB1t = bary(ti, H.tB1, H.t, H.v), A2t = bary(ti, H.A2, H.t, H.v);
        const double tB2t = bary(ti, H.tB2, H.t, H.v), A3t = bary(ti, H.A3, H.t, H.v);
        const double r2 = rho * rho, r3 = r2 * rho, r4 = r2 * r2, r5 = r4 * rho, r6 = r3 * r3;
        const double q2 = rho2 * rho2, q3 = q2 * rho2, q4 = q2 * q2, q5 = q4 * rho2, q6 = q3 * q3;
        double v = Ja[i] + Jb[i] * tB0 / rho + Ja[i] * A1 / r2 + Jb[i] * tB1t / r3 + Ja[i] * A2t / r4;
        double v2 = Jab[i] + Jbb[i] * tB0 / rho2 + Jab[i] * A1 / q2 + Jbb[i] * tB1t / q3 + Jab[i] * A2t / q4;
        v += Jb[i] * tB2t / r5 + Ja[i] * A3t / r6;
        v2 += Jbb[i] * tB2t / q5 + Jab[i] * A3t / q6;
        const double valstmp = C * v;
        const double denom = pow(sin(ti / 2), al + 0.5) * pow(cos(ti / 2), be + 0.5);
        vals[i] = sqrt(ti) * valstmp / denom;
        double d = (nn * (al - be - (2 * nn + al + be) * cos(ti)) * valstmp + 2 * (nn + al) * (nn + be) * C2 * v2) /
                   (2 * nn + al + be);
        d *= sqrt(ti) / (denom * sin(ti));
        ders[i] = d;
    }
}

// :363-398 — returns x ascending (towards +1) and w = 1/ders^2 (unscaled), npts = 10
void jacobi_asy2(int64_t n, double al, double be, int npts, double* x, double* w) {
    const double nn = (double)n;
    double jk[16], t[16], vals[16], ders[16];
    approx_besselroots(al, npts, jk);
    const double r = nn + 0.5 * (al + be + 1);
    for (int i = 0; i < npts; ++i) {
        const double phik = jk[i] / r;
        t[i] = phik + ((al * al - 0.25) * (1 - phik * (1.0 / tan(phik))) / (2 * phik) -
                       0.25 * (al * al - be * be) * tan(0.5 * phik)) / (r * r);
    }
    HigherTerms H;
    asy2_higherterms(al, be, t, npts, &H);
    for (int it = 0; it < 10; ++it) {
        feval_asy2(n, al, be, t, npts, H, vals, ders);
        double mx = 0.0;
        bool nan = false;
        for (int i = 0; i < npts; ++i) {
            const double dt = vals[i] / ders[i];
            t[i] += dt;
            if (dt != dt) nan = true;
            mx = std::max(mx, fabs(dt));
        }
        if (!nan && mx < sqrt(EPS) / 200) break;
    }
    feval_asy2(n, al, be, t, npts, H, vals, ders);
    for (int i = 0; i < npts; ++i) t[i] += vals[i] / ders[i];
    for (int i = 0; i < npts; ++i) {
        const int r_i = npts - 1 - i;
        x[i] = cos(t[r_i]);
        w[i] = 1.0 / (ders[r_i] * ders[r_i]);
    }
}

}  // namespace fgq
