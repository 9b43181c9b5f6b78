/* legendre_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * Float-for-float restatement in plain C of the Gauss-Legendre path of
 * FastGaussQuadrature.jl v1.3.0:
 *   src/gausslegendre.jl:22-309   (dispatch, closed forms, asy, rec)
 *   src/besselroots.jl:168-245    (besselZeroRoots, besselJ1)
 *   src/constants.jl:13-36,93-106 (J0 zeros, J1^2 at J0 zeros)
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The shipped library never does.
 *
 * Semantics honoured: muladd/evalpoly -> fma(), every
 * other a*b+c unfused (compile with -ffp-contract=off), cot = 1/tan, x^2 = x*x,
 * multiply by the rounded reciprocal 1/(n+0.5), 64-bit integers throughout.
 *
 * Parity pin: the golden values of the reference's own tests
 * (test/test_gausslegendre.jl:20-52) are checked in tests/test_oracle_legendre.py.
 * Julia cannot run in this image, so behaviour below 1e-14 absolute (the ulp
 * level of Julia's libm) is UNPINNED; see DESIGN.md.
 *
 * Three libm modes:
 *   default      sin/cos/tan from the host libm (glibc) — the independent check.
 *   -DFGQ_TWIN   sin/cos/cot from the library's own fgq_math.h, so that the
 *                kernels can be pinned bit-for-bit against this file.
 *   -DFGQ_JLIBM  sin/cos/tan from julia_libm.h: a restatement of the fdlibm-derived
 *                algorithms Julia Base uses (the functions the reference really calls).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef FGQ_TWIN
#include "../fastgaussquadrature.jl_b200/csrc/fgq_math.h"
#define O_SIN(x) fgq_sin_q1(x)
#define O_COS(x) fgq_cos_q1(x)
#define O_COT(x) fgq_cot_q1(x)
#define O_NAME(f) f##_twin
#elif defined(FGQ_JLIBM)
#include "julia_libm.h"
#define O_SIN(x) jl_sin(x)
#define O_COS(x) jl_cos(x)
#define O_COT(x) jl_cot(x)
#define O_NAME(f) f##_jl
#else
#define O_SIN(x) sin(x)
#define O_COS(x) cos(x)
#define O_COT(x) (1.0 / tan(x))
#define O_NAME(f) f
#endif

#define PI_F64 3.141592653589793

/* constants.jl:13-36 */
static const double J0_ROOTS[20] = {
    2.4048255576957728, 5.5200781102863106, 8.6537279129110122, 11.791534439014281,
    14.930917708487785, 18.071063967910922, 21.211636629879258, 24.352471530749302,
    27.493479132040254, 30.634606468431975, 33.775820213573568, 36.917098353664044,
    40.058425764628239, 43.199791713176730, 46.341188371661814, 49.482609897397817,
    52.624051841114996, 55.765510755019979, 58.906983926080942, 62.048469190227170};
/* constants.jl:93-106 */
static const double J1SQ_AT_J0[10] = {
    0.2695141239419169,  0.1157801385822037,  0.07368635113640822, 0.05403757319811628,
    0.04266142901724309, 0.03524210349099610, 0.03002107010305467, 0.02614739149530809,
    0.02315912182469139, 0.02078382912226786};


/* besselroots.jl:168-200; fills jk[k - k0] for k in [k0, k1] (1-based, inclusive). */
static void bessel_zero_roots(int64_t k0, int64_t k1, double* jk) {
    const double p3 = -401743168.0 / 105.0, p5 = 120928.0 / 15.0, p7 = -124.0 / 3.0;
#pragma omp parallel for schedule(static)
    for (int64_t jj = k0; jj <= k1; ++jj) {
        double r;
        if (jj <= 20) {
            r = J0_ROOTS[jj - 1];
        } else {
            double ak = PI_F64 * ((double)jj - 0.25);
            if (jj <= 47) {
                double q = 0.125 / ak;
                double ak82 = q * q;
                double e = fma(ak82, fma(ak82, fma(ak82, p3, p5), p7), 1.0);
                r = ak + 0.125 / ak * e;
            } else if (jj <= 344) {
                double q = 0.125 / ak;
                double ak82 = q * q;
                double e = fma(ak82, fma(ak82, p5, p7), 1.0);
                r = ak + 0.125 / ak * e;
            } else if (jj <= 13191) {
                double q = 0.125 / ak;
                double ak82 = q * q;
                r = ak + 0.125 / ak * fma(ak82, p7, 1.0);
            } else {
                r = ak + 0.125 / ak;
            }
        }
        jk[jj - k0] = r;
    }
}

/* besselroots.jl:202-245; J1(j0k)^2 for k in [k0, k1]. */
static void bessel_j1sq(int64_t k0, int64_t k1, double* out) {
    const double c1 = -171497088497.0 / 15206400.0, c2 = 461797.0 / 1152.0,
                 c3 = -172913.0 / 8064.0, c4 = 151.0 / 80.0, c5 = -7.0 / 24.0, c7 = 2.0;
#pragma omp parallel for schedule(static)
    for (int64_t jj = k0; jj <= k1; ++jj) {
        double r;
        if (jj <= 10) {
            r = J1SQ_AT_J0[jj - 1];
        } else {
            double ak = PI_F64 * ((double)jj - 0.25);
            double lead = 1.0 / (PI_F64 * ak);
            if (jj <= 2279) {
                double ia = 1.0 / ak;
                double ak2 = ia * ia;
                double e;
                if (jj <= 15)
                    e = fma(ak2, fma(ak2, fma(ak2, fma(ak2, c1, c2), c3), c4), c5);
                else if (jj <= 21)
                    e = fma(ak2, fma(ak2, fma(ak2, c2, c3), c4), c5);
                else if (jj <= 55)
                    e = fma(ak2, fma(ak2, c3, c4), c5);
                else if (jj <= 279)
                    e = fma(ak2, c4, c5);
                else
                    e = c5;
                r = lead * fma(e, ak2 * ak2, c7);
            } else {
                r = lead * c7;
            }
        }
        out[jj - k0] = r;
    }
}

/* gausslegendre.jl:96-149; in: a[i] (angles), out: nodes[i] = cos(theta_i). */
static void legpts_nodes(int64_t n, int64_t cnt, const double* a, double* nodes) {
    double vn = 1.0 / ((double)n + 0.5);
    double vn2 = vn * vn;
    double vn4 = vn2 * vn2;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < cnt; ++i) nodes[i] = O_COT(a[i]);
    if (n <= 255) {
        double vn6 = vn4 * vn2;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < cnt; ++i) {
            double u = nodes[i], u2 = u * u, ai = a[i];
            double ai2 = ai * ai, ai3 = ai2 * ai, ai5 = ai2 * ai3;
            double node = ai + (u - 1.0 / ai) / 8.0 * vn2;
            double v1 = (6.0 * (1.0 + u2) / ai + 25.0 / ai3 - u * fma(31.0, u2, 33.0)) / 384.0;
            double v2 = u * fma(u2, fma(u2, 3779.0 / 15360.0, 6350.0 / 15360.0), 2595.0 / 15360.0);
            double v3 = (1.0 + u2) * (-fma(31.0 / 1024.0, u2, 11.0 / 1024.0) / ai + u / 512.0 / ai2 +
                                      -25.0 / 3072.0 / ai3);
            double v4 = (v2 - 1073.0 / 5120.0 / ai5 + v3);
            node = fma(v1, vn4, node);
            node = fma(v4, vn6, node);
            nodes[i] = node;
        }
    } else if (n <= 3950) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < cnt; ++i) {
            double u = nodes[i], u2 = u * u, ai = a[i];
            double ai2 = ai * ai, ai3 = ai2 * ai;
            double node = ai + (u - 1.0 / ai) / 8.0 * vn2;
            double v1 = (6.0 * (1.0 + u2) / ai + 25.0 / ai3 - u * fma(31.0, u2, 33.0)) / 384.0;
            nodes[i] = fma(v1, vn4, node);
        }
    } else {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < cnt; ++i) {
            double u = nodes[i], ai = a[i];
            nodes[i] = ai + (u - 1.0 / ai) / 8.0 * vn2;
        }
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < cnt; ++i) nodes[i] = O_COS(nodes[i]);
}

/* gausslegendre.jl:151-231; k0 = 1-based half-index of a[0]. */
static void legpts_weights(int64_t n, int64_t k0, int64_t cnt, const double* a, double* weights,
                           double* bj1) {
    double vn = 1.0 / ((double)n + 0.5);
    double vn2 = vn * vn;
    if (n <= 850000) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < cnt; ++i) weights[i] = O_COT(a[i]);
    }
    if (n <= 170) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < cnt; ++i) {
            double u = weights[i], u2 = u * u, ai = a[i];
            double aim1 = 1.0 / ai, ai2 = ai * ai, aim2 = 1.0 / ai2, ua = u * ai;
            double W1 = fma(ua - 1.0, aim2, 1.0) / 8.0;
            double E1 = fma(u2, fma(u2, -56.0, -84.0), -27.0);
            double E2 = fma(-3.0, fma(u2, -2.0, 1.0), 6.0 * ua);
            double E3 = fma(ua, -31.0, 81.0);
            double W2 = fma(aim2, fma(aim2, E3, E2), E1) / 384.0;
            double F1 = fma(u2, fma(u2, fma(u2, 151.0 / 160.0, 187.0 / 96.0), 295.0 / 256.0),
                            153.0 / 1024.0);
            double F2 = fma(u2, fma(u2, -35.0 / 384.0, -119.0 / 768.0), -65.0 / 1024.0) * u;
            double F3 = fma(u2, fma(u2, 7.0 / 384.0, 15.0 / 512.0), 5.0 / 512.0);
            double F4 = fma(u2, 1.0 / 512.0, -13.0 / 1536.0) * u;
            double F5 = fma(u2, -7.0 / 384.0, 53.0 / 3072.0);
            double F6 = 3749.0 / 15360.0 * u;
            double F7 = -1125.0 / 1024.0;
            double W3 = fma(aim1, fma(aim1, fma(aim1, fma(aim1, fma(aim1, fma(aim1, F7, F6), F5), F4), F3), F2), F1);
            weights[i] = fma(vn2, fma(vn2, W3, W2), 1.0 / vn2 + W1);
        }
    } else if (n <= 1500) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < cnt; ++i) {
            double u = weights[i], u2 = u * u, ai = a[i];
            double ai2 = ai * ai, aim2 = 1.0 / ai2, ua = u * ai;
            double W1 = fma(ua - 1.0, aim2, 1.0) / 8.0;
            double E1 = fma(u2, fma(u2, -56.0, -84.0), -27.0);
            double E2 = fma(-3.0, fma(u2, -2.0, 1.0), 6.0 * ua);
            double E3 = fma(ua, -31.0, 81.0);
            double W2 = fma(aim2, fma(aim2, E3, E2), E1) / 384.0;
            weights[i] = fma(vn2, W2, 1.0 / vn2 + W1);
        }
    } else if (n <= 850000) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < cnt; ++i) {
            double u = weights[i], ai = a[i];
            double ai2 = ai * ai, aim2 = 1.0 / ai2, ua = u * ai;
            double W1 = fma(ua - 1.0, aim2, 1.0) / 8.0;
            weights[i] = 1.0 / vn2 + W1;
        }
    } else {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < cnt; ++i) weights[i] = 1.0 / vn2;
    }
    bessel_j1sq(k0, k0 + cnt - 1, bj1);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < cnt; ++i)
        weights[i] = 2.0 / (bj1[i] * (a[i] / O_SIN(a[i])) * weights[i]);
}

/* Half-range values for half-indices k in [k0, k1] of the n-point rule, exactly as
 * asy() computes them before mirroring (gausslegendre.jl:76-80): xh = cos(theta_k) > 0,
 * wh = weight.  Pass structure (one loop per sub-formula, temporaries in DRAM) follows
 * the reference.  Returns 0, or -1 when out of memory. */
int O_NAME(oracle_legendre_asy_half)(int64_t n, int64_t k0, int64_t k1, double* xh, double* wh) {
    int64_t cnt = k1 - k0 + 1;
    if (cnt <= 0) return 0;
    double* a = (double*)malloc(sizeof(double) * (size_t)cnt);
    double* bj1 = (double*)malloc(sizeof(double) * (size_t)cnt);
    if (!a || !bj1) {
        free(a);
        free(bj1);
        return -1;
    }
    bessel_zero_roots(k0, k1, a);
    double vn = 1.0 / ((double)n + 0.5);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < cnt; ++i) a[i] *= vn;
    legpts_nodes(n, cnt, a, xh);
    legpts_weights(n, k0, cnt, a, wh, bj1);
    free(a);
    free(bj1);
    return 0;
}

/* gausslegendre.jl:71-94 */
static int asy(int64_t n, double* x, double* w) {
    int64_t m = (n + 1) >> 1;
    if (O_NAME(oracle_legendre_asy_half)(n, 1, m, x, w)) return -1;
    for (int64_t i = 1; i <= m; ++i) {
        x[n - i] = x[i - 1];
        w[n - i] = w[i - 1];
    }
    for (int64_t i = 1; i <= m; ++i) x[i - 1] = -x[i - 1];
    if (n & 1) x[m - 1] = 0.0;
    return 0;
}

/* gausslegendre.jl:269-290 for one abscissa */
static void inner_rec(int64_t n, double xj, double* P, double* PP) {
    double Pm2 = 1.0, Pm1 = xj, PPm1 = 1.0, PPm2 = 0.0;
    for (int64_t k = 1; k <= n - 1; ++k) {
        double tk = (double)(2 * k + 1), dk = (double)k, dk1 = (double)(k + 1);
        double newP = fma(tk * Pm1, xj, -dk * Pm2) / dk1;
        Pm2 = Pm1;
        Pm1 = newP;
        double newPP = (tk * fma(xj, PPm1, Pm2) - dk * PPm2) / dk1;
        PPm2 = PPm1;
        PPm1 = newPP;
    }
    *P = Pm1;
    *PP = PPm1;
}

/* gausslegendre.jl:233-267 (+ leg_initial_guess :299-309), Float64: two Newton steps. */
static int rec(int64_t n, double* x, double* w) {
    int64_t m = n / 2 + 1;
    double a[64], x0[64], P[64], PP[64];
    if (m > 64) return -2;
    bessel_zero_roots(1, m, a);
    double vn = 1.0 / ((double)n + 0.5);
    for (int64_t i = 0; i < m; ++i) a[i] *= vn;
    legpts_nodes(n, m, a, x0);
    for (int64_t i = 0; i < m; ++i) x0[i] *= -1.0;
    for (int it = 0; it < 2; ++it) {
        for (int64_t i = 0; i < m; ++i) inner_rec(n, x0[i], &P[i], &PP[i]);
        for (int64_t i = 0; i < m; ++i) x0[i] -= P[i] / PP[i];
    }
    /* mirror (:255-261); for even n the m-th value is overwritten, as in the reference */
    for (int64_t i = 0; i < m && i < n; ++i) {
        x[i] = x0[i];
        w[i] = PP[i];
    }
    for (int64_t i = 1; i <= m - 1; ++i) {
        x[n - i] = -x0[i - 1];
        w[n - i] = -PP[i - 1];
    }
    for (int64_t i = 0; i < n; ++i) w[i] = 2.0 / ((1.0 - x[i] * x[i]) * (w[i] * w[i]));
    return 0;
}

/* gausslegendre.jl:22-68.  Returns 0 ok, 1 DomainError (n < 0), <0 internal. */
int O_NAME(oracle_gausslegendre)(int64_t n, double* x, double* w) {
    if (n < 0) return 1;
    if (n == 0) return 0;
    if (n == 1) {
        x[0] = 0.0;
        w[0] = 2.0;
    } else if (n == 2) {
        double s3 = sqrt(3.0);
        x[0] = -1.0 / s3;
        x[1] = 1.0 / s3;
        w[0] = w[1] = 1.0;
    } else if (n == 3) {
        double s = sqrt(3.0 / 5.0);
        x[0] = -s;
        x[1] = 0.0;
        x[2] = s;
        w[0] = 5.0 / 9.0;
        w[1] = 8.0 / 9.0;
        w[2] = 5.0 / 9.0;
    } else if (n == 4) {
        double a = 2.0 * sqrt(6.0 / 5.0) / 7.0, t = 3.0 / 7.0, s30 = sqrt(30.0);
        x[0] = -sqrt(t + a);
        x[1] = -sqrt(t - a);
        x[2] = sqrt(t - a);
        x[3] = sqrt(t + a);
        w[0] = w[3] = (18.0 - s30) / 36.0;
        w[1] = w[2] = (18.0 + s30) / 36.0;
    } else if (n == 5) {
        double b = 2.0 * sqrt(10.0 / 7.0), s70 = sqrt(70.0);
        x[0] = -sqrt(5.0 + b) / 3.0;
        x[1] = -sqrt(5.0 - b) / 3.0;
        x[2] = 0.0;
        x[3] = sqrt(5.0 - b) / 3.0;
        x[4] = sqrt(5.0 + b) / 3.0;
        w[0] = w[4] = (322.0 - 13.0 * s70) / 900.0;
        w[1] = w[3] = (322.0 + 13.0 * s70) / 900.0;
        w[2] = 128.0 / 225.0;
    } else if (n <= 60) {
        return rec(n, x, w);
    } else {
        return asy(n, x, w);
    }
    return 0;
}

/* CSR batch of small rules (BASELINE config C5): rule i has n_i nodes at offsets[i]. */
int O_NAME(oracle_gausslegendre_batch)(int64_t nrules, const int32_t* n_i, const int64_t* offsets,
                                       double* x, double* w) {
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(| : bad)
    for (int64_t r = 0; r < nrules; ++r)
        bad |= O_NAME(oracle_gausslegendre)(n_i[r], x + offsets[r], w + offsets[r]);
    return bad;
}

/* theta_k (the angle whose cosine is the node) for k in [k0,k1]; used to bit-compare
 * the kernel's angle (SURVEY.md §7 hard part 1). */
int O_NAME(oracle_legendre_theta)(int64_t n, int64_t k0, int64_t k1, double* theta) {
    int64_t cnt = k1 - k0 + 1;
    if (cnt <= 0) return 0;
    double* a = (double*)malloc(sizeof(double) * (size_t)cnt);
    if (!a) return -1;
    bessel_zero_roots(k0, k1, a);
    double vn = 1.0 / ((double)n + 0.5), vn2 = vn * vn, vn4 = vn2 * vn2, vn6 = vn4 * vn2;
    for (int64_t i = 0; i < cnt; ++i) {
        double ai = a[i] * vn;
        double u = O_COT(ai), u2 = u * u;
        double ai2 = ai * ai, ai3 = ai2 * ai, ai5 = ai2 * ai3;
        double node = ai + (u - 1.0 / ai) / 8.0 * vn2;
        if (n <= 3950) {
            double v1 = (6.0 * (1.0 + u2) / ai + 25.0 / ai3 - u * fma(31.0, u2, 33.0)) / 384.0;
            node = fma(v1, vn4, node);
        }
        if (n <= 255) {
            double v2 = u * fma(u2, fma(u2, 3779.0 / 15360.0, 6350.0 / 15360.0), 2595.0 / 15360.0);
            double v3 = (1.0 + u2) * (-fma(31.0 / 1024.0, u2, 11.0 / 1024.0) / ai + u / 512.0 / ai2 +
                                      -25.0 / 3072.0 / ai3);
            double v4 = (v2 - 1073.0 / 5120.0 / ai5 + v3);
            node = fma(v4, vn6, node);
        }
        theta[i] = node;
    }
    free(a);
    return 0;
}

#if !defined(FGQ_TWIN) && !defined(FGQ_JLIBM)
#ifdef _OPENMP
#include <omp.h>
#endif
/* Thread control for the CPU-baseline legs of bench.py (1 = the reference's own
 * single-threaded behaviour; 0 = all cores). Returns the thread count in effect. */
int oracle_set_threads(int nthreads) {
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_num_procs();
    omp_set_num_threads(nthreads);
    return nthreads;
#else
    (void)nthreads;
    return 1;
#endif
}
#endif

#ifdef FGQ_TWIN
/* The library's own elementary functions (fgq_math.h), exported so that tests can measure their accuracy
 * against a higher-precision libm on the host. */
void oracle_twin_sincos_q1(int64_t cnt, const double* t, double* s, double* c) {
    for (int64_t i = 0; i < cnt; ++i) fgq_sincos_q1(t[i], &s[i], &c[i]);
}
void oracle_twin_tan_q1(int64_t cnt, const double* t, double* tn) {
    for (int64_t i = 0; i < cnt; ++i) tn[i] = fgq_tan_q1(t[i]);
}
void oracle_twin_sincos_medium(int64_t cnt, const double* t, double* s, double* c) {
    for (int64_t i = 0; i < cnt; ++i) fgq_sincos_medium(t[i], &s[i], &c[i]);
}
#endif

#ifdef FGQ_JLIBM
/* The restated Julia Base sin/cos/tan, exported so that tests can measure them against a
 * higher-precision libm. */
void oracle_jl_sincostan(int64_t cnt, const double* t, double* s, double* c, double* tn) {
    for (int64_t i = 0; i < cnt; ++i) {
        s[i] = jl_sin(t[i]);
        c[i] = jl_cos(t[i]);
        tn[i] = jl_tan(t[i]);
    }
}
#endif
