/* julia_libm.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * Restatement of the sin / cos / tan that the reference actually calls.  Those functions are NOT
 * in /root/reference: FastGaussQuadrature.jl calls Julia Base (`Project.toml:18`, julia >= 1.10,
 * exact version unpinned), whose Float64 trigonometric functions are pure-Julia ports of
 * FreeBSD msun / Sun fdlibm:
 *     base/special/trig.jl      sin, cos, tan, sin_kernel, cos_kernel, tan_kernel
 *     base/special/rem_pio2.jl  rem_pio2_kernel, cody_waite_2c_pio2, cody_waite_ext_pio2
 * with the polynomial parts written through `@horner` / `evalpoly`, i.e. `muladd` -> hardware FMA,
 * and every other operation unfused.  This header restates that published algorithm operation by
 * operation (fdlibm k_sin.c, k_cos.c, k_tan.c, e_rem_pio2.c medium path; Julia's split of the
 * polynomials and its FMA placement), written from the algorithm's description — Julia is not
 * installed in this image, so the restatement itself cannot be executed against Julia:
 * PARITY UNPINNED at the ulp level, like the rest of the oracle.
 *
 * Its purpose: a THIRD elementary-function mode for the Legendre oracle (next to glibc and the
 * library's own twin) whose rounding behaviour is, as far as the published algorithm determines
 * it, the one the reference sees.  Call sites in the reference that reach these functions on the
 * Legendre path: gausslegendre.jl:100 (`cot`), :145 (`cos`), :158 (`cot`), :227 (`sin`).
 *
 * Range: |x| <= 2^20 * pi/2 (the Cody-Waite "medium" range; Payne-Hanek is not restated — the
 * Legendre angles live in (0, pi/2]).  Compile with -ffp-contract=off.
 */
#ifndef ORACLE_JULIA_LIBM_H_
#define ORACLE_JULIA_LIBM_H_

#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct {
    double hi, lo;
} jl_dd;

static inline uint32_t jl_highword(double x) {
    uint64_t u;
    memcpy(&u, &x, 8);
    return (uint32_t)(u >> 32);
}
static inline uint32_t jl_poshighword(double x) { return jl_highword(x) & 0x7fffffffu; }
static inline double jl_clear_lowword(double x) {
    uint64_t u;
    memcpy(&u, &x, 8);
    u &= 0xffffffff00000000ull;
    memcpy(&x, &u, 8);
    return x;
}

/* fdlibm k_sin.c S1..S6 / k_cos.c C1..C6 */
#define JL_DS1 -1.66666666666666324348e-01
#define JL_DS2 8.33333333332248946124e-03
#define JL_DS3 -1.98412698298579493134e-04
#define JL_DS4 2.75573137070700676789e-06
#define JL_DS5 -2.50507602534068634195e-08
#define JL_DS6 1.58969099521155010221e-10
#define JL_DC1 4.16666666666666019037e-02
#define JL_DC2 -1.38888888888741095749e-03
#define JL_DC3 2.48015872894767294178e-05
#define JL_DC4 -2.75573143513906633035e-07
#define JL_DC5 2.08757232129817482790e-09
#define JL_DC6 -1.13596475577881948265e-11

/* sin_kernel(y::DoubleFloat64): msun k_sin.c with iy = 1; the two Horner pieces are muladd chains. */
static inline double jl_sin_kernel_dd(jl_dd y) {
    double y2 = y.hi * y.hi;
    double y4 = y2 * y2;
    double r = fma(y2, fma(y2, JL_DS4, JL_DS3), JL_DS2) + y2 * y4 * fma(y2, JL_DS6, JL_DS5);
    double y3 = y2 * y.hi;
    return y.hi - ((y2 * (0.5 * y.lo - y3 * r) - y.lo) - y3 * JL_DS1);
}
/* sin_kernel(y::Float64): iy = 0 form */
static inline double jl_sin_kernel(double y) {
    double y2 = y * y;
    double y4 = y2 * y2;
    double r = fma(y2, fma(y2, JL_DS4, JL_DS3), JL_DS2) + y2 * y4 * fma(y2, JL_DS6, JL_DS5);
    double y3 = y2 * y;
    return y + y3 * (JL_DS1 + y2 * r);
}
/* cos_kernel(y::DoubleFloat64): msun k_cos.c (the 1 - z/2 split form). */
static inline double jl_cos_kernel_dd(jl_dd y) {
    double y2 = y.hi * y.hi;
    double y4 = y2 * y2;
    double r = y2 * fma(y2, fma(y2, JL_DC3, JL_DC2), JL_DC1) +
               y4 * y4 * fma(y2, fma(y2, JL_DC6, JL_DC5), JL_DC4);
    double half_y2 = 0.5 * y2;
    double w = 1.0 - half_y2;
    return w + (((1.0 - w) - half_y2) + (y2 * r - y.hi * y.lo));
}
static inline double jl_cos_kernel(double y) {
    jl_dd d = {y, 0.0};
    return jl_cos_kernel_dd(d);
}

/* fdlibm k_tan.c T[0..12], pio4, pio4lo */
static const double JL_T[13] = {
    3.33333333333334091986e-01, 1.33333333333201242699e-01, 5.39682539762260521377e-02,
    2.18694882948595424599e-02, 8.86323982359930005737e-03, 3.59207910759131235356e-03,
    1.45620945432529025516e-03, 5.88041240820264096874e-04, 2.46463134818469906812e-04,
    7.81794442939557092300e-05, 7.14072491382608190305e-05, -1.85586374855275456654e-05,
    2.59073051863633712884e-05};
#define JL_PIO4 7.85398163397448278999e-01
#define JL_PIO4LO 3.06161699786838301793e-17

/* tan_kernel(y::DoubleFloat64, k): k = 1 -> tan(y), k = -1 -> -1/tan(y). */
static inline double jl_tan_kernel_dd(jl_dd y, int k) {
    double yhi = y.hi, ylo = y.lo;
    const int negative = yhi < 0.0;
    const uint32_t hy = jl_poshighword(yhi);
    const int big = hy >= 0x3FE59428u; /* |y| >= 0.6744 */
    if (big) {
        if (negative) {
            yhi = -yhi;
            ylo = -ylo;
        }
        double z = JL_PIO4 - yhi;
        double w = JL_PIO4LO - ylo;
        yhi = z + w;
        ylo = 0.0;
    }
    double z = yhi * yhi;
    double w = z * z;
    /* odd- and even-indexed halves of the degree-27 polynomial, each a muladd chain in w = y^4 */
    double r = fma(w, fma(w, fma(w, fma(w, fma(w, JL_T[11], JL_T[9]), JL_T[7]), JL_T[5]), JL_T[3]), JL_T[1]);
    double v = z * fma(w, fma(w, fma(w, fma(w, fma(w, JL_T[12], JL_T[10]), JL_T[8]), JL_T[6]), JL_T[4]), JL_T[2]);
    double s = z * yhi;
    r = ylo + z * (s * (r + v) + ylo);
    r += JL_T[0] * s;
    double tw = yhi + r;
    if (big) {
        double vk = (double)k;
        double res = vk - 2.0 * (yhi - (tw * tw / (tw + vk) - r));
        return negative ? -res : res;
    }
    if (k == 1) return tw;
    /* -1/(y + r) to full accuracy */
    double zz = jl_clear_lowword(tw);
    double vv = r - (zz - yhi);
    double a = -1.0 / tw;
    double t = jl_clear_lowword(a);
    double ss = 1.0 + t * zz;
    return t + a * (ss + t * vv);
}

#define JL_PIO2_1 1.57079632673412561417e+00
#define JL_PIO2_1T 6.07710050650619224932e-11
#define JL_PIO2_2 6.07710050630396597660e-11
#define JL_PIO2_2T 2.02226624879595063154e-21
#define JL_PIO2_3 2.02226624871116645580e-21
#define JL_PIO2_3T 8.47842766036889956997e-32
#define JL_INVPIO2 6.36619772367581382433e-01 /* Float64(2/pi) */

/* cody_waite_2c_pio2: two-constant reduction with a known multiple fn of pi/2 */
static inline int jl_cody_waite_2c(double x, double fn, int n, jl_dd* y) {
    double z = fma(-fn, JL_PIO2_1, x);
    double y1 = fma(-fn, JL_PIO2_1T, z);
    double y2 = fma(-fn, JL_PIO2_1T, z - y1);
    y->hi = y1;
    y->lo = y2;
    return n;
}
/* cody_waite_ext_pio2: up to three rounds (85 / 118 / 151 bits), msun e_rem_pio2.c "medium" */
static inline int jl_cody_waite_ext(double x, uint32_t xhp, jl_dd* y) {
    double fn = nearbyint(x * JL_INVPIO2);
    double r = fma(-fn, JL_PIO2_1, x);
    double w = fn * JL_PIO2_1T;
    int j = (int)(xhp >> 20);
    double y1 = r - w;
    int i = j - (int)((jl_highword(y1) >> 20) & 0x7ff);
    if (i > 16) {
        double t = r;
        w = fn * JL_PIO2_2;
        r = t - w;
        w = fma(fn, JL_PIO2_2T, -((t - r) - w));
        y1 = r - w;
        i = j - (int)((jl_highword(y1) >> 20) & 0x7ff);
        if (i > 49) {
            t = r;
            w = fn * JL_PIO2_3;
            r = t - w;
            w = fma(fn, JL_PIO2_3T, -((t - r) - w));
            y1 = r - w;
        }
    }
    y->hi = y1;
    y->lo = (r - y1) - w;
    return (int)(long long)fn;
}

/* rem_pio2_kernel for |x| <= 2^20 pi/2: returns n (mod 4 matters) and y = x - n pi/2 as hi + lo */
static inline int jl_rem_pio2(double x, jl_dd* y) {
    const uint32_t xhp = jl_poshighword(x);
    if (xhp <= 0x400f6a7au) {                  /* |x| ~<= 5pi/4 */
        if ((xhp & 0xfffffu) == 0x921fbu)      /* |x| ~= pi/2 or pi: cancellation */
            return jl_cody_waite_ext(x, xhp, y);
        if (xhp <= 0x4002d97cu)                /* |x| ~<= 3pi/4 */
            return x > 0.0 ? jl_cody_waite_2c(x, 1.0, 1, y) : jl_cody_waite_2c(x, -1.0, -1, y);
        return x > 0.0 ? jl_cody_waite_2c(x, 2.0, 2, y) : jl_cody_waite_2c(x, -2.0, -2, y);
    }
    if (xhp <= 0x401c463bu) {                  /* |x| ~<= 9pi/4 */
        if (xhp <= 0x4015fdbcu) {              /* |x| ~<= 7pi/4 */
            if (xhp == 0x4012d97cu) return jl_cody_waite_ext(x, xhp, y); /* ~ 3pi/2 */
            return x > 0.0 ? jl_cody_waite_2c(x, 3.0, 3, y) : jl_cody_waite_2c(x, -3.0, -3, y);
        }
        if (xhp == 0x401921fbu) return jl_cody_waite_ext(x, xhp, y);     /* ~ 2pi */
        return x > 0.0 ? jl_cody_waite_2c(x, 4.0, 4, y) : jl_cody_waite_2c(x, -4.0, -4, y);
    }
    return jl_cody_waite_ext(x, xhp, y);       /* medium: |x| ~< 2^20 pi/2 */
}

#define JL_PIO4_F64 0.7853981633974483 /* Float64(pi)/4 */

static inline double jl_sin(double x) {
    double ax = fabs(x);
    if (ax < JL_PIO4_F64) {
        if (ax < 1.4901161193847656e-08) return x; /* sqrt(eps) */
        return jl_sin_kernel(x);
    }
    jl_dd y;
    int n = jl_rem_pio2(x, &y) & 3;
    if (n == 0) return jl_sin_kernel_dd(y);
    if (n == 1) return jl_cos_kernel_dd(y);
    if (n == 2) return -jl_sin_kernel_dd(y);
    return -jl_cos_kernel_dd(y);
}

static inline double jl_cos(double x) {
    double ax = fabs(x);
    if (ax < JL_PIO4_F64) {
        if (ax < 1.0536712127723509e-08) return 1.0; /* sqrt(eps/2) */
        return jl_cos_kernel(x);
    }
    jl_dd y;
    int n = jl_rem_pio2(x, &y) & 3;
    if (n == 0) return jl_cos_kernel_dd(y);
    if (n == 1) return -jl_sin_kernel_dd(y);
    if (n == 2) return -jl_cos_kernel_dd(y);
    return jl_sin_kernel_dd(y);
}

static inline double jl_tan(double x) {
    double ax = fabs(x);
    if (ax < JL_PIO4_F64) {
        if (ax < 7.450580596923828e-09) return x; /* sqrt(eps)/2 */
        jl_dd d = {x, 0.0};
        return jl_tan_kernel_dd(d, 1);
    }
    jl_dd y;
    int n = jl_rem_pio2(x, &y);
    return jl_tan_kernel_dd(y, (n & 1) ? -1 : 1);
}

/* cot(x) = inv(tan(x)) (Julia Base: cot(z) = inv(tan(z))) */
static inline double jl_cot(double x) { return 1.0 / jl_tan(x); }

#endif /* ORACLE_JULIA_LIBM_H_ */
