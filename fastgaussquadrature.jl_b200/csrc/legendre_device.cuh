// legendre_device.cuh — per-half-index device code of the Gauss-Legendre asymptotic path:
//   besselZeroRoots / besselJ1   src/besselroots.jl:168-245, tables src/constants.jl:13-36,93-106
//   legpts_nodes / legpts_weights   src/gausslegendre.jl:96-231
// Shared by the store kernels (legendre.cu), the fixed-menu consumer kernel and — as SOURCE TEXT compiled at run time
// with NVRTC — the user-integrand consumer kernels (integrate.cu): the file must stay free of host-only code.
#ifndef FGQ_LEGENDRE_DEVICE_CUH_
#define FGQ_LEGENDRE_DEVICE_CUH_

#include "fgq_math.h"

namespace fgq {

// src/constants.jl:13-36 — first twenty zeros of J0
static __constant__ double c_j0_roots[20] = {
    2.4048255576957728, 5.5200781102863106, 8.6537279129110122, 11.791534439014281,
    14.930917708487785, 18.071063967910922, 21.211636629879258, 24.352471530749302,
    27.493479132040254, 30.634606468431975, 33.775820213573568, 36.917098353664044,
    40.058425764628239, 43.199791713176730, 46.341188371661814, 49.482609897397817,
    52.624051841114996, 55.765510755019979, 58.906983926080942, 62.048469190227170};
// src/constants.jl:93-106 — J1(j0k)^2, k = 1..10
static __constant__ double c_j1sq[10] = {
    0.2695141239419169,  0.1157801385822037,  0.07368635113640822, 0.05403757319811628,
    0.04266142901724309, 0.03524210349099610, 0.03002107010305467, 0.02614739149530809,
    0.02315912182469139, 0.02078382912226786};

struct LegConsts {
    double vn, vn2, vn4, vn6, inv_vn2;
    double vn_q, inv_vn2_x4;  // vn / 4, 4 / vn^2 (exact scalings, see leg_bulk)
};

#if !defined(__CUDACC_RTC__)   // host side of the launches
static LegConsts make_consts(int64_t n) {
    LegConsts c;
    c.vn = 1.0 / ((double)n + 0.5);
    c.vn2 = c.vn * c.vn;
    c.vn4 = c.vn2 * c.vn2;
    c.vn6 = c.vn4 * c.vn2;
    c.inv_vn2 = 1.0 / c.vn2;
    c.vn_q = 0.25 * c.vn;
    c.inv_vn2_x4 = 4.0 * c.inv_vn2;
    return c;
}
#endif

// j0k and J1(j0k)^2 from the half-index k.  HEAD = the table / truncation bands of
// besselroots.jl:177-194 and :212-239 (k <= 13191); !HEAD = the pure McMahon tail
// (:195-198, :240-243), which is all a 5e8-index launch ever sees past its first warps.
template <bool HEAD>
__device__ __forceinline__ void bessel_pair(int64_t k, double& jk, double& j1sq) {
    const double ak = FGQ_PI_K * fgq_k_minus_quarter(k);
    if (!HEAD) {
        // k >= 13192 here: ak in [4e4, 2e9], no slow path needed
        const double q = 0.125 * fgq_rcp_fast(ak);
        jk = ak + q;
        j1sq = fgq_rcp_fast(FGQ_PI_K * ak) * 2.0;
        return;
    }
    const double rak = 1.0 / ak;
    const double q = 0.125 * rak;  // == 0.125 / ak bit-for-bit (power-of-two scaling)
    const double lead = 1.0 / (FGQ_PI * ak);
    const double p3 = -401743168.0 / 105.0, p5 = 120928.0 / 15.0, p7 = -124.0 / 3.0;
    if (k <= 20) {
        jk = c_j0_roots[k - 1];
    } else {
        const double ak82 = q * q;
        double e;
        if (k <= 47)
            e = fma(ak82, fma(ak82, fma(ak82, p3, p5), p7), 1.0);
        else if (k <= 344)
            e = fma(ak82, fma(ak82, p5, p7), 1.0);
        else if (k <= 13191)
            e = fma(ak82, p7, 1.0);
        else
            e = 1.0;
        jk = ak + q * e;
    }
    const double c1 = -171497088497.0 / 15206400.0, c2 = 461797.0 / 1152.0,
                 c3 = -172913.0 / 8064.0, c4 = 151.0 / 80.0, c5 = -7.0 / 24.0;
    if (k <= 10) {
        j1sq = c_j1sq[k - 1];
    } else if (k <= 2279) {
        const double ak2 = rak * rak;
        double e;
        if (k <= 15)
            e = fma(ak2, fma(ak2, fma(ak2, fma(ak2, c1, c2), c3), c4), c5);
        else if (k <= 21)
            e = fma(ak2, fma(ak2, fma(ak2, c2, c3), c4), c5);
        else if (k <= 55)
            e = fma(ak2, fma(ak2, c3, c4), c5);
        else if (k <= 279)
            e = fma(ak2, c4, c5);
        else
            e = c5;
        j1sq = lead * fma(e, ak2 * ak2, 2.0);
    } else {
        j1sq = lead * 2.0;
    }
}

// a/b and 1/b: IEEE operators on the head (tables, tiny launch), slow-path-free versions on the
// tail, where every operand is far inside the normal range (same bits: fgq_math.h).
template <bool FAST>
__device__ __forceinline__ double dv(double a, double b) { return FAST ? fgq_div_fast(a, b) : a / b; }
template <bool FAST>
__device__ __forceinline__ double rc(double b) { return FAST ? fgq_rcp_fast(b) : 1.0 / b; }

// theta-series, gausslegendre.jl:103-143.  NT = number of correction terms kept
// (3: n<=255, 2: n<=3950, 1: otherwise).  NT = 0 is the n >= 2^25 shortcut: there the
// single remaining correction (cot(a) - 1/a)/8 * vn^2 is below ulp(a)/2 for every k
// (|corr|/a <= vn^2/24 < 2^-54), so a + corr rounds back to a and theta == a exactly.
template <int NT, bool FAST = false>
__device__ __forceinline__ double leg_theta(double ai, double u, const LegConsts& c) {
    if (NT == 0) return ai;
    double node = ai + (u - rc<FAST>(ai)) / 8.0 * c.vn2;
    if (NT >= 2) {
        const double u2 = u * u;
        const double ai2 = ai * ai;
        const double ai3 = ai2 * ai;
        const double v1 =
            (6.0 * (1.0 + u2) / ai + 25.0 / ai3 - u * fma(31.0, u2, 33.0)) / 384.0;
        if (NT >= 3) {
            const double ai5 = ai2 * ai3;
            const double v2 =
                u * fma(u2, fma(u2, 3779.0 / 15360.0, 6350.0 / 15360.0), 2595.0 / 15360.0);
            const double v3 = (1.0 + u2) * (-fma(31.0 / 1024.0, u2, 11.0 / 1024.0) / ai +
                                            u / 512.0 / ai2 + -25.0 / 3072.0 / ai3);
            const double v4 = (v2 - 1073.0 / 5120.0 / ai5 + v3);
            node = fma(v1, c.vn4, node);
            node = fma(v4, c.vn6, node);
        } else {
            node = fma(v1, c.vn4, node);
        }
    }
    return node;
}

// W-series, gausslegendre.jl:162-224.  WT = number of terms (3: n<=170, 2: n<=1500,
// 1: n<=850000, 0: otherwise).
template <int WT, bool FAST = false>
__device__ __forceinline__ double leg_wseries(double ai, double u, const LegConsts& c) {
    if (WT == 0) return c.inv_vn2;
    const double ai2 = ai * ai;
    const double aim2 = rc<FAST>(ai2);
    const double ua = u * ai;
    const double W1 = fma(ua - 1.0, aim2, 1.0) / 8.0;
    if (WT == 1) return c.inv_vn2 + W1;
    const double u2 = u * u;
    const double E1 = fma(u2, fma(u2, -56.0, -84.0), -27.0);
    const double E2 = fma(-3.0, fma(u2, -2.0, 1.0), 6.0 * ua);
    const double E3 = fma(ua, -31.0, 81.0);
    const double W2 = fma(aim2, fma(aim2, E3, E2), E1) / 384.0;
    if (WT == 2) return fma(c.vn2, W2, c.inv_vn2 + W1);
    const double aim1 = 1.0 / ai;
    const double F1 =
        fma(u2, fma(u2, fma(u2, 151.0 / 160.0, 187.0 / 96.0), 295.0 / 256.0), 153.0 / 1024.0);
    const double F2 = fma(u2, fma(u2, -35.0 / 384.0, -119.0 / 768.0), -65.0 / 1024.0) * u;
    const double F3 = fma(u2, fma(u2, 7.0 / 384.0, 15.0 / 512.0), 5.0 / 512.0);
    const double F4 = fma(u2, 1.0 / 512.0, -13.0 / 1536.0) * u;
    const double F5 = fma(u2, -7.0 / 384.0, 53.0 / 3072.0);
    const double F6 = 3749.0 / 15360.0 * u;
    const double F7 = -1125.0 / 1024.0;
    const double W3 = fma(
        aim1, fma(aim1, fma(aim1, fma(aim1, fma(aim1, fma(aim1, F7, F6), F5), F4), F3), F2), F1);
    return fma(c.vn2, fma(c.vn2, W3, W2), c.inv_vn2 + W1);
}

// One half-index: cx = cos(theta_k) > 0, w = weight.  (asy :76-80, legpts_weights :225-228)
template <int NT, int WT, bool HEAD>
__device__ __forceinline__ void leg_half_index(int64_t k, const LegConsts& c, double& cx,
                                               double& w, double* theta_out = nullptr) {
    double jk, j1sq;
    bessel_pair<HEAD>(k, jk, j1sq);
    const double a = jk * c.vn;
    double s, ca;
    fgq_sincos_q1(a, &s, &ca);
    constexpr bool FAST = !HEAD && NT <= 1 && WT <= 1;  // tail launches of the large-n branches
    double u = 0.0;
    if (NT > 0 || WT > 0) u = rc<FAST>(fgq_tan_q1(a));  // cot(a) = inv(tan(a)), :100,158
    if (NT == 0) {
        cx = ca;
        if (theta_out) *theta_out = a;
    } else {
        const double th = leg_theta<NT, FAST>(a, u, c);
        if (theta_out) *theta_out = th;
        cx = fgq_cos_q1(th);
    }
    const double W = leg_wseries<WT, FAST>(a, u, c);
    if (FAST) {
        // 2/t == 2*(1/t) bit-for-bit (power-of-two scaling); operands far inside the normal range
        w = 2.0 * fgq_rcp_fast(j1sq * fgq_div_fast(a, s) * W);
    } else {
        w = 2.0 / (j1sq * (a / s) * W);
    }
}

// -x through the integer pipe (the FP64 pipe is the scarce one here)
__device__ __forceinline__ double neg_bits(double v) {
    return __hiloint2double(__double2hiint(v) ^ 0x80000000, __double2loint(v));
}

// x * 2^e through the integer pipe (x normal, result normal)
__device__ __forceinline__ double scale2(double x, int e) {
    return __hiloint2double(__double2hiint(x) + (e << 20), __double2loint(x));
}

// The bulk of an n >= 2^25 launch: half-indices k >= K_JK_EXACT, where
//   * 0.125/ak < ulp(ak)/2 (ak >= 2^25), so j0k = ak + 0.125/ak rounds to ak itself and the
//     first reciprocal disappears (bit-identical; tests/test_full_size_gpu.py checks every k);
//   * the whole block lies on one side of pi/4 (QUAD 0: theta < pi/4, no reduction at all;
//     QUAD 1: theta >= pi/4, always the two-constant reduction — the blocks next to the centre,
//     where theta ~= pi/2 takes the extended reduction, and the first half-index past pi/4, which
//     may be Float64(pi)/4 itself (returned unreduced by Julia's rem_pio2_kernel), stay on the
//     general path), so the quadrant selects disappear too.
// Same bits as leg_half_index<0,0,false>, ~20% fewer FP64 instructions.
constexpr int64_t K_JK_EXACT = 10680708;  // pi*(k - 1/4) >= 2^25

// All power-of-two factors of the formulas are folded into constants: scaling by 2^e commutes with rounding (no
// intermediate leaves the normal range), so with v = 4k - 1 (an exact double)
//   4 ak   = RN(pi v)                       a = RN(ak vn)          = RN(RN(pi v) (vn/4))
//   4 X    = RN(pi RN(pi v))                J1^2 = 2 RN(1/X)       = 8 RN(1/(4X))
//   w = RN(2 / RN(RN(J1^2 q) V)),  q = RN(a/s), V = 1/vn^2        = RN(1 / RN(RN(RN(1/(4X)) q) (4V)))
// bit for bit, without any exponent adjustments.
template <int QUAD>
__device__ __forceinline__ void leg_bulk_v(double v, const LegConsts& c, double& cx, double& w) {
    const double pv = FGQ_PI_K * v;
    const double t4 = fgq_rcp_fast(FGQ_PI_K * pv);
    const double a = pv * c.vn_q;
    double s;
    if (QUAD == 0) {
        s = fgq_jsin(a);
        cx = fgq_jcos(a);
    } else {
        double r, rt;
        fgq_rem_pio2_2c(a, &r, &rt);
        s = fgq_jcos_dd(r, rt);
        cx = neg_bits(fgq_jsin_dd(r, rt));
    }
    w = fgq_rcp_fast(t4 * fgq_div_fast(a, s) * c.inv_vn2_x4);
}

template <int QUAD>
__device__ __forceinline__ void leg_bulk(int64_t k, const LegConsts& c, double& cx, double& w) {
    leg_bulk_v<QUAD>(__ll2double_rn(4 * k - 1), c, cx, w);
}

}  // namespace fgq
#endif  // FGQ_LEGENDRE_DEVICE_CUH_
