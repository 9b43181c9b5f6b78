// laguerre.cu — Gauss-Laguerre nodes/weights by explicit asymptotic expansions, n >= 128, alpha < 5.
//
// Replaces, on the B200:
//   gausslaguerre dispatch              src/gausslaguerre.jl:59-96
//   gausslaguerre_asy (T = -1)          src/gausslaguerre.jl:110-261
//   gl_bulk / gl_bessel / gl_airy       src/gausslaguerre.jl:273-382
//   sum_adaptive, gl_bulk_solve_t       src/gausslaguerre.jl:400-427
//   region evaluators, compute_airy_root    src/gausslaguerre.jl:429-529
//   gl_rec_newton, evalLaguerreRec (recompute hook only)   src/gausslaguerre.jl:547-579,625-641
//
// The reference walks k = 1..n in one scalar loop and hands over Bessel -> bulk -> Airy at the first
// k where the next region's error estimate wins.  Every node's value depends only on (n, alpha, k)
// and on the region it falls in, so here the two hand-over indices are found by "first index
// where" reductions (atomicMin) over candidate ranges, then one thread per node evaluates the
// region(s) it needs.  Nodes whose error estimate exceeds 1e-13 are appended to a list and redone
// by Newton on the O(n) recurrence, one WARP per flagged node (about 35 nodes at n = 1e6).
// Launch structure of the asymptotic rule: the searches evaluate the expansions the edge nodes need anyway and run
// on a high-priority side stream beside the bulk-only nodes (90 % of the rule, independent of the searches); see
// laguerre_edge_eval_kernel.  FGQ_LAGUERRE_FAST=0 selects the sequential chain search, search, nodes.
#include <math.h>

#include <algorithm>
#include <memory>
#include <cooperative_groups.h>

#include <vector>

#include "airy.cuh"
#include "bessel.cuh"
#include "common.cuh"
#include "fgq_math.h"
#include "gw.cuh"
#include "jacobi_plan.h"

namespace cg = cooperative_groups;

namespace fgq {

struct LaguerreParams {
    int64_t n;
    double alpha;
    int alpha0;          // alpha == 0: the reference switches to its specialised formulas
    double d;            // 1 / (4n + 2 alpha + 2)
    int64_t k_airy;      // floor(0.9 n)
    double jak_head[20]; // approx_besselroots(alpha, .) for k <= head_count (table / Piessens)
    int head_count;
    double mc[7];        // McMahon coefficients a1..a13 for this alpha (besselroots.jl:79-111)
    double mu_m1;        // 4 alpha^2 - 1
    BesselOrder jorder;  // order alpha - 1
    ddv jinv_den[BESSEL_INV_DEN];  // besselj_inv_den_table(jorder.nu): term ratios of the ascending series
    double pn0;          // 1/sqrt(gamma(alpha+1))           (evalLaguerreRec :628)
    double wnorm;        // (n^2 + alpha n)^(-1/2)           (gl_rec_newton :575)
    double xmax;         // 4n + 2 alpha + 2
};

struct LaguerreState {
    unsigned long long k_b;   // first k where the bulk expansion beats the Bessel one (sentinel n+1)
    unsigned long long k_a;   // first k >= max(k_b+1, k_airy) where the Airy expansion beats the bulk one
    unsigned long long n_out; // reduced rule: number of nodes before the first underflowed weight
    unsigned int n_flagged;
    unsigned int warn;
    unsigned int n_flagged_b; // large rules: candidates flagged by the bulk-only launch (their own list)
    unsigned int fallback;    // large rules: the Bessel hand-over was not inside the search window
};

// large rules: the expansion a window / Airy-candidate node did NOT get into x, w (see laguerre_edge_eval_kernel)
struct __align__(32) LagSide {
    double x, w, delta;
    unsigned long long warn;
};

// Julia's max/min propagate NaN (CUDA's fmax/fmin drop it)
__device__ __forceinline__ double jl_max(double a, double b) { return (a != a || b != b) ? a + b : fmax(a, b); }
__device__ __forceinline__ double jl_min(double a, double b) { return (a != a || b != b) ? a + b : fmin(a, b); }

__device__ __forceinline__ double evalpoly2(double x, double c0, double c1) { return fma(x, c1, c0); }

template <int N>
__device__ __forceinline__ double evalpoly(double x, const double (&c)[N]) {
    double r = c[N - 1];
#pragma unroll
    for (int i = N - 2; i >= 0; --i) r = fma(x, r, c[i]);
    return r;
}

// sum_adaptive (:400-409)
template <int N>
__device__ __forceinline__ void sum_adaptive(const double (&v)[N], double& z, double& delta) {
    z = v[0];
    int i = 0;
    while (i < N - 1 && fabs(v[i + 1]) < fabs(v[i])) {
        i += 1;
        z += v[i];
    }
    delta = fabs(v[i + 1 < N ? i + 1 : N - 1]);
}

// besselroots.jl:74-116 with the alpha-only coefficients planned on the host
__device__ __forceinline__ double mcmahon_dev(const LaguerreParams& p, int64_t k) {
    const double b = 0.25 * (2 * p.alpha + 4 * (double)k - 1) * 3.141592653589793;
    const double b2 = b * b;
    return b - p.mu_m1 * (((((((p.mc[6] / b2 + p.mc[5]) / b2 + p.mc[4]) / b2 + p.mc[3]) / b2 + p.mc[2]) / b2 + p.mc[1]) / b2 + p.mc[0]) / b);
}

__device__ __forceinline__ double laguerre_jak(const LaguerreParams& p, int64_t k) {
    return k <= p.head_count ? p.jak_head[k - 1] : mcmahon_dev(p, k);
}

// approx_besselroots(nu, n) as a rule of its own (besselroots.jl:23-67): the first 20 (nu = 0, table) or 6
// (-1 <= nu <= 5, Piessens' Chebyshev series) zeros are parameter-only and planned on the host, every other
// zero is McMahon's expansion of its index — one thread per zero.  zero_all: nu < -1, where the reference
// returns zeros.
__global__ void besselroots_kernel(const LaguerreParams p, int zero_all, int64_t n, double* out) {
    const int64_t k = 1 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k > n) return;
    out[k - 1] = zero_all ? 0.0 : laguerre_jak(p, k);
}

// gl_bulk_solve_t (:411-427)
__device__ __forceinline__ double gl_bulk_solve_t(const LaguerreParams& p, int64_t k, unsigned int* warn) {
    const double PI = 3.141592653589793;
    const double pt = (double)(4 * p.n - 4 * k + 3) * p.d;
    double t = PI * PI / 16 * ((pt - 1) * (pt - 1));
    double diff = 100;
    int iter = 0;
    while (fabs(diff) > 100 * 2.220446049250313e-16 && iter < 20) {
        iter += 1;
        diff = (pt * PI + 2 * sqrt(t - t * t) - acos(2 * t - 1)) * sqrt(t / (1 - t)) / 2;
        t -= diff;
    }
    if (iter == 20 && warn) *warn |= FGQ_WARN_NEWTON_MAXITER;
    return t;
}

// gausslaguerre_asy_bulk / asy0_bulk (:429-465) with gl_bulk (:273-322)
__device__ void laguerre_bulk(const LaguerreParams& p, int64_t k, double& xk, double& wk, double& delta,
                              unsigned int* warn) {
    const double PI = 3.141592653589793;
    const double d = p.d, al = p.alpha;
    const double t = gl_bulk_solve_t(p, k, warn);
    const double tinv = 1.0 / (1 - t);
    const double _t = (1 - t) / t, _t2 = _t * _t, _t3 = _t2 * _t;
    const double d2 = d * d, d4 = d2 * d2, d6 = d4 * d2;
    double xs[4], ws[3];
    const double tm1 = t - 1;
    ws[0] = FGQ_DIV_LIT(d2, 6.0) * (2 * t + 3) / (tm1 * tm1 * tm1);
    if (p.alpha0) {
        { const double c[3] = {-4, -4, 5}; xs[0] = FGQ_DIV_LIT(-d, 12.0) * evalpoly(tinv, c); }
        { const double c[7] = {224, -336, -16, -576, 2814, -3815, 1600}; xs[1] = FGQ_DIV_LIT(d2 * d * _t, 720.0) * evalpoly(tinv, c); }
        { const double c[11] = {-285696, 714240, -516864, -13760, -17680, -1727136, 16131880, -48469876, 66424575, -43122800, 10797500};
          xs[2] = FGQ_DIV_LIT(-d4 * d, 181440.0) * _t2 * evalpoly(tinv, c); }
        { const double c[15] = {210677760, -737372160, 916264704, -426423552, -10338048, 2863616, -26285280, -3062050944.0,
                                43651480992.0, -212307298152.0, 518401904799.0, -714465642135.0, 566519158800.0, -241928673000.0, 43222750000.0};
          xs[3] = FGQ_DIV_LIT(d6 * d, 10886400.0) * _t3 * evalpoly(tinv, c); }
        { const double c[9] = {0, 0, 112, 32, 1712, -12408, 27517, -24860, 8000}; ws[1] = FGQ_DIV_LIT(d4, 720.0) * _t2 * evalpoly(tinv, c); }
        { const double c[13] = {0, 0, -71424, 159744, 20640, 28480, 4300160, -50986344, 201908326, -386872990, 393326325, -204917300, 43190000};
          ws[2] = FGQ_DIV_LIT(-d6 * _t3, 90720.0) * evalpoly(tinv, c); }
    } else {
        const double a2 = al * al;
        {
            const double c1 = evalpoly2(a2, -4, 12);
            const double c[3] = {c1, -4, 5};
            xs[0] = FGQ_DIV_LIT(-evalpoly(tinv, c) * d, 12.0);
        }
        {
            const double q1[3] = {224, -960, 480}, q2[3] = {7, -30, 15};
            const double c1 = evalpoly(a2, q1), c2 = evalpoly(a2, q2);
            const double c[7] = {c1, -48 * c2, -16, -576, 2814, -3815, 1600};
            xs[1] = FGQ_DIV_LIT(d2 * d * _t, 720.0) * evalpoly(tinv, c);
        }
        {
            const double q1[4] = {-285696, 1354752, -967680, 193536}, q2[4] = {-31, 147, -105, 21}, q3[4] = {-1346, 6405, -4620, 945};
            const double q4[3] = {43, -126, 63}, q5[3] = {-221, -630, 315};
            const double c1 = evalpoly(a2, q1), c2 = evalpoly(a2, q2), c3 = evalpoly(a2, q3), c4 = evalpoly(a2, q4), c5 = evalpoly(a2, q5);
            const double c[11] = {c1, -23040 * c2, 384 * c3, -320 * c4, 80 * c5, -1727136, 16131880, -48469876, 66424575, -43122800, 10797500};
            xs[2] = FGQ_DIV_LIT(-d4 * d, 181440.0) * _t2 * evalpoly(tinv, c);
        }
        {
            const double q1[5] = {210677760, -1028505600, 812851200, -232243200, 24883200}, q2[5] = {127, -620, 490, -140, 15};
            const double q3[5] = {1193053, -5826660, 4613070, -1324260, 143325}, q4[5] = {555239, -2716980, 2163630, -631260, 70875};
            const double q5[4] = {-641, 2960, -2155, 450}, q6[4] = {-1598, 17685, -13905, 3375};
            const double q7[3] = {-7823, -9042, 4521}, q8[3] = {15948182, -206850, 103425}, q9[3] = {64957561, -24000, 12000};
            const double c1 = evalpoly(a2, q1), c2 = evalpoly(a2, q2), c3 = evalpoly(a2, q3), c4 = evalpoly(a2, q4), c5 = evalpoly(a2, q5);
            const double c6 = evalpoly(a2, q6), c7 = evalpoly(a2, q7), c8 = evalpoly(a2, q8), c9 = evalpoly(a2, q9);
            const double c[15] = {c1, -5806080 * c2, 768 * c3, -768 * c4, 16128 * c5, -1792 * c6, 3360 * c7, -192 * c8, 672 * c9,
                                  -212307298152.0, 518401904799.0, -714465642135.0, 566519158800.0, -241928673000.0, 43222750000.0};
            xs[3] = FGQ_DIV_LIT(d6 * d, 10886400.0) * _t3 * evalpoly(tinv, c);
        }
        {
            const double q1[3] = {7, -30, 15};
            const double c1 = evalpoly(a2, q1);
            const double c[9] = {0, 0, 16 * c1, 32, 1712, -12408, 27517, -24860, 8000};
            ws[1] = FGQ_DIV_LIT(d4 * _t2, 720.0) * evalpoly(tinv, c);
        }
        {
            const double q1[4] = {-31, 147, -105, 21}, q2[4] = {-416, 1995, -1470, 315}, q3[3] = {43, -126, 63};
            const double q4[3] = {-89, -378, 189}, q5[3] = {53752, -630, 315};
            const double c1 = evalpoly(a2, q1), c2 = evalpoly(a2, q2), c3 = evalpoly(a2, q3), c4 = evalpoly(a2, q4), c5 = evalpoly(a2, q5);
            const double c[13] = {0, 0, 2304 * c1, -384 * c2, 480 * c3, -320 * c4, 80 * c5, -50986344, 201908326, -386872990, 393326325, -204917300, 43190000};
            ws[2] = FGQ_DIV_LIT(-d6 * _t3, 90720.0) * evalpoly(tinv, c);
        }
    }
    double xdelta, wdelta;
    sum_adaptive(xs, xk, xdelta);
    sum_adaptive(ws, wk, wdelta);
    xk += t / d;
    double wfactor = exp(-xk) * 2 * PI * sqrt(t / (1 - t));
    if (!p.alpha0) wfactor = pow(xk, al) * exp(-xk) * 2 * PI * sqrt(t / (1 - t));
    wk = wfactor * (1 + wk);
    wdelta *= wfactor;
    delta = jl_max(xdelta, wdelta);
}

// gausslaguerre_asy_bessel / asy0_bessel (:468-506) with gl_bessel (:325-367)
__device__ void laguerre_bessel(const LaguerreParams& p, double jak, double& xk, double& wk, double& delta) {
    const double d = p.d, al = p.alpha;
    const double j2 = jak * jak;
    const double d2 = d * d, d4 = d2 * d2, d8 = d4 * d4;
    double xdelta, wdelta;
    if (p.alpha0) {
        double xs[5], ws[5];
        { const double c[2] = {-2, 1}; xs[0] = FGQ_DIV_LIT(d2 * evalpoly(j2, c), 3.0); }
        { const double c[3] = {94, -57, 11}; xs[1] = FGQ_DIV_LIT(d4 * evalpoly(j2, c), 45.0); }
        { const double c[4] = {-48308, 28102, -6516, 657}; xs[2] = FGQ_DIV_LIT(d4 * d2 * evalpoly(j2, c), 2835.0); }
        { const double c[5] = {12059918, -6605817, 1456807, -172740, 10644}; xs[3] = FGQ_DIV_LIT(d8 * evalpoly(j2, c), 42525.0); }
        { const double c[6] = {-11427291076.0, 6028914206.0, -1248722004, 138902061, -9908262, 410649}; xs[4] = FGQ_DIV_LIT(d8 * d2 * evalpoly(j2, c), 1403325.0); }
        { const double c[2] = {-1, 1}; ws[0] = FGQ_DIV_LIT(2 * d2 * evalpoly(j2, c), 3.0); }
        { const double c[3] = {94, -114, 33}; ws[1] = FGQ_DIV_LIT(d4 * evalpoly(j2, c), 45.0); }
        { const double c[4] = {-12077, 14051, -4887, 657}; ws[2] = FGQ_DIV_LIT(4 * d4 * d2 * evalpoly(j2, c), 2835.0); }
        { const double c[5] = {12059918, -13211634, 4370421, -690960, 53220}; ws[3] = FGQ_DIV_LIT(d8 * evalpoly(j2, c), 42525.0); }
        { const double c[6] = {-5713645538.0, 6028914206.0, -1873083006, 277804122, -24770655, 1231947}; ws[4] = FGQ_DIV_LIT(2 * d8 * d2 * evalpoly(j2, c), 1403325.0); }
        sum_adaptive(xs, xk, xdelta);
        sum_adaptive(ws, wk, wdelta);
    } else {
        const double a2 = al * al;
        double xs[4], ws[4];
        {
            const double c1 = evalpoly2(a2, -2, 2);
            const double cx[2] = {c1, 1}, cw[2] = {c1, 2};
            xs[0] = FGQ_DIV_LIT(d2 * evalpoly(j2, cx), 3.0);
            ws[0] = FGQ_DIV_LIT(d2 * evalpoly(j2, cw), 3.0);
        }
        {
            const double q1[3] = {94, -140, 46};
            const double c1 = evalpoly(a2, q1), c2 = evalpoly2(a2, -19, 11);
            const double cx[3] = {c1, 3 * c2, 11}, cw[3] = {c1, 6 * c2, 33};
            xs[1] = FGQ_DIV_LIT(d4 * evalpoly(j2, cx), 45.0);
            ws[1] = FGQ_DIV_LIT(d4 * evalpoly(j2, cw), 45.0);
        }
        {
            const double q1[4] = {-12077, 19887, -9303, 1493}, q2[3] = {14051, -10750, 2459};
            const double c1 = evalpoly(a2, q1), c2 = evalpoly(a2, q2), c3 = evalpoly2(a2, -181, 73);
            const double cx[4] = {4 * c1, 2 * c2, 36 * c3, 657}, cw[4] = {c1, c2, 27 * c3, 657};
            xs[2] = FGQ_DIV_LIT(d4 * d2 * evalpoly(j2, cx), 2835.0);
            ws[2] = FGQ_DIV_LIT(4 * d4 * d2 * evalpoly(j2, cw), 2835.0);
        }
        {
            const double q1[5] = {6029959, -10087180, 5095482, -1146220, 107959}, q2[4] = {-2201939, 1678761, -507801, 63299};
            const double q3[3] = {1456807, -729422, 125671};
            const double c1 = evalpoly(a2, q1), c2 = evalpoly(a2, q2), c3 = evalpoly(a2, q3), c4 = evalpoly2(a2, -2879, 887);
            const double cx[5] = {2 * c1, 3 * c2, c3, 60 * c4, 10644}, cw[5] = {2 * c1, 6 * c2, 3 * c3, 240 * c4, 53220};
            xs[3] = FGQ_DIV_LIT(d8 * evalpoly(j2, cx), 42525.0);
            ws[3] = FGQ_DIV_LIT(d8 * evalpoly(j2, cw), 42525.0);
        }
        sum_adaptive(xs, xk, xdelta);
        sum_adaptive(ws, wk, wdelta);
    }
    const double xfactor = jak * jak * d;
    xk = xfactor * (1 + xk);
    xdelta *= xfactor;
    const double J = besselj_dev(p.jorder, jak, p.jinv_den);
    double wfactor = 4 * d * exp(-xk) / (J * J);
    if (!p.alpha0) wfactor = 4 * d * pow(xk, al) * exp(-xk) / (J * J);
    wk = wfactor * (1 + wk);
    wdelta *= wfactor;
    delta = jl_max(xdelta, wdelta);
}

__constant__ double c_lag_airy_roots[11] = {
    -2.338107410459767, -4.08794944413097, -5.520559828095551, -6.786708090071759,
    -7.944133587120853, -9.022650853340981, -10.04017434155809, -11.00852430373326,
    -11.93601556323626, -12.828776752865757, -13.69148903521072};

// gausslaguerre_asy_airy (:519-529), compute_airy_root (:508-517), gl_airy (:370-382)
__device__ void laguerre_airy(const LaguerreParams& p, int64_t k, double& xk, double& wk, double& delta) {
    const double PI = 3.141592653589793;
    const double d = p.d, al = p.alpha;
    const int64_t index = p.n - k + 1;
    double ak;
    if (index <= 11) {
        ak = c_lag_airy_roots[index - 1];
    } else {
        const double t = 3 * PI / 2 * ((double)index - 0.25);
        const double c[5] = {6967296, 725760, -967680, 6478500, -10856875};
        ak = -pow(t, 2.0 / 3.0) * evalpoly(1.0 / (t * t), c) / 6967296;
    }
    const double cbrt_d = cbrt(d), cbrt_d2 = cbrt_d * cbrt_d;
    const double ak2 = ak * ak, ak3 = ak2 * ak, ak4 = ak2 * ak2;
    double xs[3];
    xs[0] = 1 / d + ak * cbrt(4.0) / cbrt_d;
    xs[1] = ak2 * cbrt(16.0) / 5 * cbrt_d + (11.0 / 35 - al * al - 12.0 / 175 * ak3) * d +
            (16.0 / 1575 * ak + 92.0 / 7875 * ak4) * cbrt(4.0) * d * cbrt_d2;
    xs[2] = -(15152.0 / 3031875 * ak4 * ak + 1088.0 / 121275 * ak2) * cbrt(2.0) * d * d * cbrt_d;
    double xdelta;
    sum_adaptive(xs, xk, xdelta);
    double ai, aip;
    airy_neg(ak, ai, aip);
    wk = 1.5874010519681994 * pow(xk, al + 1.0 / 3.0) * exp(-xk) / (aip * aip);  // 4^(1/3)
    const double wdelta = fabs(wk);
    delta = jl_max(xdelta, wdelta);
}

constexpr int LAG_THREADS = 128;

__global__ void laguerre_state_init_kernel(LaguerreState* st, int64_t n) {
    st->k_b = (unsigned long long)(n + 1);
    st->k_a = (unsigned long long)(n + 1);
    st->n_out = (unsigned long long)n;
    st->n_flagged = 0;
    st->warn = 0;
    st->n_flagged_b = 0;
    st->fallback = 0;
}

// first k where delta_bulk < delta_bessel (:138-151), searched over k in [k_first, k_last] with a grid
// stride.  The hand-over happens at k = O(sqrt(n)) (the reference itself only prepares ceil(sqrt(n)) Bessel
// zeros before falling back to McMahon's expansion), so the search is staged: a first launch over a window
// of a few sqrt(n) — one evaluation per thread, few enough blocks that each warp has an SM sub-partition
// to itself — and a second launch over the rest, whose threads return at once when the window held it.
// (A single launch over all k had every resident thread, 150 000 of them, evaluate both expansions before
// the first result could stop anybody: 124 us at n = 1e6.)
__global__ void __launch_bounds__(LAG_THREADS)
laguerre_handover_bessel_kernel(const LaguerreParams p, LaguerreState* st, int64_t k_first, int64_t k_last) {
    const int64_t stride = (int64_t)gridDim.x * LAG_THREADS;
    for (int64_t k = k_first + (int64_t)blockIdx.x * LAG_THREADS + threadIdx.x; k <= k_last; k += stride) {
        if ((unsigned long long)k >= *((volatile unsigned long long*)&st->k_b)) return;  // already beaten by a smaller k
        double xb, wb, db, xu, wu, du;
        laguerre_bessel(p, laguerre_jak(p, k), xb, wb, db);
        laguerre_bulk(p, k, xu, wu, du, nullptr);
        if (du < db) atomicMin(&st->k_b, (unsigned long long)k);
    }
}

// first k >= max(k_b + 1, k_airy) where delta_airy < delta_bulk (:201-209)
__global__ void __launch_bounds__(LAG_THREADS)
laguerre_handover_airy_kernel(const LaguerreParams p, LaguerreState* st) {
    const int64_t kb = (int64_t)st->k_b;
    const int64_t start = kb + 1 > p.k_airy ? kb + 1 : p.k_airy;
    const int64_t k = (start < 1 ? 1 : start) + (int64_t)blockIdx.x * LAG_THREADS + threadIdx.x;
    if (k > p.n) return;
    double xu, wu, du, xa, wa, da;
    laguerre_bulk(p, k, xu, wu, du, nullptr);
    laguerre_airy(p, k, xa, wa, da);
    if (da < du) atomicMin(&st->k_a, (unsigned long long)k);
}

// The region walk of gausslaguerre_asy (:138-253) for node k with the hand-over indices known.
__device__ __forceinline__ void laguerre_node_walk(const LaguerreParams& p, int64_t k, int64_t kb, int64_t ka, double& xk,
                                                   double& wk, double& delta, unsigned int& warn) {
    if (k <= kb) {
        // loop A: both expansions are evaluated; Bessel is used except at k_b itself
        double xb, wb, db, xu, wu, du;
        laguerre_bessel(p, laguerre_jak(p, k), xb, wb, db);
        laguerre_bulk(p, k, xu, wu, du, &warn);
        const bool bulk = (k == kb);
        xk = bulk ? xu : xb;
        wk = bulk ? wu : wb;
        delta = jl_min(db, du);
    } else if (k < p.k_airy) {
        laguerre_bulk(p, k, xk, wk, delta, &warn);  // loop B
    } else if (k <= ka) {
        // loop C: bulk against Airy; Airy is used at k_a itself
        double xu, wu, du, xa, wa, da;
        laguerre_bulk(p, k, xu, wu, du, &warn);
        laguerre_airy(p, k, xa, wa, da);
        const bool airy = (k == ka);
        xk = airy ? xa : xu;
        wk = airy ? wa : wu;
        delta = jl_min(da, du);
    } else {
        laguerre_airy(p, k, xk, wk, delta);  // loop D
    }
}

// One thread per node.
__global__ void __launch_bounds__(LAG_THREADS)
laguerre_nodes_kernel(const LaguerreParams p, LaguerreState* st, double* x, double* w, double* delta_out,
                      unsigned int* flagged) {
    const int64_t k = 1 + (int64_t)blockIdx.x * LAG_THREADS + threadIdx.x;
    if (k > p.n) return;
    const int64_t kb = (int64_t)st->k_b;  // n + 1 when the Bessel expansion never loses
    const int64_t ka = (int64_t)st->k_a;  // n + 1 when the bulk expansion never loses
    unsigned int warn = 0;
    double xk, wk, delta;
    laguerre_node_walk(p, k, kb, ka, xk, wk, delta, warn);
    x[k - 1] = xk;
    w[k - 1] = wk;
    delta_out[k - 1] = delta;
    if (delta > 1.0e-13) flagged[atomicAdd(&st->n_flagged, 1u)] = (unsigned int)(k - 1);
    if (warn) atomicOr(&st->warn, warn);
}

// evalLaguerreRec (:625-641).  The loop stops early once everything is NaN (absorbing).
__device__ void eval_laguerre_rec(int64_t n, double alpha, double pn0, double x, double& pn_out, double& pnd_out) {
    double pnprev = 0.0, pn = pn0, pndprev = 0.0, pnd = 0.0;
    for (int64_t k = 1; k <= n; ++k) {
        const double dk = (double)k;
        const double pnold = pn;
        pn = (x - 2 * dk - alpha + 1) / sqrt(dk * (alpha + dk)) * pnold - sqrt((dk - 1 + alpha) * (dk - 1) / dk / (dk + alpha)) * pnprev;
        pnprev = pnold;
        const double pndold = pnd;
        pnd = (pnold + (x - 2 * dk - alpha + 1) * pndold) / sqrt(dk * (alpha + dk)) - sqrt((dk - 1 + alpha) * (dk - 1) / dk / (alpha + dk)) * pndprev;
        pndprev = pndold;
        if (pn != pn && pnprev != pnprev && pnd != pnd && pndprev != pndprev) break;
    }
    pn_out = pn;
    pnd_out = pnd;
}

// gl_rec_newton (:547-579)
template <class P>
__device__ void gl_rec_newton_dev(const P& p, double x0, bool computeweight, double& xk_out,
                                  double& wk_out, unsigned int* warn) {
    const double eps = 2.220446049250313e-16;
    double step = x0, xk = x0, xk_prev = x0, pn_prev = 1.7976931348623157e308, pn_deriv = 0.0;
    int iter = 0;
    while (fabs(step) > 40 * eps * xk && iter < 20) {
        iter += 1;
        double pn;
        eval_laguerre_rec(p.n, p.alpha, p.pn0, xk, pn, pn_deriv);
        if (fabs(pn) >= fabs(pn_prev) * (1 - 50 * eps)) {
            xk = xk_prev;
            break;
        }
        step = pn / pn_deriv;
        xk_prev = xk;
        xk -= step;
        pn_prev = pn;
    }
    if (xk < 0 || xk > p.xmax || iter == 20) *warn |= (unsigned int)FGQ_WARN_NEWTON_MAXITER;
    xk_out = xk;
    wk_out = 0.0;
    if (computeweight) {
        double pn_min1, dummy;
        eval_laguerre_rec(p.n - 1, p.alpha, p.pn0, xk, pn_min1, dummy);
        wk_out = p.wnorm / pn_min1 / pn_deriv;
    }
}

// evalLaguerreRec with a WARP per evaluation (every lane holds the same x and returns the same values).  One thread
// spends ~500 cycles per recurrence step, nearly all of it on the two square roots and four quotients of the step's
// coefficients, which do not depend on the running values: here lane j forms the coefficients of step k0 + j, and the
// 32 steps are then advanced (by every lane alike) with broadcasts — what remains in the dependent chain is two
// products and a difference for p_n, and a product, a sum, the quotient by sqrt(k (alpha + k)) and a difference for
// p_n'.  Operation for operation the expressions of eval_laguerre_rec, so the same bits.
__device__ void eval_laguerre_rec_warp(int64_t n, double alpha, double pn0, double x, double& pn_out, double& pnd_out) {
    const int lane = threadIdx.x & 31;
    double pnprev = 0.0, pn = pn0, pndprev = 0.0, pnd = 0.0;
    bool dead = false;
    for (int64_t k0 = 1; k0 <= n && !dead; k0 += 32) {
        const double dk = (double)(k0 + lane);
        const double c = x - 2 * dk - alpha + 1;
        const double sq = sqrt(dk * (alpha + dk));
        const double a = c / sq;
        const double b = sqrt((dk - 1 + alpha) * (dk - 1) / dk / (dk + alpha));
        const int m = n - k0 + 1 < 32 ? (int)(n - k0 + 1) : 32;
        for (int j = 0; j < m; ++j) {
            const double aj = __shfl_sync(0xffffffffu, a, j), bj = __shfl_sync(0xffffffffu, b, j);
            const double cj = __shfl_sync(0xffffffffu, c, j), sj = __shfl_sync(0xffffffffu, sq, j);
            const double pnold = pn;
            pn = aj * pnold - bj * pnprev;
            pnprev = pnold;
            const double pndold = pnd;
            pnd = (pnold + cj * pndold) / sj - bj * pndprev;
            pndprev = pndold;
            if (pn != pn && pnprev != pnprev && pnd != pnd && pndprev != pndprev) {
                dead = true;
                break;
            }
        }
    }
    pn_out = pn;
    pnd_out = pnd;
}

// gl_rec_newton (:547-579) on eval_laguerre_rec_warp: called by whole warps with lane-uniform arguments
template <class P>
__device__ void gl_rec_newton_warp(const P& p, double x0, double& xk_out, double& wk_out, unsigned int* warn) {
    const double eps = 2.220446049250313e-16;
    double step = x0, xk = x0, xk_prev = x0, pn_prev = 1.7976931348623157e308, pn_deriv = 0.0;
    int iter = 0;
    while (fabs(step) > 40 * eps * xk && iter < 20) {
        iter += 1;
        double pn;
        eval_laguerre_rec_warp(p.n, p.alpha, p.pn0, xk, pn, pn_deriv);
        if (fabs(pn) >= fabs(pn_prev) * (1 - 50 * eps)) {
            xk = xk_prev;
            break;
        }
        step = pn / pn_deriv;
        xk_prev = xk;
        xk -= step;
        pn_prev = pn;
    }
    if (xk < 0 || xk > p.xmax || iter == 20) *warn |= (unsigned int)FGQ_WARN_NEWTON_MAXITER;
    xk_out = xk;
    double pn_min1, dummy;
    eval_laguerre_rec_warp(p.n - 1, p.alpha, p.pn0, xk, pn_min1, dummy);
    wk_out = p.wnorm / pn_min1 / pn_deriv;
}

// the recompute hook (:153-162 etc.): gl_rec_newton for every flagged node, accepted only if it stays
// within 100 delta of the asymptotic value.  One warp per flagged node; first / stride count threads of whole warps.
__device__ __forceinline__ void laguerre_recompute_list(const LaguerreParams& p, LaguerreState* st, double* x, double* w,
                                                        const double* delta_in, const unsigned int* flagged,
                                                        unsigned int cnt, unsigned int first, unsigned int stride) {
    for (unsigned int f = first >> 5; f < cnt; f += stride >> 5) {
        const unsigned int i = flagged[f];
        const double x0 = x[i];
        const double delta = delta_in[i];
        unsigned int warn = 0;
        double xk, wk;
        gl_rec_newton_warp(p, x0, xk, wk, &warn);
        if ((threadIdx.x & 31) == 0) {
            if (warn) atomicOr(&st->warn, warn);
            if (fabs(xk - x0) < 100 * delta) {
                x[i] = xk;
                w[i] = wk;
            }
        }
        __syncwarp();
    }
}

// (the list of a large rule is void once the fall-back has been raised)
__global__ void laguerre_recompute_kernel(const LaguerreParams p, LaguerreState* st, double* x, double* w,
                                          const double* delta_in, const unsigned int* flagged) {
    if (st->fallback) return;
    laguerre_recompute_list(p, st, x, w, delta_in, flagged, st->n_flagged, blockIdx.x * blockDim.x + threadIdx.x,
                            gridDim.x * blockDim.x);
}

// ---- large rules: the hand-over searches beside the bulk nodes ----
// Whatever the two searches return, the nodes between the Bessel search window (a few sqrt(n)) and k_airy = 0.9 n
// are bulk nodes (loop B) — 90 % of the rule — so they do not have to wait for the searches.  Three launches replace
// search, search, nodes:
//   edge_eval    (high-priority side stream) window nodes k <= W: Bessel into x, w and bulk into `side`, k_b by
//                atomicMin; Airy candidates k >= k_airy: bulk into x, w and Airy into `side`, k_a by atomicMin — what
//                the two search launches computed and threw away
//   edge_select  (same stream) picks per edge node what the region walk takes, with k_b, k_a known; its flagged list
//                is recomputed on that stream at once
//   bulk_nodes   (the caller's stream, concurrently) nodes W < k < k_airy, their own flagged list
// and laguerre_finish_kernel closes the rule.  If the window did not hold k_b (never seen; FGQ_LAGUERRE_WINDOW forces
// it in the tests), edge_select raises `fallback`, the dependent launches return, and the finish kernel redoes the
// rule the sequential way.
__global__ void __launch_bounds__(LAG_THREADS)
laguerre_edge_eval_kernel(const LaguerreParams p, LaguerreState* st, int64_t window, unsigned window_blocks,
                          double* x, double* w, double* delta_out, LagSide* side) {
    if (blockIdx.x < window_blocks) {
        const int64_t k = 1 + (int64_t)blockIdx.x * LAG_THREADS + threadIdx.x;
        if (k > window) return;
        unsigned int warn = 0;
        double xb, wb, db, xu, wu, du;
        laguerre_bessel(p, laguerre_jak(p, k), xb, wb, db);
        laguerre_bulk(p, k, xu, wu, du, &warn);
        if (du < db) atomicMin(&st->k_b, (unsigned long long)k);
        x[k - 1] = xb;
        w[k - 1] = wb;
        delta_out[k - 1] = jl_min(db, du);
        LagSide s;
        s.x = xu; s.w = wu; s.delta = du; s.warn = warn;
        side[k - 1] = s;
    } else {
        const int64_t j = (int64_t)(blockIdx.x - window_blocks) * LAG_THREADS + threadIdx.x;
        const int64_t k = p.k_airy + j;  // k_airy > window >= 1 on this path
        if (k > p.n) return;
        unsigned int warn = 0;
        double xu, wu, du, xa, wa, da;
        laguerre_bulk(p, k, xu, wu, du, &warn);
        laguerre_airy(p, k, xa, wa, da);
        if (da < du) atomicMin(&st->k_a, (unsigned long long)k);
        x[k - 1] = xu;
        w[k - 1] = wu;
        delta_out[k - 1] = jl_min(da, du);
        LagSide s;
        s.x = xa; s.w = wa; s.delta = da; s.warn = warn;
        side[window + j] = s;
    }
}

__global__ void __launch_bounds__(256)
laguerre_edge_select_kernel(const LaguerreParams p, LaguerreState* st, int64_t window, double* x, double* w,
                            double* delta_out, const LagSide* side, unsigned int* flagged) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t kb = (int64_t)st->k_b, ka = (int64_t)st->k_a;
    if (kb > window) {  // the sentinel n + 1: not inside the window
        if (i == 0) st->fallback = 1;
        return;
    }
    const int64_t k = i < window ? i + 1 : p.k_airy + (i - window);
    if (k > p.n) return;
    const LagSide s = side[i];
    unsigned int warn = (unsigned int)s.warn;
    double delta = delta_out[k - 1];
    if (i < window) {
        // k < k_b: Bessel stays (loop A); k_b itself: bulk, delta = the minimum; beyond: loop B
        if (k >= kb) {
            x[k - 1] = s.x;
            w[k - 1] = s.w;
            if (k > kb) delta_out[k - 1] = delta = s.delta;
        }
    } else {
        // k < k_a: bulk stays (loop C); k_a itself: Airy, delta = the minimum; beyond: loop D, which never
        // evaluates the bulk expansion (so its Newton warning does not count)
        if (k >= ka) {
            x[k - 1] = s.x;
            w[k - 1] = s.w;
            if (k > ka) {
                delta_out[k - 1] = delta = s.delta;
                warn = 0;
            }
        }
    }
    if (delta > 1.0e-13) flagged[atomicAdd(&st->n_flagged, 1u)] = (unsigned int)(k - 1);
    if (warn) atomicOr(&st->warn, warn);
}

__global__ void __launch_bounds__(LAG_THREADS)
laguerre_bulk_nodes_kernel(const LaguerreParams p, LaguerreState* st, int64_t k_first, int64_t k_last, double* x,
                           double* w, double* delta_out, unsigned int* flagged_b) {
    const int64_t k = k_first + (int64_t)blockIdx.x * LAG_THREADS + threadIdx.x;
    if (k > k_last) return;
    unsigned int warn = 0;
    double xk, wk, delta;
    laguerre_bulk(p, k, xk, wk, delta, &warn);
    x[k - 1] = xk;
    w[k - 1] = wk;
    delta_out[k - 1] = delta;
    if (delta > 1.0e-13) flagged_b[atomicAdd(&st->n_flagged_b, 1u)] = (unsigned int)(k - 1);
    if (warn) atomicOr(&st->warn, warn);
}

// Last launch of a large rule (cooperative, one co-resident grid), three steps in one:
//   * only after edge_select raised `fallback`: the sequential form of the rule, with grid barriers where the
//     separate launches of the small-n path are;
//   * the recompute hook for the bulk-only launch's flagged list (normally empty);
//   * the reduced length (:163-169 ...): first index whose weight is below 10 floatmin.  Grid-stride from the low
//     end with a poll of the running minimum: ~98 % of the weights of a large rule underflow, so only the first
//     pass over the grid reads anything.
__global__ void __launch_bounds__(LAG_THREADS)
laguerre_finish_kernel(const LaguerreParams p, LaguerreState* st, int64_t window, double* x, double* w,
                       double* delta_out, unsigned int* flagged, const unsigned int* flagged_b) {
    cg::grid_group grid = cg::this_grid();
    const int64_t tid = (int64_t)blockIdx.x * LAG_THREADS + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * LAG_THREADS;
    if (st->fallback) {  // written by an earlier launch: the same for every thread
        if (tid == 0) {
            st->k_a = (unsigned long long)(p.n + 1);  // the speculative Airy search assumed k_b < k_airy
            st->n_flagged = 0;
            st->n_flagged_b = 0;
            st->warn = 0;
        }
        grid.sync();
        for (int64_t k = window + 1 + tid; k <= p.n; k += stride) {
            if ((unsigned long long)k >= *((volatile unsigned long long*)&st->k_b)) break;
            double xb, wb, db, xu, wu, du;
            laguerre_bessel(p, laguerre_jak(p, k), xb, wb, db);
            laguerre_bulk(p, k, xu, wu, du, nullptr);
            if (du < db) atomicMin(&st->k_b, (unsigned long long)k);
        }
        grid.sync();
        const int64_t kb = (int64_t)st->k_b;
        {
            const int64_t start = kb + 1 > p.k_airy ? kb + 1 : p.k_airy;
            for (int64_t k = start + tid; k <= p.n; k += stride) {
                double xu, wu, du, xa, wa, da;
                laguerre_bulk(p, k, xu, wu, du, nullptr);
                laguerre_airy(p, k, xa, wa, da);
                if (da < du) atomicMin(&st->k_a, (unsigned long long)k);
            }
        }
        grid.sync();
        const int64_t ka = (int64_t)st->k_a;
        for (int64_t k = 1 + tid; k <= p.n; k += stride) {
            unsigned int warn = 0;
            double xk, wk, delta;
            laguerre_node_walk(p, k, kb, ka, xk, wk, delta, warn);
            x[k - 1] = xk;
            w[k - 1] = wk;
            delta_out[k - 1] = delta;
            if (delta > 1.0e-13) flagged[atomicAdd(&st->n_flagged, 1u)] = (unsigned int)(k - 1);
            if (warn) atomicOr(&st->warn, warn);
        }
        grid.sync();
        laguerre_recompute_list(p, st, x, w, delta_out, flagged, st->n_flagged, (unsigned int)tid, (unsigned int)stride);
        grid.sync();
    } else if (st->n_flagged_b) {  // written by the bulk-only launch: the same for every thread
        laguerre_recompute_list(p, st, x, w, delta_out, flagged_b, st->n_flagged_b, (unsigned int)tid, (unsigned int)stride);
        grid.sync();
    }
    unsigned long long cand = ~0ull;
    for (int64_t i = tid; i < p.n; i += stride) {
        if ((unsigned long long)i >= *((volatile unsigned long long*)&st->n_out)) break;
        if (fabs(w[i]) < 10 * 2.2250738585072014e-308) {
            cand = (unsigned long long)i;
            break;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_down_sync(0xffffffffu, cand, o);
        cand = other < cand ? other : cand;
    }
    if ((threadIdx.x & 31) == 0 && cand < *((volatile unsigned long long*)&st->n_out)) atomicMin(&st->n_out, cand);
}

// Small rules, one thread (the work is inherently sequential and tiny):
//   n = 1, 2      closed forms                                   (:70-78)
//   n < 15        Golub-Welsch on the Jacobi matrix              (gausslaguerre_GW :535-542)
//   otherwise     gausslaguerre_rec (:582-618): Newton on the recurrence, node after node, each
//                 initial guess extrapolated from the seven previous nodes
struct LaguerreSmall {
    double g1;          // gamma(1 + alpha)
    double g2;          // gamma(alpha + 2)
    double jak_pre[7];  // approx_besselroots(alpha, min(n, 7))
};

template <class P>
__device__ void laguerre_small_rule(const P& p, const LaguerreSmall& q, unsigned int* warn_word, double* x, double* w) {
    const int n = (int)p.n;
    const double al = p.alpha;
    if (n == 1) {
        x[0] = 1 + al;
        w[0] = q.g1;
        return;
    }
    if (n == 2) {
        const double s = sqrt(al + 2);
        x[0] = al + 2 - s;
        x[1] = al + 2 + s;
        w[0] = ((al - s + 2) * q.g2) / (2 * (al + 2) * ((s - 1) * (s - 1)));
        w[1] = ((al + s + 2) * q.g2) / (2 * (al + 2) * ((s + 1) * (s + 1)));
        return;
    }
    if (n < 15) {
        double d[GW_MAX], e[GW_MAX], z[GW_MAX];
        for (int k = 1; k <= n; ++k) d[k - 1] = 2 * (double)k + (al - 1);
        for (int k = 1; k <= n - 1; ++k) e[k - 1] = sqrt((double)k * (al + (double)k));
        tridiag_ql_first_row(n, d, e, z);
        for (int i = 0; i < n; ++i) {
            x[i] = d[i];
            w[i] = q.g1 * (z[i] * z[i]);
        }
        return;
    }
    const int n_pre = n < 7 ? n : 7;
    const double nu = 4 * (double)n + 2 * al + 2;
    bool no_underflow = true;
    unsigned int warn = 0;
    for (int k = 1; k <= n; ++k) {
        double xk;
        if (k <= n_pre) {
            xk = q.jak_pre[k - 1] * q.jak_pre[k - 1] / nu;
        } else {
            xk = 7 * x[k - 2] - 21 * x[k - 3] + 35 * x[k - 4] - 35 * x[k - 5] + 21 * x[k - 6] - 7 * x[k - 7] + x[k - 8];
        }
        double wk;
        gl_rec_newton_dev(p, xk, no_underflow, xk, wk, &warn);
        if (no_underflow && fabs(wk) < 10 * 2.2250738585072014e-308) no_underflow = false;
        x[k - 1] = xk;
        w[k - 1] = wk;
    }
    if (warn) atomicOr(warn_word, warn);
}

__global__ void laguerre_small_kernel(const LaguerreParams p, const LaguerreSmall q, LaguerreState* st, double* x, double* w) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    laguerre_small_rule(p, q, &st->warn, x, w);
}

// Batched small rules (n < 128): the small-rule branches are sequential per rule (each
// Newton start is extrapolated from the previous nodes), so the batch runs one THREAD per rule through
// the same device function as the single-rule launch: bit-identical to the concatenation of single
// calls.  Only the fields the small-rule branches read travel with a rule.
struct LaguerreBatchRule {
    int64_t n;
    int64_t off;
    double alpha, pn0, wnorm, xmax;
    LaguerreSmall q;
};

constexpr int LAG_BATCH_THREADS = 32;

__global__ void __launch_bounds__(LAG_BATCH_THREADS)
laguerre_small_batch_kernel(int64_t nrules, const LaguerreBatchRule* __restrict__ rules, unsigned int* warn_word,
                            double* __restrict__ x, double* __restrict__ w) {
    const int64_t i = (int64_t)blockIdx.x * LAG_BATCH_THREADS + threadIdx.x;
    if (i >= nrules) return;
    const LaguerreBatchRule r = rules[i];
    if (r.n <= 0) return;
    laguerre_small_rule(r, r.q, warn_word, x + r.off, w + r.off);
}

// ---------------------------------------------------------------------------------------------
// The batched recurrence rule, table-driven (replaces one-thread-per-rule calls of laguerre_small_rule for n >= 15).
//
// gausslaguerre_rec is a strictly sequential chain per rule (each Newton start is extrapolated from the seven
// previous nodes; each Newton step is an n-step recurrence), ~5.5 n^2 recurrence steps, so the parallelism is across
// rules: one lane per rule.  What made the first version slow (25 ms per 1e5 rules, FP64 pipe 32 %) is that every
// recurrence step recomputed its abscissa-independent coefficients — two square roots and three of its five divisions.
// Here a lane tabulates c1_k = sqrt(k (alpha + k)), c2_k = sqrt((k-1+alpha)(k-1)/k/(k+alpha)) and the correctly
// rounded 1/c1_k of ITS rule once (same IEEE expressions as evalLaguerreRec, gausslaguerre.jl:631-637) in shared
// memory, laid out [k][lane] (conflict-free), and a step becomes ~21 FP64 operations: the two remaining quotients
// are the last three operations of the IEEE division's fast path against the correctly rounded reciprocal (same
// bits; an evaluation that leaves the safe range is repeated with the IEEE operators).  Lanes run a small state machine so that every trip of
// the warp-wide loop is one polynomial evaluation for every lane (no lane waits for another lane's Newton loop).
// The batch is cut into buckets of similar n (rules arrive sorted by n): shared memory per warp is 512 B x n_max of
// the bucket, so short rules get many resident warps and long ones as many as fit.
// ---------------------------------------------------------------------------------------------
struct LagTab {
    double* c1;   // [k][lane], k = 1..nmax
    double* c2;
};

// (pn, pnd) of evalLaguerreRec(deg, alpha, x) for this lane from its tables; straight-line per step (no branches in
// the dependent chain): a lane whose rule is shorter than the warp's longest keeps its state by selects.
// SAFE = false: the two quotients by c1 are the last three operations of the IEEE division's fast path against the
// correctly rounded 1/c1 (same bits as u / c1, v / c1 while v stays inside ~[1e-260, 1e255]); the exponent range of
// v is tracked on the integer pipe and reported through `out_of_range`, in which case the caller repeats the
// evaluation with SAFE = true (IEEE operators).
template <bool SAFE>
__device__ __forceinline__ void lag_eval_tab(const LagTab& t, int lane, int deg, double alpha, double pn0, double x,
                                             int deg_warp, double& pn_out, double& pnd_out, bool& out_of_range) {
    double pnprev = 0.0, pn = pn0, pndprev = 0.0, pnd = 0.0;
    unsigned bad = 0;
#pragma unroll 4
    for (int k = 1; k <= deg_warp; ++k) {
        const double dk = (double)k;
        const double c1 = t.c1[k * 32 + lane], c2 = t.c2[k * 32 + lane];
        const double u = x - 2 * dk - alpha + 1;
        const double pnold = pn, pndold = pnd;
        const double v = pnold + u * pndold;
        double q, qv;
        if (SAFE) {
            q = u / c1;
            qv = v / c1;
        } else {
            const double r1 = fgq_rcp_fast(c1);   // == 1.0 / c1 (c1 in [~0.3, ~200])
            q = u * r1;
            q = fma(r1, fma(-c1, q, u), q);
            qv = v * r1;
            qv = fma(r1, fma(-c1, qv, v), qv);
            const unsigned ev = ((unsigned)__double2hiint(v) >> 20) & 0x7ffu;
            bad |= (ev - 0x0a1u > 0x74fu - 0x0a1u) && (k <= deg) ? 1u : 0u;
        }
        const double pn_new = q * pnold - c2 * pnprev;
        const double pnd_new = qv - c2 * pndprev;
        const bool live = k <= deg;
        pnprev = live ? pnold : pnprev;
        pn = live ? pn_new : pn;
        pndprev = live ? pndold : pndprev;
        pnd = live ? pnd_new : pnd;
    }
    pn_out = pn;
    pnd_out = pnd;
    out_of_range = bad != 0;
}

__global__ void __launch_bounds__(32)
laguerre_rec_batch_kernel(int64_t first, int64_t count, int nmax, const LaguerreBatchRule* __restrict__ rules,
                          unsigned int* warn_word, double* __restrict__ x, double* __restrict__ w) {
    extern __shared__ double s_tab[];
    const int lane = threadIdx.x;
    const int64_t i = first + (int64_t)blockIdx.x * 32 + lane;
    const bool have = i < first + count;
    LaguerreBatchRule r = {};
    if (have) r = rules[i];
    const int n = have ? (int)r.n : 0;
    double* xr = x + (have ? r.off : 0);
    double* wr = w + (have ? r.off : 0);
    if (have && n > 0 && n < 15) {   // closed forms / Golub-Welsch: no recurrence rule, no tables
        laguerre_small_rule(r, r.q, warn_word, xr, wr);
    }
    const bool rec = have && n >= 15;
    LagTab t{s_tab, s_tab + (size_t)(nmax + 1) * 32};
    const double al = r.alpha;
    // (a bucket of closed-form / Golub-Welsch rules only, nmax < 15, is launched without shared memory)
    for (int k = 1; nmax >= 15 && k <= nmax; ++k) {   // entries beyond a lane's own n are read but never used: keep them finite
        const double dk = (double)k;
        const bool mine = rec && k <= n;
        t.c1[k * 32 + lane] = mine ? sqrt(dk * (al + dk)) : 1.0;
        t.c2[k * 32 + lane] = mine ? sqrt((dk - 1 + al) * (dk - 1) / dk / (dk + al)) : 0.0;
    }
    __syncwarp();
    const int n_warp = __reduce_max_sync(0xffffffffu, rec ? n : 0);
    if (n_warp == 0) return;
    // gausslaguerre_rec (:582-618) + gl_rec_newton (:547-579) as a per-lane state machine
    const double eps = 2.220446049250313e-16;
    const double nu = 4 * (double)n + 2 * al + 2;
    int k = 1;                       // node being solved (1-based)
    int phase = rec ? 0 : 3;         // 0 = Newton evaluation pending, 1 = weight evaluation pending, 3 = finished
    bool no_underflow = true;
    unsigned int warn = 0;
    double step = 0, xk = 0, xk_prev = 0, pn_prev = 0, pn_deriv = 0;
    int iter = 0;
    double h1 = 0, h2 = 0, h3 = 0, h4 = 0, h5 = 0, h6 = 0, h7 = 0;   // the seven previous nodes, h1 = x[k-1]
    auto start_node = [&]() {        // initial guess and the entry test of the Newton loop
        double x0;
        if (k <= 7) {
            x0 = r.q.jak_pre[k - 1] * r.q.jak_pre[k - 1] / nu;
        } else {
            x0 = 7 * h1 - 21 * h2 + 35 * h3 - 35 * h4 + 21 * h5 - 7 * h6 + h7;
        }
        step = x0; xk = x0; xk_prev = x0; pn_prev = 1.7976931348623157e308; pn_deriv = 0.0; iter = 0;
    };
    // after the Newton loop: the warning test, then either a weight evaluation or the end of the node
    auto finish_node = [&](double wk) {
        if (no_underflow && fabs(wk) < 10 * 2.2250738585072014e-308) no_underflow = false;
        xr[k - 1] = xk;
        wr[k - 1] = wk;
        h7 = h6; h6 = h5; h5 = h4; h4 = h3; h3 = h2; h2 = h1; h1 = xk;
        k += 1;
        if (k > n) { phase = 3; return; }
        start_node();
        phase = 0;
    };
    auto after_newton = [&]() {
        if (xk < 0 || xk > r.xmax || iter == 20) warn |= (unsigned int)FGQ_WARN_NEWTON_MAXITER;
        if (no_underflow) phase = 1; else finish_node(0.0);
    };
    // phase 0 requires the loop test to hold; settle states that need no evaluation
    auto settle = [&]() {
        while (phase == 0 && !(fabs(step) > 40 * eps * xk && iter < 20)) after_newton();
    };
    if (rec) {
        start_node();
        settle();
    }
    while (__any_sync(0xffffffffu, phase != 3)) {
        const bool act = phase != 3;
        const int deg = phase == 1 ? n - 1 : n;
        double pn, pnd;
        bool oor;
        lag_eval_tab<false>(t, lane, act ? deg : 0, al, r.pn0, xk, n_warp, pn, pnd, oor);
        if (__any_sync(0xffffffffu, act && oor)) {   // a value outside the fast quotient's range: IEEE operators
            double pn_s, pnd_s;
            bool dummy;
            lag_eval_tab<true>(t, lane, act ? deg : 0, al, r.pn0, xk, n_warp, pn_s, pnd_s, dummy);
            if (oor) {
                pn = pn_s;
                pnd = pnd_s;
            }
        }
        if (act) {
            if (phase == 0) {
                iter += 1;
                pn_deriv = pnd;
                if (fabs(pn) >= fabs(pn_prev) * (1 - 50 * eps)) {
                    xk = xk_prev;
                    after_newton();
                } else {
                    step = pn / pn_deriv;
                    xk_prev = xk;
                    xk -= step;
                    pn_prev = pn;
                }
                settle();
            } else {
                finish_node(r.wnorm / pn / pn_deriv);
                settle();
            }
        }
    }
    if (warn) atomicOr(warn_word, warn);
}

// reduced rule (:163-169 ...): the rule ends before the first weight below 10 floatmin
__global__ void laguerre_reduced_kernel(int64_t n, const double* w, LaguerreState* st) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // ~98 % of the weights underflow at n = 1e6: one candidate per BLOCK (warp shuffles, then shared memory),
    // and blocks that cannot lower the minimum any more skip the atomic — instead of a million atomics, or
    // 31 000 polls, on one address
    unsigned long long cand = ~0ull;
    if (i < n && fabs(w[i]) < 10 * 2.2250738585072014e-308) cand = (unsigned long long)i;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_down_sync(0xffffffffu, cand, o);
        cand = other < cand ? other : cand;
    }
    __shared__ unsigned long long s_cand[32];
    if ((threadIdx.x & 31) == 0) s_cand[threadIdx.x >> 5] = cand;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (unsigned q = 1; q < (blockDim.x + 31) / 32; ++q) cand = s_cand[q] < cand ? s_cand[q] : cand;
        if (cand < *((volatile unsigned long long*)&st->n_out)) atomicMin(&st->n_out, cand);
    }
}

}  // namespace fgq

using namespace fgq;

constexpr int MAX_ATTR_DEV_LAG = 16;
constexpr int64_t LAG_REC_MAX = 4096;  // sequential recurrence rule: one thread, O(n^2) dependent steps

static int laguerre_check(int64_t n, double alpha) {
    if (!(alpha > -1.0)) {
        set_error("DomainError(%g): The Laguerre parameter alpha <= -1 corresponds to a nonintegrable weight function", alpha);
        return FGQ_EDOMAIN;
    }
    if (n < 0) {
        set_error("DomainError(%lld): gausslaguerre(%lld,%g) not defined: n must be positive.", (long long)n, (long long)n, alpha);
        return FGQ_EDOMAIN;
    }
    if (alpha >= 5.0 && n > LAG_REC_MAX) {
        set_error("gausslaguerre(%lld,%g): alpha >= 5 is served by the O(n^2) recurrence rule in the reference "
                  "(gausslaguerre.jl:93-94); the device path takes it up to n = %lld", (long long)n, alpha, (long long)LAG_REC_MAX);
        return FGQ_ENOTIMPL;
    }
    if (n > 4000000000LL) {
        set_error("gausslaguerre: n too large");
        return FGQ_EINVAL;
    }
    return FGQ_OK;
}

// for_rule: the parameters of gausslaguerre_asy (false: of a plain approx_besselroots(nu, .) call)
static LaguerreParams laguerre_params(int64_t n, double alpha, bool for_rule = true) {
    LaguerreParams p;
    p.n = n;
    p.alpha = alpha;
    p.alpha0 = alpha == 0.0;
    p.d = 1.0 / (4 * (double)n + 2 * alpha + 2);
    p.k_airy = (int64_t)floor(0.9 * (double)n);
    p.xmax = 4 * (double)n + 2 * alpha + 2;
    // approx_besselroots(alpha, k_bessel): only the first 20 (alpha = 0) / 6 entries are not McMahon
    p.head_count = alpha == 0.0 ? 20 : ((-1 <= alpha && alpha <= 5) ? 6 : 0);
    if (p.head_count) approx_besselroots(alpha, p.head_count, p.jak_head);
    if (for_rule) {
        // gausslaguerre_asy takes jak_vector[k] only for k < k_bessel = max(ceil(sqrt(n)), 7) and McMahon(alpha, k)
        // from k_bessel on (gausslaguerre.jl:131,144): for alpha = 0 and n < 441 that cuts into the 20 tabulated zeros
        const int64_t k_bessel = std::max<int64_t>((int64_t)ceil(sqrt((double)n)), 7);
        if (p.head_count > k_bessel - 1) p.head_count = (int)(k_bessel - 1);
    }
    const double mu = 4 * alpha * alpha;
    p.mu_m1 = mu - 1;
    p.mc[0] = 1.0 / 8;
    p.mc[1] = (7 * mu - 31) / 384;
    p.mc[2] = 4 * (3779 + mu * (-982 + 83 * mu)) / 61440;
    p.mc[3] = 6 * (-6277237 + mu * (1585743 + mu * (-153855 + 6949 * mu))) / 20643840;
    p.mc[4] = 144 * (2092163573.0 + mu * (-512062548 + mu * (48010494 + mu * (-2479316 + 70197 * mu)))) / 11890851840.0;
    p.mc[5] = 720 * (-8249725736393.0 + mu * (1982611456181.0 + mu * (-179289628602.0 + mu * (8903961290.0 + mu * (-287149133 + mu * 5592657))))) / 10463949619200.0;
    p.mc[6] = 576 * (423748443625564327.0 + mu * (-100847472093088506.0 + mu * (8929489333108377.0 + mu * (-426353946885548.0 + mu * (13172003634537.0 + mu * (-291245357370.0 + mu * 4148944183.0)))))) / 13059009124761600.0;
    // order alpha - 1 of besselj(alpha - 1, jak) (:484); alpha = 0 -> J_{-1} = -J_1 (:501)
    double nu = alpha - 1;
    p.jorder.sign = 1.0;
    if (nu < 0 && nu == floor(nu)) {
        const double m = -nu;
        p.jorder.sign = fmod(m, 2.0) == 0.0 ? 1.0 : -1.0;
        nu = m;
    }
    p.jorder.nu = nu;
    p.jorder.rgamma = 1.0 / tgamma(nu + 1);
    const double c = (nu / 2 + 0.25) * 3.141592653589793;
    p.jorder.cosc = cos(c);
    p.jorder.sinc = sin(c);
    besselj_inv_den_table(nu, p.jinv_den);
    p.pn0 = 1 / sqrt(tgamma(alpha + 1));
    p.wnorm = pow((double)n * (double)n + alpha * (double)n, -0.5);
    return p;
}

// The asymptotic rule (n >= 128): the Bessel hand-over is searched in a window of a few sqrt(n) (it sits at
// ~1.8 sqrt(n)) that ends below k_airy = floor(0.9 n); for small n the window is everything below k_airy and there
// are no bulk-only nodes.
constexpr int64_t LAG_FAST_MIN = 128;

static int64_t laguerre_window(int64_t n) {
    int64_t window = 8 * (int64_t)ceil(sqrt((double)n));
    if (window < 1024) window = 1024;
    const int64_t k_airy = (int64_t)floor(0.9 * (double)n);
    if (window > k_airy - 1) window = k_airy - 1;
    if (window < 1) window = 1;
    return window;
}

// FGQ_LAGUERRE_FAST=0: the sequential launch chain for every n (regression checks)
static bool laguerre_use_fast() {
    static const bool on = []() {
        const char* e = getenv("FGQ_LAGUERRE_FAST");
        if (e && atoi(e) == 0) return false;
        int dev = 0, coop = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return false;
        if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess) return false;
        return coop != 0;
    }();
    return on;
}

extern "C" int64_t fgq_gausslaguerre_workspace_bytes(int64_t n) {
    if (n < 0) return 0;
    int64_t bytes = 256 + 8 * n + 4 * n + 64;
    if (n >= LAG_FAST_MIN) {
        const int64_t k_airy = (int64_t)floor(0.9 * (double)n);
        bytes += 32 + (int64_t)sizeof(LagSide) * (laguerre_window(n) + (n - k_airy + 1));
    }
    return bytes;
}

extern "C" int fgq_gausslaguerre_device(int64_t n, double alpha, double* x_dev, double* w_dev,
                                        void* workspace_dev, void* stream) {
    clear_status();
    FGQ_TRY(laguerre_check(n, alpha));
    if (n == 0) return FGQ_OK;
    if (!x_dev || !w_dev || !workspace_dev) {
        set_error("null pointer");
        return FGQ_EINVAL;
    }
    if (alpha * alpha / (double)n > 1) add_warning(FGQ_WARN_LARGE_ALPHA);  // :117-119
    cudaStream_t st = as_stream(stream);
    const LaguerreParams p = laguerre_params(n, alpha);
    LaguerreState* state = (LaguerreState*)workspace_dev;
    double* delta = (double*)((char*)workspace_dev + 256);
    unsigned int* flagged = (unsigned int*)(delta + n);
    laguerre_state_init_kernel<<<1, 1, 0, st>>>(state, n);  // a kernel, not a memcpy: stays graph-capturable
    FGQ_LAUNCH_CHECK();
    if (n == 0) return FGQ_OK;
    if (n < 128 || alpha >= 5.0) {  // gausslaguerre.jl:70-95: closed forms, Golub-Welsch, recurrence rule
        LaguerreSmall q;
        q.g1 = tgamma(1 + alpha);
        q.g2 = tgamma(alpha + 2);
        for (int i = 0; i < 7; ++i) q.jak_pre[i] = 0.0;
        approx_besselroots(alpha, (int)(n < 7 ? n : 7), q.jak_pre);
        laguerre_small_kernel<<<1, 32, 0, st>>>(p, q, state, x_dev, w_dev);
        FGQ_LAUNCH_CHECK();
        laguerre_reduced_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, w_dev, state);
        FGQ_LAUNCH_CHECK();
        return FGQ_OK;
    }
    const unsigned blocks = (unsigned)((n + LAG_THREADS - 1) / LAG_THREADS);
    int64_t window = laguerre_window(n);
    if (n >= LAG_FAST_MIN && laguerre_use_fast()) {
        // FGQ_LAGUERRE_WINDOW=<count>: a smaller search window (the tests use it to reach the fall-back)
        static const int64_t window_env = []() {
            const char* e = getenv("FGQ_LAGUERRE_WINDOW");
            return e ? (int64_t)atoll(e) : (int64_t)0;
        }();
        if (window_env > 0 && window_env < window) window = window_env;
        const int64_t cand = n - p.k_airy + 1;
        LagSide* side = (LagSide*)(((uintptr_t)(flagged + n) + 31) & ~(uintptr_t)31);
        unsigned int* flagged_b = flagged + (window + cand);
        ScopedStream edge_holder;
        ScopedEvents evs;
        FGQ_TRY(edge_holder.acquire(true));
        FGQ_TRY(evs.acquire(2));
        cudaStream_t edge = edge_holder.st;
        FGQ_CUDA_TRY(cudaEventRecord(evs.ev[0], st));  // after the state reset
        FGQ_CUDA_TRY(cudaStreamWaitEvent(edge, evs.ev[0], 0));
        const unsigned window_blocks = (unsigned)((window + LAG_THREADS - 1) / LAG_THREADS);
        const unsigned cand_blocks = (unsigned)((cand + LAG_THREADS - 1) / LAG_THREADS);
        laguerre_edge_eval_kernel<<<window_blocks + cand_blocks, LAG_THREADS, 0, edge>>>(p, state, window, window_blocks,
                                                                                      x_dev, w_dev, delta, side);
        FGQ_LAUNCH_CHECK();
        const int64_t nb = p.k_airy - 1 - window;  // bulk-only nodes window < k < k_airy
        if (nb > 0) {
            laguerre_bulk_nodes_kernel<<<(unsigned)((nb + LAG_THREADS - 1) / LAG_THREADS), LAG_THREADS, 0, st>>>(
                p, state, window + 1, p.k_airy - 1, x_dev, w_dev, delta, flagged_b);
            FGQ_LAUNCH_CHECK();
        }
        laguerre_edge_select_kernel<<<(unsigned)((window + cand + 255) / 256), 256, 0, edge>>>(p, state, window, x_dev, w_dev,
                                                                                            delta, side, flagged);
        FGQ_LAUNCH_CHECK();
        laguerre_recompute_kernel<<<64, 64, 0, edge>>>(p, state, x_dev, w_dev, delta, flagged);
        FGQ_LAUNCH_CHECK();
        FGQ_CUDA_TRY(cudaEventRecord(evs.ev[1], edge));
        FGQ_CUDA_TRY(cudaStreamWaitEvent(st, evs.ev[1], 0));
        {
            static int grid_cache[MAX_ATTR_DEV_LAG] = {0};
            int dev = 0;
            FGQ_CUDA_TRY(cudaGetDevice(&dev));
            int grid = dev < MAX_ATTR_DEV_LAG ? grid_cache[dev] : 0;
            if (grid == 0) {
                int per_sm = 0, sms = 0;
                FGQ_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, laguerre_finish_kernel, LAG_THREADS, 0));
                FGQ_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
                grid = per_sm * sms;
                if (grid < 1) grid = 1;
                if (dev < MAX_ATTR_DEV_LAG) grid_cache[dev] = grid;
            }
            const LaguerreParams* pp = &p;
            void* args[] = {(void*)pp, (void*)&state, (void*)&window, (void*)&x_dev, (void*)&w_dev, (void*)&delta,
                            (void*)&flagged, (void*)&flagged_b};
            FGQ_CUDA_TRY(cudaLaunchCooperativeKernel((void*)laguerre_finish_kernel, dim3((unsigned)grid),
                                                     dim3(LAG_THREADS), args, 0, st));
            count_launch();
        }
        return FGQ_OK;
    }
    {
        laguerre_handover_bessel_kernel<<<(unsigned)((window + LAG_THREADS - 1) / LAG_THREADS), LAG_THREADS, 0, st>>>(
            p, state, 1, window);
        FGQ_LAUNCH_CHECK();
        if (window < n) {
            const int64_t rest = (n - window + LAG_THREADS - 1) / LAG_THREADS;
            laguerre_handover_bessel_kernel<<<(unsigned)(rest < 148 * 6 ? rest : 148 * 6), LAG_THREADS, 0, st>>>(
                p, state, window + 1, n);
            FGQ_LAUNCH_CHECK();
        }
    }
    // candidates k >= k_airy (the kernel skips those <= k_b)
    const int64_t cand = n - (p.k_airy < 1 ? 1 : p.k_airy) + 1;
    laguerre_handover_airy_kernel<<<(unsigned)((cand + LAG_THREADS - 1) / LAG_THREADS), LAG_THREADS, 0, st>>>(p, state);
    FGQ_LAUNCH_CHECK();
    laguerre_nodes_kernel<<<blocks, LAG_THREADS, 0, st>>>(p, state, x_dev, w_dev, delta, flagged);
    FGQ_LAUNCH_CHECK();
    laguerre_recompute_kernel<<<64, 64, 0, st>>>(p, state, x_dev, w_dev, delta, flagged);
    FGQ_LAUNCH_CHECK();
    laguerre_reduced_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, w_dev, state);
    FGQ_LAUNCH_CHECK();
    return FGQ_OK;
}

// k_b, k_a (n+1 = never), number of recomputed candidates, reduced length; synchronises.
extern "C" int fgq_gausslaguerre_stats(const void* workspace_dev, int64_t* k_b, int64_t* k_a, int64_t* n_flagged,
                                       int64_t* n_reduced) {
    clear_status();
    if (!workspace_dev) return FGQ_EINVAL;
    LaguerreState h;
    FGQ_CUDA_TRY(cudaMemcpy(&h, workspace_dev, sizeof(h), cudaMemcpyDeviceToHost));
    if (k_b) *k_b = (int64_t)h.k_b;
    if (k_a) *k_a = (int64_t)h.k_a;
    if (n_flagged) *n_flagged = (int64_t)h.n_flagged + (int64_t)h.n_flagged_b;
    if (n_reduced) *n_reduced = (int64_t)h.n_out;
    if (h.warn) add_warning((int)h.warn);
    return FGQ_OK;
}

extern "C" int fgq_gausslaguerre_host(int64_t n, double alpha, int reduced, double* x, double* w, int64_t* n_out) {
    clear_status();
    FGQ_TRY(laguerre_check(n, alpha));
    if (n == 0) {
        if (n_out) *n_out = 0;
        return FGQ_OK;
    }
    if (!x || !w) {
        set_error("null output pointer");
        return FGQ_EINVAL;
    }
    HostPathLock lock;
    void *ws, *out;
    FGQ_TRY(scratch_device((size_t)fgq_gausslaguerre_workspace_bytes(n), 6, &ws));
    FGQ_TRY(scratch_device((size_t)n * 16, 0, &out));
    cudaStream_t st;
    FGQ_TRY(scratch_stream(0, &st));
    double* xd = (double*)out;
    double* wd = xd + n;
    FGQ_TRY(fgq_gausslaguerre_device(n, alpha, xd, wd, ws, (void*)st));
    const int warn0 = fgq_warnings();
    LaguerreState h;
    FGQ_CUDA_TRY(cudaMemcpyAsync(&h, ws, sizeof(h), cudaMemcpyDeviceToHost, st));
    FGQ_CUDA_TRY(cudaStreamSynchronize(st));
    // `reduced` exists only inside gausslaguerre_asy (gausslaguerre.jl:88-95, 163-169): the closed-form, Golub-Welsch
    // and recurrence branches (n < 128 or alpha >= 5) always return all n entries, underflowed weights included
    const bool asy_branch = n >= 128 && alpha < 5;
    const int64_t cnt = (reduced && asy_branch) ? (int64_t)h.n_out : n;
    FGQ_TRY(drain_to_host(x, w, xd, wd, cnt, st));
    if (n_out) *n_out = cnt;
    add_warning(warn0 | (int)h.warn);
    return FGQ_OK;
}

// ---- batched small rules: n < 128 ----
constexpr int LAGUERRE_BATCH_MAX_N = 127;
constexpr int MAX_ATTR_DEV = 16;

extern "C" int64_t fgq_gausslaguerre_batch_workspace_bytes(int64_t nrules) {
    if (nrules < 0) return 0;
    return 256 + nrules * (int64_t)sizeof(LaguerreBatchRule);
}

// Buckets of the batch kernel over rules sorted by n (descending): a bucket holds the rules with n in
// (16 floor((n_top - 1)/16), n_top] (everything up to 16 in the last one); rules with n <= 0 sort last and get none.
struct LagBucket {
    int64_t first, count;
    int n_top;
};
static std::vector<LagBucket> laguerre_batch_buckets(const std::vector<int>& sorted_n) {
    std::vector<LagBucket> out;
    const int64_t nr = (int64_t)sorted_n.size();
    int64_t j = 0;
    while (j < nr) {
        const int n_top = sorted_n[(size_t)j];
        if (n_top <= 0) break;
        const int n_floor = n_top <= 16 ? 0 : ((n_top - 1) / 16) * 16;
        int64_t e = j;
        while (e < nr && sorted_n[(size_t)e] > n_floor) ++e;
        out.push_back(LagBucket{j, e - j, n_top});
        j = e;
    }
    return out;
}

// test hook (host only): OrderedPlanner's contract
extern "C" int64_t fgq_debug_ordered_planner(int64_t n, int64_t wait_step) {
    if (n <= 0 || wait_step <= 0) return -1;
    std::vector<std::atomic<int>> hits((size_t)n);
    for (auto& h : hits) h.store(0, std::memory_order_relaxed);
    int64_t bad = 0;
    {
        std::atomic<int>* hp = hits.data();
        OrderedPlanner planner(n, [hp](int64_t lo, int64_t hi) {
            for (int64_t s = lo; s < hi; ++s) hp[s].fetch_add(1, std::memory_order_relaxed);
        });
        for (int64_t hi = wait_step; hi < n + wait_step; hi += wait_step) {
            const int64_t bound = hi < n ? hi : n;
            planner.wait(bound);
            for (int64_t s = bound - wait_step < 0 ? 0 : bound - wait_step; s < bound; ++s)
                if (hits[(size_t)s].load(std::memory_order_relaxed) != 1) ++bad;
        }
    }
    for (int64_t s = 0; s < n; ++s)
        if (hits[(size_t)s].load(std::memory_order_relaxed) != 1) ++bad;
    return bad;
}

// test hook (host only): the buckets for an UNSORTED list of rule sizes; returns their number
extern "C" int fgq_debug_laguerre_batch_buckets(int64_t nrules, const int32_t* n_i, int64_t* first, int64_t* count,
                                                int32_t* n_top, int max_buckets) {
    std::vector<int> v;
    for (int64_t i = 0; i < nrules; ++i) v.push_back(n_i[i]);
    std::sort(v.begin(), v.end(), [](int a, int b) { return a > b; });
    const std::vector<LagBucket> b = laguerre_batch_buckets(v);
    for (size_t i = 0; i < b.size() && (int)i < max_buckets; ++i) {
        first[i] = b[i].first;
        count[i] = b[i].count;
        n_top[i] = b[i].n_top;
    }
    return (int)b.size();
}

// the per-rule constants (3 tgamma, pow, the Bessel-zero guesses): the same expressions as laguerre_params /
// fgq_gausslaguerre_device
static LaguerreBatchRule laguerre_batch_rule(int64_t n, double alpha, int64_t off) {
    LaguerreBatchRule r;
    r.n = n;
    r.off = off;
    r.alpha = alpha;
    r.pn0 = 1 / sqrt(tgamma(alpha + 1));
    r.wnorm = pow((double)n * (double)n + alpha * (double)n, -0.5);
    r.xmax = 4 * (double)n + 2 * alpha + 2;
    r.q.g1 = tgamma(1 + alpha);
    r.q.g2 = tgamma(alpha + 2);
    for (int k = 0; k < 7; ++k) r.q.jak_pre[k] = 0.0;
    if (n >= 15) approx_besselroots(alpha, 7, r.q.jak_pre);
    return r;
}

// Validates the batch and fixes the launch order.  order (may be null): order[s] = index of the rule processed at
// position s.  rules (may be null): filled with the constants of every rule at its position, on host threads for a
// large batch — unless `defer_constants`, where only the vector is sized (the caller plans with an OrderedPlanner).
static int laguerre_batch_plan(int64_t nrules, const int32_t* n_i, const double* alpha_i, const int64_t* offsets,
                               std::vector<LaguerreBatchRule>* rules, int64_t* total, int* warn,
                               std::vector<int64_t>* order = nullptr, bool defer_constants = false) {
    FGQ_TRY(small_batch_rules(nrules, n_i, offsets, LAGUERRE_BATCH_MAX_N, "gausslaguerre", nullptr, total));
    if (nrules > 0 && !alpha_i) {
        set_error("invalid batch description");
        return FGQ_EINVAL;
    }
    for (int64_t i = 0; i < nrules; ++i) {
        const int64_t n = n_i[i];
        const double alpha = alpha_i[i];
        FGQ_TRY(laguerre_check(n, alpha));  // alpha <= -1 is a DomainError for every n, as in the reference (:61)
        if (n > 0 && alpha * alpha / (double)n > 1) *warn |= FGQ_WARN_LARGE_ALPHA;
    }
    if (!rules) return FGQ_OK;
    rules->resize((size_t)nrules);
    // One thread per rule on the device, and the cost of a rule grows like n^2: rules of equal n share a
    // warp, longest first (the output position travels with the rule, the processing order is free).
    // Counting sort over n: slot[i] = position of rule i in the launch order.
    std::vector<int64_t> start(LAGUERRE_BATCH_MAX_N + 2, 0), slot((size_t)nrules);
    for (int64_t i = 0; i < nrules; ++i) start[LAGUERRE_BATCH_MAX_N - n_i[i] + 1]++;
    for (int k = 0; k <= LAGUERRE_BATCH_MAX_N; ++k) start[k + 1] += start[k];
    for (int64_t i = 0; i < nrules; ++i) slot[(size_t)i] = start[LAGUERRE_BATCH_MAX_N - n_i[i]]++;
    if (order) {
        order->resize((size_t)nrules);
        for (int64_t i = 0; i < nrules; ++i) (*order)[(size_t)slot[(size_t)i]] = i;
    }
    if (defer_constants) return FGQ_OK;
    host_parallel_for(nrules, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i)
            (*rules)[(size_t)slot[(size_t)i]] = laguerre_batch_rule(n_i[i], alpha_i[i], offsets[i]);
    });
    return FGQ_OK;
}

extern "C" int fgq_gausslaguerre_batch_device(int64_t nrules, const int32_t* n_i, const double* alpha_i,
                                              const int64_t* offsets, double* x_dev, double* w_dev,
                                              void* workspace_dev, void* stream) {
    clear_status();
    std::vector<LaguerreBatchRule> rules;
    std::vector<int64_t> order;
    int64_t total = 0;
    int warn = 0;
    // FGQ_LAGUERRE_BATCH=thread selects the first version (one thread per rule, coefficients recomputed per step).
    static const bool per_thread = []() {
        const char* e = getenv("FGQ_LAGUERRE_BATCH");
        return e && strcmp(e, "thread") == 0;
    }();
    const bool pipelined = !per_thread && nrules >= 16384;  // plan beside the kernels (OrderedPlanner)
    FGQ_TRY(laguerre_batch_plan(nrules, n_i, alpha_i, offsets, &rules, &total, &warn, pipelined ? &order : nullptr, pipelined));
    if (warn) add_warning(warn);
    if (nrules == 0 || total == 0) return FGQ_OK;
    if (!x_dev || !w_dev || !workspace_dev) {
        set_error("null pointer");
        return FGQ_EINVAL;
    }
    cudaStream_t st = as_stream(stream);
    char* ws = (char*)workspace_dev;
    laguerre_state_init_kernel<<<1, 1, 0, st>>>((LaguerreState*)ws, 0);
    FGQ_LAUNCH_CHECK();
    LaguerreBatchRule* drules = (LaguerreBatchRule*)(ws + 256);
    unsigned int* warn_word = &((LaguerreState*)ws)->warn;
    if (!pipelined)
        FGQ_CUDA_TRY(cudaMemcpyAsync(drules, rules.data(), rules.size() * sizeof(LaguerreBatchRule), cudaMemcpyHostToDevice, st));
    if (per_thread) {
        const int64_t blocks = (nrules + LAG_BATCH_THREADS - 1) / LAG_BATCH_THREADS;
        laguerre_small_batch_kernel<<<(unsigned)blocks, LAG_BATCH_THREADS, 0, st>>>(nrules, drules, warn_word, x_dev, w_dev);
        FGQ_LAUNCH_CHECK();
        return FGQ_OK;
    }
    static bool attr_set[MAX_ATTR_DEV] = {};
    int dev = 0;
    FGQ_CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 0 && dev < MAX_ATTR_DEV && !attr_set[dev]) {
        FGQ_CUDA_TRY(cudaFuncSetAttribute(laguerre_rec_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          512 * (LAGUERRE_BATCH_MAX_N + 1)));
        attr_set[dev] = true;
    }
    // Buckets of similar n over the sorted rules (longest first): shared memory per warp follows the bucket's n_max.
    std::vector<int> sorted_n((size_t)nrules);
    if (pipelined) {
        for (int64_t q = 0; q < nrules; ++q) sorted_n[(size_t)q] = (int)n_i[order[(size_t)q]];
    } else {
        for (int64_t q = 0; q < nrules; ++q) sorted_n[(size_t)q] = (int)rules[(size_t)q].n;
    }
    const std::vector<LagBucket> buckets = laguerre_batch_buckets(sorted_n);
    // The long-n buckets run at a few warps per SM (their tables fill the shared memory) and every bucket ends in a
    // partially filled wave: the buckets are spread over several streams (forked and joined with events: still
    // stream-ordered for the caller, graph-capturable) so that they fill each other's gaps.  Pipelined form: the
    // caller's stream carries only the uploads, bucket by bucket as the planner delivers them, and every bucket
    // runs on a side stream behind the event of its own upload — an upload never queues behind a kernel.
    constexpr int NSIDE = 4;
    ScopedStream side[NSIDE];
    ScopedEvents evs, ev_up;
    const bool fork = pipelined || buckets.size() > 1;
    const int nside = pipelined ? NSIDE : NSIDE - 1;
    if (fork) {
        for (int q = 0; q < nside; ++q) FGQ_TRY(side[q].acquire());
        FGQ_TRY(evs.acquire(1 + nside));
        FGQ_CUDA_TRY(cudaEventRecord(evs.ev[0], st));
        for (int q = 0; q < nside; ++q) FGQ_CUDA_TRY(cudaStreamWaitEvent(side[q].st, evs.ev[0], 0));
    }
    if (buckets.size() > 60) {
        set_error("internal: too many buckets");
        return FGQ_EINVAL;
    }
    {
        // declared after the streams and events: joined (destructor) before they are handed back
        std::unique_ptr<OrderedPlanner> planner;
        if (pipelined) {
            FGQ_TRY(ev_up.acquire((int)buckets.size()));
            const int64_t* ord = order.data();
            LaguerreBatchRule* out = rules.data();
            planner.reset(new OrderedPlanner(nrules, [=](int64_t lo, int64_t hi) {
                for (int64_t s = lo; s < hi; ++s) {
                    const int64_t i = ord[s];
                    out[s] = laguerre_batch_rule(n_i[i], alpha_i[i], offsets[i]);
                }
            }));
        }
        int64_t uploaded = 0;
        for (size_t bi = 0; bi < buckets.size(); ++bi) {
            const LagBucket& bk = buckets[bi];
            const size_t smem = bk.n_top >= 15 ? (size_t)512 * (size_t)(bk.n_top + 1) : 0;
            cudaStream_t s_b;
            if (pipelined) {
                // the last bucket also takes the rules without nodes that sort behind it (never read by a kernel)
                const int64_t up_hi = bi + 1 == buckets.size() ? nrules : bk.first + bk.count;
                planner->wait(up_hi);
                FGQ_CUDA_TRY(cudaMemcpyAsync(drules + uploaded, rules.data() + uploaded,
                                             (size_t)(up_hi - uploaded) * sizeof(LaguerreBatchRule), cudaMemcpyHostToDevice, st));
                uploaded = up_hi;
                FGQ_CUDA_TRY(cudaEventRecord(ev_up.ev[bi], st));
                s_b = side[bi % NSIDE].st;
                FGQ_CUDA_TRY(cudaStreamWaitEvent(s_b, ev_up.ev[bi], 0));
            } else {
                s_b = (fork && bi % NSIDE != 0) ? side[bi % NSIDE - 1].st : st;
            }
            laguerre_rec_batch_kernel<<<(unsigned)((bk.count + 31) / 32), 32, smem, s_b>>>(bk.first, bk.count, bk.n_top, drules,
                                                                                          warn_word, x_dev, w_dev);
            FGQ_LAUNCH_CHECK();
        }
    }
    if (fork) {
        for (int q = 0; q < nside; ++q) {
            FGQ_CUDA_TRY(cudaEventRecord(evs.ev[1 + q], side[q].st));
            FGQ_CUDA_TRY(cudaStreamWaitEvent(st, evs.ev[1 + q], 0));
        }
    }
    return FGQ_OK;
}

extern "C" int fgq_gausslaguerre_batch_host(int64_t nrules, const int32_t* n_i, const double* alpha_i,
                                            const int64_t* offsets, double* x, double* w) {
    clear_status();
    int64_t total = 0;
    int warn = 0;
    FGQ_TRY(laguerre_batch_plan(nrules, n_i, alpha_i, offsets, nullptr, &total, &warn));
    if (total == 0) return FGQ_OK;
    if (!x || !w) {
        set_error("null output pointer");
        return FGQ_EINVAL;
    }
    HostPathLock lock;
    void *ws, *out;
    FGQ_TRY(scratch_device((size_t)fgq_gausslaguerre_batch_workspace_bytes(nrules), 6, &ws));
    FGQ_TRY(scratch_device((size_t)total * 16, 0, &out));
    cudaStream_t st;
    FGQ_TRY(scratch_stream(0, &st));
    double* xd = (double*)out;
    double* wd = xd + total;
    FGQ_TRY(fgq_gausslaguerre_batch_device(nrules, n_i, alpha_i, offsets, xd, wd, ws, (void*)st));
    const int warn0 = fgq_warnings();
    LaguerreState h;
    FGQ_CUDA_TRY(cudaMemcpyAsync(&h, ws, sizeof(h), cudaMemcpyDeviceToHost, st));
    FGQ_TRY(drain_to_host(x, w, xd, wd, total, st));
    add_warning(warn0 | (int)h.warn);
    return FGQ_OK;
}

// ---- approx_besselroots(nu, n) on the device ----
extern "C" int fgq_approx_besselroots_device(double nu, int64_t n, double* x_dev, void* stream) {
    clear_status();
    if (n < 0) {
        set_error("DomainError(%lld): Input n must be a non-negative integer", (long long)n);  // besselroots.jl:41-43
        return FGQ_EDOMAIN;
    }
    if (n == 0) return FGQ_OK;
    if (!x_dev) {
        set_error("null output pointer");
        return FGQ_EINVAL;
    }
    // laguerre_params plans exactly the nu-only data this needs: the table / Piessens head and McMahon's
    // coefficients (its n-dependent fields are not read by the kernel)
    const bool zero_all = !(nu >= -1.0);
    const LaguerreParams p = laguerre_params(1, zero_all ? 0.0 : nu, false);
    besselroots_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(p, zero_all ? 1 : 0, n, x_dev);
    FGQ_LAUNCH_CHECK();
    return FGQ_OK;
}

extern "C" int fgq_approx_besselroots(double nu, int64_t n, double* x) {
    clear_status();
    if (n < 0) {
        set_error("DomainError(%lld): Input n must be a non-negative integer", (long long)n);
        return FGQ_EDOMAIN;
    }
    if (n == 0) return FGQ_OK;
    if (!x) {
        set_error("null output pointer");
        return FGQ_EINVAL;
    }
    HostPathLock lock;
    void* out;
    FGQ_TRY(scratch_device((size_t)n * 8, 0, &out));
    cudaStream_t st;
    FGQ_TRY(scratch_stream(0, &st));
    FGQ_TRY(fgq_approx_besselroots_device(nu, n, (double*)out, (void*)st));
    FGQ_CUDA_TRY(cudaMemcpyAsync(x, out, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    FGQ_CUDA_TRY(cudaStreamSynchronize(st));
    return FGQ_OK;
}
