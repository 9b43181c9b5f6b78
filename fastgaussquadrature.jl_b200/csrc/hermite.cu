// hermite.cu — Gauss-Hermite nodes/weights, asymptotic (Airy) + Newton path, n > 200.
//
// Replaces, on the B200:
//   gausshermite / unweightedgausshermite   src/gausshermite.jl:28-65
//   hermite_asy                             src/gausshermite.jl:68-89
//   hermpoly_asy_airy                       src/gausshermite.jl:171-250
//   hermite_initialguess                    src/gausshermite.jl:252-323   (AIRY_ROOTS: constants.jl:121-135)
//
// The reference runs ~40 broadcast passes per Newton sweep over half-range vectors; here one thread
// owns one half-range node: initial guess, every Newton sweep and the weight stay in registers, the
// only state between sweeps is theta (8 B per node).  The reference's GLOBAL stopping rule
// (||dtheta||_inf < sqrt(eps)/10 over all nodes, :79) is kept without a host round trip: each sweep
// kernel reduces its max into the workspace, the last block to finish sets a `done` flag, and the
// remaining (pre-enqueued) sweep launches exit immediately.  Then a reduction for the global
// normalisation sum (:62) and a fold/scale kernel that also applies exp(-x^2) (:30).  That is the multi-launch
// form (FGQ_HERMITE_FUSED=0); by default the whole rule is ONE cooperative launch, hermite_fused_kernel below.
#include <cooperative_groups.h>

#include "airy.cuh"
#include "common.cuh"
#include "fgq_math.h"
#include "gw.cuh"

#include <string.h>

#include <mutex>
#include <vector>

namespace cg = cooperative_groups;

namespace fgq {

__constant__ double c_airy_roots[10] = {
    -2.338107410459767, -4.08794944413097, -5.520559828095551, -6.786708090071759,
    -7.944133587120853, -9.022650853340981, -10.04017434155809, -11.00852430373326,
    -11.93601556323626, -12.828776752865757};

constexpr int HERMITE_SUM_BLOCKS = 64;
constexpr int HERMITE_REC_MAX_N = 200;   // hermite_rec serves 21 <= n <= 200 (gausshermite.jl:46-48)

struct HermiteState {   // lives at the start of the workspace
    unsigned long long maxdt_bits[20];
    unsigned int ticket[20];
    int done;
    int sweeps;
    double norm_sum;
    unsigned int sum_ticket;
    double norm_part[HERMITE_SUM_BLOCKS];  // per-block partial sums, added in block order by the last block
};
static_assert(sizeof(HermiteState) <= 1024, "state must fit the workspace header");
// workspace: [0, 1024) HermiteState; [1024, HERM_WS_HEADER) two arrays of per-block |dtheta| maxima of the fused
// kernel (alternating between sweeps); then theta, xh, wh
constexpr int HERM_FUSED_MAX_GRID = 1024;
constexpr int64_t HERM_WS_HEADER = 1024 + 2 * 8 * HERM_FUSED_MAX_GRID;

struct HermiteParams {
    int64_t n;
    int64_t m;        // (n-1)>>1 (odd) or n>>1 (even): count of Tricomi/Airy guesses
    int64_t nh;       // half-range length: m+1 (odd, leading 0) or m
    int64_t nact;     // half-range nodes actually computed, counted from the centre: nh, or fewer for the
                      // truncated rule (fgq_gausshermite_reduced_*)
    int odd;
    double a;         // +-0.5
    double nu;        // 4m + 2a + 2
    double musq;      // 2n+1
    double sq_musq;   // sqrt(2n+1)
    double musq16, musq13, musq23, musq43, musq2, musq103, musq83;  // powers of musq
    double r_musq23, r_musq43, r_musq2, r_musq103, r_musq83;         // their correctly rounded reciprocals (herm_div_by)
    double nu13, num13, num53, num73;
    int64_t n_sin;    // floor(p n): guesses 1..n_sin are Tricomi's
    int64_t airy_from;  // ceil(p n): guesses taken from the reversed Airy list start here (1-based)
    double tol;
};

// gausshermite.jl:252-323 for one position i (1-based) of the m-vector
__device__ double hermite_guess(const HermiteParams& p, int64_t i) {
    const double PI = 3.141592653589793;
    // the patched vector is [x_sin[1:n_sin]; x_airy[airy_from:m]]
    int64_t src;
    bool use_sin;
    if (i <= p.n_sin) {
        use_sin = true;
        src = i;
    } else {
        use_sin = false;
        src = p.airy_from + (i - p.n_sin - 1);
    }
    if (use_sin) {
        const double rhs = (double)((4 * p.m + 3) - 4 * src) / p.nu * PI;
        double T = PI / 2;
        for (int k = 0; k < 7; ++k) {
            const double val = T - sin(T) - rhs;
            const double dval = 1.0 - cos(T);
            T = T - val / dval;
        }
        const double ch = cos(T / 2);
        const double tnk = ch * ch;
        const double tm1 = tnk - 1.0;
        return sqrt(p.nu * tnk - ((tnk + 0.25) / (tm1 * tm1) + (3.0 * p.a * p.a - 1.0)) / (3.0 * p.nu));
    }
    // x_init_airy = reverse(x_init): position src <- Airy root index j = m + 1 - src
    const int64_t j = p.m + 1 - src;
    double r;
    if (j <= 10) {
        r = c_airy_roots[j - 1];
    } else {
        const double t = 3.0 * PI / 8.0 * (double)(4 * j - 1);
        const double it2 = 1.0 / (t * t);
        const double it4 = it2 * it2, it6 = it4 * it2, it8 = it4 * it4, it10 = it8 * it2;
        r = -(cbrt(t * t) * (1.0 + 5.0 / 48.0 * it2 - 5.0 / 36.0 * it4 + (77125.0 / 82944.0) * it6 -
                                   108056875.0 / 6967296.0 * it8 + 162375596875.0 / 334430208.0 * it10));
    }
    const double r2 = r * r, r3 = r2 * r, r4 = r2 * r2, r5 = r4 * r;
    const double c223 = 1.5874010519681994;   // 2^(2/3)
    const double c213 = 1.2599210498948732;   // 2^(1/3)
    const double c243 = 2.5198420997897464;   // 2^(4/3)
    const double v = p.nu + c223 * r * p.nu13 + (1.0 / 5.0 * c243) * r2 * p.num13 +
                     (11.0 / 35.0 - p.a * p.a - 12.0 / 175.0) * r3 / p.nu +
                     ((16.0 / 1575.0) * r + (92.0 / 7875.0) * r4) * c223 * p.num53 -
                     ((15152.0 / 3031875.0) * r5 + (1088.0 / 121275.0) * r2) * c213 * p.num73;
    return sqrt(fabs(v));
}

// initial theta of half-range node t (0-based, from the centre)
__device__ __forceinline__ double hermite_theta0(const HermiteParams& p, int64_t t) {
    double x0;
    if (p.odd) {
        x0 = (t == 0) ? 0.0 : hermite_guess(p, t);  // [0; x_init][1:m+1]
    } else {
        x0 = hermite_guess(p, t + 1);
    }
    return acos(x0 / p.sq_musq);
}

__device__ __forceinline__ void hermite_state_reset(HermiteState* st) {
    for (int i = 0; i < 20; ++i) {
        st->maxdt_bits[i] = 0ull;
        st->ticket[i] = 0u;
    }
    st->done = 0;
    st->sweeps = 0;
    st->norm_sum = 0.0;
    st->sum_ticket = 0u;
}

__global__ void hermite_init_kernel(HermiteParams p, HermiteState* st, double* theta) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) hermite_state_reset(st);
    if (t >= p.nact) return;
    theta[t] = hermite_theta0(p, t);
}

// gausshermite.jl:171-250 at one theta.  Its 20 quotients were more than half of the FP64 instructions of a sweep
// (CUDA's a / b: ~25 instructions with its range checks).  The twelve by literals or per-rule constants use
// fgq_div_by_inrange (fgq_math.h): exact for the literal divisors; for the per-rule constants a quotient within
// 2^-53 ulp of a rounding boundary — a 1e-16-probability event, in a term of relative size (2n+1)^-2/3 or less —
// would come out one ulp off.  The eight by powers of chi / phi use fgq_div_fast (CUDA's own fast path without the
// range checks).  Every operand here is far inside the normal range.
__device__ __forceinline__ void hermpoly_asy_airy(const HermiteParams& p, double th, double sinT,
                                                  double cosT, double& val, double& dval) {
    const double sin2T = 2.0 * cosT * sinT;
    const double eta = 0.5 * th - 0.25 * sin2T;
    // (3 eta / 2)^(2/3) and (.)^(1/4) (gausshermite.jl:176-177) as cbrt of the square and sqrt of sqrt: within 1.2 and
    // 0.75 ulp of the exact powers (CUDA's pow: 2 ulp) at a fifth of pow's ~250 instructions each
    const double y32 = 3.0 * eta / 2.0;
    const double chi = -cbrt(y32 * y32);
    const double phi = sqrt(sqrt(fgq_div_fast(-chi, sinT * sinT)));
    const double SQPI = 1.7724538509055159;  // sqrt(pi)
    double C = 2.0 * SQPI * p.musq16 * phi;
    double Airy0, Airy1;
    airy_neg(p.musq23 * chi, Airy0, Airy1);

    const double a1 = 15.0 / 144.0, b1 = -7.0 / 5.0 * a1;
    const double a2 = 5.0 * 7.0 * 9.0 * 11.0 / 2.0 / (144.0 * 144.0), b2 = -13.0 / 11.0 * a2;
    const double a3 = 7.0 * 9.0 * 11.0 * 13.0 * 15.0 * 17.0 / 6.0 / (144.0 * 144.0 * 144.0), b3 = -19.0 / 17.0 * a3;
    constexpr double R24 = 1.0 / 24.0, R1152 = 1.0 / 1152.0, R414720 = 1.0 / 414720.0;  // correctly rounded by the compiler
    const double c2 = cosT * cosT, c3 = c2 * cosT, c4 = c2 * c2, c5 = c4 * cosT, c7 = c5 * c2, c9 = c7 * c2;
    const double u1 = fgq_div_by_inrange(c3 - 6.0 * cosT, 24.0, R24);
    const double u2 = fgq_div_by_inrange(-9.0 * c4 + 249.0 * c2 + 145.0, 1152.0, R1152);
    const double u3 = fgq_div_by_inrange(-4042.0 * c9 + 18189.0 * c7 - 28287.0 * c5 - 151995.0 * c3 - 259290.0 * cosT, 414720.0, R414720);
    const double phi2 = phi * phi, phi6 = phi2 * phi2 * phi2, phi12 = phi6 * phi6, phi18 = phi12 * phi6;
    const double chi2 = chi * chi, chi3 = chi2 * chi, chi4 = chi2 * chi2, chi5 = chi4 * chi;

    val = Airy0;
    const double B0 = fgq_div_fast(-(u1 * phi6 + a1), chi2);
    val = val + fgq_div_by_inrange(B0 * Airy1, p.musq43, p.r_musq43);
    const double A1 = fgq_div_fast(u2 * phi12 + b1 * u1 * phi6 + b2, chi3);
    val = val + fgq_div_by_inrange(A1 * Airy0, p.musq2, p.r_musq2);
    const double B1 = fgq_div_fast(-(u3 * phi18 + a1 * u2 * phi12 + a2 * u1 * phi6 + a3), chi5);
    val = val + fgq_div_by_inrange(B1 * Airy1, p.musq103, p.r_musq103);
    val = C * val;

    C = fgq_div_fast(2.5066282746310002 * p.musq13, phi);  // sqrt(2 pi)
    const double v1 = fgq_div_by_inrange(c3 + 6.0 * cosT, 24.0, R24);
    const double v2 = fgq_div_by_inrange(15.0 * c4 - 327.0 * c2 - 143.0, 1152.0, R1152);
    const double v3 = fgq_div_by_inrange(259290.0 * cosT + 238425.0 * c3 - 36387.0 * c5 + 18189.0 * c7 - 4042.0 * c9, 414720.0, R414720);
    const double C0 = fgq_div_fast(-(phi6 * v1 + b1), chi);
    dval = fgq_div_by_inrange(C0 * Airy0, p.musq23, p.r_musq23);
    dval = dval + Airy1;
    const double C1 = fgq_div_fast(-(v3 * phi18 + b1 * v2 * phi12 + b2 * v1 * phi6 + b3), chi4);
    dval = dval + fgq_div_by_inrange(C1 * Airy0, p.musq83, p.r_musq83);
    const double D1 = fgq_div_fast(v2 * phi12 + a1 * v1 * phi6 + a2, chi3);
    dval = dval + fgq_div_by_inrange(D1 * Airy1, p.musq2, p.r_musq2);
    dval = C * dval;
}

constexpr int HERM_THREADS = 128;

// One Newton step (gausshermite.jl:75-86) of half-range node t; returns the bit pattern of |dtheta| (non-negative
// doubles order like integers; a NaN update returns the NaN pattern, which sorts last: it poisons the norm, as
// norm() does)
__device__ __forceinline__ unsigned long long hermite_sweep_node(const HermiteParams& p, int64_t t, double* theta,
                                                                 double* xh, double* wh) {
    double th = theta[t];
    double s, c;
    sincos(th, &s, &c);
    double val, dval;
    hermpoly_asy_airy(p, th, s, c, val, dval);
    const double dth = -val / (1.4142135623730951 * p.sq_musq * dval * s);
    th = th - dth;
    theta[t] = th;
    double adt = fabs(dth);
    if (dth != dth) adt = __longlong_as_double(0x7ff8000000000000LL);
    const double x = p.sq_musq * cos(th);
    const double w = x * val + 1.4142135623730951 * dval;
    xh[t] = x;
    wh[t] = 1.0 / (w * w);
    return (unsigned long long)__double_as_longlong(adt);
}

// One Newton sweep (gausshermite.jl:75-86) over the half range.  Persistent grid (grid stride over the
// node tiles): the pre-enqueued sweeps that find the rule converged then cost ~2 us instead of the time to
// schedule and retire thousands of empty blocks.
constexpr int64_t HERM_GRID = 148 * 4;

__global__ void __launch_bounds__(HERM_THREADS, 4)
hermite_sweep_kernel(HermiteParams p, int sweep, HermiteState* st, double* theta, double* xh, double* wh) {
    if (*((volatile int*)&st->done)) return;
    unsigned long long bits = 0ull;
    for (int64_t t = (int64_t)blockIdx.x * HERM_THREADS + threadIdx.x; t < p.nact; t += (int64_t)gridDim.x * HERM_THREADS) {
        const unsigned long long b = hermite_sweep_node(p, t, theta, xh, wh);
        bits = b > bits ? b : bits;
    }
    // block max of |dtheta| through the bit pattern (non-negative doubles order like integers; NaN sorts last)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_down_sync(0xffffffffu, bits, o);
        bits = other > bits ? other : bits;
    }
    __shared__ unsigned long long sh[HERM_THREADS / 32];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = bits;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < HERM_THREADS / 32; ++i) bits = sh[i] > bits ? sh[i] : bits;
        atomicMax(&st->maxdt_bits[sweep], bits);
        __threadfence();
        const unsigned int tk = atomicAdd(&st->ticket[sweep], 1u);
        if (tk == gridDim.x - 1) {  // last block: apply the reference's stopping rule
            const double mx = __longlong_as_double((long long)atomicMax(&st->maxdt_bits[sweep], 0ull));
            st->sweeps = sweep + 1;
            if (mx < p.tol) st->done = 1;
            __threadfence();
        }
    }
}

// hermpoly_rec(n, x0) (gausshermite.jl:113-137): scaled Hermite polynomial and derivative by the
// recurrence, regularised by powers of exp(-x0^2/(4n)) whenever |H| >= 100.
// The divisors sqrt(k+1) and sqrt((k+1)/k) of a step do not depend on x0: the block tabulates them once,
// together with their correctly rounded reciprocals, and every quotient a / b becomes the last three
// operations of the IEEE quotient's fast path (fgq_div_fast) — the same bits as a / b for every operand
// here (|H| stays within [1e-200, 1e4]; a signed zero passes through as IEEE division would return it) —
// instead of two square roots and two full divisions per step and thread.
struct HermiteRecTables {
    double s1[HERMITE_REC_MAX_N], r1[HERMITE_REC_MAX_N];  // sqrt(k+1), 1/sqrt(k+1)         at [k]
    double s2[HERMITE_REC_MAX_N], r2[HERMITE_REC_MAX_N];  // sqrt((k+1)/k), its reciprocal  at [k]
};

__device__ __forceinline__ double div_tab(double a, double b, double r) {
    const double q = a * r;
    const double rem = fma(-b, q, a);
    const double res = fma(r, rem, q);
    return a == 0.0 ? a : res;
}

__device__ __forceinline__ void hermite_rec_tables(int n, HermiteRecTables& T) {
    for (int k = 1 + (int)threadIdx.x; k <= n - 1; k += (int)blockDim.x) {
        const double k1 = (double)(k + 1);
        const double a = sqrt(k1), b = sqrt(k1 / (double)k);
        T.s1[k] = a;
        T.r1[k] = fgq_rcp_fast(a);
        T.s2[k] = b;
        T.r2[k] = fgq_rcp_fast(b);
    }
}

__device__ void hermpoly_rec(int n, double x0, const HermiteRecTables& T, double& val, double& dval) {
    const double w = exp(-(x0 * x0) / (4 * (double)n));
    int wc = 0;
    double Hold = 1.0, H = x0;
    for (int k = 1; k <= n - 1; ++k) {
        const double Hnew = div_tab(x0 * H, T.s1[k], T.r1[k]) - div_tab(Hold, T.s2[k], T.r2[k]);
        Hold = H;
        H = Hnew;
        while (fabs(H) >= 100 && wc < n) {
            H *= w;
            Hold *= w;
            wc += 1;
        }
    }
    for (int q = wc + 1; q <= n; ++q) {
        H *= w;
        Hold *= w;
    }
    val = H;
    dval = -x0 * H + sqrt((double)n) * Hold;
}

// hermite_rec (gausshermite.jl:91-111), 21 <= n <= 200: one block, one thread per half-range node,
// at most 10 Newton steps with the global stopping rule ||dx||_inf < sqrt(eps).
__device__ __forceinline__ void hermite_rec_block(const HermiteParams& p, double* xh, double* wh) {
    const int t = threadIdx.x;
    const bool active = t < p.nh;
    const double SQ2 = 1.4142135623730951;
    double x0 = 0.0, val = 0.0, dval = 1.0;
    if (active) {
        x0 = p.odd ? (t == 0 ? 0.0 : hermite_guess(p, t)) : hermite_guess(p, t + 1);
        x0 *= SQ2;
    }
    __shared__ double s_mx[4];
    __shared__ HermiteRecTables s_tab;
    if (t < 4) s_mx[t] = 0.0;  // a block may have fewer than four warps (batch launches)
    hermite_rec_tables((int)p.n, s_tab);
    __syncthreads();
    for (int it = 0; it < 10; ++it) {
        double adx = 0.0;
        if (active) {
            hermpoly_rec((int)p.n, x0, s_tab, val, dval);
            double dx = val / dval;
            if (dx != dx) dx = 0.0;
            x0 = x0 - dx;
            adx = fabs(dx);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) adx = fmax(adx, __shfl_down_sync(0xffffffffu, adx, o));
        __syncthreads();
        if ((t & 31) == 0) s_mx[t >> 5] = adx;
        __syncthreads();
        const double mx = fmax(fmax(s_mx[0], s_mx[1]), fmax(s_mx[2], s_mx[3]));
        if (mx < 1.4901161193847656e-08) break;  // sqrt(eps)
    }
    if (active) {
        xh[t] = x0 / SQ2;
        wh[t] = 1.0 / (dval * dval);
    }
}

__global__ void __launch_bounds__(128) hermite_rec_kernel(HermiteParams p, double* xh, double* wh) {
    hermite_rec_block(p, xh, wh);
}

// hermite_gw (gausshermite.jl:325-341), 2 <= n <= 20: eigenvalues of the Jacobi matrix, upper half.
__device__ void hermite_gw_thread(const HermiteParams& p, double* xh, double* wh) {
    const int n = (int)p.n;
    double d[GW_MAX], e[GW_MAX], z[GW_MAX];
    for (int i = 0; i < n; ++i) d[i] = 0.0;
    for (int k = 1; k <= n - 1; ++k) e[k - 1] = sqrt((double)k / 2);
    tridiag_ql_first_row(n, d, e, z);
    const int first = n / 2;  // i = floor(n/2)+1 : n (1-based)
    for (int i = first; i < n; ++i) {
        const double x = d[i];
        const double w = 1.7724538509055159 * (z[i] * z[i]);
        xh[i - first] = x;
        wh[i - first] = exp(x * x) * w;
    }
}

__global__ void hermite_gw_kernel(HermiteParams p, double* xh, double* wh) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    hermite_gw_thread(p, xh, wh);
}

// sum over the FOLDED rule of exp(-x^2) w (gausshermite.jl:62): twice the half-range sum, the centre
// node of an odd rule counted once.
// Deterministic: fixed grid-stride assignment, per-block partials summed in block order by the last block
// to arrive (an atomicAdd of the partials would make the last bit of every weight depend on the order in
// which the blocks finish).
__device__ __forceinline__ double hermite_norm_term(const HermiteParams& p, int64_t t, const double* xh,
                                                    const double* wh) {
    const double x = xh[t];
    const double term = exp(-(x * x)) * wh[t];
    return (p.odd && t == 0) ? term : 2.0 * term;
}

__global__ void __launch_bounds__(512)
hermite_normsum_kernel(HermiteParams p, HermiteState* st, const double* xh, const double* wh) {
    double s = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < p.nact; t += stride) {
        s += hermite_norm_term(p, t, xh, wh);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    __shared__ double sh[16];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    __shared__ bool s_last;
    __shared__ double s_part[HERMITE_SUM_BLOCKS];
    if (threadIdx.x == 0) {
        double a = 0.0;
        for (int i = 0; i < 16; ++i) a += sh[i];
        st->norm_part[blockIdx.x] = a;
        __threadfence();
        s_last = atomicAdd(&st->sum_ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    // last block: fetch the partials in parallel (64 dependent L2 round trips by one thread cost ~10 us),
    // add them in block order as before
    __threadfence();
    if (threadIdx.x < gridDim.x) s_part[threadIdx.x] = ((volatile double*)st->norm_part)[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (unsigned i = 0; i < gridDim.x; ++i) tot += s_part[i];
        st->norm_sum = tot;
    }
}

// fold out (:55-61), global scale (:62), exp(-x^2) (:30), optional normalize (:31)
// output entry i of the folded rule from the half-range arrays
__device__ __forceinline__ void hermite_fold_one(const HermiteParams& p, int normalize, double norm_sum,
                                                 const double* xh, const double* wh, int64_t i, double& xo,
                                                 double& wo) {
    // half index: odd n: centre at nh-1 (output) <-> xh[0]; even: outputs nh-1 and nh <-> xh[0]
    int64_t h;
    bool neg;
    if (i < p.nh) {
        h = p.nh - 1 - i;
        neg = true;
    } else {
        h = p.odd ? i - p.nh + 1 : i - p.nh;
        neg = false;
    }
    const double xv = xh[h];
    const double scale = 1.7724538509055159 / norm_sum;
    double wv = wh[h] * scale;
    wv = wv * exp(-(xv * xv));
    xo = neg ? -xv : xv;
    if (normalize) {
        xo = 1.4142135623730951 * xo;
        wv = wv / 1.7724538509055159;
    }
    wo = wv;
}

__global__ void __launch_bounds__(256)
hermite_fold_kernel(HermiteParams p, int normalize, const HermiteState* st, const double* xh,
                    const double* wh, int64_t lo, int64_t hi, double* x, double* w) {
    const int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // output index
    if (i >= hi) return;
    double xo, wo;
    hermite_fold_one(p, normalize, st->norm_sum, xh, wh, i, xo, wo);
    x[i - lo] = xo;
    w[i - lo] = wo;
}

// ---------------------------------------------------------------------------------------------
// The whole asymptotic rule in ONE cooperative launch (n > 200).  The multi-launch form above pre-enqueues
// init, 20 sweeps, the normalisation sum and the fold — 23 launches of which 17 find the rule converged and
// return, and at n = 1e6 those empty launches plus the kernel boundaries are a quarter of the call.  Here a
// persistent grid (every block co-resident: cudaLaunchCooperativeKernel) keeps each half-range node with one
// thread from the initial guess to the last sweep; the only global steps are the reference's stopping rule
// (a grid barrier per sweep, then everybody reads the same max) and the normalisation sum.  Same device
// functions per node and the same summation order as hermite_normsum_kernel (virtual blocks of 512 threads,
// grid-stride terms, shuffle tree per warp, warp partials and block partials added in index order), so the
// results are bit-identical to the multi-launch form (FGQ_HERMITE_FUSED=0 selects it; the digest tool
// compares the two).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HERM_THREADS, 4)
hermite_fused_kernel(HermiteParams p, int normalize, HermiteState* st, unsigned long long* block_max, double* theta,
                     double* xh, double* wh, int64_t lo, int64_t hi, double* x, double* w) {
    cg::grid_group grid = cg::this_grid();
    const int64_t gid = (int64_t)blockIdx.x * HERM_THREADS + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * HERM_THREADS;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (gid == 0) hermite_state_reset(st);  // nothing below reduces into it: only the last writer of each field counts
    __shared__ unsigned long long sh_bits[HERM_THREADS / 32];
    for (int sweep = 0; sweep < 20; ++sweep) {
        unsigned long long bits = 0ull;
        for (int64_t t = gid; t < p.nact; t += stride) {
            if (sweep == 0) theta[t] = hermite_theta0(p, t);  // the initial guess feeds its own node's first step
            const unsigned long long b = hermite_sweep_node(p, t, theta, xh, wh);
            bits = b > bits ? b : bits;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_down_sync(0xffffffffu, bits, o);
            bits = other > bits ? other : bits;
        }
        if (lane == 0) sh_bits[wid] = bits;
        __syncthreads();
        // The reference's stopping rule (:79) needs max |dtheta| over the whole rule: every block publishes its
        // maximum in its own slot (no atomics, hence no reset and no barrier before the first sweep), and after
        // the grid barrier every block reduces the slots itself.  Two slot arrays alternate: a block can only
        // overwrite the array of sweep s in sweep s + 2, after everybody has passed the barrier of sweep s + 1.
        unsigned long long* slots = block_max + (sweep & 1) * HERM_FUSED_MAX_GRID;
        if (tid == 0) {
            for (int i = 1; i < HERM_THREADS / 32; ++i) bits = sh_bits[i] > bits ? sh_bits[i] : bits;
            slots[blockIdx.x] = bits;
        }
        grid.sync();
        unsigned long long mb = 0ull;
        for (unsigned i = tid; i < gridDim.x; i += HERM_THREADS) {
            const unsigned long long v = ((volatile unsigned long long*)slots)[i];
            mb = v > mb ? v : mb;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_down_sync(0xffffffffu, mb, o);
            mb = other > mb ? other : mb;
        }
        __syncthreads();  // sh_bits of this sweep has been read by thread 0
        if (lane == 0) sh_bits[wid] = mb;
        __syncthreads();
        mb = sh_bits[0];
        for (int i = 1; i < HERM_THREADS / 32; ++i) mb = sh_bits[i] > mb ? sh_bits[i] : mb;
        __syncthreads();
        const double mx = __longlong_as_double((long long)mb);
        if (gid == 0) st->sweeps = sweep + 1;
        if (mx < p.tol) break;  // the same value everywhere
    }
    // normalisation sum: real block b plays virtual block b of hermite_normsum_kernel (512 threads = 16 warps;
    // each of the 4 real warps takes 4 of them)
    int64_t vb = (p.nact + 511) / 512;
    if (vb > HERMITE_SUM_BLOCKS) vb = HERMITE_SUM_BLOCKS;
    __shared__ double sh_sum[16];
    __shared__ double s_part[HERMITE_SUM_BLOCKS];
    __shared__ double s_tot;
    // the terms themselves (an exp each) are computed by the whole grid, into theta, which has served its
    // purpose; the fixed-order sums below then only add
    for (int64_t t = gid; t < p.nact; t += stride) theta[t] = hermite_norm_term(p, t, xh, wh);
    grid.sync();
    for (int64_t b = blockIdx.x; b < vb; b += gridDim.x) {
        // real warp `wid` plays virtual warps wid, wid + 4, wid + 8, wid + 12 at once, four strides per round:
        // 16 independent loads in flight, then the adds of each virtual thread in its fixed order (a term
        // past the end is replaced by +0.0, which leaves a non-negative sum unchanged)
        const int64_t vstride = vb * 512;
        const int64_t t0 = b * 512 + wid * 32 + lane;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int64_t base = 0; b * 512 + base < p.nact; base += 4 * vstride) {
            double v[4][4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int64_t t = t0 + 128 * k + base + j * vstride;
                    v[k][j] = t < p.nact ? theta[t] : 0.0;
                }
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[k] += v[k][j];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double s = acc[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
            if (lane == 0) sh_sum[wid + 4 * k] = s;
        }
        __syncthreads();
        if (tid == 0) {
            double a = 0.0;
            for (int i = 0; i < 16; ++i) a += sh_sum[i];
            st->norm_part[b] = a;
        }
        __syncthreads();
    }
    grid.sync();
    if (tid < vb) s_part[tid] = ((volatile double*)st->norm_part)[tid];
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int64_t i = 0; i < vb; ++i) tot += s_part[i];
        s_tot = tot;
        if (blockIdx.x == 0) st->norm_sum = tot;
    }
    __syncthreads();
    const double tot = s_tot;
    for (int64_t i = lo + gid; i < hi; i += stride) {
        double xo, wo;
        hermite_fold_one(p, normalize, tot, xh, wh, i, xo, wo);
        x[i - lo] = xo;
        w[i - lo] = wo;
    }
}

// ---------------------------------------------------------------------------------------------
// Batched small rules (n <= 200): one block per rule runs the whole single-rule
// pipeline — hermite_gw / hermite_rec, the normalisation sum, the fold — out of shared memory, with the
// same device functions and the same summation order as the single-rule launches above, so a batch
// is bit-identical to the concatenation of single calls.  The parameter sets depend on n only and are
// read from a table of the 201 possible ones.  A rule keeps ceil(nh / 32) warps busy (nh <= 100 half-range
// nodes; one lane for n <= 20): the host sorts the batch by that count and launches each class with blocks of
// exactly that many warps (hermite_batch_warps), so no warp sits idle in a resident block — with 128 threads for
// every rule 57 % of the resident warps of a uniform n in [2, 200] batch only waited at the barriers.
// ---------------------------------------------------------------------------------------------
constexpr int HERMITE_BATCH_MAX_N = 200;

__global__ void __launch_bounds__(128)
hermite_small_batch_kernel(const SmallRule* __restrict__ rules, const HermiteParams* __restrict__ table,
                           int normalize, double* __restrict__ x, double* __restrict__ w) {
    __shared__ double xh[128], wh[128];
    __shared__ double sh[4];
    const SmallRule r = rules[blockIdx.x];
    if (r.n <= 0 || r.n > HERMITE_BATCH_MAX_N) return;
    const HermiteParams p = table[r.n];
    const int t = threadIdx.x;
    if (t < 4) sh[t] = 0.0;  // warps this block does not have contribute +0.0 to the sum below
    if (r.n <= 20) {
        if (t == 0) hermite_gw_thread(p, xh, wh);
    } else {
        hermite_rec_block(p, xh, wh);
    }
    __syncthreads();
    // hermite_normsum_kernel with one block: thread t holds term t, warp sums, partials in warp order
    double s = 0.0;
    if (t < p.nh) s += hermite_norm_term(p, t, xh, wh);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((t & 31) == 0) sh[t >> 5] = s;
    __syncthreads();
    double a = 0.0;
    for (int i = 0; i < 4; ++i) a += sh[i];
    double tot = 0.0;
    tot += a;
    for (int i = t; i < r.n; i += (int)blockDim.x) {
        double xo, wo;
        hermite_fold_one(p, normalize, tot, xh, wh, i, xo, wo);
        x[r.off + i] = xo;
        w[r.off + i] = wo;
    }
}

// warps a rule of the batch keeps busy: one lane (hermite_gw) for n <= 20, one thread per half-range node beyond
__host__ __device__ inline int hermite_batch_warps(int n) {
    if (n <= 20) return 1;
    const int nh = (n + 1) >> 1;
    return (nh + 31) >> 5;
}

}  // namespace fgq

using namespace fgq;

static HermiteParams hermite_params(int64_t n) {
    HermiteParams p;
    p.n = n;
    p.odd = (int)(n & 1);
    if (p.odd) {
        p.m = (n - 1) >> 1;
        p.a = 0.5;
        p.nh = p.m + 1;
    } else {
        p.m = n >> 1;
        p.a = -0.5;
        p.nh = p.m;
    }
    p.nact = p.nh;
    p.nu = (double)(4 * p.m) + 2.0 * p.a + 2.0;
    p.musq = (double)(2 * n + 1);
    p.sq_musq = sqrt(p.musq);
    p.musq16 = pow(p.musq, 1.0 / 6.0);
    p.musq13 = pow(p.musq, 1.0 / 3.0);
    p.musq23 = pow(p.musq, 2.0 / 3.0);
    p.musq43 = pow(p.musq, 4.0 / 3.0);
    p.musq2 = p.musq * p.musq;
    p.musq103 = pow(p.musq, 4.0 / 3.0 + 2.0);
    p.musq83 = pow(p.musq, 2.0 / 3.0 + 2.0);
    p.r_musq23 = 1.0 / p.musq23;
    p.r_musq43 = 1.0 / p.musq43;
    p.r_musq2 = 1.0 / p.musq2;
    p.r_musq103 = 1.0 / p.musq103;
    p.r_musq83 = 1.0 / p.musq83;
    p.nu13 = pow(p.nu, 1.0 / 3.0);
    p.num13 = pow(p.nu, -1.0 / 3.0);
    p.num53 = pow(p.nu, -5.0 / 3.0);
    p.num73 = pow(p.nu, -7.0 / 3.0);
    const double pp = 0.4985 + 2.220446049250313e-16;
    p.n_sin = (int64_t)floor(pp * (double)n);
    p.airy_from = (int64_t)ceil(pp * (double)n);
    if (p.n_sin > p.m) p.n_sin = p.m;
    p.tol = sqrt(2.220446049250313e-16) / 10.0;
    return p;
}

extern "C" int64_t fgq_gausshermite_workspace_bytes(int64_t n) {
    if (n < 0) return 0;
    const int64_t nh = (n >> 1) + 1;
    return HERM_WS_HEADER + 3 * 8 * nh;
}

// FGQ_HERMITE_FUSED=0: the multi-launch form (regression checks); default: one cooperative launch
static bool hermite_use_fused() {
    static const bool on = []() {
        const char* e = getenv("FGQ_HERMITE_FUSED");
        if (e && atoi(e) == 0) return false;
        int dev = 0, coop = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return false;
        if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess) return false;
        return coop != 0;
    }();
    return on;
}

// nact: half-range nodes to compute (from the centre); the output slice [lo, hi) must lie within them
static int hermite_device_impl(int64_t n, int normalize, int64_t lo, int64_t hi, int64_t nact, double* x_dev,
                               double* w_dev, void* workspace_dev, cudaStream_t st) {
    HermiteParams p = hermite_params(n);
    p.nact = nact;
    HermiteState* state = (HermiteState*)workspace_dev;
    double* theta = (double*)((char*)workspace_dev + HERM_WS_HEADER);
    unsigned long long* block_max = (unsigned long long*)((char*)workspace_dev + 1024);
    double* xh = theta + p.nh;
    double* wh = xh + p.nh;
    const unsigned b128 = (unsigned)((p.nact + HERM_THREADS - 1) / HERM_THREADS);
    if (n > 200 && hermite_use_fused()) {
        int dev = 0, per_sm = 0, sms = 0;
        FGQ_CUDA_TRY(cudaGetDevice(&dev));
        FGQ_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hermite_fused_kernel, HERM_THREADS, 0));
        FGQ_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        int64_t grid = (int64_t)per_sm * sms;
        if (grid > b128) grid = b128;
        if (grid < 1) grid = 1;
        if (grid > HERM_FUSED_MAX_GRID) grid = HERM_FUSED_MAX_GRID;
        void* args[] = {(void*)&p, (void*)&normalize, (void*)&state, (void*)&block_max, (void*)&theta, (void*)&xh,
                        (void*)&wh, (void*)&lo, (void*)&hi, (void*)&x_dev, (void*)&w_dev};
        FGQ_CUDA_TRY(cudaLaunchCooperativeKernel((void*)hermite_fused_kernel, dim3((unsigned)grid), dim3(HERM_THREADS),
                                                 args, 0, st));
        count_launch();
        return FGQ_OK;
    }
    hermite_init_kernel<<<b128, HERM_THREADS, 0, st>>>(p, state, theta);  // also resets the state
    FGQ_LAUNCH_CHECK();
    if (n <= 20) {  // :41-45; n = 1 (x = 0, w = sqrt(pi), :41-42) is the 1 x 1 case of the same kernel
        hermite_gw_kernel<<<1, 32, 0, st>>>(p, xh, wh);
        FGQ_LAUNCH_CHECK();
    } else if (n <= 200) {
        hermite_rec_kernel<<<1, 128, 0, st>>>(p, xh, wh);  // :46-48
        FGQ_LAUNCH_CHECK();
    } else {
        const unsigned bsweep = (unsigned)(b128 < HERM_GRID ? b128 : HERM_GRID);
        for (int s = 0; s < 20; ++s) {
            hermite_sweep_kernel<<<bsweep, HERM_THREADS, 0, st>>>(p, s, state, theta, xh, wh);
            FGQ_LAUNCH_CHECK();
        }
    }
    int64_t rb = (p.nact + 511) / 512;
    if (rb > HERMITE_SUM_BLOCKS) rb = HERMITE_SUM_BLOCKS;
    hermite_normsum_kernel<<<(unsigned)rb, 512, 0, st>>>(p, state, xh, wh);
    FGQ_LAUNCH_CHECK();
    if (hi > lo) {
        hermite_fold_kernel<<<(unsigned)((hi - lo + 255) / 256), 256, 0, st>>>(p, normalize, state, xh, wh,
                                                                             lo, hi, x_dev, w_dev);
        FGQ_LAUNCH_CHECK();
    }
    return FGQ_OK;
}

extern "C" int fgq_gausshermite_device(int64_t n, int normalize, int64_t lo, int64_t hi, double* x_dev,
                                       double* w_dev, void* workspace_dev, void* stream) {
    clear_status();
    if (n < 0) {
        set_error("DomainError(%lld): Input n must be a non-negative integer", (long long)n);
        return FGQ_EDOMAIN;
    }
    if (lo < 0 || hi < lo || hi > n) {
        set_error("invalid shard [%lld, %lld) of n=%lld", (long long)lo, (long long)hi, (long long)n);
        return FGQ_EINVAL;
    }
    if (n == 0) return FGQ_OK;
    if (!workspace_dev || (hi > lo && (!x_dev || !w_dev))) {
        set_error("null pointer");
        return FGQ_EINVAL;
    }
    return hermite_device_impl(n, normalize, lo, hi, hermite_params(n).nh, x_dev, w_dev, workspace_dev,
                               as_stream(stream));
}

// ---- truncated rule ----
// exp(-x^2) is exactly 0 in double precision for x^2 > 745.14, and the unweighted weights it multiplies
// (gausshermite.jl:30) are below 1: every node beyond |x| = 27.5 carries the weight 0.0.  The positive
// zeros of H_n are spaced at least pi / sqrt(2n+1) apart and the first one lies at least half that from
// the origin (Sturm comparison of y'' + (2n + 1 - x^2) y = 0 with the constant-coefficient equation), so
// the half-range nodes with index >= 27.5 sqrt(2n+1) / pi + 1 all lie beyond the cut.
static int64_t hermite_active_half(int64_t n) {
    const int64_t nh = hermite_params(n).nh;
    if (n <= 200) return nh;  // the small-n branches always compute the whole rule
    const double bound = 27.5 * sqrt(2.0 * (double)n + 1.0) / 3.141592653589793 + 2.0;
    return bound < (double)nh ? (int64_t)bound : nh;
}

// output range [first, first + count) of the full rule covered by the computed half-range nodes
static void hermite_reduced_range(int64_t n, int64_t* first, int64_t* count) {
    const HermiteParams p = hermite_params(n);
    const int64_t nact = hermite_active_half(n);
    *first = p.nh - nact;
    *count = n - 2 * (p.nh - nact);
}

extern "C" int64_t fgq_gausshermite_reduced_bound(int64_t n) {
    if (n <= 0) return 0;
    int64_t first, count;
    hermite_reduced_range(n, &first, &count);
    return count;
}

extern "C" int fgq_gausshermite_reduced_device(int64_t n, int normalize, double* x_dev, double* w_dev,
                                               void* workspace_dev, void* stream, int64_t* first,
                                               int64_t* count) {
    clear_status();
    if (n < 0) {
        set_error("DomainError(%lld): Input n must be a non-negative integer", (long long)n);
        return FGQ_EDOMAIN;
    }
    int64_t f = 0, c = 0;
    if (n > 0) hermite_reduced_range(n, &f, &c);
    if (first) *first = f;
    if (count) *count = c;
    if (n == 0) return FGQ_OK;
    if (!workspace_dev || !x_dev || !w_dev) {
        set_error("null pointer");
        return FGQ_EINVAL;
    }
    return hermite_device_impl(n, normalize, f, f + c, hermite_active_half(n), x_dev, w_dev, workspace_dev,
                               as_stream(stream));
}

extern "C" int fgq_gausshermite_reduced_host(int64_t n, int normalize, double* x, double* w, int64_t* first,
                                             int64_t* n_out) {
    clear_status();
    if (n < 0) {
        set_error("DomainError(%lld): Input n must be a non-negative integer", (long long)n);
        return FGQ_EDOMAIN;
    }
    if (first) *first = 0;
    if (n_out) *n_out = 0;
    if (n == 0) return FGQ_OK;
    if (!x || !w) {
        set_error("null output pointer");
        return FGQ_EINVAL;
    }
    HostPathLock lock;
    int64_t f, c;
    hermite_reduced_range(n, &f, &c);
    void *ws, *out;
    FGQ_TRY(scratch_device((size_t)fgq_gausshermite_workspace_bytes(n), 6, &ws));
    FGQ_TRY(scratch_device((size_t)c * 16, 0, &out));
    cudaStream_t st;
    FGQ_TRY(scratch_stream(0, &st));
    double* xd = (double*)out;
    double* wd = xd + c;
    FGQ_TRY(hermite_device_impl(n, normalize, f, f + c, hermite_active_half(n), xd, wd, ws, st));
    FGQ_TRY(drain_to_host(x, w, xd, wd, c, st));
    // keep the centre block of weights >= 10 floatmin (the threshold of the reference's reduced Laguerre rule,
    // gausslaguerre.jl:163-169); the rule is symmetric, so the same number of entries leaves at both ends
    int64_t drop = 0;
    while (2 * drop < c && fabs(w[drop]) < 10 * 2.2250738585072014e-308) ++drop;
    if (2 * drop >= c) drop = 0;  // cannot happen for a Hermite rule (the centre weights are O(1/sqrt(n)))
    if (drop > 0) {
        memmove(x, x + drop, (size_t)(c - 2 * drop) * 8);
        memmove(w, w + drop, (size_t)(c - 2 * drop) * 8);
    }
    if (first) *first = f + drop;
    if (n_out) *n_out = c - 2 * drop;
    return FGQ_OK;
}

extern "C" int fgq_gausshermite_host(int64_t n, int normalize, double* x, double* w) {
    clear_status();
    if (n < 0) {
        set_error("DomainError(%lld): Input n must be a non-negative integer", (long long)n);
        return FGQ_EDOMAIN;
    }
    if (n == 0) return FGQ_OK;
    if (!x || !w) {
        set_error("null output pointer");
        return FGQ_EINVAL;
    }
    HostPathLock lock;
    void *ws, *out;
    FGQ_TRY(scratch_device((size_t)fgq_gausshermite_workspace_bytes(n), 6, &ws));
    FGQ_TRY(scratch_device((size_t)n * 16, 0, &out));
    cudaStream_t st;
    FGQ_TRY(scratch_stream(0, &st));
    double* xd = (double*)out;
    double* wd = xd + n;
    FGQ_TRY(fgq_gausshermite_device(n, normalize, 0, n, xd, wd, ws, (void*)st));
    FGQ_TRY(drain_to_host(x, w, xd, wd, n, st));
    return FGQ_OK;
}

// number of Newton sweeps the last fgq_gausshermite_device call on this workspace needed
extern "C" int fgq_gausshermite_sweeps(const void* workspace_dev, int* sweeps) {
    clear_status();
    if (!workspace_dev || !sweeps) return FGQ_EINVAL;
    HermiteState h;
    FGQ_CUDA_TRY(cudaMemcpy(&h, workspace_dev, sizeof(h), cudaMemcpyDeviceToHost));
    *sweeps = h.sweeps;
    return FGQ_OK;
}

// ---- batched small rules ----
static const HermiteParams* hermite_param_table() {
    static HermiteParams table[HERMITE_BATCH_MAX_N + 1];
    static std::once_flag once;
    std::call_once(once, [] {
        for (int n = 0; n <= HERMITE_BATCH_MAX_N; ++n) table[n] = hermite_params(n);
    });
    return table;
}

constexpr size_t HERMITE_TABLE_BYTES = ((HERMITE_BATCH_MAX_N + 1) * sizeof(HermiteParams) + 255) / 256 * 256;

extern "C" int64_t fgq_gausshermite_batch_workspace_bytes(int64_t nrules) {
    if (nrules < 0) return 0;
    return (int64_t)HERMITE_TABLE_BYTES + nrules * (int64_t)sizeof(SmallRule);
}

extern "C" int fgq_gausshermite_batch_device(int64_t nrules, const int32_t* n_i, const int64_t* offsets,
                                             int normalize, double* x_dev, double* w_dev, void* workspace_dev,
                                             void* stream) {
    clear_status();
    std::vector<SmallRule> rules;
    int64_t total = 0;
    FGQ_TRY(small_batch_rules(nrules, n_i, offsets, HERMITE_BATCH_MAX_N, "gausshermite", &rules, &total));
    if (nrules == 0 || total == 0) return FGQ_OK;
    if (!x_dev || !w_dev || !workspace_dev) {
        set_error("null pointer");
        return FGQ_EINVAL;
    }
    cudaStream_t st = as_stream(stream);
    char* ws = (char*)workspace_dev;
    // counting sort by warps per rule, widest class first (each rule carries its own output offset)
    int64_t first[6] = {0, 0, 0, 0, 0, 0};
    for (const SmallRule& r : rules) first[4 - hermite_batch_warps(r.n) + 1] += 1;  // class c = 4 - warps
    for (int c = 0; c < 4; ++c) first[c + 1] += first[c];
    std::vector<SmallRule> sorted(rules.size());
    {
        int64_t fill[4] = {first[0], first[1], first[2], first[3]};
        for (const SmallRule& r : rules) sorted[(size_t)fill[4 - hermite_batch_warps(r.n)]++] = r;
    }
    FGQ_CUDA_TRY(cudaMemcpyAsync(ws, hermite_param_table(), (HERMITE_BATCH_MAX_N + 1) * sizeof(HermiteParams),
                                 cudaMemcpyHostToDevice, st));
    FGQ_CUDA_TRY(cudaMemcpyAsync(ws + HERMITE_TABLE_BYTES, sorted.data(), sorted.size() * sizeof(SmallRule),
                                 cudaMemcpyHostToDevice, st));
    for (int c = 0; c < 4; ++c) {
        const int64_t count = first[c + 1] - first[c];
        if (count == 0) continue;
        hermite_small_batch_kernel<<<(unsigned)count, 32 * (4 - c), 0, st>>>(
            (const SmallRule*)(ws + HERMITE_TABLE_BYTES) + first[c], (const HermiteParams*)ws, normalize, x_dev, w_dev);
        FGQ_LAUNCH_CHECK();
    }
    return FGQ_OK;
}

extern "C" int fgq_gausshermite_batch_host(int64_t nrules, const int32_t* n_i, const int64_t* offsets, int normalize,
                                           double* x, double* w) {
    clear_status();
    int64_t total = 0;
    FGQ_TRY(small_batch_rules(nrules, n_i, offsets, HERMITE_BATCH_MAX_N, "gausshermite", nullptr, &total));
    if (total == 0) return FGQ_OK;
    if (!x || !w) {
        set_error("null output pointer");
        return FGQ_EINVAL;
    }
    HostPathLock lock;
    void *ws, *out;
    FGQ_TRY(scratch_device((size_t)fgq_gausshermite_batch_workspace_bytes(nrules), 6, &ws));
    FGQ_TRY(scratch_device((size_t)total * 16, 0, &out));
    cudaStream_t st;
    FGQ_TRY(scratch_stream(0, &st));
    double* xd = (double*)out;
    double* wd = xd + total;
    FGQ_TRY(fgq_gausshermite_batch_device(nrules, n_i, offsets, normalize, xd, wd, ws, (void*)st));
    FGQ_TRY(drain_to_host(x, w, xd, wd, total, st));
    return FGQ_OK;
}
