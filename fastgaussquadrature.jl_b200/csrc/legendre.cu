// legendre.cu — Gauss-Legendre node/weight generation for sm_100a.
//
// Replaces, on the B200, the reference's array-pass implementation
//   asy / legpts_nodes / legpts_weights      src/gausslegendre.jl:71-231
//   besselZeroRoots / besselJ1               src/besselroots.jl:168-245
//   rec / innerRec! / newt_step! / leg_initial_guess   src/gausslegendre.jl:233-309
//   closed forms n <= 5                      src/gausslegendre.jl:27-60
// with ONE fused kernel per (node-series, weight-series) branch: a thread derives
// everything for two consecutive half-indices k from k itself, in registers, and only
// ever WRITES 16 bytes per node (8 B node + 8 B weight); each half-index is evaluated
// once and stored twice (+-x) with 16-byte streaming stores.  No global loads at all.
//
// Numerics: the reference's operation order is kept exactly (muladd/evalpoly -> fma,
// everything else unfused; the file is compiled with --fmad=false), so the angle theta
// is reproduced bit-for-bit and x, w differ from a libm-based evaluation only through
// the final cos/sin (fgq_math.h).
#include "common.cuh"
#include "fgq_math.h"
#include "legendre_device.cuh"

namespace fgq {

struct LegLaunch {
    int64_t kbase;     // half-index of thread 0's first element
    int64_t kmin, kmax;  // half-indices evaluated (inclusive)
    int64_t kL0, kL1;  // left stores:  xl[k - kL0] = -cos(theta_k)      (empty if kL0 > kL1)
    int64_t kR0, kR1;  // right stores: xr[kR1 - k] = +cos(theta_k)
    int64_t kin0, kin1;  // half-indices stored on BOTH sides with vector stores (interior fast path)
    int64_t kq0_last;  // NT=0 only: last half-index with theta_k < pi/4 (host bisection)
    int64_t kext_first;  // NT=0 only: first half-index whose theta_k takes the extended pi/2 reduction
    int64_t kcentre;   // (n+1)/2 when n is odd (node forced to 0.0, gausslegendre.jl:91), else 0
    double *xl, *wl, *xr, *wr;
    int vecL, vecR;    // 16-byte stores legal on that side
    // Block-uniform bulk ranges of an NT = 0 tail launch with both sides 16-byte aligned (host-planned): every
    // half-index of blocks [bq0_lo, bq0_hi) has theta < pi/4, of blocks [bq1_lo, bq1_hi) pi/4 < theta below the
    // band next to pi/2; all of them are interior (stored on both sides, no centre) and have k >= K_JK_EXACT.
    // Pair t = blockIdx.x * LEG_THREADS + threadIdx.x of the launch lives at xl2[t], wl2[t], xr2[-t], wr2[-t].
    // Odd n: the mirrored pair of a thread straddles a 16-byte boundary; shiftR = 1 selects the variant that stores the
    // aligned pair (k0, k0 - 1) instead, the second element coming from the previous lane (one shuffle), with 8-byte
    // stores at the two ends of each warp; xr2 / wr2 then address the pair (k0, k0 - 1) of pair t at [-t].
    unsigned bq0_lo, bq0_hi, bq1_lo, bq1_hi;
    int shiftR;
    double2 *xl2, *wl2, *xr2, *wr2;
    int64_t v_base;    // 4 kbase - 1: pair t holds half-indices with 4k - 1 = v_base + 8t and v_base + 8t + 4
    LegConsts c;
};

constexpr int LEG_THREADS = 256;
#ifndef FGQ_LEG_MIN_BLOCKS
#define FGQ_LEG_MIN_BLOCKS 1   // resident blocks per SM the compiler must allow for (register cap); tuning aid
#endif


// One block of a bulk range: no index clamping, no per-element predicates, 32-bit pair index, four address
// computations (one multiply-add each) and four 16-byte streaming stores per thread.
template <int QUAD>
__device__ __forceinline__ void leg_bulk_block(const LegLaunch& p) {
    const unsigned t = blockIdx.x * LEG_THREADS + threadIdx.x;
    const int64_t v0 = p.v_base + 8 * (int64_t)t;
    double cx0, w0, cx1, w1;
    leg_bulk_v<QUAD>(__ll2double_rn(v0), p.c, cx0, w0);
    leg_bulk_v<QUAD>(__ll2double_rn(v0 + 4), p.c, cx1, w1);
    __stcs(p.xl2 + t, make_double2(neg_bits(cx0), neg_bits(cx1)));
    __stcs(p.wl2 + t, make_double2(w0, w1));
    if (!p.shiftR) {
        __stcs(p.xr2 - t, make_double2(cx1, cx0));
        __stcs(p.wr2 - t, make_double2(w1, w0));
    } else {
        const double pc1 = __shfl_up_sync(0xffffffffu, cx1, 1);
        const double pw1 = __shfl_up_sync(0xffffffffu, w1, 1);
        const int lane = threadIdx.x & 31;
        double* xr = reinterpret_cast<double*>(p.xr2 - t);   // element k0; k0 - 1 follows, k1 precedes
        double* wr = reinterpret_cast<double*>(p.wr2 - t);
        if (lane > 0) {
            __stcs(p.xr2 - t, make_double2(cx0, pc1));
            __stcs(p.wr2 - t, make_double2(w0, pw1));
        } else {
            st_stream(xr, cx0);
            st_stream(wr, w0);
        }
        if (lane == 31) {
            st_stream(xr - 1, cx1);
            st_stream(wr - 1, w1);
        }
    }
}

template <int NT, int WT, bool HEAD>
__global__ void __launch_bounds__(LEG_THREADS, FGQ_LEG_MIN_BLOCKS)
legendre_asy_kernel(const LegLaunch p) {
    if (NT == 0 && WT == 0 && !HEAD) {
        if (blockIdx.x - p.bq0_lo < p.bq0_hi - p.bq0_lo) {
            leg_bulk_block<0>(p);
            return;
        }
        if (blockIdx.x - p.bq1_lo < p.bq1_hi - p.bq1_lo) {
            leg_bulk_block<1>(p);
            return;
        }
    }
    const int64_t t = (int64_t)blockIdx.x * LEG_THREADS + threadIdx.x;
    const int64_t k0 = p.kbase + 2 * t;
    const int64_t k1 = k0 + 1;
    // Interior blocks (every half-index of the block is stored on both sides with 16-byte
    // stores, no centre node): no clamping, no per-element predicates.  All but the first and
    // last block of a launch take this path.
    const int64_t kb0 = p.kbase + 2 * (int64_t)blockIdx.x * LEG_THREADS;
    const int64_t kb1 = kb0 + 2 * LEG_THREADS - 1;
    if (kb0 >= p.kin0 && kb1 <= p.kin1) {
        double cx0, w0, cx1, w1;
        leg_half_index<NT, WT, HEAD>(k0, p.c, cx0, w0);
        leg_half_index<NT, WT, HEAD>(k1, p.c, cx1, w1);
        st_stream2(p.xl + (k0 - p.kL0), neg_bits(cx0), neg_bits(cx1));
        st_stream2(p.wl + (k0 - p.kL0), w0, w1);
        if (p.vecR) {
            st_stream2(p.xr + (p.kR1 - k1), cx1, cx0);
            st_stream2(p.wr + (p.kR1 - k1), w1, w0);
        } else {
            // odd n: the mirrored pair of this thread straddles a 16-byte boundary; the aligned pair is
            // (k0, k0-1), whose second element is the previous lane's k1 -> one shuffle, and only the
            // two ends of each warp fall back to 8-byte stores
            const double pc1 = __shfl_up_sync(0xffffffffu, cx1, 1);
            const double pw1 = __shfl_up_sync(0xffffffffu, w1, 1);
            const int lane = threadIdx.x & 31;
            if (lane > 0) {
                st_stream2(p.xr + (p.kR1 - k0), cx0, pc1);
                st_stream2(p.wr + (p.kR1 - k0), w0, pw1);
            } else {
                st_stream(p.xr + (p.kR1 - k0), cx0);
                st_stream(p.wr + (p.kR1 - k0), w0);
            }
            if (lane == 31) {
                st_stream(p.xr + (p.kR1 - k1), cx1);
                st_stream(p.wr + (p.kR1 - k1), w1);
            }
        }
        return;
    }
    if (k0 > p.kmax) return;
    double cx[2], w[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        int64_t k = k0 + e;
        k = k < p.kmin ? p.kmin : (k > p.kmax ? p.kmax : k);  // clamp: result unused if outside
        leg_half_index<NT, WT, HEAD>(k, p.c, cx[e], w[e]);
    }
    // left side: ascending addresses with k, nodes negated; odd-n centre forced to 0.0
    {
        const bool in0 = k0 >= p.kL0 && k0 <= p.kL1, in1 = k1 >= p.kL0 && k1 <= p.kL1;
        const double x0 = (k0 == p.kcentre) ? 0.0 : neg_bits(cx[0]);
        const double x1 = (k1 == p.kcentre) ? 0.0 : neg_bits(cx[1]);
        if (p.vecL && in0 && in1) {
            st_stream2(p.xl + (k0 - p.kL0), x0, x1);
            st_stream2(p.wl + (k0 - p.kL0), w[0], w[1]);
        } else {
            if (in0) { st_stream(p.xl + (k0 - p.kL0), x0); st_stream(p.wl + (k0 - p.kL0), w[0]); }
            if (in1) { st_stream(p.xl + (k1 - p.kL0), x1); st_stream(p.wl + (k1 - p.kL0), w[1]); }
        }
    }
    // right side: descending addresses with k
    {
        const bool in0 = k0 >= p.kR0 && k0 <= p.kR1, in1 = k1 >= p.kR0 && k1 <= p.kR1;
        if (p.vecR && in0 && in1) {
            st_stream2(p.xr + (p.kR1 - k1), cx[1], cx[0]);
            st_stream2(p.wr + (p.kR1 - k1), w[1], w[0]);
        } else {
            if (in0) { st_stream(p.xr + (p.kR1 - k0), cx[0]); st_stream(p.wr + (p.kR1 - k0), w[0]); }
            if (in1) { st_stream(p.xr + (p.kR1 - k1), cx[1]); st_stream(p.wr + (p.kR1 - k1), w[1]); }
        }
    }
}

// Counts half-indices in [k_lo, k_hi) for which the slow-path-free reciprocal/quotient of the
// tail kernel gives a different node or weight than the IEEE operators (expected: 0).
__global__ void __launch_bounds__(256)
legendre_fastdiv_check_kernel(int64_t k_lo, int64_t k_hi, LegConsts c,
                              unsigned long long* mismatches) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long bad = 0;
    for (int64_t k = k_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < k_hi; k += stride) {
        double cf, wf, ci, wi;
        leg_half_index<0, 0, false>(k, c, cf, wf);
        leg_half_index<0, 0, true>(k, c, ci, wi);
        bad += (__double_as_longlong(cf) != __double_as_longlong(ci)) ||
               (__double_as_longlong(wf) != __double_as_longlong(wi));
        if (k >= K_JK_EXACT) {  // the specialised bulk path must give the same bits as well
            const double ak = FGQ_PI * fgq_k_minus_quarter(k);
            const double a = ak * c.vn;
            if (a < FGQ_PIO4) leg_bulk<0>(k, c, cf, wf);
            else if (!fgq_near_pio2(a)) leg_bulk<1>(k, c, cf, wf);
            bad += (__double_as_longlong(cf) != __double_as_longlong(ci)) ||
                   (__double_as_longlong(wf) != __double_as_longlong(wi));
        }
    }
    if (bad) atomicAdd(mismatches, bad);
}

// ---------------------------------------------------------------------------------------------
// Consumer fusion: the README's only use of a rule is dot(w, f.(x))
// (README.md:27-30).  This kernel generates the nodes in registers and reduces on the fly —
// sum w, sum w x^2 and optionally sum w exp(x) — so the 16 B/node store disappears and the path
// is purely FP64-bound.  Each half-index contributes both +-x.
// ---------------------------------------------------------------------------------------------
struct FusedLaunch {
    int64_t kmin, kmax;   // half-indices evaluated (inclusive)
    int64_t kq0_last, kext_first;  // as in LegLaunch
    int64_t kcentre;      // (n+1)/2 for odd n: that half-index is the single node x = 0
    int with_exp;
    LegConsts c;
};

__device__ __forceinline__ void kahan_add(double& s, double& comp, double v) {
    const double y = v - comp;
    const double t = s + y;
    comp = (t - s) - y;
    s = t;
}

template <int NT, int WT, bool HEAD>
__global__ void __launch_bounds__(LEG_THREADS)
legendre_fused_moments_kernel(const FusedLaunch p, double* sums) {
    double s0 = 0, c0 = 0, s2 = 0, c2 = 0, se = 0, ce = 0;
    const int64_t ntiles = (p.kmax - p.kmin) / (2 * LEG_THREADS) + 1;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t kb0 = p.kmin + tile * 2 * LEG_THREADS;
        const int64_t kb1 = kb0 + 2 * LEG_THREADS - 1;
        const int64_t k0 = kb0 + 2 * threadIdx.x;
        double cx[2], w[2];
        const bool full = kb1 <= p.kmax;
        if (NT == 0 && WT == 0 && !HEAD && full && kb0 >= K_JK_EXACT && kb1 <= p.kq0_last) {
            leg_bulk<0>(k0, p.c, cx[0], w[0]);
            leg_bulk<0>(k0 + 1, p.c, cx[1], w[1]);
        } else if (NT == 0 && WT == 0 && !HEAD && full && kb0 >= K_JK_EXACT && kb0 > p.kq0_last + 1 &&
                   kb1 < p.kext_first) {
            leg_bulk<1>(k0, p.c, cx[0], w[0]);
            leg_bulk<1>(k0 + 1, p.c, cx[1], w[1]);
        } else {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int64_t k = k0 + e > p.kmax ? p.kmax : k0 + e;
                leg_half_index<NT, WT, HEAD>(k, p.c, cx[e], w[e]);
            }
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int64_t k = k0 + e;
            if (k > p.kmax) continue;
            const bool centre = (k == p.kcentre);
            const double x = centre ? 0.0 : cx[e];
            const double mult = centre ? 1.0 : 2.0;
            kahan_add(s0, c0, mult * w[e]);
            kahan_add(s2, c2, mult * (w[e] * x) * x);
            if (p.with_exp) {
                const double ex = exp(x);
                kahan_add(se, ce, centre ? w[e] : w[e] * (ex + 1.0 / ex));
            }
        }
    }
    double v[3] = {s0, s2, se};
    __shared__ double sh[3][LEG_THREADS / 32];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_down_sync(0xffffffffu, v[q], o);
        if ((threadIdx.x & 31) == 0) sh[q][threadIdx.x >> 5] = v[q];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double a = 0.0;
        for (int i = 0; i < LEG_THREADS / 32; ++i) a += sh[threadIdx.x][i];
        atomicAdd(sums + threadIdx.x, a);
    }
}

typedef void (*FusedKernel)(const FusedLaunch, double*);
static FusedKernel pick_fused(int nt, int wt, bool head);

template <int NT>
__global__ void legendre_theta_kernel(int64_t k_lo, int64_t k_hi, LegConsts c, double* theta) {
    const int64_t k = k_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= k_hi) return;
    double cx, w, th;
    if (NT == 0) leg_half_index<0, 0, true>(k, c, cx, w, &th);
    if (NT == 1) leg_half_index<1, 0, true>(k, c, cx, w, &th);
    if (NT == 2) leg_half_index<2, 1, true>(k, c, cx, w, &th);
    if (NT == 3) leg_half_index<3, 2, true>(k, c, cx, w, &th);
    theta[k - k_lo] = th;
}

int legendre_branch(int64_t n, int* nt, int* wt);
void legendre_bulk_split(int64_t n, int64_t* kq0_last, int64_t* kext_first);
static int node_terms(int64_t n) { return n <= 255 ? 3 : n <= 3950 ? 2 : n < (1LL << 25) ? 1 : 0; }
static int weight_terms(int64_t n) { return n <= 170 ? 3 : n <= 1500 ? 2 : n <= 850000 ? 1 : 0; }

typedef void (*LegKernel)(const LegLaunch);
static double tail_angle(int64_t k, double vn);

// last k with a_k < pi/4 (Julia's sin/cos reduce from |x| >= pi/4 on), by bisection on the very
// arithmetic the kernel uses (a_k is monotone in k)
static int64_t quarter_split(int64_t n, double vn) {
    int64_t lo_k = 10680708, hi_k = (n + 1) / 2 + 1;  // K_JK_EXACT; invariant: a(lo_k) < pi/4 <= a(hi_k)
    if (!(tail_angle(lo_k, vn) < FGQ_PIO4 && tail_angle(hi_k, vn) >= FGQ_PIO4)) return 0;
    while (hi_k - lo_k > 1) {
        const int64_t mid = lo_k + (hi_k - lo_k) / 2;
        if (tail_angle(mid, vn) < FGQ_PIO4) lo_k = mid; else hi_k = mid;
    }
    return lo_k;
}

// first k >= K_JK_EXACT whose angle has the high word of pi/2 (0x3ff921fb: rem_pio2_kernel's "x ~= pi/2"
// test, which switches to the extended Cody-Waite reduction); a_k <= pi/2 (1 + 2^-50) on a rule, so
// that is a_k >= the double 0x3ff921fb00000000.
static int64_t near_pio2_first(int64_t n, double vn) {
    const double band_lo = 1.57079601287841796875;  // 0x3ff921fb00000000
    int64_t lo_k = 10680708, hi_k = (n + 1) / 2 + 2;  // invariant: a(lo_k) < band_lo <= a(hi_k) (or hi_k past the rule)
    if (!(tail_angle(lo_k, vn) < band_lo)) return lo_k;
    if (tail_angle(hi_k, vn) < band_lo) return INT64_MAX / 4;
    while (hi_k - lo_k > 1) {
        const int64_t mid = lo_k + (hi_k - lo_k) / 2;
        if (tail_angle(mid, vn) < band_lo) lo_k = mid; else hi_k = mid;
    }
    return hi_k;
}

// a_k for k >= K_JK_EXACT exactly as leg_bulk computes it (host IEEE arithmetic, no contraction
// possible: a product of a difference, then a product)
static double tail_angle(int64_t k, double vn) {
    volatile double ak = FGQ_PI * ((double)k - 0.25);
    volatile double a = ak * vn;
    return a;
}

// series branch of an n-point rule and the block-uniform split points of its n >= 2^25 tail (for integrate.cu)
int legendre_branch(int64_t n, int* nt, int* wt) {
    *nt = node_terms(n);
    *wt = weight_terms(n);
    return FGQ_OK;
}
void legendre_bulk_split(int64_t n, int64_t* kq0_last, int64_t* kext_first) {
    const double vn = make_consts(n).vn;
    *kq0_last = quarter_split(n, vn);
    *kext_first = near_pio2_first(n, vn);
}

static inline int64_t imax(int64_t a, int64_t b) { return a > b ? a : b; }
static inline int64_t imin(int64_t a, int64_t b) { return a < b ? a : b; }

static LegKernel pick_kernel(int nt, int wt, bool head) {
#define FGQ_PICK(NT, WT)                                                   \
    if (nt == NT && wt == WT)                                              \
        return head ? legendre_asy_kernel<NT, WT, true> : legendre_asy_kernel<NT, WT, false>;
    FGQ_PICK(3, 3)
    FGQ_PICK(3, 2)
    FGQ_PICK(2, 2)
    FGQ_PICK(2, 1)
    FGQ_PICK(1, 1)
    FGQ_PICK(1, 0)
    FGQ_PICK(0, 0)
#undef FGQ_PICK
    return nullptr;
}

static FusedKernel pick_fused(int nt, int wt, bool head) {
#define FGQ_PICKF(NT, WT)                                                  \
    if (nt == NT && wt == WT)                                              \
        return head ? legendre_fused_moments_kernel<NT, WT, true> : legendre_fused_moments_kernel<NT, WT, false>;
    FGQ_PICKF(3, 3)
    FGQ_PICKF(3, 2)
    FGQ_PICKF(2, 2)
    FGQ_PICKF(2, 1)
    FGQ_PICKF(1, 1)
    FGQ_PICKF(1, 0)
    FGQ_PICKF(0, 0)
#undef FGQ_PICKF
    return nullptr;
}

// Evaluate half-indices [kmin, kmax]; store left for k in [kL0,kL1], right for [kR0,kR1].
static int launch_one(int64_t n, bool head, int64_t kmin, int64_t kmax, int64_t kL0, int64_t kL1,
                      double* xl, double* wl, int64_t kR0, int64_t kR1, double* xr, double* wr,
                      cudaStream_t st) {
    if (kmin > kmax) return FGQ_OK;
    LegLaunch p;
    p.c = make_consts(n);
    p.kmin = kmin;
    p.kmax = kmax;
    p.kL0 = kL0; p.kL1 = kL1; p.kR0 = kR0; p.kR1 = kR1;
    p.xl = xl; p.wl = wl; p.xr = xr; p.wr = wr;
    p.kcentre = (n & 1) ? (n + 1) / 2 : 0;
    const bool hasL = kL0 <= kL1, hasR = kR0 <= kR1;
    // Pair parity.  Element k of the left side lives at xl + (k - kL0), of the right side at
    // xr + (kR1 - k); a 16-byte store needs an even offset from a 16-byte boundary, so the parity of
    // the pointer itself (in units of 8 bytes) counts too (e.g. the right segment of an odd rule
    // starts one element into its allocation).  Prefer aligned left pairs; the right side then
    // either has its own pair aligned (vecR) or the pair shifted by one element (vecR_shifted).
    auto par8 = [](const double* q) { return (int64_t)((reinterpret_cast<uintptr_t>(q) >> 3) & 1); };
    const bool okL = hasL && par8(xl) == par8(wl);
    const bool okR = hasR && par8(xr) == par8(wr);
    p.kbase = kmin;
    if (okL) {
        if (((par8(xl) + p.kbase - kL0) & 1) != 0) p.kbase -= 1;
    } else if (okR) {
        if (((par8(xr) + kR1 - p.kbase - 1) & 1) != 0) p.kbase -= 1;
    }
    p.vecL = okL && (((par8(xl) + p.kbase - kL0) & 1) == 0);
    p.vecR = okR && (((par8(xr) + kR1 - p.kbase - 1) & 1) == 0);
    // interior = both sides present, both vectorisable, centre excluded
    p.kin0 = 1;
    p.kin1 = 0;
    const bool vecR_shifted = okR && !p.vecR;
    if (p.vecL && (p.vecR || vecR_shifted)) {
        p.kin0 = imax(kL0, kR0);
        p.kin1 = imin(kL1, kR1);
        if (p.kcentre && p.kin1 >= p.kcentre) p.kin1 = p.kcentre - 1;
    }
    // only the n >= 2^25 kernel reads the pi/4 split
    p.kq0_last = (node_terms(n) == 0 && !head) ? quarter_split(n, p.c.vn) : 0;
    p.kext_first = (node_terms(n) == 0 && !head) ? near_pio2_first(n, p.c.vn) : 0;
    const int64_t npairs = (kmax - p.kbase) / 2 + 1;
    const int64_t blocks = (npairs + LEG_THREADS - 1) / LEG_THREADS;
    // bulk block ranges (see LegLaunch): block b covers half-indices kbase + 512 b .. kbase + 512 b + 511
    p.bq0_lo = p.bq0_hi = p.bq1_lo = p.bq1_hi = 0;
    p.xl2 = p.wl2 = p.xr2 = p.wr2 = nullptr;
    p.v_base = 4 * p.kbase - 1;
    p.shiftR = 0;
    if (node_terms(n) == 0 && weight_terms(n) == 0 && !head && p.vecL && (p.vecR || vecR_shifted) &&
        blocks * LEG_THREADS < (1LL << 32)) {
        const int64_t B = 2 * LEG_THREADS;
        auto ceil_div = [](int64_t a, int64_t b) { return a <= 0 ? 0 : (a + b - 1) / b; };
        auto floor_div = [](int64_t a, int64_t b) { return a < 0 ? 0 : a / b; };
        const int64_t first = ceil_div(imax(p.kin0, K_JK_EXACT) - p.kbase, B);   // kb0 >= max(kin0, K_JK_EXACT)
        const int64_t last = floor_div(p.kin1 + 1 - p.kbase, B);                  // kb1 <= kin1 (exclusive bound)
        const int64_t q0_hi = imin(last, floor_div(p.kq0_last + 1 - p.kbase, B));               // kb1 <= kq0_last
        const int64_t q1_lo = imax(first, ceil_div(p.kq0_last + 2 - p.kbase, B));               // kb0 > kq0_last + 1
        const int64_t q1_hi = imin(last, floor_div(p.kext_first - p.kbase, B));                 // kb1 < kext_first
        if (q0_hi > first) { p.bq0_lo = (unsigned)first; p.bq0_hi = (unsigned)q0_hi; }
        if (q1_hi > q1_lo) { p.bq1_lo = (unsigned)q1_lo; p.bq1_hi = (unsigned)q1_hi; }
        // element k of the left side lives at xl + (k - kL0); pair t starts at k = kbase + 2t
        p.xl2 = reinterpret_cast<double2*>(xl + (p.kbase - kL0));
        p.wl2 = reinterpret_cast<double2*>(wl + (p.kbase - kL0));
        // right side: elements k1 = kbase + 2t + 1, k0 at xr + (kR1 - k1), + 1; shifted: k0, k0 - 1 at xr + (kR1 - k0), + 1
        p.shiftR = p.vecR ? 0 : 1;
        p.xr2 = reinterpret_cast<double2*>(xr + (kR1 - p.kbase - 1 + p.shiftR));
        p.wr2 = reinterpret_cast<double2*>(wr + (kR1 - p.kbase - 1 + p.shiftR));
    }
    if (blocks > 0x7fffffffLL) {
        set_error("legendre launch too large (%lld blocks)", (long long)blocks);
        return FGQ_EINVAL;
    }
    LegKernel kern = pick_kernel(node_terms(n), weight_terms(n), head);
    if (!kern) {
        set_error("internal: no kernel for n=%lld", (long long)n);
        return FGQ_EINVAL;
    }
    kern<<<(unsigned)blocks, LEG_THREADS, 0, st>>>(p);
    FGQ_LAUNCH_CHECK();
    return FGQ_OK;
}

constexpr int64_t K_TAIL = 13192;  // first half-index the pure McMahon tail may serve

// Left stores xl[k-kL0] for k in [kL0,kL1]; right stores xr[kR1-k] for k in [kR0,kR1].
int legendre_asy_launch(int64_t n, int64_t kL0, int64_t kL1, double* xl, double* wl, int64_t kR0,
                        int64_t kR1, double* xr, double* wr, cudaStream_t st) {
    // Head (tables / truncation bands, k < split) and tail (k >= split) launches.  The split is
    // moved by one when that keeps the tail's left pointer 16-byte aligned (the head kernel is
    // generic in k, so it may take k = 13192 as well).
    int64_t split = K_TAIL;
    if (kL0 <= kL1 && kL0 < split &&
        ((((reinterpret_cast<uintptr_t>(xl) >> 3) & 1) + (uintptr_t)(split - kL0)) & 1))
        split += 1;
    for (int part = 0; part < 2; ++part) {
        const bool head = part == 0;
        const int64_t a = head ? 1 : split;
        const int64_t b = head ? split - 1 : INT64_MAX / 4;
        // clip both store ranges to this part
        const int64_t L0 = imax(kL0, a), L1 = imin(kL1, b);
        const int64_t R0 = imax(kR0, a), R1 = imin(kR1, b);
        const bool hasL = L0 <= L1, hasR = R0 <= R1;
        double* pxl = hasL ? xl + (L0 - kL0) : nullptr;
        double* pwl = hasL ? wl + (L0 - kL0) : nullptr;
        double* pxr = hasR ? xr + (kR1 - R1) : nullptr;
        double* pwr = hasR ? wr + (kR1 - R1) : nullptr;
        if (hasL && hasR && imax(L0, R0) <= imin(L1, R1) + 1) {
            FGQ_TRY(launch_one(n, head, imin(L0, R0), imax(L1, R1), L0, L1, pxl, pwl, R0, R1, pxr,
                               pwr, st));
        } else {
            if (hasL) FGQ_TRY(launch_one(n, head, L0, L1, L0, L1, pxl, pwl, 1, 0, nullptr, nullptr, st));
            if (hasR) FGQ_TRY(launch_one(n, head, R0, R1, 1, 0, nullptr, nullptr, R0, R1, pxr, pwr, st));
        }
    }
    return FGQ_OK;
}

// ---------------------------------------------------------------------------------------------
// Small rules: one warp per rule.  Lane j owns node j+1 of the m = n/2+1 Newton iterates
// (rec, gausslegendre.jl:233-267); n <= 5 are the closed forms (:27-60).
// ---------------------------------------------------------------------------------------------

// Per-step constants of the recurrence (gausslegendre.jl:269-290), shared by every lane, rule and Newton
// step of a block and therefore tabulated once per block in shared memory: the counters 2k+1, k, k+1 as
// doubles (no I2F conversions, no FP64 adds to advance them) and r = the correctly rounded 1/(k+1).
// A lane reads step k with two 16-byte broadcast loads: the FP64 pipe, which bounds this kernel, only sees
// the recurrence itself.
struct __align__(16) RecStep {
    double r, tk;    // 1/(k+1), 2k+1
    double dk, dk1;  // k, k+1
};
constexpr int REC_STEPS = 64;

__device__ __forceinline__ void rec_step_table(RecStep* tab) {  // call with >= 64 threads, then __syncthreads()
    if (threadIdx.x < REC_STEPS) {
        const int k = threadIdx.x;
        RecStep s;
        s.r = fgq_rcp_fast((double)(k + 1));
        s.tk = (double)(2 * k + 1);
        s.dk = (double)k;
        s.dk1 = (double)(k + 1);
        tab[k] = s;
    }
}

// a / b for b = k + 1 (an exact small integer), given r = the correctly rounded 1/b: the last three
// operations of the IEEE quotient's fast path (fgq_div_fast), so the same bits as a / b for every
// normal a (|P_k|, |P'_k| stay within [1e-300, 1e6] here).  A zero numerator comes out as +0 whatever its
// sign, where IEEE returns -0 for -0: the sign of a zero P_k or P'_k never reaches an output (it is only
// ever added to non-zero terms or subtracted from the abscissa), which tools/small_rule_digest.py confirms
// over the whole input space of this path (n = 0..60) — so no FP64 compare-and-select guards it.
__device__ __forceinline__ double div_by_count(double a, double b, double r) {
    const double q = a * r;
    const double rem = fma(-b, q, a);
    return fma(r, rem, q);
}

// gausslegendre.jl:269-290 for one abscissa.
__device__ __forceinline__ void leg_inner_rec(int n, double xj, const RecStep* __restrict__ tab, double& P,
                                              double& PP) {
    double Pm2 = 1.0, Pm1 = xj, PPm1 = 1.0, PPm2 = 0.0;
    for (int k = 1; k <= n - 1; ++k) {
        const RecStep s = tab[k];
        const double newP = div_by_count(fma(s.tk * Pm1, xj, -s.dk * Pm2), s.dk1, s.r);
        Pm2 = Pm1;
        Pm1 = newP;
        const double newPP = div_by_count(s.tk * fma(xj, PPm1, Pm2) - s.dk * PPm2, s.dk1, s.r);
        PPm2 = PPm1;
        PPm1 = newPP;
    }
    P = Pm1;
    PP = PPm1;
}

__device__ __forceinline__ void leg_closed_form(int n, int i, double& x, double& w) {
    x = 0.0;
    w = 0.0;
    if (n == 1) {
        x = 0.0;
        w = 2.0;
    } else if (n == 2) {
        const double s3 = sqrt(3.0);
        x = (i == 0) ? -1.0 / s3 : 1.0 / s3;
        w = 1.0;
    } else if (n == 3) {
        const double s = sqrt(3.0 / 5.0);
        x = (i == 0) ? -s : (i == 1 ? 0.0 : s);
        w = (i == 1) ? 8.0 / 9.0 : 5.0 / 9.0;
    } else if (n == 4) {
        const double a = 2.0 * sqrt(6.0 / 5.0) / 7.0, t = 3.0 / 7.0, s30 = sqrt(30.0);
        const bool outer = (i == 0 || i == 3);
        const double r = outer ? sqrt(t + a) : sqrt(t - a);
        x = (i < 2) ? -r : r;
        w = outer ? (18.0 - s30) / 36.0 : (18.0 + s30) / 36.0;
    } else if (n == 5) {
        const double b = 2.0 * sqrt(10.0 / 7.0), s70 = sqrt(70.0);
        if (i == 2) {
            x = 0.0;
            w = 128.0 / 225.0;
        } else {
            const bool outer = (i == 0 || i == 4);
            const double r = (outer ? sqrt(5.0 + b) : sqrt(5.0 - b)) / 3.0;
            x = (i < 2) ? -r : r;
            w = outer ? (322.0 - 13.0 * s70) / 900.0 : (322.0 + 13.0 * s70) / 900.0;
        }
    }
}

// One Newton iterate of rec (gausslegendre.jl:233-267): initial guess j (0-based) of the n-point rule,
// two Newton steps on the recurrence, weight from P' of the LAST evaluation, i.e. at the pre-update
// iterate (:262-264).
__device__ __forceinline__ void leg_rec_iterate(int n, int j, const RecStep* __restrict__ rcp_tab, double& xj,
                                                double& wv) {
    // leg_initial_guess (:299-309): asymptotic nodes with the n <= 255 series, negated
    LegConsts c;
    c.vn = 1.0 / ((double)n + 0.5);
    c.vn2 = c.vn * c.vn;
    c.vn4 = c.vn2 * c.vn2;
    c.vn6 = c.vn4 * c.vn2;
    c.inv_vn2 = 0.0;
    double jk, j1sq;
    bessel_pair<true>(j + 1, jk, j1sq);
    const double a = jk * c.vn;
    const double u = fgq_cot_q1(a);
    xj = fgq_cos_q1(leg_theta<3>(a, u, c)) * -1.0;
    double P, PP;
#pragma unroll 1
    for (int it = 0; it < 2; ++it) {
        leg_inner_rec(n, xj, rcp_tab, P, PP);
        xj -= P / PP;
    }
    wv = 2.0 / ((1.0 - xj * xj) * (PP * PP));
}

constexpr int REC_THREADS = 128;

// n_i == nullptr: a single rule of n_single nodes, of which output indices [lo, hi) are
// stored at x[i - lo]; otherwise CSR batch (lo = 0, hi = n).
__global__ void __launch_bounds__(REC_THREADS)
legendre_rec_batch_kernel(int64_t nrules, const int32_t* __restrict__ n_i,
                          const int64_t* __restrict__ offsets, int n_single, int64_t lo, int64_t hi,
                          double* __restrict__ x, double* __restrict__ w) {
    __shared__ RecStep rcp_tab[REC_STEPS];
    rec_step_table(rcp_tab);
    __syncthreads();
    const int64_t rule = ((int64_t)blockIdx.x * REC_THREADS + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (rule >= nrules) return;
    int n;
    double *xr, *wr;
    if (n_i) {
        n = n_i[rule];
        const int64_t off = offsets[rule];
        xr = x + off;
        wr = w + off;
        lo = 0;
        hi = n;
    } else {
        n = n_single;
        xr = x - lo;
        wr = w - lo;
    }
    if (n <= 0 || n > 60) return;
    if (n <= 5) {
        if (lane < n && lane >= lo && lane < hi) {
            double xv, wv;
            leg_closed_form(n, lane, xv, wv);
            xr[lane] = xv;
            wr[lane] = wv;
        }
        return;
    }
    const int m = n / 2 + 1;
    // for even n the m-th iterate is overwritten by the mirror (:258-261): never computed here
    if (lane >= (n + 1) / 2) return;
    double xj, wv;
    leg_rec_iterate(n, lane, rcp_tab, xj, wv);
    const int i = lane;  // 0-based output index of the left copy
    if (i < m - 1) {
        if (i >= lo && i < hi) {
            xr[i] = xj;
            wr[i] = wv;
        }
        const int j = n - 1 - i;
        if (j >= lo && j < hi) {
            xr[j] = -xj;
            wr[j] = wv;
        }
    } else {  // i == m-1: the centre of an odd rule
        if (i >= lo && i < hi) {
            xr[i] = xj;
            wr[i] = wv;
        }
    }
}

// The batched form.  One warp per rule leaves 32 - (n/2+1) lanes idle (37 % of them for n uniform on
// [2, 60]); here a block takes a TILE of 128 consecutive rules, sorts them by n in shared memory
// (counting sort, longest first) and deals the (rule, iterate) pairs of the sorted list out to its
// threads as one flat index space: every lane works, and the lanes of a warp hold iterates of rules
// with (nearly) equal n, so the recurrence loops of a warp end together.  Per-iterate arithmetic is
// leg_rec_iterate, the same as above: results do not depend on the packing.
constexpr int TILE = 128;

__global__ void __launch_bounds__(TILE)
legendre_rec_tile_kernel(int64_t nrules, const int32_t* __restrict__ n_i, const int64_t* __restrict__ offsets,
                         double* __restrict__ x, double* __restrict__ w) {
    __shared__ RecStep rcp_tab[REC_STEPS];
    __shared__ int s_bucket[64];         // histogram over key = 60 - n, then bucket starts
    __shared__ int s_n[TILE];            // n of the s-th rule of the sorted tile
    __shared__ int s_start[TILE + 1];    // first flat index of the s-th rule
    __shared__ long long s_off[TILE];    // its output offset
    __shared__ int s_warp[TILE / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    rec_step_table(rcp_tab);
    if (tid < 64) s_bucket[tid] = 0;
    __syncthreads();
    const int64_t rule = (int64_t)blockIdx.x * TILE + tid;
    int n = 0;
    long long off = 0;
    if (rule < nrules) {
        n = n_i[rule];
        off = offsets[rule];
        if (n < 0 || n > 60) n = 0;
    }
    const int key = 60 - n;
    const int rank = atomicAdd(&s_bucket[key], 1);
    __syncthreads();
    if (tid < 32) {  // exclusive scan of the 64 bucket counts, two per lane
        const int v0 = s_bucket[2 * tid], v1 = s_bucket[2 * tid + 1];
        int incl = v0 + v1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        const int excl = incl - (v0 + v1);
        s_bucket[2 * tid] = excl;
        s_bucket[2 * tid + 1] = excl + v0;
    }
    __syncthreads();
    const int pos = s_bucket[key] + rank;
    s_n[pos] = n;
    s_off[pos] = off;
    __syncthreads();
    // threads a rule needs: n <= 5 closed forms, one per node; otherwise one per kept Newton iterate
    const int ns = s_n[tid];
    const int need = ns <= 5 ? ns : (ns + 1) / 2;
    int incl = need;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    int before = 0;
    for (int q = 0; q < wid; ++q) before += s_warp[q];
    s_start[tid] = before + incl - need;
    if (tid == TILE - 1) s_start[TILE] = before + incl;
    __syncthreads();
    const int total = s_start[TILE];
    for (int f = tid; f < total; f += TILE) {
        int s = 0;  // largest s with s_start[s] <= f (empty rules sort last, at s_start == total)
#pragma unroll
        for (int step = TILE / 2; step > 0; step >>= 1)
            if (s_start[s + step] <= f) s += step;
        const int nr = s_n[s];
        const int i = f - s_start[s];
        double* xr = x + s_off[s];
        double* wr = w + s_off[s];
        if (nr <= 5) {
            double xv, wv;
            leg_closed_form(nr, i, xv, wv);
            xr[i] = xv;
            wr[i] = wv;
            continue;
        }
        double xj, wv;
        leg_rec_iterate(nr, i, rcp_tab, xj, wv);
        xr[i] = xj;
        wr[i] = wv;
        if (i < nr / 2) {  // not the centre of an odd rule
            xr[nr - 1 - i] = -xj;
            wr[nr - 1 - i] = wv;
        }
    }
}

int legendre_small_launch(int n, int64_t lo, int64_t hi, double* x, double* w, cudaStream_t st) {
    legendre_rec_batch_kernel<<<1, REC_THREADS, 0, st>>>(1, nullptr, nullptr, n, lo, hi, x, w);
    FGQ_LAUNCH_CHECK();
    return FGQ_OK;
}

}  // namespace fgq

using namespace fgq;

extern "C" int fgq_gausslegendre_device(int64_t n, int64_t lo, int64_t hi, double* x_dev,
                                        double* w_dev, void* stream) {
    clear_status();
    if (n < 0) {
        set_error("DomainError(%lld): Input n must be a non-negative integer", (long long)n);
        return FGQ_EDOMAIN;
    }
    if (lo < 0 || hi < lo || hi > n) {
        set_error("invalid shard [%lld, %lld) of n=%lld", (long long)lo, (long long)hi, (long long)n);
        return FGQ_EINVAL;
    }
    if (hi == lo) return FGQ_OK;
    if (!x_dev || !w_dev) {
        set_error("null output pointer");
        return FGQ_EINVAL;
    }
    cudaStream_t st = as_stream(stream);
    if (n <= 60) return legendre_small_launch((int)n, lo, hi, x_dev, w_dev, st);
    const int64_t m = (n + 1) >> 1;  // left side incl. centre: indices [0, m) <-> k = i + 1
    const int64_t mr = n >> 1;       // right side: indices [m, n) <-> k = n - i in [1, mr]
    const int64_t kL0 = lo + 1, kL1 = imin(hi, m);
    const int64_t rlo = imax(lo, m);  // first right-side output index of this shard
    const int64_t kR0 = n - hi + 1, kR1 = imin(n - rlo, mr);
    double* xr = x_dev + (rlo - lo);
    double* wr = w_dev + (rlo - lo);
    return legendre_asy_launch(n, kL0, kL1, x_dev, w_dev, kR0, kR1, xr, wr, st);
}

extern "C" int fgq_gausslegendre_device_pair(int64_t n, int64_t k_lo, int64_t k_hi, double* xl_dev,
                                             double* wl_dev, double* xr_dev, double* wr_dev,
                                             void* stream) {
    clear_status();
    if (n <= 60) {
        set_error("mirror-pair shards need the asymptotic path (n > 60), got n=%lld", (long long)n);
        return FGQ_EINVAL;
    }
    const int64_t m = (n + 1) >> 1, mr = n >> 1;
    if (k_lo < 1 || k_hi < k_lo || k_hi > m + 1) {
        set_error("invalid half-index shard [%lld, %lld) of n=%lld", (long long)k_lo,
                  (long long)k_hi, (long long)n);
        return FGQ_EINVAL;
    }
    if (k_hi == k_lo) return FGQ_OK;
    if (!xl_dev || !wl_dev || !xr_dev || !wr_dev) {
        set_error("null output pointer");
        return FGQ_EINVAL;
    }
    const int64_t kL1 = imin(k_hi - 1, m), kR1 = imin(k_hi - 1, mr);
    const int64_t skip = (k_hi - 1) - kR1;  // 1 iff this shard holds the centre of an odd rule
    return legendre_asy_launch(n, k_lo, kL1, xl_dev, wl_dev, k_lo, kR1, xr_dev + skip,
                               wr_dev + skip, as_stream(stream));
}

extern "C" int fgq_gausslegendre_batch_device(int64_t nrules, const int32_t* n_i_dev,
                                              const int64_t* offsets_dev, double* x_dev,
                                              double* w_dev, void* stream) {
    clear_status();
    if (nrules < 0 || (nrules > 0 && (!n_i_dev || !offsets_dev))) {
        set_error("invalid batch description");
        return FGQ_EINVAL;
    }
    if (nrules == 0) return FGQ_OK;
    const int64_t blocks = (nrules + TILE - 1) / TILE;
    if (blocks > 0x7fffffffLL) {
        set_error("batch too large");
        return FGQ_EINVAL;
    }
    legendre_rec_tile_kernel<<<(unsigned)blocks, TILE, 0, as_stream(stream)>>>(nrules, n_i_dev, offsets_dev,
                                                                                 x_dev, w_dev);
    FGQ_LAUNCH_CHECK();
    return FGQ_OK;
}

extern "C" int fgq_debug_legendre_theta_device(int64_t n, int64_t k_lo, int64_t k_hi,
                                               double* theta_dev, void* stream) {
    clear_status();
    if (n <= 5 || k_lo < 1 || k_hi < k_lo) {
        set_error("invalid theta request");
        return FGQ_EINVAL;
    }
    if (k_hi == k_lo) return FGQ_OK;
    const LegConsts c = make_consts(n);
    const int64_t cnt = k_hi - k_lo;
    const unsigned blocks = (unsigned)((cnt + 255) / 256);
    cudaStream_t st = as_stream(stream);
    switch (node_terms(n)) {
        case 0: legendre_theta_kernel<0><<<blocks, 256, 0, st>>>(k_lo, k_hi, c, theta_dev); break;
        case 1: legendre_theta_kernel<1><<<blocks, 256, 0, st>>>(k_lo, k_hi, c, theta_dev); break;
        case 2: legendre_theta_kernel<2><<<blocks, 256, 0, st>>>(k_lo, k_hi, c, theta_dev); break;
        default: legendre_theta_kernel<3><<<blocks, 256, 0, st>>>(k_lo, k_hi, c, theta_dev); break;
    }
    FGQ_LAUNCH_CHECK();
    return FGQ_OK;
}

extern "C" int fgq_debug_fastdiv_check(int64_t n, int64_t k_lo, int64_t k_hi,
                                       unsigned long long* mismatches_dev, void* stream) {
    clear_status();
    if (n < (1LL << 25) || k_lo < 13192 || k_hi < k_lo || !mismatches_dev) {
        set_error("invalid fastdiv check request");
        return FGQ_EINVAL;
    }
    if (k_hi == k_lo) return FGQ_OK;
    legendre_fastdiv_check_kernel<<<148 * 8, 256, 0, as_stream(stream)>>>(k_lo, k_hi, make_consts(n),
                                                                          mismatches_dev);
    FGQ_LAUNCH_CHECK();
    return FGQ_OK;
}

extern "C" int fgq_gausslegendre_moments_fused_device(int64_t n, int64_t k_lo, int64_t k_hi, int with_exp,
                                                      double* sums_dev, void* stream) {
    clear_status();
    if (n <= 60) {
        set_error("fused moments need the asymptotic path (n > 60), got n=%lld", (long long)n);
        return FGQ_EINVAL;
    }
    const int64_t m = (n + 1) >> 1;
    if (k_lo < 1 || k_hi < k_lo || k_hi > m + 1 || !sums_dev) {
        set_error("invalid half-index shard [%lld, %lld) of n=%lld", (long long)k_lo, (long long)k_hi, (long long)n);
        return FGQ_EINVAL;
    }
    if (k_hi == k_lo) return FGQ_OK;
    cudaStream_t st = as_stream(stream);
    for (int part = 0; part < 2; ++part) {
        const bool head = part == 0;
        FusedLaunch p;
        p.kmin = head ? k_lo : imax(k_lo, K_TAIL);
        p.kmax = head ? imin(k_hi - 1, K_TAIL - 1) : k_hi - 1;
        if (p.kmin > p.kmax) continue;
        p.c = make_consts(n);
        p.kcentre = (n & 1) ? m : 0;
        p.with_exp = with_exp;
        p.kq0_last = (node_terms(n) == 0 && !head) ? quarter_split(n, p.c.vn) : 0;
        p.kext_first = (node_terms(n) == 0 && !head) ? near_pio2_first(n, p.c.vn) : 0;
        FusedKernel kern = pick_fused(node_terms(n), weight_terms(n), head);
        const int64_t ntiles = (p.kmax - p.kmin) / (2 * LEG_THREADS) + 1;
        const int64_t blocks = imin(ntiles, 148 * 8);
        kern<<<(unsigned)blocks, LEG_THREADS, 0, st>>>(p, sums_dev);
        FGQ_LAUNCH_CHECK();
    }
    return FGQ_OK;
}
