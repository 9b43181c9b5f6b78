// jacobi.cu — Gauss-Jacobi nodes/weights for n > 100, max(alpha, beta) < 5, plus the
// Gauss-Radau / Gauss-Lobatto epilogues built on it.
//
// Replaces, on the B200:
//   gaussjacobi dispatch, jacobi_asy       src/gaussjacobi.jl:23-55, 170-194
//   asy1, feval_asy1 (interior)            src/gaussjacobi.jl:196-361
//   gaussradau(n), gausslobatto(n)         src/gaussradau.jl:31-49, src/gausslobatto.jl:31-52
//   asy2 and helpers (2 x 10 boundary nodes)   src/gaussjacobi.jl:363-667   (jacobi_boundary.cuh)
// (parameter-only tables and constants are planned on the host: jacobi_boundary.cu).
//
// The reference materialises ~10 dense 20 x N arrays per evaluation and runs a BLAS gemv per
// series term; here one thread owns one node and keeps the whole 20-term expansion in registers:
// per Newton sweep it reads and writes 8 B of state (theta) per node.  The two pieces of GLOBAL
// data-dependent control are kept, evaluated on the device without host round trips:
//   * the series is truncated where the inf-norm of a term over the interior set drops below
//     eps/100 (:309): a norm pass (same device function, no stores) precedes each evaluation pass;
//   * Newton stops on the inf-norm of the update over the interior set (:213,:235), then runs once
//     more "for luck": the last block of each pass advances a per-half phase flag and the remaining
//     pre-enqueued launches return immediately.
#include <math.h>

#include <mutex>

#include "common.cuh"
#include "fgq_math.h"
#include "jacobi_boundary.cuh"
#include "jacobi_plan.h"

namespace fgq {

struct JacobiState {
    unsigned long long termnorm[2][11][9];  // [half][sweep][m-12]: max |dS1+dS2| over the interior set
    unsigned long long dtnorm[2][11];
    unsigned int ticket[2][11][2];
    int phase[2];   // 0 iterate, 1 once more for luck, 2 done
    int sweeps[2];  // Newton sweeps before the extra one
    int terms[2][11];
};

struct JacobiLaunch {
    JacobiHalfPlan hp;   // parameters of this half ((alpha,beta) swapped for the left one)
    int half;            // 0: right (x > 0), 1: left
    int sweep;
    int64_t n;
    int64_t first, count;  // global (0-based, ascending-x) index range of this half
    int64_t idx_lo, idx_hi;  // interior index set [idx_lo, idx_hi) in global indices (:206, :228)
    double wc;             // weightsConstant
    int64_t lo, hi;        // output slice
    double tol;
    double tskip;          // norm passes: nodes with theta >= tskip provably stay below the truncation threshold
    int64_t norm_tile_lo, norm_tile_hi;  // norm passes: the tiles (of this half) that can hold a node with theta < tskip
};

constexpr int JAC_THREADS = 128;
#ifndef JAC_ROTATE
#define JAC_ROTATE 1   // 0: every term's phase through its own medium-range sincos (round 1)
#endif
constexpr double JAC_SKIP_TMAX = 1.5718;  // pi/2 + 1e-3: the bound behind tskip assumes 1/(2 cos(t/2)) <= 0.7079
constexpr int JAC_NORM_STAGE1 = 4;  // truncation candidates (terms 12, 13, ...) examined by the first norm stage
constexpr int64_t JAC_GRID = 148 * 5;  // persistent grid: 5 blocks of 128 threads fit one SM at 96 registers

// gaussjacobi.jl:254-361 for one node.  MODE 0: the evaluation.  MODE 1 / 2: only the truncation norms of
// terms 12..15 / 16..20 are produced (the norm pass is staged: the second stage runs only if none of the
// first four candidates falls below the threshold; large rules are cut at term 15: n t >= ~10 pi on the
// interior set whatever n is).
template <int MODE>
__device__ __forceinline__ void feval_asy1_node(const JacobiHalfPlan& P, double nn, double t, int mbreak,
                                                double& vals, double& ders, double* tnorm, bool take_norm = false) {
    constexpr bool NORM = MODE != 0;
    constexpr int M_FIRST = MODE == 0 ? 1 : (MODE == 1 ? 12 : 12 + JAC_NORM_STAGE1);
    constexpr int M_LAST = MODE == 1 ? 11 + JAC_NORM_STAGE1 : JAC_M;
    const double PI = 3.141592653589793;
    const double a = P.a, b = P.b;
    double sinH, cosH;
    sincos(t / 2, &sinH, &cosH);
    const double c0 = 2 * nn + a + b;
    // Away from the ends (see `rotate` below) sin t, cos t follow from the half-angle pair (4 operations instead of a
    // second sincos) and the two pow calls of the normalisation become one exp of two logs: a few ulp in quantities
    // that enter the node only through vals / ders and the weight at the 1e-15 level.  Next to the ends, where the
    // derivative sum cancels, the reference's own operations are kept.
    const bool lean = JAC_ROTATE && !NORM && c0 * t >= 8192.0;
    double sinT_, cosT_;
    if (lean) {
        sinT_ = 2 * sinH * cosH;
        cosT_ = (cosH - sinH) * (cosH + sinH);
    } else {
        sincos(t, &sinT_, &cosT_);
    }
    const double csc = (1.0 / sinH) / 2;
    const double sec = (1.0 / cosH) / 2;
    double pw[JAC_M];  // sinT[:, j] of the reference (:271), rescaled by secT after every term (:316)
    pw[0] = 1.0;
#pragma unroll
    for (int j = 1; j < JAC_M; ++j) pw[j] = pw[j - 1] * csc;
    const double shift = (a + 0.5) * PI / 2;
    double S = 0.0, S2 = 0.0;
    double sA = 0.0, cA = 1.0;
    // Next to the ends (n t up to a few hundred) the first two terms of S2 nearly cancel (to 1 part in ~400 at
    // n t = 40), so the weight there is sensitive to the 1e-14 rad by which an exactly advanced phase differs from the
    // reference's individually rounded A_m: those nodes (a few hundred per end, whatever n) keep the reference's phases.
    const bool rotate = lean;
#pragma unroll
    for (int m = 1; m <= M_LAST; ++m) {
        if (m >= M_FIRST) {
            if (!NORM && m == mbreak) break;
            if (NORM || m == 1 || !rotate) {
                const double A = (c0 + m) * t / 2 - shift;
                fgq_sincos_medium(A, &sA, &cA);
            } else {
                // A_m = A_{m-1} + t/2: one plane rotation by t/2 instead of a 1e7-radian sincos per term.  The
                // rotation is more accurate than the reference's own phase (whose rounding, ulp(A)/2 ~ 1e-9 rad at
                // n = 1e7, this does not reproduce), and terms m >= 2 are smaller than the first by (n t)^-(m-1),
                // n t >= ~10 pi on the interior set: the node moves by < 1e-16 / n.  The truncation norms (NORM
                // modes) keep the reference's phases, so the truncation index is unchanged.
                const double c2 = cA * cosH - sA * sinH;
                sA = sA * cosH + cA * sinH;
                cA = c2;
            }
            double d1 = 0.0, d2 = 0.0, d12 = 0.0, d22 = 0.0;
#pragma unroll
            for (int l = 0; l < m; l += 2) {
                d1 = fma(pw[l], P.PHI[l][m - 1], d1);
                if (!NORM) d12 = fma(pw[l], P.PHI2[l][m - 1], d12);
            }
#pragma unroll
            for (int l = 1; l < m; l += 2) {
                d2 = fma(pw[l], P.PHI[l][m - 1], d2);
                if (!NORM) d22 = fma(pw[l], P.PHI2[l][m - 1], d22);
            }
            d1 *= cA;
            d2 *= sA;
            if (NORM) {
                // running max over the bit patterns (as the block/grid reduction): a NaN norm, whose
                // pattern lies above +inf, wins
                const double v = fabs(d1 + d2);
                const unsigned long long vb = (unsigned long long)__double_as_longlong(v);
                const unsigned long long cb = (unsigned long long)__double_as_longlong(tnorm[m - 12]);
                if (take_norm && vb > cb) tnorm[m - 12] = v;
            } else {
                const double cA2 = cA * cosT_ + sA * sinT_;
                const double sA2 = sA * cosT_ - cA * sinT_;
                S += d1;
                S += d2;
                S2 += d12 * cA2;
                S2 += d22 * sA2;
            }
        }
#pragma unroll
        for (int j = 0; j < m; ++j) pw[j] *= sec;
    }
    if (NORM) return;
    vals = P.C * S;
    ders = (nn * ((a - b) - c0 * cosT_) * vals + (2 * (nn + a) * (nn + b) * P.C2) * S2) / c0 / sinT_;
    const double denom = lean ? exp(-((a + 0.5) * log(sinH) + (b + 0.5) * log(cosH)))
                              : 1.0 / (pow(sin(fabs(t) / 2), a + 0.5) * pow(cos(t / 2), b + 0.5));
    vals *= denom;
    ders *= denom;
}

__device__ __forceinline__ unsigned long long block_max_bits(double v) {
    unsigned long long bits = (unsigned long long)__double_as_longlong(v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_down_sync(0xffffffffu, bits, o);
        bits = other > bits ? other : bits;
    }
    __shared__ unsigned long long sh[JAC_THREADS / 32];
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = bits;
    __syncthreads();
    if (threadIdx.x == 0)
        for (int i = 1; i < JAC_THREADS / 32; ++i) bits = sh[i] > bits ? sh[i] : bits;
    return bits;  // valid in thread 0
}

// Initial guesses (Gatteschi-Pittaluga, :200-201) for every node; theta-space of its own half.
__global__ void jacobi_init_kernel(int64_t n, double a, double b, int64_t n_left, double* theta) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // 0-based position in tt
    if (i >= n) return;
    const double PI = 3.141592653589793;
    const double j = (double)(n - i);
    const double d = 2 * (double)n + a + b + 1;
    const double K = PI * (2 * j + a - 0.5) / d;
    const double tt = K + (1 / (d * d)) * ((0.25 - a * a) * (1.0 / tan(K / 2)) - (0.25 - b * b) * tan(K / 2));
    theta[i] = (i < n_left) ? PI - tt : tt;
}

// Persistent form: a fixed grid (a few blocks per SM) walks the node tiles with a grid stride, so the
// pre-enqueued launches that find their half already converged cost ~2 us instead of the ~42 us it takes
// to schedule and retire 39 063 empty blocks (n = 1e7; those were 30 % of the call), and every block
// reduces its norms once instead of once per tile.
template <int MODE>
__global__ void __launch_bounds__(JAC_THREADS, 5)
jacobi_pass_kernel(const __grid_constant__ JacobiLaunch p, JacobiState* st, double* theta, double* x, double* w) {
    constexpr bool NORM = MODE != 0;
    // Programmatic dependent launch: the pass kernels are enqueued as a fixed sequence of 66 launches of which most
    // return at once; with the stream-serialisation attribute the launch of kernel j+1 is processed while kernel j
    // still runs.  Nothing of the predecessor's output is read before the wait (full completion + visibility).
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    const int h = p.half;
    const int phase = *((volatile int*)&st->phase[h]);
    if (phase == 2) return;
    const double nn = (double)p.n;
    const int64_t ntiles = (p.count + JAC_THREADS - 1) / JAC_THREADS;
    if (NORM) {
        if (MODE == 2) {  // second stage: only if the series is not already cut at one of terms 12..15
            for (int q = 0; q < JAC_NORM_STAGE1; ++q) {
                const double nm = __longlong_as_double((long long)st->termnorm[h][p.sweep][q]);
                if (nm < 2.220446049250313e-16 / 100) return;
            }
        }
        double tmax[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) tmax[q] = 0.0;
        // theta is monotone along a half and small only at its outer end: the host bounds the tiles that can hold a
        // node with theta < tskip (twice the angle on the initial guesses, at least 4096 nodes), the others are not
        // even read (reading all of theta was 34 of a norm pass's 37 us at n = 1e7)
        for (int64_t tile = p.norm_tile_lo + blockIdx.x; tile < p.norm_tile_hi; tile += gridDim.x) {
            const int64_t r = tile * JAC_THREADS + threadIdx.x;
            const int64_t ig = p.first + r;
            const bool active = r < p.count;
            const bool interior = active && ig >= p.idx_lo && ig < p.idx_hi;
            const double t = active ? theta[ig] : 1.0;
            // jacobi_norm_skip_angle: for these angles every candidate term is below eps/100 whatever its
            // phase, so the node cannot change the outcome of the comparison the norms are used for
            if (!interior || (t >= p.tskip && t <= JAC_SKIP_TMAX)) continue;
            double dummy_v, dummy_d;
            feval_asy1_node<MODE>(p.hp, nn, t, 0, dummy_v, dummy_d, tmax, true);
        }
#pragma unroll
        for (int q = (MODE == 1 ? 0 : JAC_NORM_STAGE1); q < (MODE == 1 ? JAC_NORM_STAGE1 : 9); ++q) {
            const unsigned long long bits = block_max_bits(tmax[q]);
            if (threadIdx.x == 0) atomicMax(&st->termnorm[h][p.sweep][q], bits);
        }
        return;
    }
    // truncation index: first m >= 12 whose interior inf-norm is below eps/100 (:309)
    int mbreak = 0;
    for (int q = 0; q < 9; ++q) {
        const double nm = __longlong_as_double((long long)st->termnorm[h][p.sweep][q]);
        if (nm < 2.220446049250313e-16 / 100) {
            mbreak = 12 + q;
            break;
        }
    }
    unsigned long long adt_bits = 0ull;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t r = tile * JAC_THREADS + threadIdx.x;
        const int64_t ig = p.first + r;
        if (r >= p.count) continue;
        const bool interior = ig >= p.idx_lo && ig < p.idx_hi;
        double t = theta[ig];
        double vals, ders;
        feval_asy1_node<0>(p.hp, nn, t, mbreak, vals, ders, nullptr);
        const double dt = vals / ders;
        t += dt;
        theta[ig] = t;
        if (interior) {
            const double adt = (dt != dt) ? __longlong_as_double(0x7ff8000000000000LL) : fabs(dt);
            const unsigned long long bits = (unsigned long long)__double_as_longlong(adt);
            adt_bits = bits > adt_bits ? bits : adt_bits;
        }
        // the 10 nodes next to each end belong to the boundary kernel, which runs concurrently
        if (ig >= p.lo && ig < p.hi && ig >= JAC_NBDY && ig < p.n - JAC_NBDY) {
            const double c = cos(t);
            x[ig - p.lo] = h ? -c : c;                   // vcat(-x_left, x_right) (:246)
            w[ig - p.lo] = (1.0 / (ders * ders)) * p.wc;  // 1 ./ ders.^2 (:222,:244), rmul! (:192)
        }
    }
    const unsigned long long bits = block_max_bits(__longlong_as_double((long long)adt_bits));
    if (threadIdx.x == 0) {
        atomicMax(&st->dtnorm[h][p.sweep], bits);
        __threadfence();
        const unsigned int tk = atomicAdd(&st->ticket[h][p.sweep][1], 1u);
        if (tk == gridDim.x - 1) {
            const double mx = __longlong_as_double((long long)atomicMax(&st->dtnorm[h][p.sweep], 0ull));
            st->terms[h][p.sweep] = mbreak ? mbreak - 1 : JAC_M;
            if (phase == 0) {
                st->sweeps[h] += 1;
                if (mx < p.tol || st->sweeps[h] == 10) st->phase[h] = 1;
            } else {
                st->phase[h] = 2;
            }
            __threadfence();
        }
    }
}

// kernel<<<grid, block, 0, st>>>(args...) with programmatic stream serialisation allowed (FGQ_JACOBI_PDL=0: plain launch)
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), unsigned grid, unsigned block, cudaStream_t st, Args... args) {
    static const bool pdl = []() {
        const char* e = getenv("FGQ_JACOBI_PDL");
        return !(e && atoi(e) == 0);
    }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

struct JacobiPatch {
    double x[2 * JAC_NBDY], w[2 * JAC_NBDY];  // [0..10): left end (ascending), [10..20): right end
    int64_t n, lo, hi;
};

// jacobi_asy :182-190: overwrite the 10 nodes next to each end with the boundary formula.
__global__ void jacobi_patch_kernel(const JacobiPatch q, double* x, double* w) {
    const int i = threadIdx.x;
    if (i >= 2 * JAC_NBDY) return;
    const int64_t ig = i < JAC_NBDY ? i : q.n - 2 * JAC_NBDY + i;
    if (ig >= q.lo && ig < q.hi) {
        x[ig - q.lo] = q.x[i];
        w[ig - q.lo] = q.w[i];
    }
}

// gaussradau.jl:41-48 / gausslobatto.jl:41-49: rescale the Jacobi weights and add fixed endpoints.
// mode 1: Radau  (inner = gaussjacobi(n-1, 0, 1)): w /= (1 + x);   x[0] = -1, w[0] = 2/n^2
// mode 2: Lobatto (inner = gaussjacobi(n-2, 1, 1)): w /= (1 - x^2); ends +-1, w = 2/(n(n-1))
__global__ void endpoint_epilogue_kernel(int mode, int64_t n, int64_t lo, int64_t hi, double* x, double* w) {
    const int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // output index of the n-point rule
    if (i >= hi) return;
    const double nn = (double)n;
    if (i == 0) {
        x[0] = -1.0;
        w[0] = mode == 1 ? 2 / (nn * nn) : 2 / (nn * (nn - 1));
    } else if (mode == 2 && i == n - 1) {
        x[i - lo] = 1.0;
        w[i - lo] = 2 / (nn * (nn - 1));
    } else {
        const double xi = x[i - lo];
        w[i - lo] = mode == 1 ? w[i - lo] / (1 + xi) : w[i - lo] / (1 - xi * xi);
    }
}

// ---------------------------------------------------------------------------------------------
// n <= 100: jacobi_rec / HalfRec / innerjacobi_rec! (gaussjacobi.jl:61-155).  One block per rule:
// threads 0..63 own the x > 0 half with (alpha, beta), threads 64..127 the mirrored half with
// (beta, alpha); each thread is one Newton iterate on the three-term recurrence.  The reference's
// per-half stopping rule (max dx^2 <= eps/1e6, :117-125) is a two-warp reduction through shared memory.
// ---------------------------------------------------------------------------------------------
struct JacobiRecParams {
    int n;
    double alpha, beta;
    double c;        // 2^(a+b+1) G(2+a) G(2+b) / (G(2+a+b)(a+1)(b+1))  (:89)
    int64_t lo, hi;  // output slice
};

// innerjacobi_rec! for one abscissa (:129-155).  The recurrence coefficients of step k
//   A = 2(k+1)(k+a+b+1)(2k+a+b), B = (2k+a+b+1)(a^2-b^2), C = (2k+a+b)(2k+a+b+1)(2k+a+b+2), D = 2(k+a)(k+b)(2k+a+b+2)
// do not depend on the abscissa: the block tabulates them once (with the correctly rounded 1/A, so that both
// quotients of a step become the last three operations of the IEEE quotient's fast path: the same bits as
// x / A, see fgq_div_fast).  The mirrored half works with (b, a): every coefficient is unchanged bit for bit
// (sums and products of two numbers commute in floating point too) except B, which changes sign.
constexpr int JACOBI_REC_MAX_N = 100;
struct JacobiRecTables {
    double A[JACOBI_REC_MAX_N], rA[JACOBI_REC_MAX_N], B[JACOBI_REC_MAX_N], C[JACOBI_REC_MAX_N], D[JACOBI_REC_MAX_N];
};

__device__ __forceinline__ void jacobi_rec_tables(int n, double a, double b, JacobiRecTables& T) {
    for (int k = 1 + (int)threadIdx.x; k <= n - 1; k += (int)blockDim.x) {
        const double dk = (double)k;
        const double k0 = fma(2.0, dk, a + b);
        const double k1 = k0 + 1;
        const double k2 = k0 + 2;
        const double A = 2 * (dk + 1) * (dk + (a + b + 1)) * k0;
        T.A[k] = A;
        T.rA[k] = fgq_rcp_fast(A);
        T.B[k] = k1 * (a * a - b * b);
        T.C[k] = k0 * k1 * k2;
        T.D[k] = 2 * (dk + a) * (dk + b) * k2;
    }
}

__device__ __forceinline__ double jacobi_div_tab(double x, double A, double rA) {
    const double q = x * rA;
    const double rem = fma(-A, q, x);
    const double res = fma(rA, rem, q);
    return x == 0.0 ? x : res;
}

// bsign = +1 for the half that works with the tabulated (a, b), -1 for the mirrored one
__device__ __forceinline__ void jacobi_inner_rec(int n, double xj, double a, double b, double bsign,
                                                 const JacobiRecTables& T, double& P, double& PP) {
    double Pj = (a - b + (a + b + 2) * xj) / 2;
    double Pm1 = 1.0;
    double PPj = (a + b + 2) / 2;
    double PPm1 = 0.0;
    for (int k = 1; k <= n - 1; ++k) {
        const double A = T.A[k], rA = T.rA[k], C = T.C[k], D = T.D[k];
        const double B = bsign * T.B[k];
        const double c1 = fma(C, xj, B);
        const double newP = jacobi_div_tab(fma(-D, Pm1, c1 * Pj), A, rA);
        const double oldP = Pj;
        Pm1 = oldP;   // (Pm1, Pj) = (Pj, ...): the right-hand side uses the OLD Pm1
        Pj = newP;
        const double newPP = jacobi_div_tab(fma(c1, PPj, fma(-D, PPm1, C * Pm1)), A, rA);  // C * Pm1 uses the NEW Pm1 (= old Pj), as in :148-149
        PPm1 = PPj;
        PPj = newPP;
    }
    P = Pj;
    PP = PPj;
}

__device__ __forceinline__ void jacobi_rec_block(const JacobiRecParams& p, double* x, double* w) {
    const double PI = 3.141592653589793;
    const int tid = threadIdx.x;
    // Blocks of 128 threads (two warps per half: any n <= 100) or of 64 (one warp per half: n <= 64, batch launches)
    const int hshift = blockDim.x == 128 ? 6 : 5;
    const int h = tid >> hshift;     // 0: HalfRec(n, alpha, beta, 1); 1: HalfRec(n, beta, alpha, 0)
    const int i0 = tid & ((1 << hshift) - 1);
    const int n = p.n;
    const int m1 = (n + 1) / 2, m2 = n / 2;
    const int m = h == 0 ? m1 : m2;
    const bool active = i0 < m;
    const double a = h == 0 ? p.alpha : p.beta;
    const double b = h == 0 ? p.beta : p.alpha;
    __shared__ double s_dx2[4];
    __shared__ int s_done[2];
    __shared__ double xs[128], wsum[128];
    __shared__ JacobiRecTables s_tab;
    jacobi_rec_tables(n, p.alpha, p.beta, s_tab);
    const double bsign = h == 0 ? 1.0 : -1.0;
    if (tid < 2) s_done[tid] = 0;
    // initial guess (:97-109)
    const double c1 = 1 / (2 * (double)n + a + b + 1);
    const double a1 = 0.25 - a * a, b1 = 0.25 - b * b, c12 = c1 * c1;
    double xj = 0.0, P1 = 0.0, P2 = 1.0;
    if (active) {
        const double r = (double)(m - i0);  // r[i] = m, m-1, ..., 1
        const double C = fma(2.0, r, a - 0.5) * (PI * c1);
        const double C_2 = C / 2;
        xj = cos(fma(c12, fma(-b1, tan(C_2), a1 * (1.0 / tan(C_2))), C));
    }
    __syncthreads();
    for (int it = 0; it < 10; ++it) {
        const bool run = !s_done[h];
        double dx2 = 0.0;
        if (run && active) {
            jacobi_inner_rec(n, xj, a, b, bsign, s_tab, P1, P2);
            const double dx = P1 / P2;
            dx2 = dx * dx;
            xj = xj - dx;
        }
        // max over the half (ifelse(_dx2 > dx2, ...) never picks up a NaN, :119-122)
        double mx = dx2 > 0.0 ? dx2 : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double other = __shfl_down_sync(0xffffffffu, mx, o);
            mx = other > mx ? other : mx;
        }
        __syncthreads();
        if ((tid & 31) == 0) s_dx2[tid >> 5] = mx;
        __syncthreads();
        if (tid < 2 && !s_done[tid]) {
            // the half's one or two warp maxima (all >= 0)
            const double m0 = hshift == 6 ? s_dx2[2 * tid] : s_dx2[tid], m1v = hshift == 6 ? s_dx2[2 * tid + 1] : 0.0;
            const double mm = m1v > m0 ? m1v : m0;
            if (!(mm > 2.220446049250313e-16 / 1.0e6)) s_done[tid] = 1;
        }
        __syncthreads();
        if (s_done[0] && s_done[1]) break;
    }
    // once more for derivatives (:127)
    if (active) jacobi_inner_rec(n, xj, a, b, bsign, s_tab, P1, P2);
    // assemble (:69-88): left half mirrored, sum in the reference's order
    if (active) {
        const double xi = h == 0 ? xj : -xj;
        const int idx = h == 0 ? m2 + i0 : m2 - 1 - i0;
        xs[idx] = xi;
        wsum[idx] = 1 / ((1 - xi * xi) * (P2 * P2));
    }
    __syncthreads();
    __shared__ double s_scale;
    if (tid == 0) {
        double sum_w = 0.0;
        for (int i = 0; i < m2; ++i) sum_w += wsum[m2 - 1 - i];
        for (int i = 0; i < m1; ++i) sum_w += wsum[m2 + i];
        s_scale = p.c / sum_w;
    }
    __syncthreads();
    for (int i = tid; i < n; i += (int)blockDim.x) {
        if (i >= p.lo && i < p.hi) {
            x[i - p.lo] = xs[i];
            w[i - p.lo] = wsum[i] * s_scale;
        }
    }
}

__global__ void __launch_bounds__(128) jacobi_rec_kernel(const JacobiRecParams p, double* x, double* w) {
    jacobi_rec_block(p, x, w);
}

// Batched small rules: one block per rule, the same device function as the single-rule
// launch, so a batch is bit-identical to the concatenation of single calls.  kind 1: the one-node rule
// (gaussjacobi.jl:42-43; node in `alpha`, weight in `c`); kind 2: jacobi_rec (:61-155); kind 0: nothing
// to do here (empty, or served by the Legendre / Chebyshev / Golub-Welsch routes from the host loop).
// The host sorts the batch into two classes, n > 64 (blocks of 128 threads, two warps per half) and n <= 64
// (blocks of 64, one warp per half), so that a resident block has no warp without nodes.
struct JacobiBatchRule {
    int32_t n;
    int32_t kind;
    int64_t off;
    double alpha, beta, c;
};

__global__ void __launch_bounds__(128)
jacobi_small_batch_kernel(const JacobiBatchRule* __restrict__ rules, double* __restrict__ x, double* __restrict__ w) {
    const JacobiBatchRule r = rules[blockIdx.x];
    if (r.kind == 1) {
        if (threadIdx.x == 0) {
            x[r.off] = r.alpha;
            w[r.off] = r.c;
        }
        return;
    }
    if (r.kind != 2) return;
    JacobiRecParams p;
    p.n = r.n;
    p.alpha = r.alpha;
    p.beta = r.beta;
    p.c = r.c;
    p.lo = 0;
    p.hi = r.n;
    jacobi_rec_block(p, x + r.off, w + r.off);
}

// ---------------------------------------------------------------------------------------------
// max(alpha, beta) >= 5, n <= 4000: the reference's jacobi_gw (gaussjacobi.jl:570-599) is a dense
// LAPACK eigen-solve of the n x n Jacobi matrix.  The per-node-parallel equivalent used here:
//   * thread k finds the k-th eigenvalue of the symmetric tridiagonal matrix by bisection on
//     Sturm counts (every eigenvalue independently, O(n) per count, ~60 counts);
//   * the weight is mu0 / sum_j p_j(x_k)^2 with the orthonormal polynomials p_j from the three-term
//     recurrence — the squared first component of the normalised eigenvector, without eigenvectors.
// The same kernel serves the general Gauss-Radau rule gaussradau(n, a, b) (gaussradau.jl:52-69), whose
// Jacobi matrix differs in its last diagonal entry.
// ---------------------------------------------------------------------------------------------
// jacobi_jacobimatrix (:570-584); radau != 0 replaces the last diagonal entry (gaussradau.jl:63-64)
__global__ void jacobi_matrix_kernel(int n, double a, double b, int radau, double* diag, double* off) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // 0-based row
    if (i >= n) return;
    const double s = a + b;
    const double nn = (double)n;
    const double ii = (double)(i + 1);  // 1-based index
    double d;
    if (i == 0) {
        d = (b - a) / (2 + s);
    } else if (i < n - 1) {
        const double si = 2 * ii + s;
        d = (b * b - a * a) / ((si - 2) * si);
    } else {
        d = (b * b - a * a) / ((2 * nn - 2 + s) * (2 * nn + s));
    }
    if (radau && i == n - 1) {
        const double m = nn - 1;
        d = -1 + 2 * m * (m + a) / ((2 * m + s) * (2 * m + s + 1));
    }
    diag[i] = d;
    if (i < n - 1) {
        double e;
        if (i == 0) {
            e = 2 * sqrt((1 + a) * (1 + b) / (s + 3)) / (s + 2);
        } else {
            const double si = 2 * ii + s;
            e = 2 * sqrt(ii * (ii + a) * (ii + b) * (ii + s) / (si * si - 1)) / si;
        }
        off[i] = e;
    }
}

__device__ __forceinline__ int sturm_count(int n, const double* __restrict__ d, const double* __restrict__ e,
                                           double x, double pivmin) {
    // LAPACK dlaebz convention: a pivot inside (-pivmin, pivmin) is replaced by -pivmin BEFORE it is counted
    double q = d[0] - x;
    if (fabs(q) < pivmin) q = -pivmin;
    int cnt = q <= 0.0;
    for (int i = 1; i < n; ++i) {
        q = (d[i] - x) - e[i - 1] * e[i - 1] / q;
        if (fabs(q) < pivmin) q = -pivmin;
        cnt += q <= 0.0;
    }
    return cnt;
}

__global__ void __launch_bounds__(128)
tridiag_gw_kernel(int n, const double* __restrict__ d, const double* __restrict__ e, double mu0, int fix_first,
                  double* x, double* w) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    // Gershgorin interval and pivmin
    double lo = 1e300, hi = -1e300, emax2 = 0.0;
    for (int i = 0; i < n; ++i) {
        const double r = (i > 0 ? fabs(e[i - 1]) : 0.0) + (i < n - 1 ? fabs(e[i]) : 0.0);
        lo = fmin(lo, d[i] - r);
        hi = fmax(hi, d[i] + r);
        if (i < n - 1) emax2 = fmax(emax2, e[i] * e[i]);
    }
    const double pivmin = 2.2250738585072014e-308 * fmax(1.0, emax2) * 4;
    const double span = hi - lo;
    lo -= 2.220446049250313e-16 * span * n + pivmin;
    hi += 2.220446049250313e-16 * span * n + pivmin;
    for (int it = 0; it < 200; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (mid <= lo || mid >= hi) break;  // interval exhausted to the last ulp
        if (sturm_count(n, d, e, mid, pivmin) >= k + 1) hi = mid; else lo = mid;
    }
    const double xk = 0.5 * (lo + hi);
    // orthonormal recurrence: e_j p_{j+1} = (x - d_j) p_j - e_{j-1} p_{j-1}, p_0 = 1
    double pm1 = 0.0, p = 1.0, ssum = 1.0;
    for (int j = 0; j < n - 1; ++j) {
        const double pn = ((xk - d[j]) * p - (j > 0 ? e[j - 1] : 0.0) * pm1) / e[j];
        pm1 = p;
        p = pn;
        ssum += p * p;
    }
    x[k] = (fix_first && k == 0) ? -1.0 : xk;  // gaussradau.jl:67
    w[k] = mu0 / ssum;
}

static double jacobi_tt_host(int64_t n, double a, double b, int64_t i) {
    const double PI = 3.141592653589793;
    const double j = (double)(n - i);
    const double d = 2 * (double)n + a + b + 1;
    const double K = PI * (2 * j + a - 0.5) / d;
    return K + (1 / (d * d)) * ((0.25 - a * a) * (1.0 / tan(K / 2)) - (0.25 - b * b) * tan(K / 2));
}

// Norm passes: only nodes with theta < tskip are evaluated (see jacobi_norm_skip_angle), and theta = tt(i) (right
// half, h = 0) / pi - tt(i) (left half) is monotone in i, small at the outer end of the half.  Bound the node range
// generously on the initial guesses — angle 2 tskip (Newton moves an interior theta by a 1e-6 fraction at most),
// never fewer than 4096 nodes — and return the tiles (of JAC_THREADS nodes, counted within the half) that cover it.
// A half whose inner end could exceed JAC_SKIP_TMAX (it cannot: the split is at pi/2) keeps every tile.
static void jacobi_norm_tiles(int64_t n, double a, double b, int64_t n_left, int h, int64_t count, double tskip,
                              int64_t* tile_lo, int64_t* tile_hi) {
    const double PI = 3.141592653589793;
    const int64_t ntiles = (count + JAC_THREADS - 1) / JAC_THREADS;
    *tile_lo = 0;
    *tile_hi = ntiles;
    if (count <= 8192 || !(tskip > 0)) return;
    const double inner = h == 0 ? jacobi_tt_host(n, a, b, n_left) : PI - jacobi_tt_host(n, a, b, n_left - 1);
    if (!(inner < JAC_SKIP_TMAX - 5e-4)) return;
    const double bound = 2 * tskip;
    auto theta_at = [&](int64_t c) {  // initial theta of the c-th node from the outer end of the half
        return h == 0 ? jacobi_tt_host(n, a, b, n - 1 - c) : PI - jacobi_tt_host(n, a, b, c);
    };
    if (!(theta_at(count - 1) >= bound)) return;  // the whole half is below the bound: every tile
    int64_t lo_c = 0, hi_c = count - 1;           // first c with theta_at(c) >= bound, by bisection
    while (hi_c - lo_c > 1) {
        const int64_t mid = lo_c + (hi_c - lo_c) / 2;
        if (theta_at(mid) >= bound) hi_c = mid; else lo_c = mid;
    }
    int64_t keep = hi_c + 1;
    if (keep < 4096) keep = 4096;
    if (keep >= count) return;
    if (h == 0) {
        *tile_lo = (count - keep) / JAC_THREADS;                // nodes r >= count - keep of the half
    } else {
        *tile_hi = (keep + JAC_THREADS - 1) / JAC_THREADS;      // nodes r < keep of the half
    }
}


}  // namespace fgq

using namespace fgq;

// Dispatch of gaussjacobi.jl:23-55 for the general (alpha, beta): 0 ok (asy), -1 ok (rec / closed form).
static int jacobi_check_args(int64_t n, double a, double b) {
    if (n < 0) {
        set_error("DomainError(%lld): gaussjacobi(%lld,%g,%g) not defined: n must be non-negative.",
                  (long long)n, (long long)n, a, b);
        return FGQ_EDOMAIN;
    }
    if (n <= 1) return FGQ_OK;  // :40-43 come before the integrability check in the reference
    if (!(a > -1.0) || !(b > -1.0)) {
        set_error("DomainError((%g, %g)): The Jacobi parameters correspond to a nonintegrable weight function", a, b);
        return FGQ_EDOMAIN;
    }
    const double mx = a > b ? a : b;
    if (mx < 5.0) return FGQ_OK;
    if (n > 4000) {  // :52-53: ErrorException in the reference
        set_error("gaussjacobi(%lld,%g,%g) is not implemented: n must be <= 4000 for max(a,b) >= 5.", (long long)n, a, b);
        return FGQ_ENOTIMPL;
    }
    return FGQ_OK;
}

// jacobimoment (:586-591)
static double jacobi_moment(double a, double b) {
    const double s = a + b;
    return exp((s + 1) * log(2.0) + lgamma(a + 1) + lgamma(b + 1) - lgamma(2 + s));
}

// jacobi_gw (:593-599) / gaussradau(n, a, b) (gaussradau.jl:52-69) through bisection + recurrence
static int jacobi_gw_device(int64_t n, double a, double b, int radau, int64_t lo, int64_t hi, double* x_dev,
                            double* w_dev, void* workspace_dev, cudaStream_t st) {
    double* diag = (double*)((char*)workspace_dev + 4096);
    double* off = diag + n;
    double* xfull = x_dev;
    double* wfull = w_dev;
    if (lo != 0 || hi != n) {
        set_error("the Golub-Welsch branch writes whole rules only (lo = 0, hi = n)");
        return FGQ_EINVAL;
    }
    jacobi_matrix_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>((int)n, a, b, radau, diag, off);
    FGQ_LAUNCH_CHECK();
    tridiag_gw_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>((int)n, diag, off, jacobi_moment(a, b), radau, xfull, wfull);
    FGQ_LAUNCH_CHECK();
    return FGQ_OK;
}

// :42-43: x = (b-a)/(a+b+2), w = 2^(a+b+1) beta(a+1, b+1)
static void jacobi_one_node(double a, double b, double* x, double* w) {
    *x = (b - a) / (a + b + 2);
    *w = pow(2.0, a + b + 1) * (tgamma(a + 1) * tgamma(b + 1) / tgamma(a + b + 2));
}

// the weight normalisation of jacobi_rec (:89)
static double jacobi_rec_const(double a, double b) {
    return pow(2.0, a + b + 1) * tgamma(2 + a) * tgamma(2 + b) / (tgamma(2 + a + b) * (a + 1) * (b + 1));
}

static int jacobi_rec_device(int64_t n, double a, double b, int64_t lo, int64_t hi, double* x_dev, double* w_dev,
                             cudaStream_t st) {
    JacobiRecParams p;
    p.n = (int)n;
    p.alpha = a;
    p.beta = b;
    p.lo = lo;
    p.hi = hi;
    if (n == 1) {
        // :42-43: x = (b-a)/(a+b+2), w = 2^(a+b+1) beta(a+1, b+1); reuse the patch kernel for the single node
        JacobiPatch q;
        q.n = 2 * JAC_NBDY;  // maps patch slot 0 to output index 0
        q.lo = lo;
        q.hi = hi;
        for (int i = 0; i < 2 * JAC_NBDY; ++i) q.x[i] = q.w[i] = 0.0;
        jacobi_one_node(a, b, &q.x[0], &q.w[0]);
        if (hi > 1) q.hi = 1;
        jacobi_patch_kernel<<<1, 32, 0, st>>>(q, x_dev, w_dev);
        FGQ_LAUNCH_CHECK();
        return FGQ_OK;
    }
    p.c = jacobi_rec_const(a, b);
    jacobi_rec_kernel<<<1, 128, 0, st>>>(p, x_dev, w_dev);
    FGQ_LAUNCH_CHECK();
    return FGQ_OK;
}

extern "C" int64_t fgq_gaussjacobi_workspace_bytes(int64_t n) {
    if (n < 0) return 0;
    return 4096 + 16 * n;
}

// The truncation rule (:309) only asks whether max over the interior set of |dS1 + dS2| of a term m >= 12 is
// below eps/100.  |dS1 + dS2| <= sum_l |PHI[l][m-1]| csc^l sec^(m-1-l) with csc = 1/(2 sin(t/2)) decreasing in t
// and sec = 1/(2 cos(t/2)) <= 0.7079 up to t = pi/2 + 1e-3: beyond the angle returned here every candidate term
// is below a QUARTER of the threshold (the margin covers the rounding of the device sums by orders of
// magnitude), so such a node can never make a norm reach the threshold and the norm passes skip it.  The
// comparison, hence the truncation index and every result bit, is unchanged; the stored norms become the
// maxima over the nodes that were looked at.  +inf: no node may be skipped (small n).
static double jacobi_norm_skip_angle(const JacobiHalfPlan& hp) {
    const double thr = 2.220446049250313e-16 / 100 / 4;
    auto bound = [&](double t) {
        const double csc = (1.0 / sin(t / 2)) / 2 * (1 + 1e-9);
        const double sec = 0.7079;
        double worst = 0.0;
        for (int m = 12; m <= JAC_M; ++m) {
            double sum = 0.0;
            for (int l = 0; l < m; ++l) sum += fabs(hp.PHI[l][m - 1]) * pow(csc, l) * pow(sec, m - 1 - l);
            if (!(sum <= worst)) worst = sum;  // a NaN table entry disables skipping
        }
        return worst * (1 + 1e-9);
    };
    const double PIO2 = 3.141592653589793 / 2;
    if (!(bound(PIO2) < thr)) return HUGE_VAL;
    double lo = 1e-300, hi = PIO2;  // bound(lo) >= thr (or overflow), bound(hi) < thr
    if (bound(lo) < thr) return lo;
    for (int it = 0; it < 200 && hi - lo > 1e-12 * hi; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (bound(mid) < thr) hi = mid; else lo = mid;
    }
    return hi;
}

// Core: the (alpha, beta) != (0,0), (+-1/2, +-1/2) asymptotic branch.
static int jacobi_asy_device(int64_t n, double a, double b, int64_t lo, int64_t hi, double* x_dev, double* w_dev,
                             void* workspace_dev, cudaStream_t st) {
    JacobiState* state = (JacobiState*)workspace_dev;
    static_assert(sizeof(JacobiState) <= 4096, "state must fit the workspace header");
    double* theta = (double*)((char*)workspace_dev + 4096);
    FGQ_CUDA_TRY(cudaMemsetAsync(state, 0, sizeof(JacobiState), st));
    // split: tt is decreasing in the index; n_left = number of tt > pi/2 (:204)
    const double PIO2 = 3.141592653589793 / 2;
    int64_t n_left;
    if (!(jacobi_tt_host(n, a, b, 0) > PIO2)) {
        n_left = 0;
    } else if (jacobi_tt_host(n, a, b, n - 1) > PIO2) {
        n_left = n;
    } else {
        int64_t l = 0, r = n - 1;  // tt(l) > pi/2 >= tt(r)
        while (r - l > 1) {
            const int64_t mid = l + (r - l) / 2;
            if (jacobi_tt_host(n, a, b, mid) > PIO2) l = mid; else r = mid;
        }
        n_left = r;
    }
    if (n_left < 2 * JAC_NBDY || n - n_left < 2 * JAC_NBDY) {
        set_error("internal: degenerate interior split (%lld of %lld)", (long long)n_left, (long long)n);
        return FGQ_EINVAL;
    }
    const double wc = jacobi_weights_constant(n, a, b);
    // Boundary nodes (:182-190): asy2 with (a, b) at the right end, with (b, a) mirrored at the left end (for
    // a == b the reference reuses the right-end result; the same solve gives the same numbers).  One
    // launch of two blocks (one per end), independent of the interior: forked onto a side stream and joined at the end
    // (event dependencies only, so the call stays stream-ordered and graph-capturable).
    ScopedStream side_holder;   // a side stream and two events of this call's own (returned on every exit path)
    FGQ_TRY(side_holder.acquire());
    cudaStream_t side = side_holder.st;
    ScopedEvents evs;
    FGQ_TRY(evs.acquire(2));
    cudaEvent_t ev_fork = evs.ev[0], ev_join = evs.ev[1];
    FGQ_CUDA_TRY(cudaEventRecord(ev_fork, st));
    FGQ_CUDA_TRY(cudaStreamWaitEvent(side, ev_fork, 0));
    {
        JacobiBoundaryPlans plans;
        for (int sd = 0; sd < 2; ++sd) {
            JacobiBoundaryPlan& bp = plans.end[sd];
            jacobi_boundary_plan(n, sd == 0 ? a : b, sd == 0 ? b : a, &bp);
            bp.wc = wc;
            bp.side = sd;
            bp.lo = lo;
            bp.hi = hi;
        }
        static const cudaError_t smem_attr = cudaFuncSetAttribute(
            jacobi_boundary_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)JB_TABLE_BYTES);
        FGQ_CUDA_TRY(smem_attr);
        jacobi_boundary_kernel<<<2, JB_THREADS, JB_TABLE_BYTES, side>>>(plans, x_dev, w_dev);  // one block per end
        FGQ_LAUNCH_CHECK();
    }
    FGQ_CUDA_TRY(cudaEventRecord(ev_join, side));
    jacobi_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, a, b, n_left, theta);
    FGQ_LAUNCH_CHECK();

    JacobiLaunch L[2];
    jacobi_interior_plan(n, a, b, &L[0].hp);
    jacobi_interior_plan(n, b, a, &L[1].hp);
    for (int h = 0; h < 2; ++h) {
        L[h].half = h;
        L[h].n = n;
        L[h].wc = wc;
        L[h].lo = lo;
        L[h].hi = hi;
        L[h].tol = sqrt(2.220446049250313e-16) / 100;
        L[h].tskip = jacobi_norm_skip_angle(L[h].hp);
    }
    L[0].first = n_left;          L[0].count = n - n_left;
    L[0].idx_lo = n_left;         L[0].idx_hi = n - (JAC_NBDY - 1);   // all but the last 9 (:206)
    L[1].first = 0;               L[1].count = n_left;
    L[1].idx_lo = JAC_NBDY;       L[1].idx_hi = n_left;              // all but the first 10 (:228)
    for (int h = 0; h < 2; ++h)
        jacobi_norm_tiles(n, a, b, n_left, h, L[h].count, L[h].tskip, &L[h].norm_tile_lo, &L[h].norm_tile_hi);
    // The two halves are independent chains (their own phase flags, norms, tickets, theta ranges and outputs): the
    // left half's 33 launches go to a second side stream, forked after the initial guesses and joined at the end, so
    // that the launch gaps and the converged-and-return launches of one chain are hidden behind the other's work
    // (FGQ_JACOBI_TWO_STREAMS=0: one stream, as in round 1).
    static const bool two_streams = []() {
        const char* e = getenv("FGQ_JACOBI_TWO_STREAMS");
        return !(e && atoi(e) == 0);
    }();
    ScopedStream left_holder;
    ScopedEvents evs2;
    cudaStream_t chain[2] = {st, st};
    // (measured: n = 1e7 2.54 -> 2.42 ms, n = 1e6 0.73 -> 0.68 ms, but n = 1e5 0.52 -> 0.81 ms: for small rules every
    // launch is a no-op-sized kernel and the second stream only adds cross-stream scheduling)
    const bool split = two_streams && n >= 500000;
    if (split) {
        FGQ_TRY(left_holder.acquire());
        FGQ_TRY(evs2.acquire(2));
        chain[1] = left_holder.st;
        FGQ_CUDA_TRY(cudaEventRecord(evs2.ev[0], st));              // after jacobi_init_kernel
        FGQ_CUDA_TRY(cudaStreamWaitEvent(chain[1], evs2.ev[0], 0));
    }
    for (int s = 0; s < 11; ++s) {
        for (int h = 0; h < 2; ++h) {
            L[h].sweep = s;
            const int64_t tiles = (L[h].count + JAC_THREADS - 1) / JAC_THREADS;
            const unsigned blocks = (unsigned)(tiles < JAC_GRID ? tiles : JAC_GRID);
            FGQ_CUDA_TRY(launch_pdl(jacobi_pass_kernel<1>, blocks, JAC_THREADS, chain[h], L[h], state, theta, x_dev, w_dev));
            count_launch();
            FGQ_CUDA_TRY(launch_pdl(jacobi_pass_kernel<2>, blocks, JAC_THREADS, chain[h], L[h], state, theta, x_dev, w_dev));
            count_launch();
            FGQ_CUDA_TRY(launch_pdl(jacobi_pass_kernel<0>, blocks, JAC_THREADS, chain[h], L[h], state, theta, x_dev, w_dev));
            count_launch();
        }
    }
    if (split) {
        FGQ_CUDA_TRY(cudaEventRecord(evs2.ev[1], chain[1]));
        FGQ_CUDA_TRY(cudaStreamWaitEvent(st, evs2.ev[1], 0));
    }
    FGQ_CUDA_TRY(cudaStreamWaitEvent(st, ev_join, 0));
    return FGQ_OK;
}

extern "C" int fgq_gausslegendre_device(int64_t n, int64_t lo, int64_t hi, double* x_dev, double* w_dev,
                                        void* stream);
extern "C" int fgq_gausschebyshev_device(int kind, int64_t n, int64_t lo, int64_t hi, double* x_dev,
                                         double* w_dev, void* stream);

extern "C" int fgq_gaussjacobi_device(int64_t n, double a, double b, int64_t lo, int64_t hi, double* x_dev,
                                      double* w_dev, void* workspace_dev, void* stream) {
    clear_status();
    if (n < 0) return jacobi_check_args(n, a, b);
    if (lo < 0 || hi < lo || hi > n) {
        set_error("invalid shard [%lld, %lld) of n=%lld", (long long)lo, (long long)hi, (long long)n);
        return FGQ_EINVAL;
    }
    if (a == 0.0 && b == 0.0) return fgq_gausslegendre_device(n, lo, hi, x_dev, w_dev, stream);  // :29-30
    if (fabs(a) == 0.5 && fabs(b) == 0.5) {  // :30-39
        const int kind = (a < 0 && b < 0) ? 1 : (a > 0 && b > 0) ? 2 : (a < 0 ? 3 : 4);
        return fgq_gausschebyshev_device(kind, n, lo, hi, x_dev, w_dev, stream);
    }
    if (n == 0) return FGQ_OK;
    FGQ_TRY(jacobi_check_args(n, a, b));
    if (hi > lo && (!x_dev || !w_dev)) {
        set_error("null pointer");
        return FGQ_EINVAL;
    }
    if (n == 1) return jacobi_rec_device(n, a, b, lo, hi, x_dev, w_dev, as_stream(stream));   // :42-43
    if ((a > b ? a : b) >= 5.0) {                                                              // :50-51
        if (!workspace_dev) {
            set_error("null workspace pointer");
            return FGQ_EINVAL;
        }
        return jacobi_gw_device(n, a, b, 0, lo, hi, x_dev, w_dev, workspace_dev, as_stream(stream));
    }
    if (n <= 100) return jacobi_rec_device(n, a, b, lo, hi, x_dev, w_dev, as_stream(stream));  // :46-47
    if (!workspace_dev) {
        set_error("null workspace pointer");
        return FGQ_EINVAL;
    }
    return jacobi_asy_device(n, a, b, lo, hi, x_dev, w_dev, workspace_dev, as_stream(stream));
}

// ---- batched small rules: n <= 100 ----
constexpr int JACOBI_BATCH_MAX_N = 100;
constexpr size_t JACOBI_BATCH_GW_BYTES = 8192;  // workspace of one Golub-Welsch rule (4096 + 16 n)

extern "C" int64_t fgq_gaussjacobi_batch_workspace_bytes(int64_t nrules) {
    if (nrules < 0) return 0;
    return (int64_t)JACOBI_BATCH_GW_BYTES + nrules * (int64_t)sizeof(JacobiBatchRule);
}

// Classifies every rule the way fgq_gaussjacobi_device dispatches it.  `others` receives the rules served
// by single-rule launches (the Legendre / Chebyshev routes of (0,0), (+-1/2,+-1/2) and the Golub-Welsch
// region max(a,b) >= 5).
static int jacobi_batch_plan(int64_t nrules, const int32_t* n_i, const double* alpha_i, const double* beta_i,
                             const int64_t* offsets, std::vector<JacobiBatchRule>* rules,
                             std::vector<int64_t>* others, int64_t* total) {
    std::vector<SmallRule> base;
    FGQ_TRY(small_batch_rules(nrules, n_i, offsets, JACOBI_BATCH_MAX_N, "gaussjacobi", rules ? &base : nullptr, total));
    if (nrules > 0 && (!alpha_i || !beta_i)) {
        set_error("invalid batch description");
        return FGQ_EINVAL;
    }
    if (rules) rules->resize((size_t)nrules);
    for (int64_t i = 0; i < nrules; ++i) {
        const int n = n_i[i];
        const double a = alpha_i[i], b = beta_i[i];
        JacobiBatchRule r{n, 0, offsets[i], a, b, 0.0};
        const bool special = (a == 0.0 && b == 0.0) || (fabs(a) == 0.5 && fabs(b) == 0.5);
        if (special) {
            if (n > 0 && others) others->push_back(i);
        } else if (n > 0) {
            const int rc = jacobi_check_args(n, a, b);
            if (rc != FGQ_OK) return rc;
            if (n == 1) {
                r.kind = 1;
            } else if ((a > b ? a : b) >= 5.0) {
                if (others) others->push_back(i);
            } else {
                r.kind = 2;
            }
        }
        if (rules) (*rules)[(size_t)i] = r;
    }
    // the gamma-function constants (4 tgamma + pow per rule), on host threads for a large batch
    if (rules)
        host_parallel_for(nrules, [&](int64_t lo, int64_t hi) {
            for (int64_t i = lo; i < hi; ++i) {
                JacobiBatchRule& r = (*rules)[(size_t)i];
                const double a = r.alpha, b = r.beta;
                if (r.kind == 1) jacobi_one_node(a, b, &r.alpha, &r.c);
                if (r.kind == 2) r.c = jacobi_rec_const(a, b);
            }
        });
    return FGQ_OK;
}

extern "C" int fgq_gaussjacobi_batch_device(int64_t nrules, const int32_t* n_i, const double* alpha_i,
                                            const double* beta_i, const int64_t* offsets, double* x_dev,
                                            double* w_dev, void* workspace_dev, void* stream) {
    clear_status();
    std::vector<JacobiBatchRule> rules;
    std::vector<int64_t> others;
    int64_t total = 0;
    FGQ_TRY(jacobi_batch_plan(nrules, n_i, alpha_i, beta_i, offsets, &rules, &others, &total));
    if (nrules == 0 || total == 0) return FGQ_OK;
    if (!x_dev || !w_dev || !workspace_dev) {
        set_error("null pointer");
        return FGQ_EINVAL;
    }
    cudaStream_t st = as_stream(stream);
    char* ws = (char*)workspace_dev;
    // rules with work for the batch kernel, wide class (n > 64) first; each carries its own output offset
    std::vector<JacobiBatchRule> sorted;
    sorted.reserve(rules.size());
    for (const JacobiBatchRule& r : rules)
        if (r.kind != 0 && r.n > 64) sorted.push_back(r);
    const size_t n_wide = sorted.size();
    for (const JacobiBatchRule& r : rules)
        if (r.kind != 0 && r.n <= 64) sorted.push_back(r);
    if (!sorted.empty())
        FGQ_CUDA_TRY(cudaMemcpyAsync(ws + JACOBI_BATCH_GW_BYTES, sorted.data(), sorted.size() * sizeof(JacobiBatchRule),
                                     cudaMemcpyHostToDevice, st));
    const JacobiBatchRule* drules = (const JacobiBatchRule*)(ws + JACOBI_BATCH_GW_BYTES);
    if (n_wide) {
        jacobi_small_batch_kernel<<<(unsigned)n_wide, 128, 0, st>>>(drules, x_dev, w_dev);
        FGQ_LAUNCH_CHECK();
    }
    if (sorted.size() > n_wide) {
        jacobi_small_batch_kernel<<<(unsigned)(sorted.size() - n_wide), 64, 0, st>>>(drules + n_wide, x_dev, w_dev);
        FGQ_LAUNCH_CHECK();
    }
    for (int64_t i : others) {
        const JacobiBatchRule& r = rules[(size_t)i];
        FGQ_TRY(fgq_gaussjacobi_device(r.n, r.alpha, r.beta, 0, r.n, x_dev + r.off, w_dev + r.off, ws, stream));
    }
    return FGQ_OK;
}

extern "C" int fgq_gaussjacobi_batch_host(int64_t nrules, const int32_t* n_i, const double* alpha_i,
                                          const double* beta_i, const int64_t* offsets, double* x, double* w) {
    clear_status();
    int64_t total = 0;
    FGQ_TRY(jacobi_batch_plan(nrules, n_i, alpha_i, beta_i, offsets, nullptr, nullptr, &total));
    if (total == 0) return FGQ_OK;
    if (!x || !w) {
        set_error("null output pointer");
        return FGQ_EINVAL;
    }
    HostPathLock lock;
    void *ws, *out;
    FGQ_TRY(scratch_device((size_t)fgq_gaussjacobi_batch_workspace_bytes(nrules), 6, &ws));
    FGQ_TRY(scratch_device((size_t)total * 16, 0, &out));
    cudaStream_t st;
    FGQ_TRY(scratch_stream(0, &st));
    double* xd = (double*)out;
    double* wd = xd + total;
    FGQ_TRY(fgq_gaussjacobi_batch_device(nrules, n_i, alpha_i, beta_i, offsets, xd, wd, ws, (void*)st));
    FGQ_TRY(drain_to_host(x, w, xd, wd, total, st));
    return FGQ_OK;
}

// mode 1 = gaussradau(n), mode 2 = gausslobatto(n); the inner Jacobi rule has n - mode nodes and is
// written one slot to the right so that the epilogue can fill the fixed endpoint(s) in place.
static int endpoint_rule_device(int mode, int64_t n, double* x_dev, double* w_dev, void* workspace_dev,
                                cudaStream_t st) {
    // closed forms: gaussradau.jl:34-37, gausslobatto.jl:36-39
    if ((mode == 1 && n <= 2) || (mode == 2 && n <= 3)) {
        JacobiPatch q;
        q.n = 2 * JAC_NBDY;
        q.lo = 0;
        q.hi = n;
        for (int i = 0; i < 2 * JAC_NBDY; ++i) q.x[i] = q.w[i] = 0.0;
        if (mode == 1 && n == 1) { q.x[0] = -1; q.w[0] = 2; }
        if (mode == 1 && n == 2) { q.x[0] = -1; q.x[1] = 1.0 / 3; q.w[0] = 0.5; q.w[1] = 1.5; }
        if (mode == 2 && n == 2) { q.x[0] = -1; q.x[1] = 1; q.w[0] = 1; q.w[1] = 1; }
        if (mode == 2 && n == 3) { q.x[0] = -1; q.x[1] = 0; q.x[2] = 1; q.w[0] = 1.0 / 3; q.w[1] = 4.0 / 3; q.w[2] = 1.0 / 3; }
        jacobi_patch_kernel<<<1, 32, 0, st>>>(q, x_dev, w_dev);
        FGQ_LAUNCH_CHECK();
        return FGQ_OK;
    }
    const int64_t inner = n - mode;
    const double a = mode == 1 ? 0.0 : 1.0, b = 1.0;
    FGQ_TRY(jacobi_check_args(inner, a, b));
    if (inner <= 100) FGQ_TRY(jacobi_rec_device(inner, a, b, 0, inner, x_dev + 1, w_dev + 1, st));
    else FGQ_TRY(jacobi_asy_device(inner, a, b, 0, inner, x_dev + 1, w_dev + 1, workspace_dev, st));
    endpoint_epilogue_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(mode, n, 0, n, x_dev, w_dev);
    FGQ_LAUNCH_CHECK();
    return FGQ_OK;
}

extern "C" int fgq_gaussradau_device(int64_t n, double* x_dev, double* w_dev, void* workspace_dev, void* stream) {
    clear_status();
    if (n <= 0) {
        set_error("DomainError(%lld): Input n must be a positive integer", (long long)n);  // gaussradau.jl:34
        return FGQ_EDOMAIN;
    }
    if (!x_dev || !w_dev || !workspace_dev) {
        set_error("null pointer");
        return FGQ_EINVAL;
    }
    return endpoint_rule_device(1, n, x_dev, w_dev, workspace_dev, as_stream(stream));
}

extern "C" int fgq_gausslobatto_device(int64_t n, double* x_dev, double* w_dev, void* workspace_dev, void* stream) {
    clear_status();
    if (n <= 1) {
        set_error("DomainError(%lld): Lobatto undefined for n <= 1.", (long long)n);  // gausslobatto.jl:34
        return FGQ_EDOMAIN;
    }
    if (!x_dev || !w_dev || !workspace_dev) {
        set_error("null pointer");
        return FGQ_EINVAL;
    }
    return endpoint_rule_device(2, n, x_dev, w_dev, workspace_dev, as_stream(stream));
}

// ---- host-returning drop-ins --------------------------------------------------------------------

extern "C" int fgq_gaussradau_jacobi_device(int64_t n, double a, double b, double* x_dev, double* w_dev,
                                            void* workspace_dev, void* stream);

static int jacobi_family_host(int which, int64_t n, double a, double b, double* x, double* w) {
    if (n == 0) return FGQ_OK;
    if (!x || !w) {
        set_error("null output pointer");
        return FGQ_EINVAL;
    }
    HostPathLock lock;
    void *ws, *out;
    FGQ_TRY(scratch_device((size_t)fgq_gaussjacobi_workspace_bytes(n), 6, &ws));
    FGQ_TRY(scratch_device((size_t)n * 16 + 64, 0, &out));
    cudaStream_t st;
    FGQ_TRY(scratch_stream(0, &st));
    double* xd = (double*)out;
    double* wd = xd + n + (n & 1);
    int rc;
    if (which == 0) rc = fgq_gaussjacobi_device(n, a, b, 0, n, xd, wd, ws, (void*)st);
    else if (which == 1) rc = fgq_gaussradau_device(n, xd, wd, ws, (void*)st);
    else if (which == 3) rc = fgq_gaussradau_jacobi_device(n, a, b, xd, wd, ws, (void*)st);
    else rc = fgq_gausslobatto_device(n, xd, wd, ws, (void*)st);
    if (rc != FGQ_OK) return rc;
    FGQ_TRY(drain_to_host(x, w, xd, wd, n, st));
    return FGQ_OK;
}

extern "C" int fgq_gaussjacobi_host(int64_t n, double a, double b, double* x, double* w) {
    clear_status();
    if (n < 0) return jacobi_check_args(n, a, b);
    return jacobi_family_host(0, n, a, b, x, w);
}
extern "C" int fgq_gaussradau_host(int64_t n, double* x, double* w) {
    clear_status();
    if (n <= 0) {
        set_error("DomainError(%lld): Input n must be a positive integer", (long long)n);
        return FGQ_EDOMAIN;
    }
    return jacobi_family_host(1, n, 0, 0, x, w);
}
extern "C" int fgq_gausslobatto_host(int64_t n, double* x, double* w) {
    clear_status();
    if (n <= 1) {
        set_error("DomainError(%lld): Lobatto undefined for n <= 1.", (long long)n);
        return FGQ_EDOMAIN;
    }
    return jacobi_family_host(2, n, 0, 0, x, w);
}

// gaussradau(n, a, b) — gaussradau.jl:52-69
extern "C" int fgq_gaussradau_jacobi_device(int64_t n, double a, double b, double* x_dev, double* w_dev,
                                            void* workspace_dev, void* stream) {
    clear_status();
    if (n <= 0) {
        set_error("DomainError(%lld): Input N must be a positive integer", (long long)n);  // :53-55
        return FGQ_EDOMAIN;
    }
    if (n > 4000) {
        set_error("gaussradau(%lld,%g,%g): the eigenvalue kernel is sized for n <= 4000", (long long)n, a, b);
        return FGQ_ENOTIMPL;
    }
    if (!x_dev || !w_dev || !workspace_dev) {
        set_error("null pointer");
        return FGQ_EINVAL;
    }
    if (n == 1) {  // :61
        JacobiPatch q;
        q.n = 2 * JAC_NBDY;
        q.lo = 0;
        q.hi = 1;
        for (int i = 0; i < 2 * JAC_NBDY; ++i) q.x[i] = q.w[i] = 0.0;
        q.x[0] = -1.0;
        q.w[0] = jacobi_moment(a, b);
        jacobi_patch_kernel<<<1, 32, 0, as_stream(stream)>>>(q, x_dev, w_dev);
        FGQ_LAUNCH_CHECK();
        return FGQ_OK;
    }
    return jacobi_gw_device(n, a, b, 1, 0, n, x_dev, w_dev, workspace_dev, as_stream(stream));
}

extern "C" int fgq_gaussradau_jacobi_host(int64_t n, double a, double b, double* x, double* w) {
    clear_status();
    if (n <= 0) {
        set_error("DomainError(%lld): Input N must be a positive integer", (long long)n);
        return FGQ_EDOMAIN;
    }
    return jacobi_family_host(3, n, a, b, x, w);
}

// sweeps[2] (right, left), terms[2] (series terms kept in the last evaluation): for tests/reporting
// The angles beyond which the norm passes skip a node (jacobi_norm_skip_angle), for the right half (a, b) and the
// mirrored left half (b, a) of gaussjacobi(n, a, b): host-only, exported so that the bound can be checked against
// the oracle's series terms without a device.  +inf = nothing may be skipped.
extern "C" int fgq_debug_jacobi_skip_angles(int64_t n, double a, double b, double* t_right, double* t_left) {
    clear_status();
    if (n < 2 || !t_right || !t_left) return FGQ_EINVAL;
    static JacobiHalfPlan hp;  // 6.5 KB: off the stack
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    jacobi_interior_plan(n, a, b, &hp);
    *t_right = jacobi_norm_skip_angle(hp);
    jacobi_interior_plan(n, b, a, &hp);
    *t_left = jacobi_norm_skip_angle(hp);
    return FGQ_OK;
}

// The split index and, per half (right, left), the node range [first, first + count) and the tiles its norm passes
// walk: out = {n_left, lo_r, hi_r, lo_l, hi_l} (host only; tests check the bound against the initial guesses).
extern "C" int fgq_debug_jacobi_norm_tiles(int64_t n, double a, double b, int64_t* out) {
    clear_status();
    if (n < 101 || !out) return FGQ_EINVAL;
    const double PIO2 = 3.141592653589793 / 2;
    int64_t n_left;
    if (!(jacobi_tt_host(n, a, b, 0) > PIO2)) {
        n_left = 0;
    } else if (jacobi_tt_host(n, a, b, n - 1) > PIO2) {
        n_left = n;
    } else {
        int64_t l = 0, r = n - 1;
        while (r - l > 1) {
            const int64_t mid = l + (r - l) / 2;
            if (jacobi_tt_host(n, a, b, mid) > PIO2) l = mid; else r = mid;
        }
        n_left = r;
    }
    double tr, tl;
    FGQ_TRY(fgq_debug_jacobi_skip_angles(n, a, b, &tr, &tl));
    out[0] = n_left;
    jacobi_norm_tiles(n, a, b, n_left, 0, n - n_left, tr, &out[1], &out[2]);
    jacobi_norm_tiles(n, a, b, n_left, 1, n_left, tl, &out[3], &out[4]);
    return FGQ_OK;
}

extern "C" int fgq_gaussjacobi_stats(const void* workspace_dev, int* sweeps, int* terms) {
    clear_status();
    if (!workspace_dev || !sweeps || !terms) return FGQ_EINVAL;
    JacobiState h;
    FGQ_CUDA_TRY(cudaMemcpy(&h, workspace_dev, sizeof(h), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 2; ++k) {
        sweeps[k] = h.sweeps[k];
        terms[k] = h.terms[k][h.sweeps[k] < 11 ? h.sweeps[k] : 10];
    }
    return FGQ_OK;
}

// host-planned Bessel pieces, exported for tests (test_besselroots.jl:7-12)
extern "C" double fgq_besselj(double nu, double x) { return besselj_real(nu, x); }
