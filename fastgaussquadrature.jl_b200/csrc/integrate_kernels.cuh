// integrate_kernels.cuh — consumer kernels for a USER integrand (README.md:27-30: the only thing the
// reference's README does with a rule is dot(w, f.(x))).  This text is compiled at run time by NVRTC (integrate.cu)
// after fgq_math.h + legendre_device.cuh and a line
//     __device__ __forceinline__ double fgq_user_f(double x) { return (<the caller's expression in x>); }
// with -DFGQ_NT / -DFGQ_WT / -DFGQ_HEAD selecting the series branch of the rule (gausslegendre.jl:103-143,162-224).
//
//   fgq_integrate_legendre_kernel   generate-and-reduce: nodes and weights of half-indices [kmin, kmax] are produced in
//                                   registers and never stored; each half-index contributes w (f(-x) + f(x))
//   fgq_apply_rule_kernel           dot(w, f.(x)) over a stored rule of any family (Jacobi, Hermite, Laguerre, ...)
// sums[0] += sum w f(x), sums[1] += sum w (the rule's own check); Kahan accumulation per thread, block tree, one
// atomicAdd pair per block.

struct FgqIntegrateLaunch {
    long long kmin, kmax;       // half-indices evaluated (inclusive)
    long long kq0_last;         // last half-index with theta < pi/4 (host bisection), 0 = unknown
    long long kext_first;       // first half-index taking the extended pi/2 reduction
    long long kcentre;          // (n+1)/2 for odd n: that half-index is the single node x = 0
    fgq::LegConsts c;
};

__device__ __forceinline__ void fgq_kahan(double& s, double& comp, double v) {
    const double y = v - comp;
    const double t = s + y;
    comp = (t - s) - y;
    s = t;
}

__device__ __forceinline__ void fgq_block_sum2(double a, double b, double* sums) {
    __shared__ double sh[2][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        b += __shfl_down_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0) {
        sh[0][threadIdx.x >> 5] = a;
        sh[1][threadIdx.x >> 5] = b;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += sh[threadIdx.x][i];
        atomicAdd(sums + threadIdx.x, t);
    }
}

#ifndef FGQ_NT
#define FGQ_NT 0
#define FGQ_WT 0
#define FGQ_HEAD 0
#endif

extern "C" __global__ void __launch_bounds__(256) fgq_integrate_legendre_kernel(const FgqIntegrateLaunch p, double* sums) {
    using namespace fgq;
    constexpr bool HEAD = FGQ_HEAD != 0;
    double s0 = 0, c0 = 0, s1 = 0, c1 = 0;
    const long long ntiles = (p.kmax - p.kmin) / 512 + 1;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long kb0 = p.kmin + tile * 512, kb1 = kb0 + 511;
        const long long k0 = kb0 + 2 * threadIdx.x;
        double cx[2], w[2];
        const bool full = kb1 <= p.kmax;
        if (FGQ_NT == 0 && FGQ_WT == 0 && !HEAD && full && kb0 >= K_JK_EXACT && kb1 <= p.kq0_last) {
            leg_bulk<0>(k0, p.c, cx[0], w[0]);
            leg_bulk<0>(k0 + 1, p.c, cx[1], w[1]);
        } else if (FGQ_NT == 0 && FGQ_WT == 0 && !HEAD && full && kb0 >= K_JK_EXACT && kb0 > p.kq0_last + 1 &&
                   kb1 < p.kext_first) {
            leg_bulk<1>(k0, p.c, cx[0], w[0]);
            leg_bulk<1>(k0 + 1, p.c, cx[1], w[1]);
        } else {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const long long k = k0 + e > p.kmax ? p.kmax : k0 + e;
                leg_half_index<FGQ_NT, FGQ_WT, HEAD>(k, p.c, cx[e], w[e]);
            }
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const long long k = k0 + e;
            if (k > p.kmax) continue;
            if (k == p.kcentre) {
                fgq_kahan(s0, c0, w[e] * fgq_user_f(0.0));
                fgq_kahan(s1, c1, w[e]);
            } else {
                fgq_kahan(s0, c0, w[e] * (fgq_user_f(-cx[e]) + fgq_user_f(cx[e])));
                fgq_kahan(s1, c1, 2.0 * w[e]);
            }
        }
    }
    fgq_block_sum2(s0, s1, sums);
}

extern "C" __global__ void __launch_bounds__(256) fgq_apply_rule_kernel(const double* __restrict__ x,
                                                                         const double* __restrict__ w, long long count,
                                                                         double* sums) {
    double s0 = 0, c0 = 0, s1 = 0, c1 = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const double xi = __ldcs(x + i), wi = __ldcs(w + i);
        fgq_kahan(s0, c0, wi * fgq_user_f(xi));
        fgq_kahan(s1, c1, wi);
    }
    fgq_block_sum2(s0, s1, sums);
}
