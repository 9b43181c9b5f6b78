// chebyshev.cu — Gauss-Chebyshev rules of the four kinds: closed forms in cospi / sinpi.
//
// Replaces gausschebyshevt/u/v/w (src/gausschebyshev.jl:22-26, 51-55, 80-84, 109-113) and the
// (alpha, beta) = (+-1/2, +-1/2) routes of gaussjacobi (src/gaussjacobi.jl:30-39).  One thread
// per node, two 8-byte stores; output index i (ascending x) corresponds to k = n - i.
#include <math.h>

#include "common.cuh"

namespace fgq {

__global__ void __launch_bounds__(256)
chebyshev_kernel(int kind, int64_t n, int64_t lo, int64_t hi, double* x, double* w) {
    const int64_t i = lo + (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= hi) return;
    const double PI = 3.141592653589793;
    const double k = (double)(n - i);
    const double nn = (double)n;
    double xv, wv;
    if (kind == 1) {
        xv = cospi((2 * k - 1) / (2 * nn));
        wv = PI / nn;
    } else if (kind == 2) {
        const double s = sinpi(k / (nn + 1));
        xv = cospi(k / (nn + 1));
        wv = PI / (nn + 1) * (s * s);
    } else if (kind == 3) {
        const double c = cospi((2 * k - 1) / (4 * nn + 2));
        xv = cospi((2 * k - 1) / (2 * nn + 1));
        wv = 4 * PI / (2 * nn + 1) * (c * c);
    } else {
        const double s = sinpi(k / (2 * nn + 1));
        xv = cospi(2 * k / (2 * nn + 1));
        wv = 4 * PI / (2 * nn + 1) * (s * s);
    }
    st_stream(x + (i - lo), xv);
    st_stream(w + (i - lo), wv);
}

}  // namespace fgq

using namespace fgq;

extern "C" int fgq_gausschebyshev_device(int kind, int64_t n, int64_t lo, int64_t hi, double* x_dev,
                                         double* w_dev, void* stream) {
    clear_status();
    if (kind < 1 || kind > 4) {
        set_error("ArgumentError: Chebyshev kind should be 1, 2, 3, or 4");  // gausschebyshev.jl:141
        return FGQ_EINVAL;
    }
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
    chebyshev_kernel<<<(unsigned)((hi - lo + 255) / 256), 256, 0, as_stream(stream)>>>(kind, n, lo, hi, x_dev, w_dev);
    FGQ_LAUNCH_CHECK();
    return FGQ_OK;
}

struct ChebCtx {
    int kind;
    int64_t n;
};
static int cheb_slice(void* ctx, int64_t lo, int64_t hi, double* xd, double* wd, cudaStream_t st) {
    const ChebCtx* c = (const ChebCtx*)ctx;
    return fgq_gausschebyshev_device(c->kind, c->n, lo, hi, xd, wd, (void*)st);
}

extern "C" int fgq_gausschebyshev_host(int kind, int64_t n, double* x, double* w) {
    clear_status();
    if (kind < 1 || kind > 4) {
        set_error("ArgumentError: Chebyshev kind should be 1, 2, 3, or 4");
        return FGQ_EINVAL;
    }
    if (n < 0) {
        set_error("DomainError(%lld): Input n must be a non-negative integer", (long long)n);
        return FGQ_EDOMAIN;
    }
    ChebCtx c{kind, n};
    return host_pipeline(0, n, x, w, cheb_slice, &c);
}
