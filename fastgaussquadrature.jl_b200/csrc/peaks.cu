// peaks.cu — roofline denominators measured on the current device:
// a DFMA issue-rate microbenchmark and a store-only HBM streaming kernel.
#include "common.cuh"

namespace fgq {

constexpr int DFMA_ITERS = 4096;
constexpr int DFMA_CHAINS = 8;

// 8 independent FMA chains per thread; nothing but DFMA in the loop body.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, double a, double b) {
    double v[DFMA_CHAINS];
#pragma unroll
    for (int i = 0; i < DFMA_CHAINS; ++i) v[i] = (double)(threadIdx.x + i);
#pragma unroll 1
    for (int it = 0; it < DFMA_ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < DFMA_CHAINS; ++i) v[i] = fma(v[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < DFMA_CHAINS; ++i) s += v[i];
    if (s == 123.456) out[0] = s;  // never true; keeps the chains alive
}

// Same store shape as the Legendre kernels: one 16-byte streaming store per thread into each
// of four streams (x-left, w-left, x-right, w-right), a non-persistent grid.
__global__ void __launch_bounds__(256) store_peak_kernel(double2* out, int64_t n16) {
    const int64_t q = n16 / 4;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q) return;
    const double2 v = make_double2(1.0, 2.0);
    __stcs(out + i, v);
    __stcs(out + q + i, v);
    __stcs(out + 2 * q + (q - 1 - i), v);
    __stcs(out + 3 * q + (q - 1 - i), v);
}

}  // namespace fgq

using namespace fgq;

extern "C" int fgq_measure_dfma_peak(double* tflops) {
    clear_status();
    if (!tflops) return FGQ_EINVAL;
    HostPathLock lock;
    void* scratch;
    FGQ_TRY(scratch_device(1024, 4, &scratch));
    cudaDeviceProp prop;
    int dev;
    FGQ_CUDA_TRY(cudaGetDevice(&dev));
    FGQ_CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    const int blocks = prop.multiProcessorCount * 8;
    cudaEvent_t e0, e1;
    FGQ_CUDA_TRY(cudaEventCreate(&e0));
    FGQ_CUDA_TRY(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        FGQ_CUDA_TRY(cudaEventRecord(e0, 0));
        dfma_peak_kernel<<<blocks, 256>>>((double*)scratch, 1.0000001, 1e-9);
        FGQ_LAUNCH_CHECK();
        FGQ_CUDA_TRY(cudaEventRecord(e1, 0));
        FGQ_CUDA_TRY(cudaEventSynchronize(e1));
        float ms;
        FGQ_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double fma_count = (double)blocks * 256.0 * DFMA_ITERS * DFMA_CHAINS;
        const double tf = 2.0 * fma_count / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return FGQ_OK;
}

extern "C" int fgq_measure_store_peak(int64_t bytes, double* gbs) {
    clear_status();
    if (!gbs || bytes < (1 << 20)) return FGQ_EINVAL;
    HostPathLock lock;
    void* scratch;
    FGQ_TRY(scratch_device((size_t)bytes, 5, &scratch));
    cudaEvent_t e0, e1;
    FGQ_CUDA_TRY(cudaEventCreate(&e0));
    FGQ_CUDA_TRY(cudaEventCreate(&e1));
    const int64_t n16 = bytes / 16;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        FGQ_CUDA_TRY(cudaEventRecord(e0, 0));
        store_peak_kernel<<<(unsigned)((n16 / 4 + 255) / 256), 256>>>((double2*)scratch, n16);
        FGQ_LAUNCH_CHECK();
        FGQ_CUDA_TRY(cudaEventRecord(e1, 0));
        FGQ_CUDA_TRY(cudaEventSynchronize(e1));
        float ms;
        FGQ_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double g = (double)n16 * 16.0 / (ms * 1e-3) / 1e9;
        if (rep > 0 && g > best) best = g;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *gbs = best;
    return FGQ_OK;
}
