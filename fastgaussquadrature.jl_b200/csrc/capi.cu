// capi.cu — status plumbing, library-owned scratch, host-returning entry points and the
// rule-moments consumer of libfgq_cuda.so (see include/fgq.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "common.cuh"

namespace fgq {

static thread_local char t_err[512] = "";
static thread_local int t_warn = 0;
static std::atomic<int64_t> g_launches{0};
static std::mutex g_host_mutex;
static std::mutex g_res_mutex;  // lazily created per-device streams / events / scratch

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}
void clear_status() {
    t_err[0] = 0;
    t_warn = 0;
}
void add_warning(int bits) { t_warn |= bits; }
void count_launch(int64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

HostPathLock::HostPathLock() { g_host_mutex.lock(); }
HostPathLock::~HostPathLock() { g_host_mutex.unlock(); }

constexpr int MAX_DEV = 16, MAX_SLOT = 8;
static void* g_scratch[MAX_DEV][MAX_SLOT];
static size_t g_scratch_bytes[MAX_DEV][MAX_SLOT];
static cudaStream_t g_streams[MAX_DEV][2];
static cudaEvent_t g_events[MAX_DEV][4];

static int cur_dev(int* dev) {
    FGQ_CUDA_TRY(cudaGetDevice(dev));
    if (*dev < 0 || *dev >= MAX_DEV) {
        set_error("device ordinal %d out of range", *dev);
        return FGQ_EINVAL;
    }
    return FGQ_OK;
}

int scratch_device(size_t bytes, int slot, void** out) {
    int dev;
    FGQ_TRY(cur_dev(&dev));
    std::lock_guard<std::mutex> guard(g_res_mutex);
    if (g_scratch_bytes[dev][slot] < bytes) {
        if (g_scratch[dev][slot]) {
            FGQ_CUDA_TRY(cudaFree(g_scratch[dev][slot]));
            g_scratch[dev][slot] = nullptr;
            g_scratch_bytes[dev][slot] = 0;
        }
        FGQ_CUDA_TRY(cudaMalloc(&g_scratch[dev][slot], bytes));
        g_scratch_bytes[dev][slot] = bytes;
    }
    *out = g_scratch[dev][slot];
    return FGQ_OK;
}

static void* g_pinned[3];
static size_t g_pinned_bytes[3];
// library-owned pinned staging (grow-only), for destinations that are ordinary pageable memory
static int scratch_pinned(size_t bytes, int slot, void** out) {
    if (g_pinned_bytes[slot] < bytes) {
        if (g_pinned[slot]) {
            FGQ_CUDA_TRY(cudaFreeHost(g_pinned[slot]));
            g_pinned[slot] = nullptr;
            g_pinned_bytes[slot] = 0;
        }
        FGQ_CUDA_TRY(cudaHostAlloc(&g_pinned[slot], bytes, cudaHostAllocDefault));
        g_pinned_bytes[slot] = bytes;
    }
    *out = g_pinned[slot];
    return FGQ_OK;
}

static bool is_pinned_host(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

int scratch_stream(int which, cudaStream_t* out) {
    int dev;
    FGQ_TRY(cur_dev(&dev));
    std::lock_guard<std::mutex> guard(g_res_mutex);
    if (!g_streams[dev][which])
        FGQ_CUDA_TRY(cudaStreamCreateWithFlags(&g_streams[dev][which], cudaStreamNonBlocking));
    *out = g_streams[dev][which];
    return FGQ_OK;
}

int scratch_event(int which, cudaEvent_t* out) {
    int dev;
    FGQ_TRY(cur_dev(&dev));
    std::lock_guard<std::mutex> guard(g_res_mutex);
    if (!g_events[dev][which])
        FGQ_CUDA_TRY(cudaEventCreateWithFlags(&g_events[dev][which], cudaEventDisableTiming));
    *out = g_events[dev][which];
    return FGQ_OK;
}

// sum w, sum w x^2: block-level tree + one atomicAdd pair per block
__global__ void __launch_bounds__(256)
rule_moments_kernel(const double* __restrict__ x, const double* __restrict__ w, int64_t count,
                    double* __restrict__ sums) {
    double s0 = 0.0, s2 = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const double xi = __ldcs(x + i), wi = __ldcs(w + i);
        s0 += wi;
        s2 = fma(wi * xi, xi, s2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_down_sync(0xffffffffu, s0, o);
        s2 += __shfl_down_sync(0xffffffffu, s2, o);
    }
    __shared__ double sh0[8], sh2[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        sh0[wid] = s0;
        sh2[wid] = s2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int i = 0; i < 8; ++i) {
            a += sh0[i];
            b += sh2[i];
        }
        atomicAdd(sums, a);
        atomicAdd(sums + 1, b);
    }
}

// Mirror copy of the host-returning Legendre path: dst[cnt-1-e] = (negate ? -src[e] : src[e]) for e in
// [a, b).  A pure data movement over up to 8 GB per array; the destination is written once and not read
// again by this call, so the AVX2 form uses non-temporal stores (no write-allocate traffic competing
// with the PCIe DMA for host memory bandwidth).
#if defined(__x86_64__)
__attribute__((target("avx2"))) static void mirror_avx2(const double* src, double* dst, int64_t cnt, int64_t a,
                                                         int64_t b, bool negate) {
    // destination positions cnt-1-e descend as e ascends: walk e upwards, write blocks of 4 downwards
    int64_t e = a;
    // scalar until the 4-block [cnt-4-e, cnt-e) starts on a 32-byte boundary
    while (e < b && ((uintptr_t)(dst + (cnt - e)) & 31) != 0) {
        dst[cnt - 1 - e] = negate ? -src[e] : src[e];
        ++e;
    }
    const __m256d sign = negate ? _mm256_set1_pd(-0.0) : _mm256_setzero_pd();
    for (; e + 4 <= b; e += 4) {
        __m256d v = _mm256_loadu_pd(src + e);
        v = _mm256_permute4x64_pd(v, 0x1B);  // reverse the four lanes
        v = _mm256_xor_pd(v, sign);
        _mm256_stream_pd(dst + (cnt - 4 - e), v);
    }
    for (; e < b; ++e) dst[cnt - 1 - e] = negate ? -src[e] : src[e];
    _mm_sfence();
}
#endif

static void mirror_copy(const double* src, double* dst, int64_t cnt, int64_t a, int64_t b, bool negate) {
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) {
        mirror_avx2(src, dst, cnt, a, b, negate);
        return;
    }
#endif
    for (int64_t e = a; e < b; ++e) dst[cnt - 1 - e] = negate ? -src[e] : src[e];
}

}  // namespace fgq

using namespace fgq;

extern "C" int fgq_version(void) { return 100; }
extern "C" const char* fgq_last_error(void) { return t_err; }
extern "C" int fgq_warnings(void) { return t_warn; }
extern "C" int64_t fgq_launch_count(void) { return g_launches.load(); }

extern "C" int fgq_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int fgq_rule_moments_device(const double* x_dev, const double* w_dev, int64_t count,
                                       double* sums_dev, void* stream) {
    clear_status();
    if (count < 0 || !sums_dev || (count > 0 && (!x_dev || !w_dev))) {
        set_error("invalid moments request");
        return FGQ_EINVAL;
    }
    if (count == 0) return FGQ_OK;
    int64_t blocks = (count + 256 * 8 - 1) / (256 * 8);
    if (blocks > 148 * 8) blocks = 148 * 8;
    rule_moments_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x_dev, w_dev, count,
                                                                          sums_dev);
    FGQ_LAUNCH_CHECK();
    return FGQ_OK;
}

// ---- host-returning drop-ins ----------------------------------------------------------------
// Chunked: the kernel for chunk j+1 runs while chunk j drains over PCIe, so a 16 GB rule
// needs only 2 x CHUNK nodes of device scratch.

namespace fgq {

// Device -> host copy of a finished result (x and w, cnt entries each), ordered after the work already
// enqueued on `st`; returns when the data is in place.  A pinned destination is written by DMA directly.
// A pageable one (a fresh Julia Vector, np.empty) would be staged by the driver through its own small
// bounce buffer at 4-5 GB/s, one thread taking every first-touch page fault; instead the result is cut
// into 2 MiB pieces that are DMA'd into a library-owned pinned ring and copied out by host threads as
// they land (piece p by thread p mod T), which also spreads the page faults.
int drain_to_host(double* x, double* w, const double* xd, const double* wd, int64_t cnt, cudaStream_t st) {
    if (cnt <= 0) return FGQ_OK;
    const int64_t PIECE = 1LL << 18;  // entries per piece and array
    if (cnt < PIECE || (is_pinned_host(x) && is_pinned_host(w))) {
        FGQ_CUDA_TRY(cudaMemcpyAsync(x, xd, (size_t)cnt * 8, cudaMemcpyDeviceToHost, st));
        FGQ_CUDA_TRY(cudaMemcpyAsync(w, wd, (size_t)cnt * 8, cudaMemcpyDeviceToHost, st));
        FGQ_CUDA_TRY(cudaStreamSynchronize(st));
        return FGQ_OK;
    }
    constexpr int RING = 16;
    void* ring_p;
    FGQ_TRY(scratch_pinned((size_t)RING * 2 * PIECE * 8, 2, &ring_p));
    double* ring = (double*)ring_p;
    const int64_t pieces = (cnt + PIECE - 1) / PIECE;
    unsigned hw = std::thread::hardware_concurrency();
    int T = (int)(hw == 0 ? 4 : (hw > 16 ? 16 : hw));
    if (T > pieces) T = (int)pieces;
    int dev;
    FGQ_CUDA_TRY(cudaGetDevice(&dev));
    cudaEvent_t ev[RING];
    for (int i = 0; i < RING; ++i) FGQ_CUDA_TRY(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    std::mutex mu;
    std::condition_variable cv;
    int64_t published = 0;            // pieces whose copies and event have been enqueued
    std::vector<char> consumed((size_t)pieces, 0);
    bool failed = false;
    std::vector<std::thread> workers;
    for (int t = 0; t < T; ++t) {
        workers.emplace_back([&, t]() {
            cudaSetDevice(dev);
            for (int64_t p = t; p < pieces; p += T) {
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return published > p || failed; });
                    if (failed) return;
                }
                const bool ok = cudaEventSynchronize(ev[p % RING]) == cudaSuccess;
                if (ok) {
                    const int64_t lo = p * PIECE, len = (lo + PIECE < cnt ? PIECE : cnt - lo);
                    const double* sx = ring + (p % RING) * 2 * PIECE;
                    memcpy(x + lo, sx, (size_t)len * 8);
                    memcpy(w + lo, sx + PIECE, (size_t)len * 8);
                }
                std::lock_guard<std::mutex> lk(mu);
                consumed[(size_t)p] = 1;
                if (!ok) failed = true;
                cv.notify_all();
            }
        });
    }
    cudaError_t err = cudaSuccess;
    for (int64_t p = 0; p < pieces && err == cudaSuccess; ++p) {
        if (p >= RING) {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return consumed[(size_t)(p - RING)] != 0 || failed; });
            if (failed) break;
        }
        const int64_t lo = p * PIECE, len = (lo + PIECE < cnt ? PIECE : cnt - lo);
        double* sx = ring + (p % RING) * 2 * PIECE;
        err = cudaMemcpyAsync(sx, xd + lo, (size_t)len * 8, cudaMemcpyDeviceToHost, st);
        if (err == cudaSuccess) err = cudaMemcpyAsync(sx + PIECE, wd + lo, (size_t)len * 8, cudaMemcpyDeviceToHost, st);
        if (err == cudaSuccess) err = cudaEventRecord(ev[p % RING], st);
        std::lock_guard<std::mutex> lk(mu);
        if (err == cudaSuccess) published = p + 1; else failed = true;
        cv.notify_all();
    }
    for (auto& th : workers) th.join();
    cudaError_t e2 = cudaStreamSynchronize(st);
    for (int i = 0; i < RING; ++i) cudaEventDestroy(ev[i]);
    FGQ_CUDA_TRY(err);
    FGQ_CUDA_TRY(e2);
    if (failed) {
        set_error("device-to-host copy failed");
        return FGQ_ECUDA;
    }
    return FGQ_OK;
}

int host_pipeline(int64_t lo_all, int64_t hi_all, double* x, double* w, SliceFn fn, void* ctx) {
    const int64_t n = hi_all - lo_all;
    if (n <= 0) return FGQ_OK;
    if (!x || !w) {
        set_error("null output pointer");
        return FGQ_EINVAL;
    }
    HostPathLock lock;
    const int64_t CHUNK = 1LL << 24;  // nodes per chunk: 128 MiB of x + 128 MiB of w
    const int64_t chunk = n < CHUNK ? n : CHUNK;
    const int nbuf = n > CHUNK ? 2 : 1;
    double* buf[2] = {nullptr, nullptr};
    for (int b = 0; b < nbuf; ++b) {
        void* p;
        FGQ_TRY(scratch_device((size_t)chunk * 16, b, &p));
        buf[b] = (double*)p;
    }
    cudaStream_t s_compute, s_copy;
    FGQ_TRY(scratch_stream(0, &s_compute));
    FGQ_TRY(scratch_stream(1, &s_copy));
    cudaEvent_t computed[2], drained[2];
    for (int b = 0; b < 2; ++b) {
        FGQ_TRY(scratch_event(b, &computed[b]));
        FGQ_TRY(scratch_event(2 + b, &drained[b]));
    }
    int rc = FGQ_OK;
    if (is_pinned_host(x) && is_pinned_host(w)) {
        // pinned destination: DMA straight into it, everything enqueued up front
        int64_t j = 0;
        for (int64_t lo = 0; lo < n && rc == FGQ_OK; lo += chunk, ++j) {  // lo, hi relative to lo_all
            const int64_t hi = lo + chunk < n ? lo + chunk : n;
            const int b = (int)(j % nbuf);
            double* xd = buf[b];
            double* wd = buf[b] + chunk;
            if (j >= nbuf) FGQ_CUDA_TRY(cudaStreamWaitEvent(s_compute, drained[b], 0));
            rc = fn(ctx, lo_all + lo, lo_all + hi, xd, wd, s_compute);
            if (rc != FGQ_OK) break;
            FGQ_CUDA_TRY(cudaEventRecord(computed[b], s_compute));
            FGQ_CUDA_TRY(cudaStreamWaitEvent(s_copy, computed[b], 0));
            FGQ_CUDA_TRY(cudaMemcpyAsync(x + lo, xd, (size_t)(hi - lo) * 8, cudaMemcpyDeviceToHost, s_copy));
            FGQ_CUDA_TRY(cudaMemcpyAsync(w + lo, wd, (size_t)(hi - lo) * 8, cudaMemcpyDeviceToHost, s_copy));
            FGQ_CUDA_TRY(cudaEventRecord(drained[b], s_copy));
        }
    } else {
        // pageable destination: chunk j+1 is computed while chunk j is drained through the pinned ring
        // (drain_to_host returns when chunk j is in place, so its device buffer is free again)
        const int64_t nchunks = (n + chunk - 1) / chunk;
        auto launch = [&](int64_t j) -> int {
            const int64_t lo = j * chunk, hi = lo + chunk < n ? lo + chunk : n;
            const int b = (int)(j % nbuf);
            const int r = fn(ctx, lo_all + lo, lo_all + hi, buf[b], buf[b] + chunk, s_compute);
            if (r != FGQ_OK) return r;
            FGQ_CUDA_TRY(cudaEventRecord(computed[b], s_compute));
            return FGQ_OK;
        };
        rc = launch(0);
        for (int64_t j = 0; j < nchunks && rc == FGQ_OK; ++j) {
            const int64_t lo = j * chunk, hi = lo + chunk < n ? lo + chunk : n;
            const int b = (int)(j % nbuf);
            FGQ_CUDA_TRY(cudaStreamWaitEvent(s_copy, computed[b], 0));
            if (j + 1 < nchunks) rc = launch(j + 1);
            if (rc != FGQ_OK) break;
            rc = drain_to_host(x + lo, w + lo, buf[b], buf[b] + chunk, hi - lo, s_copy);
        }
    }
    cudaError_t e1 = cudaStreamSynchronize(s_compute);
    cudaError_t e2 = cudaStreamSynchronize(s_copy);
    if (rc != FGQ_OK) return rc;
    FGQ_CUDA_TRY(e1);
    FGQ_CUDA_TRY(e2);
    return FGQ_OK;
}

}  // namespace fgq

static int legendre_slice(void* ctx, int64_t lo, int64_t hi, double* xd, double* wd,
                          cudaStream_t st) {
    const int64_t n = *(const int64_t*)ctx;
    return fgq_gausslegendre_device(n, lo, hi, xd, wd, (void*)st);
}

// Host-returning mirror-pair shard: only the LEFT slice (half-indices [k_lo, k_hi)) crosses PCIe; the
// mirrored slice x[n-1-i] = -x[i], w[n-1-i] = w[i] is a data movement, done by host threads chunk by
// chunk while the next chunk is still in flight.  Halves the D2H bytes of a rule, which is what bounds
// the host-returning call (16 B/node at ~57 GB/s against 7 TB/s of generation).
extern "C" int fgq_gausslegendre_host_pair(int64_t n, int64_t k_lo, int64_t k_hi, double* xl, double* wl,
                                           double* xr, double* wr) {
    clear_status();
    if (n <= 60) {
        set_error("mirror-pair shards need the asymptotic path (n > 60), got n=%lld", (long long)n);
        return FGQ_EINVAL;
    }
    const int64_t m = (n + 1) >> 1, mr = n >> 1;
    if (k_lo < 1 || k_hi < k_lo || k_hi > m + 1) {
        set_error("invalid half-index shard [%lld, %lld) of n=%lld", (long long)k_lo, (long long)k_hi, (long long)n);
        return FGQ_EINVAL;
    }
    const int64_t cnt = k_hi - k_lo;
    if (cnt == 0) return FGQ_OK;
    if (!xl || !wl || !xr || !wr) {
        set_error("null output pointer");
        return FGQ_EINVAL;
    }
    HostPathLock lock;
    int dev;
    FGQ_CUDA_TRY(cudaGetDevice(&dev));
    // half-indices per pipeline stage; FGQ_HOST_CHUNK_LOG2 is a tuning aid (tools/host_call_timing.py)
    static const int chunk_log2 = []() {
        const char* e = getenv("FGQ_HOST_CHUNK_LOG2");
        const int v = e ? atoi(e) : 0;
        return (v >= 16 && v <= 28) ? v : 24;
    }();
    const int64_t CHUNK = 1LL << chunk_log2;
    const int64_t chunk = cnt < CHUNK ? cnt : CHUNK;
    const int64_t nchunks = (cnt + chunk - 1) / chunk;
    const int nbuf = nchunks > 1 ? 2 : 1;
    double* buf[2] = {nullptr, nullptr};
    for (int b = 0; b < nbuf; ++b) {
        void* p;
        FGQ_TRY(scratch_device((size_t)chunk * 16, b, &p));
        buf[b] = (double*)p;
    }
    cudaStream_t s_compute, s_copy;
    FGQ_TRY(scratch_stream(0, &s_compute));
    FGQ_TRY(scratch_stream(1, &s_copy));
    // Pageable destination (a Julia Vector, a numpy array): DMA into library-owned pinned staging and let
    // the host threads do the final copy as well as the mirroring — that also spreads the first-touch page
    // faults of a freshly allocated 16 GB result over all threads instead of one driver thread.
    const bool staged = !(is_pinned_host(xl) && is_pinned_host(wl));
    double* stage[2] = {nullptr, nullptr};
    if (staged) {
        for (int b = 0; b < nbuf; ++b) {
            void* p;
            FGQ_TRY(scratch_pinned((size_t)chunk * 16, b, &p));
            stage[b] = (double*)p;
        }
    }
    std::vector<cudaEvent_t> computed(nchunks), landed(nchunks);
    for (int64_t j = 0; j < nchunks; ++j) {
        FGQ_CUDA_TRY(cudaEventCreateWithFlags(&computed[j], cudaEventDisableTiming));
        FGQ_CUDA_TRY(cudaEventCreateWithFlags(&landed[j], cudaEventDisableTiming));
    }
    int rc = FGQ_OK;
    auto enqueue = [&](int64_t j) -> int {
        const int64_t lo = j * chunk, hi = lo + chunk < cnt ? lo + chunk : cnt;  // relative to k_lo
        const int b = (int)(j % nbuf);
        double* xd = buf[b];
        double* wd = buf[b] + chunk;
        if (j >= nbuf) cudaStreamWaitEvent(s_compute, landed[j - nbuf], 0);
        // left slice = output indices [k_lo-1+lo, k_lo-1+hi)
        const int r = fgq_gausslegendre_device(n, k_lo - 1 + lo, k_lo - 1 + hi, xd, wd, (void*)s_compute);
        if (r != FGQ_OK) return r;
        cudaEventRecord(computed[j], s_compute);
        cudaStreamWaitEvent(s_copy, computed[j], 0);
        double* dx = staged ? stage[b] : xl + lo;
        double* dw = staged ? stage[b] + chunk : wl + lo;
        cudaMemcpyAsync(dx, xd, (size_t)(hi - lo) * 8, cudaMemcpyDeviceToHost, s_copy);
        cudaMemcpyAsync(dw, wd, (size_t)(hi - lo) * 8, cudaMemcpyDeviceToHost, s_copy);
        cudaEventRecord(landed[j], s_copy);
        return FGQ_OK;
    };
    // element e of the left slice (half-index k = k_lo + e) mirrors to right-slice position cnt-1-e;
    // the centre of an odd rule (k = m) has no mirror image
    const int64_t mirror_cnt = (k_hi - 1 > mr) ? cnt - 1 : cnt;
    unsigned hw = std::thread::hardware_concurrency();
    const int T = (int)(hw == 0 ? 4 : (hw > 16 ? 16 : hw));
    auto host_part = [&](int64_t j, int t) {  // thread t's share of chunk j
        const int64_t lo = j * chunk, hi0 = lo + chunk < cnt ? lo + chunk : cnt;
        const int b = (int)(j % nbuf);
        const double* sx = staged ? stage[b] - lo : xl;          // source indexed by e
        const double* sw = staged ? stage[b] + chunk - lo : wl;
        const int64_t len0 = hi0 - lo;
        const int64_t a0 = lo + len0 * t / T, b0 = lo + len0 * (t + 1) / T;
        if (staged) {
            memcpy(xl + a0, sx + a0, (size_t)(b0 - a0) * 8);
            memcpy(wl + a0, sw + a0, (size_t)(b0 - a0) * 8);
        }
        const int64_t b1 = b0 < mirror_cnt ? b0 : mirror_cnt;
        mirror_copy(sx, xr, cnt, a0, b1, true);
        mirror_copy(sw, wr, cnt, a0, b1, false);
    };
    if (!staged) {
        // pinned destination: everything is enqueued up front, the workers only follow the copy events
        for (int64_t j = 0; j < nchunks && rc == FGQ_OK; ++j) rc = enqueue(j);
        if (rc == FGQ_OK) {
            std::vector<std::thread> workers;
            std::atomic<int> failed{0};
            for (int t = 0; t < T; ++t) {
                workers.emplace_back([&, t]() {
                    cudaSetDevice(dev);
                    for (int64_t j = 0; j < nchunks; ++j) {
                        if (cudaEventSynchronize(landed[j]) != cudaSuccess) {
                            failed = 1;
                            return;
                        }
                        host_part(j, t);
                    }
                });
            }
            for (auto& th : workers) th.join();
            if (failed) {
                set_error("device-to-host copy failed");
                rc = FGQ_ECUDA;
            }
        }
    } else {
        // staged: a staging buffer may only be refilled after the host threads have emptied it
        for (int64_t j = 0; j < nchunks && j < nbuf && rc == FGQ_OK; ++j) rc = enqueue(j);
        for (int64_t j = 0; j < nchunks && rc == FGQ_OK; ++j) {
            if (cudaEventSynchronize(landed[j]) != cudaSuccess) {
                set_error("device-to-host copy failed");
                rc = FGQ_ECUDA;
                break;
            }
            std::vector<std::thread> workers;
            for (int t = 0; t < T; ++t) workers.emplace_back([&, t]() { host_part(j, t); });
            for (auto& th : workers) th.join();
            if (j + nbuf < nchunks) rc = enqueue(j + nbuf);
        }
    }
    cudaError_t e1 = cudaStreamSynchronize(s_compute);
    cudaError_t e2 = cudaStreamSynchronize(s_copy);
    for (int64_t j = 0; j < nchunks; ++j) {
        cudaEventDestroy(computed[j]);
        cudaEventDestroy(landed[j]);
    }
    if (rc != FGQ_OK) return rc;
    FGQ_CUDA_TRY(e1);
    FGQ_CUDA_TRY(e2);
    return FGQ_OK;
}

extern "C" int fgq_gausslegendre_host(int64_t n, double* x, double* w) {
    clear_status();
    if (n < 0) {
        set_error("DomainError(%lld): Input n must be a non-negative integer", (long long)n);
        return FGQ_EDOMAIN;
    }
    if (n >= (1LL << 21) && x && w) {
        // large rules: ship the left half only and mirror it on the host (see fgq_gausslegendre_host_pair)
        const int64_t m = (n + 1) >> 1, mr = n >> 1;
        // right slice = output indices [n - m, n): its element 0 is the centre itself for odd n (untouched)
        return fgq_gausslegendre_host_pair(n, 1, m + 1, x, w, x + (n - m), w + (n - m));
        (void)mr;
    }
    return host_pipeline(0, n, x, w, legendre_slice, &n);
}

extern "C" int fgq_gausslegendre_host_slice(int64_t n, int64_t lo, int64_t hi, double* x, double* w) {
    clear_status();
    if (n < 0) {
        set_error("DomainError(%lld): Input n must be a non-negative integer", (long long)n);
        return FGQ_EDOMAIN;
    }
    if (lo < 0 || hi < lo || hi > n) {
        set_error("invalid slice [%lld, %lld) of n=%lld", (long long)lo, (long long)hi, (long long)n);
        return FGQ_EINVAL;
    }
    return host_pipeline(lo, hi, x, w, legendre_slice, &n);
}

extern "C" int fgq_gausslegendre_batch_host(int64_t nrules, const int32_t* n_i,
                                            const int64_t* offsets, double* x, double* w) {
    clear_status();
    if (nrules < 0 || (nrules > 0 && (!n_i || !offsets))) {
        set_error("invalid batch description");
        return FGQ_EINVAL;
    }
    if (nrules == 0) return FGQ_OK;
    int64_t total = 0;
    for (int64_t i = 0; i < nrules; ++i) {
        if (n_i[i] < 0) {
            set_error("DomainError(%d): Input n must be a non-negative integer", n_i[i]);
            return FGQ_EDOMAIN;
        }
        if (n_i[i] > 60) {
            set_error("batched rules are limited to n <= 60 (rule %lld has n=%d)", (long long)i, n_i[i]);
            return FGQ_EINVAL;
        }
        if (offsets[i] < 0) {
            set_error("negative offset");
            return FGQ_EINVAL;
        }
        const int64_t end = offsets[i] + n_i[i];
        if (end > total) total = end;
    }
    if (total == 0) return FGQ_OK;
    if (!x || !w) {
        set_error("null output pointer");
        return FGQ_EINVAL;
    }
    HostPathLock lock;
    void *p_out, *p_n, *p_off;
    FGQ_TRY(scratch_device((size_t)total * 16, 0, &p_out));
    FGQ_TRY(scratch_device((size_t)nrules * sizeof(int32_t), 2, &p_n));
    FGQ_TRY(scratch_device((size_t)nrules * sizeof(int64_t), 3, &p_off));
    cudaStream_t st;
    FGQ_TRY(scratch_stream(0, &st));
    double* xd = (double*)p_out;
    double* wd = xd + total;
    FGQ_CUDA_TRY(cudaMemcpyAsync(p_n, n_i, (size_t)nrules * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    FGQ_CUDA_TRY(cudaMemcpyAsync(p_off, offsets, (size_t)nrules * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    FGQ_TRY(fgq_gausslegendre_batch_device(nrules, (const int32_t*)p_n, (const int64_t*)p_off, xd, wd, (void*)st));
    FGQ_TRY(drain_to_host(x, w, xd, wd, total, st));
    return FGQ_OK;
}
