// capi.cu — status plumbing, library-owned scratch, host-returning entry points and the
// rule-moments consumer of libfgq_cuda.so (see include/fgq.h).
#include <stdarg.h>
#include <stdlib.h>
#if defined(__linux__)
#include <sys/mman.h>
#endif
#include <string.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "common.cuh"
#include "delta_codec.h"

namespace fgq {

static thread_local char t_err[512] = "";
static thread_local int t_warn = 0;
static std::atomic<int64_t> g_launches{0};
static std::mutex g_host_mutex[16];  // host-returning calls serialise PER DEVICE (fgq_gausslegendre_host_multi runs one per device)
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

static int host_dev_slot() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        dev = 0;
    }
    return (dev >= 0 && dev < 16) ? dev : 0;
}
HostPathLock::HostPathLock() : slot(host_dev_slot()) { g_host_mutex[slot].lock(); }
HostPathLock::~HostPathLock() { g_host_mutex[slot].unlock(); }

// Host threads a host-returning call may use; 0 = the pool's default.  fgq_gausslegendre_host_multi sets it on the
// per-device threads it starts so that the devices share the host's cores instead of each claiming all of them.
static thread_local int t_host_thread_budget = 0;
void set_host_thread_budget(int threads) { t_host_thread_budget = threads; }

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

static void* g_pinned[MAX_DEV][3];
static size_t g_pinned_bytes[MAX_DEV][3];
// library-owned pinned staging (grow-only, per device: the caller holds that device's HostPathLock), for
// destinations that are ordinary pageable memory and for the coded streams
static int scratch_pinned(size_t bytes, int slot, void** out) {
    int dev;
    FGQ_TRY(cur_dev(&dev));
    if (g_pinned_bytes[dev][slot] < bytes) {
        if (g_pinned[dev][slot]) {
            FGQ_CUDA_TRY(cudaFreeHost(g_pinned[dev][slot]));
            g_pinned[dev][slot] = nullptr;
            g_pinned_bytes[dev][slot] = 0;
        }
        FGQ_CUDA_TRY(cudaHostAlloc(&g_pinned[dev][slot], bytes, cudaHostAllocPortable));
        g_pinned_bytes[dev][slot] = bytes;
    }
    *out = g_pinned[dev][slot];
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

static std::vector<cudaEvent_t> g_event_pool[MAX_DEV];
static std::vector<cudaStream_t> g_stream_pool[MAX_DEV];
static std::vector<cudaStream_t> g_stream_pool_hp[MAX_DEV];  // created at the greatest stream priority

int ScopedEvents::acquire(int count) {
    if (count < 0 || count > 64) return FGQ_EINVAL;
    FGQ_TRY(cur_dev(&dev));
    std::lock_guard<std::mutex> guard(g_res_mutex);
    while (n < count) {
        if (!g_event_pool[dev].empty()) {
            ev[n] = g_event_pool[dev].back();
            g_event_pool[dev].pop_back();
        } else {
            FGQ_CUDA_TRY(cudaEventCreateWithFlags(&ev[n], cudaEventDisableTiming));
        }
        ++n;
    }
    return FGQ_OK;
}
ScopedEvents::~ScopedEvents() {
    if (n == 0 || dev < 0) return;
    std::lock_guard<std::mutex> guard(g_res_mutex);
    for (int i = 0; i < n; ++i) g_event_pool[dev].push_back(ev[i]);
}
int ScopedStream::acquire(bool high_priority) {
    FGQ_TRY(cur_dev(&dev));
    high = high_priority;
    std::lock_guard<std::mutex> guard(g_res_mutex);
    std::vector<cudaStream_t>& pool = high ? g_stream_pool_hp[dev] : g_stream_pool[dev];
    if (!pool.empty()) {
        st = pool.back();
        pool.pop_back();
        return FGQ_OK;
    }
    if (high) {
        int least = 0, greatest = 0;
        FGQ_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        FGQ_CUDA_TRY(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, greatest));
    } else {
        FGQ_CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    }
    return FGQ_OK;
}
ScopedStream::~ScopedStream() {
    if (!st || dev < 0) return;
    std::lock_guard<std::mutex> guard(g_res_mutex);
    (high ? g_stream_pool_hp[dev] : g_stream_pool[dev]).push_back(st);
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

extern "C" int fgq_host_alloc(int64_t bytes, void** out) {
    clear_status();
    if (!out || bytes < 0) {
        set_error("fgq_host_alloc: invalid arguments");
        return FGQ_EINVAL;
    }
    *out = nullptr;
    if (bytes == 0) return FGQ_OK;
    FGQ_CUDA_TRY(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocPortable));
    return FGQ_OK;
}

extern "C" int fgq_host_free(void* p) {
    clear_status();
    if (!p) return FGQ_OK;
    FGQ_CUDA_TRY(cudaFreeHost(p));
    return FGQ_OK;
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

// Mirror copy of the host-returning Legendre path: dst[cnt-1-e] = (negate ? -src[e] : src[e]) for e in
// [a, b).  A pure data movement over up to 8 GB per array; the destination is written once and not read
// again by this call, so the AVX2 form uses non-temporal stores (no write-allocate traffic competing
// with the PCIe DMA for host memory bandwidth).
#if defined(__x86_64__)
__attribute__((target("avx2"))) static void mirror_avx2(const double* src, double* dst, int64_t cnt, int64_t a,
                                                         int64_t b, bool negate, bool fence) {
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
    if (fence) _mm_sfence();
}
#endif

// Ring slot -> pageable destination.  Like the mirror copy, the destination is written once and not read
// again by this call: non-temporal stores avoid the read-for-ownership of every destination line (glibc's
// memcpy only switches to them for copies far larger than a 2 MiB piece), i.e. a third of the host-memory
// traffic of this step, on a path that is host-memory bound.
#if defined(__x86_64__)
__attribute__((target("avx2"))) static void copy_stream_avx2(double* dst, const double* src, int64_t n, bool fence) {
    int64_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 31) != 0) {
        dst[i] = src[i];
        ++i;
    }
    for (; i + 4 <= n; i += 4) _mm256_stream_pd(dst + i, _mm256_loadu_pd(src + i));
    for (; i < n; ++i) dst[i] = src[i];
    if (fence) _mm_sfence();
}
#endif

// fence = false: the caller issues one store fence for a whole run of calls (stream_fence), instead of draining
// the write-combining buffers after every 8 KB
static void stream_fence() {
#if defined(__x86_64__)
    _mm_sfence();
#endif
}

static void copy_stream(double* dst, const double* src, int64_t n, bool fence = true) {
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2 && n >= 256) {
        copy_stream_avx2(dst, src, n, fence);
        return;
    }
#endif
    memcpy(dst, src, (size_t)n * 8);
}

static void mirror_copy(const double* src, double* dst, int64_t cnt, int64_t a, int64_t b, bool negate,
                        bool fence = true) {
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) {
        mirror_avx2(src, dst, cnt, a, b, negate, fence);
        return;
    }
#endif
    for (int64_t e = a; e < b; ++e) dst[cnt - 1 - e] = negate ? -src[e] : src[e];
}

// Host worker threads of the host-returning entry points: created once, parked on a condition variable
// between calls (spawning 16 threads per call costs ~0.5 ms, more than a whole gausshermite(1e6)).  Only
// used under HostPathLock, i.e. by one call at a time.  Deliberately leaked: parked threads must not
// outlive the synchronisation objects they wait on during static destruction.
class WorkerPool {
  public:
    static WorkerPool& get() {  // one pool per device slot: calls on different devices run concurrently
        static WorkerPool* pools[16] = {};
        static std::mutex mu;
        const int slot = host_dev_slot();
        std::lock_guard<std::mutex> lk(mu);
        if (!pools[slot]) pools[slot] = new WorkerPool();
        return *pools[slot];
    }
    int max_threads() const {
        if (t_host_thread_budget > 0) return t_host_thread_budget < 64 ? t_host_thread_budget : 64;
        return (int)threads_;
    }
    // runs job(t) for t in [0, T): t = 0 on the calling thread, the others on pool threads; returns when
    // all of them have finished
    void run(int T, const std::function<void(int)>& job) {
        if (T > 1) {
            std::unique_lock<std::mutex> lk(mu_);
            while ((int)spawned_ < T - 1) {
                const int id = (int)spawned_++;
                std::thread([this, id]() { loop(id); }).detach();
            }
            job_ = &job;
            active_ = T - 1;
            remaining_ = T - 1;
            ++generation_;
            cv_start_.notify_all();
        }
        job(0);
        if (T > 1) {
            std::unique_lock<std::mutex> lk(mu_);
            cv_done_.wait(lk, [&] { return remaining_ == 0; });
            job_ = nullptr;
        }
    }

  private:
    WorkerPool() {
        unsigned hw = std::thread::hardware_concurrency();
        threads_ = hw == 0 ? 4 : (hw > 16 ? 16 : hw);
        // several processes on one host (one per GPU) share its cores: FGQ_HOST_THREADS caps this process
        if (const char* e = getenv("FGQ_HOST_THREADS")) {
            const int v = atoi(e);
            if (v >= 1 && v <= 64) threads_ = (unsigned)v;
        }
    }
    void loop(int id) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int)>* job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_start_.wait(lk, [&] { return generation_ != seen; });
                seen = generation_;
                if (id >= active_) continue;
                job = job_;
            }
            (*job)(id + 1);
            std::lock_guard<std::mutex> lk(mu_);
            if (--remaining_ == 0) cv_done_.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_start_, cv_done_;
    const std::function<void(int)>* job_ = nullptr;
    uint64_t generation_ = 0;
    int active_ = 0, remaining_ = 0;
    unsigned threads_ = 4, spawned_ = 0;
};

// A freshly allocated multi-GB result (Julia `Vector{Float64}(undef, n)`, np.empty) is untouched
// anonymous memory: with 4 KiB pages its first touch costs ~4e6 page faults per 16 GB.  Ask for
// transparent huge pages on the interior of the range (a hint; a no-op where THP is off or the memory
// is already populated).
static void hint_huge_pages(void* p, size_t bytes) {
#if defined(__linux__) && defined(MADV_HUGEPAGE)
    if (bytes < (64u << 20)) return;
    const uintptr_t HP = 2u << 20;
    const uintptr_t lo = ((uintptr_t)p + HP - 1) & ~(HP - 1), hi = ((uintptr_t)p + bytes) & ~(HP - 1);
    if (hi > lo) madvise((void*)lo, hi - lo, MADV_HUGEPAGE);
#else
    (void)p;
    (void)bytes;
#endif
}

static std::atomic<int64_t> g_drain_pieces{0}, g_drain_dma{0}, g_drain_landed{0};
static std::atomic<int> g_drain_share_milli{0};

// Device -> host copy of a result (x and w, cnt entries each) on stream `st`; returns when the data is in
// place.  The result is cut into pieces; before the copies of piece [lo, hi) are enqueued `src` names the
// device addresses holding it (and may enqueue the kernels that produce it: see ChunkedProducer), so the
// whole transfer is one uninterrupted queue of DMA copies.
//  * Pinned destination: DMA straight into it.
//  * Pageable destination (a fresh Julia Vector, np.empty): the driver would stage it through its own
//    small bounce buffer at 4-5 GB/s, one thread taking every first-touch page fault; instead 2 MiB
//    pieces are DMA'd into a library-owned pinned ring and copied out by host threads as they land
//    (piece p by thread p mod T), which also spreads the page faults.
//  * Optional mirror (the Legendre +-x symmetry): entry i of this call is element e = m.e0 + i of a left
//    slice of m.cnt half-indices; the threads also write m.xr[m.cnt-1-e] = -x, m.wr[m.cnt-1-e] = w for
//    e < m.mirror_cnt.
int drain_to_host(double* x, double* w, int64_t cnt, cudaStream_t st, const PieceSource& src, const MirrorSpec* m) {
    if (cnt <= 0) return FGQ_OK;
    if (!m && cnt < DRAIN_PIECE) {  // small result: one plain copy, no threads
        PieceDev pd{};
        FGQ_TRY(src(0, cnt, &pd));
        FGQ_CUDA_TRY(cudaMemcpyAsync(x, pd.x, (size_t)cnt * 8, cudaMemcpyDeviceToHost, st));
        FGQ_CUDA_TRY(cudaMemcpyAsync(w, pd.w, (size_t)cnt * 8, cudaMemcpyDeviceToHost, st));
        FGQ_CUDA_TRY(cudaStreamSynchronize(st));
        return FGQ_OK;
    }
    const bool direct = is_pinned_host(x) && is_pinned_host(w);
    // the mirror image of a piece may travel by DMA instead (see below) when everything is pinned
    const bool dma_mirror_ok = m && direct && is_pinned_host(m->xr) && is_pinned_host(m->wr);
    const int64_t PIECE = direct ? DRAIN_PIECE_PINNED : DRAIN_PIECE;  // entries per piece and array
    // pieces in flight: 16 ring slots for a pageable destination; 64 when the DMA goes straight to pinned
    // memory (no ring), so that PCIe can run well ahead of the host threads and a thread can tell a piece
    // that has been waiting for it from one it has to wait for
    constexpr int MAX_RING = 64;
    const int RING = direct ? MAX_RING : 16;
    double* ring = nullptr;
    if (!direct) {
        void* ring_p;
        FGQ_TRY(scratch_pinned((size_t)RING * 2 * PIECE * 8, 2, &ring_p));
        ring = (double*)ring_p;
    }
    const int64_t pieces = (cnt + PIECE - 1) / PIECE;
    const bool threaded = m || !direct;  // a pinned destination without mirror needs no host work
    WorkerPool& pool = WorkerPool::get();
    int T = threaded ? pool.max_threads() : 0;
    if (T > pieces) T = (int)pieces;
    int dev;
    FGQ_CUDA_TRY(cudaGetDevice(&dev));
    ScopedEvents evs;
    if (threaded) FGQ_TRY(evs.acquire(RING));
    cudaEvent_t* ev = evs.ev;
    std::mutex mu;
    std::condition_variable cv;
    int64_t published = 0;            // pieces whose copies and event have been enqueued
    int64_t claimed = 0;              // pieces taken by a host thread (in order)
    std::vector<char> consumed((size_t)pieces, 0);
    std::vector<char> by_dma((size_t)pieces, 0);
    bool failed = false;
    int rc = FGQ_OK;
    cudaError_t err = cudaSuccess;
    // Who writes the mirror image of a piece?  The host threads (8 B/node over PCIe, but 24 B/node of host
    // memory traffic) or a second pair of DMA copies of the image the kernel left in device memory
    // (16 B/node over PCIe, 16 B/node of host traffic, no host work).  One GPU on its own is PCIe-bound and
    // wants the former; several processes sharing one host's memory and cores want more of the latter.
    // `dma_share` is the fraction of pieces sent the second way, steered by what the host threads see:
    // a piece that had already landed when a thread picked it up means the host side is behind (raise the
    // share); a piece the thread had to wait for means PCIe is behind (lower it).
    // The share carries over from call to call (the caller holds HostPathLock): the feeder runs up to RING
    // pieces ahead of the host threads, so a short drain is over before its own observations can act.
    static double s_dma_share_dev[16] = {};
    double& s_dma_share = s_dma_share_dev[host_dev_slot()];
    double dma_share = s_dma_share, dma_acc = 0.0;
    bool adapt = true;
    if (const char* e = getenv("FGQ_HOST_MIRROR_DMA_SHARE")) {  // tests / tuning: a fixed share in [0, 1]
        const double v = atof(e);
        if (v >= 0.0 && v <= 1.0) {
            dma_share = v;
            adapt = false;
        }
    }
    // the calling thread feeds the copy queue, the pool threads drain it
    auto feeder = [&]() {
        for (int64_t p = 0; p < pieces && err == cudaSuccess && rc == FGQ_OK; ++p) {
            bool use_dma = false;
            if (threaded) {
                std::unique_lock<std::mutex> lk(mu);
                if (p >= RING) cv.wait(lk, [&] { return consumed[(size_t)(p - RING)] != 0 || failed; });
                if (failed) break;
                if (dma_mirror_ok) {
                    dma_acc += dma_share;
                    if (dma_acc >= 1.0) {
                        dma_acc -= 1.0;
                        use_dma = true;
                    }
                }
            }
            const int64_t lo = p * PIECE, len = (lo + PIECE < cnt ? PIECE : cnt - lo);
            PieceDev pd{};
            rc = src(lo, lo + len, &pd);
            if (rc == FGQ_OK) {
                double* dx = direct ? x + lo : ring + (p % RING) * 2 * PIECE;
                double* dw = direct ? w + lo : dx + PIECE;
                err = cudaMemcpyAsync(dx, pd.x, (size_t)len * 8, cudaMemcpyDeviceToHost, st);
                if (err == cudaSuccess) err = cudaMemcpyAsync(dw, pd.w, (size_t)len * 8, cudaMemcpyDeviceToHost, st);
                use_dma = use_dma && pd.xr && pd.wr;
                if (use_dma && err == cudaSuccess) {
                    // elements [e0, e1) of the left slice <-> right positions [cnt - e1, cnt - e0), ascending;
                    // pd.xr points at the image of position cnt - (e0 + len)
                    const int64_t e0 = m->e0 + lo;
                    const int64_t e1 = e0 + len < m->mirror_cnt ? e0 + len : m->mirror_cnt;
                    const int64_t skip = e0 + len - e1;  // the centre of an odd rule has no image
                    if (e1 > e0) {
                        err = cudaMemcpyAsync(m->xr + (m->cnt - e1), pd.xr + skip, (size_t)(e1 - e0) * 8,
                                              cudaMemcpyDeviceToHost, st);
                        if (err == cudaSuccess)
                            err = cudaMemcpyAsync(m->wr + (m->cnt - e1), pd.wr + skip, (size_t)(e1 - e0) * 8,
                                                  cudaMemcpyDeviceToHost, st);
                    }
                }
                if (err == cudaSuccess && threaded) err = cudaEventRecord(ev[p % RING], st);
            }
            if (threaded) {
                std::lock_guard<std::mutex> lk(mu);
                by_dma[(size_t)p] = use_dma;
                if (err == cudaSuccess && rc == FGQ_OK) published = p + 1; else failed = true;
                cv.notify_all();
            }
        }
    };
    auto drainer = [&](int) {
        cudaSetDevice(dev);
        for (;;) {
            int64_t p;
            bool dma;
            {
                std::unique_lock<std::mutex> lk(mu);
                if (claimed >= pieces || failed) return;
                p = claimed++;
                cv.wait(lk, [&] { return published > p || failed; });
                if (failed) return;
                dma = by_dma[(size_t)p] != 0;
            }
            const bool landed = cudaEventQuery(ev[p % RING]) == cudaSuccess;
            const bool ok = landed || cudaEventSynchronize(ev[p % RING]) == cudaSuccess;
            if (ok && !dma) {
                const int64_t lo = p * PIECE, len = (lo + PIECE < cnt ? PIECE : cnt - lo);
                const double* sx = direct ? x + lo : ring + (p % RING) * 2 * PIECE;
                const double* sw = direct ? w + lo : sx + PIECE;
                if (!direct) {
                    copy_stream(x + lo, sx, len);
                    copy_stream(w + lo, sw, len);
                }
                if (m) {
                    const int64_t e0 = m->e0 + lo;  // element index of sx[0]
                    const int64_t e1 = e0 + len < m->mirror_cnt ? e0 + len : m->mirror_cnt;
                    if (e1 > e0) {
                        mirror_copy(sx - e0, m->xr, m->cnt, e0, e1, true);
                        mirror_copy(sw - e0, m->wr, m->cnt, e0, e1, false);
                    }
                }
            }
            std::lock_guard<std::mutex> lk(mu);
            consumed[(size_t)p] = 1;
            if (!ok) failed = true;
            g_drain_pieces++;
            g_drain_dma += dma;
            g_drain_landed += landed;
            if (adapt) {
                dma_share += landed ? 1.0 / 16 : -1.0 / 16;
                dma_share = dma_share < 0.0 ? 0.0 : (dma_share > 1.0 ? 1.0 : dma_share);
            }
            cv.notify_all();
        }
    };
    pool.run(T + 1, [&](int t) {
        if (t == 0) feeder(); else drainer(t - 1);
    });
    cudaError_t e2 = cudaStreamSynchronize(st);
    g_drain_share_milli = (int)(dma_share * 1000);
    if (adapt && dma_mirror_ok) s_dma_share = dma_share;
    if (rc != FGQ_OK) return rc;
    FGQ_CUDA_TRY(err);
    FGQ_CUDA_TRY(e2);
    if (failed) {
        set_error("device-to-host copy failed");
        return FGQ_ECUDA;
    }
    return FGQ_OK;
}

// A finished result in device memory.
int drain_to_host(double* x, double* w, const double* xd, const double* wd, int64_t cnt, cudaStream_t st) {
    if (cnt <= 0) return FGQ_OK;
    if (cnt < DRAIN_PIECE) {
        FGQ_CUDA_TRY(cudaMemcpyAsync(x, xd, (size_t)cnt * 8, cudaMemcpyDeviceToHost, st));
        FGQ_CUDA_TRY(cudaMemcpyAsync(w, wd, (size_t)cnt * 8, cudaMemcpyDeviceToHost, st));
        FGQ_CUDA_TRY(cudaStreamSynchronize(st));
        return FGQ_OK;
    }
    hint_huge_pages(x, (size_t)cnt * 8);
    hint_huge_pages(w, (size_t)cnt * 8);
    return drain_to_host(x, w, cnt, st, [&](int64_t lo, int64_t, PieceDev* pd) {
        pd->x = xd + lo;
        pd->w = wd + lo;
        return (int)FGQ_OK;
    }, nullptr);
}

// A result produced chunk by chunk into two library-owned device buffers: produce(lo, hi, x_dev, w_dev, st)
// enqueues the kernels for entries [lo, hi).  As a PieceSource it launches chunk j+1 when the first piece of
// chunk j is about to be copied, with event dependencies only (a buffer is refilled after its copies have
// run), so kernels, DMA and host threads all overlap and nothing blocks between chunks.
struct ChunkedProducer {
    int64_t cnt, chunk, nchunks;
    int nbuf;
    double* buf[2];
    cudaStream_t s_compute, s_copy;
    cudaEvent_t computed[2], drained[2];
    // produce(lo, hi, x_dev, w_dev, xr_dev, wr_dev, st): xr_dev/wr_dev (null unless `mirrored`) receive the
    // +-x image of the chunk in ascending order, xr_dev[(hi-lo)-1-i] = -x_dev[i]
    std::function<int(int64_t, int64_t, double*, double*, double*, double*, cudaStream_t)> produce;
    int64_t launched = 0;
    bool mirrored = false;

    int init(int64_t cnt_, int64_t max_chunk, bool mirrored_ = false) {
        mirrored = mirrored_;
        cnt = cnt_;
        chunk = cnt < max_chunk ? cnt : max_chunk;
        nchunks = (cnt + chunk - 1) / chunk;
        nbuf = nchunks > 1 ? 2 : 1;
        buf[0] = buf[1] = nullptr;
        for (int b = 0; b < nbuf; ++b) {
            void* p;
            FGQ_TRY(scratch_device((size_t)chunk * (mirrored ? 32 : 16), b, &p));
            buf[b] = (double*)p;
        }
        FGQ_TRY(scratch_stream(0, &s_compute));
        FGQ_TRY(scratch_stream(1, &s_copy));
        for (int b = 0; b < 2; ++b) {
            FGQ_TRY(scratch_event(b, &computed[b]));
            FGQ_TRY(scratch_event(2 + b, &drained[b]));
        }
        return FGQ_OK;
    }
    int launch(int64_t j) {
        const int64_t lo = j * chunk, hi = lo + chunk < cnt ? lo + chunk : cnt;
        const int b = (int)(j % nbuf);
        if (j >= nbuf) FGQ_CUDA_TRY(cudaStreamWaitEvent(s_compute, drained[b], 0));
        FGQ_TRY(produce(lo, hi, buf[b], buf[b] + chunk, mirrored ? buf[b] + 2 * chunk : nullptr,
                        mirrored ? buf[b] + 3 * chunk : nullptr, s_compute));
        FGQ_CUDA_TRY(cudaEventRecord(computed[b], s_compute));
        return FGQ_OK;
    }
    int piece(int64_t lo, int64_t hi, PieceDev* pd) {
        const int64_t j = lo / chunk;
        const int b = (int)(j % nbuf);
        if (lo == j * chunk) {  // first piece of chunk j: every copy of chunk j-1 has been enqueued
            if (j >= 1) FGQ_CUDA_TRY(cudaEventRecord(drained[(j - 1) % nbuf], s_copy));
            while (launched <= j) FGQ_TRY(launch(launched++));
            FGQ_CUDA_TRY(cudaStreamWaitEvent(s_copy, computed[b], 0));
            if (j + 1 < nchunks && launched <= j + 1) FGQ_TRY(launch(launched++));
        }
        pd->x = buf[b] + (lo - j * chunk);
        pd->w = buf[b] + chunk + (lo - j * chunk);
        if (mirrored) {  // image of [lo, hi) within the chunk's image of length len_j
            const int64_t len_j = (j * chunk + chunk < cnt ? chunk : cnt - j * chunk);
            pd->xr = buf[b] + 2 * chunk + (len_j - (hi - j * chunk));
            pd->wr = buf[b] + 3 * chunk + (len_j - (hi - j * chunk));
        }
        return FGQ_OK;
    }
    int finish(int rc) {
        cudaError_t e1 = cudaStreamSynchronize(s_compute);
        cudaError_t e2 = cudaStreamSynchronize(s_copy);
        if (rc != FGQ_OK) return rc;
        FGQ_CUDA_TRY(e1);
        FGQ_CUDA_TRY(e2);
        return FGQ_OK;
    }
};

int host_pipeline(int64_t lo_all, int64_t hi_all, double* x, double* w, SliceFn fn, void* ctx) {
    const int64_t n = hi_all - lo_all;
    if (n <= 0) return FGQ_OK;
    if (!x || !w) {
        set_error("null output pointer");
        return FGQ_EINVAL;
    }
    HostPathLock lock;
    hint_huge_pages(x, (size_t)n * 8);
    hint_huge_pages(w, (size_t)n * 8);
    ChunkedProducer cp;
    FGQ_TRY(cp.init(n, 1LL << 24));  // nodes per chunk: 128 MiB of x + 128 MiB of w
    cp.produce = [&](int64_t lo, int64_t hi, double* xd, double* wd, double*, double*, cudaStream_t st) {
        return fn(ctx, lo_all + lo, lo_all + hi, xd, wd, st);
    };
    const int rc = drain_to_host(x, w, n, cp.s_copy,
                                 [&](int64_t lo, int64_t hi, PieceDev* pd) { return cp.piece(lo, hi, pd); }, nullptr);
    return cp.finish(rc);
}

// ---------------------------------------------------------------------------------------------
// Compressed transport of a mirror-pair shard (delta_codec.h): the kernel's output is packed on the device
// into second-difference streams (~1 byte per value instead of 8), those cross PCIe into pinned staging, and
// the host threads unpack them straight into the caller's left slice and its mirror image.  PCIe stops being
// the bound (1-2 GB instead of 8 GB per 1e9-node rule); what remains is the host writing its own 16 GB.
// Works the same for pinned and pageable destinations (no ring copy: the decoder's output IS the result).
// Pipeline: chunk j+1 is generated and packed while chunk j is in flight / being decoded; a chunk's streams
// are cut into batches that fit a staging slot (normally one batch per chunk; incompressible data — never
// produced by the rule kernels, but the codec does not assume that — just takes more batches).
// ---------------------------------------------------------------------------------------------
static std::atomic<int64_t> g_codec_bytes{0}, g_codec_values{0};

#if defined(__x86_64__)
// Decodes a block straight into its two destinations from the vector registers (no bounce buffer): dl[i] = u_i,
// dr_end[-1 - i] = +-u_i.  Requires len >= 8, len % 4 == 0 and both dl and dr_end 32-byte aligned.
__attribute__((target("avx2"))) static void codec_unpack_fused_avx2(const uint8_t* p, int cls, int len, double* dl,
                                                                     double* dr_end, bool negate) {
    __m256i v = _mm256_loadu_si256((const __m256i*)p);
    __m256i d1 = _mm256_loadu_si256((const __m256i*)(p + 32));
    const __m256d sign = negate ? _mm256_set1_pd(-0.0) : _mm256_setzero_pd();
#define FGQ_EMIT4(i, vv)                                                                              \
    do {                                                                                              \
        const __m256d f_ = _mm256_castsi256_pd(vv);                                                   \
        _mm256_stream_pd(dl + (i), f_);                                                               \
        _mm256_stream_pd(dr_end - 4 - (i), _mm256_xor_pd(_mm256_permute4x64_pd(f_, 0x1B), sign));     \
    } while (0)
    FGQ_EMIT4(0, v);
    v = _mm256_add_epi64(v, d1);
    FGQ_EMIT4(4, v);
    const uint8_t* q = p + 64;
    for (int i = 8; i < len; i += 4) {
        __m256i d2;
        switch (cls) {
            case 0: {
                int32_t four;
                memcpy(&four, q + (i - 8), 4);
                d2 = _mm256_cvtepi8_epi64(_mm_cvtsi32_si128(four));
                break;
            }
            case 1: d2 = _mm256_cvtepi16_epi64(_mm_loadl_epi64((const __m128i*)(q + 2 * (size_t)(i - 8)))); break;
            case 2: d2 = _mm256_cvtepi32_epi64(_mm_loadu_si128((const __m128i*)(q + 4 * (size_t)(i - 8)))); break;
            default: d2 = _mm256_loadu_si256((const __m256i*)(q + 8 * (size_t)(i - 8)));
        }
        d1 = _mm256_add_epi64(d1, d2);
        v = _mm256_add_epi64(v, d1);
        FGQ_EMIT4(i, v);
    }
#undef FGQ_EMIT4
}
#endif

// One block of the left slice (elements [e0, e0 + len)) from its coded form into dst[e0 ..] and, for the elements
// below e1m, into the mirror image dstr[mcnt - 1 - e].  No store fence: the caller fences once per run of blocks.
// fused = true takes the register-to-destination path where its alignment conditions hold (off by default:
// FGQ_CODEC_FUSED=1), otherwise the block goes through the 8 KB bounce buffer `tmp`.
static void codec_unpack_block(const uint8_t* p, int cls, int len, double* dst, double* dstr, int64_t mcnt, int64_t e0,
                               int64_t e1m, bool negate, bool fused, int64_t* tmp) {
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (fused && have_avx2 && len >= 8 && (len & 3) == 0 && e1m == e0 + len &&
        ((uintptr_t)(dst + e0) & 31) == 0 && ((uintptr_t)(dstr + (mcnt - e0)) & 31) == 0) {
        codec_unpack_fused_avx2(p, cls, len, dst + e0, dstr + (mcnt - e0), negate);
        return;
    }
#endif
    codec_decode_block(p, cls, len, tmp);
    copy_stream(dst + e0, (const double*)tmp, len, false);
    if (e1m > e0) mirror_copy((const double*)tmp - e0, dstr, mcnt, e0, e1m, negate, false);
}

static bool codec_fused_default() {
    static const bool on = []() {
        const char* e = getenv("FGQ_CODEC_FUSED");
        return e && atoi(e) != 0;
    }();
    return on;
}

struct CodecBatch {
    int64_t chunk_lo;      // first element (of the left slice) of the chunk this batch belongs to
    int64_t b0, b1;        // block range within the chunk
    int64_t chunk_len;     // elements in the chunk
    const uint32_t* dir[2];
    const uint8_t* stage[2];  // staging address of the first byte of block b0's stream, per array
    uint32_t base8[2];        // stream offset (8-byte units) of block b0, per array
    int slot;
};

static int codec_pair_pipeline(int64_t n, int64_t k_lo, int64_t cnt, double* xl, double* wl, const MirrorSpec& mir) {
    constexpr int64_t CHUNK = 1LL << 24;
    constexpr int64_t NB = CHUNK / CODEC_BLOCK;
    constexpr size_t STAGE_BYTES = 64u << 20;            // per staging slot (both arrays of a batch)
    const size_t pay_bytes = (size_t)codec_worst_bytes(CHUNK);  // per array, per device slot
    const int64_t nch = (cnt + CHUNK - 1) / CHUNK;
    const int nslot = nch > 1 ? 2 : 1;
    const int64_t cap = cnt < CHUNK ? cnt : CHUNK;  // values per array in a raw slot; w starts 16-byte aligned after x
    // device memory: raw chunk (scratch slots 0/1 as the plain pipeline), codec arrays (slot 7)
    const size_t dev_per_slot = 2 * pay_bytes + 4 * (size_t)NB * 4 + 64;
    void* dev_all;
    FGQ_TRY(scratch_device(dev_per_slot * 2, 7, &dev_all));
    double* raw[2] = {nullptr, nullptr};
    for (int s = 0; s < nslot; ++s) {
        void* p;
        FGQ_TRY(scratch_device((size_t)(cap + 2) * 16, s, &p));
        raw[s] = (double*)p;
    }
    // pinned: 2 staging slots + per device slot the directories and totals
    void* pin_all;
    const size_t dir_bytes = 2 * (size_t)NB * 4 + 64;
    FGQ_TRY(scratch_pinned(2 * STAGE_BYTES + 2 * dir_bytes, 0, &pin_all));
    uint8_t* stage[2] = {(uint8_t*)pin_all, (uint8_t*)pin_all + STAGE_BYTES};
    uint8_t* dir_h[2] = {(uint8_t*)pin_all + 2 * STAGE_BYTES, (uint8_t*)pin_all + 2 * STAGE_BYTES + dir_bytes};
    CodecArrays ca[2];
    for (int s = 0; s < 2; ++s) {
        uint8_t* base = (uint8_t*)dev_all + (size_t)s * dev_per_slot;
        ca[s].payload[0] = base;
        ca[s].payload[1] = base + pay_bytes;
        ca[s].size8[0] = (uint32_t*)(base + 2 * pay_bytes);
        ca[s].size8[1] = ca[s].size8[0] + NB;
        ca[s].dir[0] = ca[s].size8[1] + NB;
        ca[s].dir[1] = ca[s].dir[0] + NB;   // dir[0], dir[1] contiguous: one copy
        ca[s].total8 = (unsigned long long*)(ca[s].dir[1] + NB);
        ca[s].capacity8 = pay_bytes / 8;
    }
    cudaStream_t s_compute, s_copy;
    FGQ_TRY(scratch_stream(0, &s_compute));
    FGQ_TRY(scratch_stream(1, &s_copy));
    ScopedEvents evs;   // per device slot: packed / all payload copied; per staging slot: landed
    FGQ_TRY(evs.acquire(6));
    cudaEvent_t* packed = evs.ev;
    cudaEvent_t* sent = evs.ev + 2;
    cudaEvent_t* landed = evs.ev + 4;
    WorkerPool& pool = WorkerPool::get();
    const int T = pool.max_threads() > 1 ? pool.max_threads() : 1;
    int dev;
    FGQ_CUDA_TRY(cudaGetDevice(&dev));

    std::mutex mu;
    std::condition_variable cv;
    std::vector<CodecBatch> batches;     // published batches, in order
    int64_t next_group = 0;              // next block group to claim within batches[cur]
    size_t cur = 0;                      // batch being decoded
    int64_t groups_done = 0;             // finished groups of batches[cur]
    size_t batches_done = 0;
    bool all_published = false, failed = false;
    int rc = FGQ_OK;
    cudaError_t err = cudaSuccess;
    constexpr int64_t GROUP = 16;        // blocks per claim

    auto launch_chunk = [&](int64_t j) -> int {
        const int s = (int)(j % nslot);
        const int64_t lo = j * CHUNK, len = lo + CHUNK < cnt ? CHUNK : cnt - lo;
        if (j >= nslot) FGQ_CUDA_TRY(cudaStreamWaitEvent(s_compute, sent[s], 0));  // device slot drained
        double* xd = raw[s];
        double* wd = raw[s] + ((cap + 1) & ~1LL);
        FGQ_TRY(fgq_gausslegendre_device(n, k_lo - 1 + lo, k_lo - 1 + lo + len, xd, wd, (void*)s_compute));
        ca[s].values[0] = xd;
        ca[s].values[1] = wd;
        FGQ_TRY(codec_encode_launch(ca[s], len, s_compute));
        FGQ_CUDA_TRY(cudaEventRecord(packed[s], s_compute));
        return FGQ_OK;
    };

    auto feeder = [&]() {
        int64_t launched = 0;
        size_t published = 0;
        auto fail = [&](int r, cudaError_t e) {
            std::lock_guard<std::mutex> lk(mu);
            if (r != FGQ_OK) rc = r;
            if (e != cudaSuccess) err = e;
            failed = true;
            cv.notify_all();
        };
        for (int64_t j = 0; j < nch; ++j) {
            const int s = (int)(j % nslot);
            while (launched <= j + 1 && launched < nch) {
                const int r = launch_chunk(launched);
                if (r != FGQ_OK) return fail(r, cudaSuccess);
                ++launched;
            }
            const int64_t lo = j * CHUNK, len = lo + CHUNK < cnt ? CHUNK : cnt - lo;
            const int64_t nb = (len + CODEC_BLOCK - 1) / CODEC_BLOCK;
            // the host copy of chunk j-2's directory is still in use until that chunk is decoded
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] {
                    if (failed) return true;
                    // every batch of chunk j - nslot (same slot) consumed?
                    for (size_t q = batches_done; q < batches.size(); ++q)
                        if (batches[q].chunk_lo == (j - nslot) * CHUNK && j >= nslot) return false;
                    return true;
                });
                if (failed) return;
            }
            cudaError_t e = cudaStreamWaitEvent(s_copy, packed[s], 0);
            uint32_t* dh = (uint32_t*)dir_h[s];
            if (e == cudaSuccess)  // dir[0], dir[1], total8[2] are contiguous on the device
                e = cudaMemcpyAsync(dh, ca[s].dir[0], 2 * (size_t)NB * 4 + 16, cudaMemcpyDeviceToHost, s_copy);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s_copy);
            if (e != cudaSuccess) return fail(FGQ_OK, e);
            const uint32_t* d[2] = {dh, dh + NB};
            unsigned long long tot8[2];
            memcpy(tot8, dh + 2 * NB, 16);
            if (tot8[0] > ca[s].capacity8 || tot8[1] > ca[s].capacity8) return fail(FGQ_EINVAL, cudaSuccess);
            g_codec_bytes += (int64_t)(tot8[0] + tot8[1]) * 8;
            g_codec_values += 2 * len;
            auto off8 = [&](int a, int64_t b) -> uint64_t { return b < nb ? (uint64_t)(d[a][b] >> 2) : (uint64_t)tot8[a]; };
            int64_t b0 = 0;
            while (b0 < nb) {
                // largest b1 with both ranges fitting one staging slot
                int64_t b1 = b0;
                while (b1 < nb) {
                    const uint64_t bytes = ((off8(0, b1 + 1) - off8(0, b0)) + (off8(1, b1 + 1) - off8(1, b0))) * 8;
                    if (bytes > STAGE_BYTES) break;
                    ++b1;
                }
                if (b1 == b0) return fail(FGQ_EINVAL, cudaSuccess);  // cannot happen: a block is <= 8.3 KB
                const int slot = (int)(published & 1);
                {   // staging slot free: the batch that used it two publications ago is decoded
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return failed || published < 2 || batches_done + 2 > published; });
                    if (failed) return;
                }
                const size_t bx = (size_t)(off8(0, b1) - off8(0, b0)) * 8, bw = (size_t)(off8(1, b1) - off8(1, b0)) * 8;
                e = cudaMemcpyAsync(stage[slot], ca[s].payload[0] + off8(0, b0) * 8, bx, cudaMemcpyDeviceToHost, s_copy);
                if (e == cudaSuccess)
                    e = cudaMemcpyAsync(stage[slot] + bx, ca[s].payload[1] + off8(1, b0) * 8, bw, cudaMemcpyDeviceToHost, s_copy);
                if (e == cudaSuccess) e = cudaEventRecord(landed[slot], s_copy);
                if (e == cudaSuccess && b1 == nb) e = cudaEventRecord(sent[s], s_copy);
                if (e != cudaSuccess) return fail(FGQ_OK, e);
                CodecBatch B;
                B.chunk_lo = lo;
                B.chunk_len = len;
                B.b0 = b0;
                B.b1 = b1;
                B.dir[0] = d[0];
                B.dir[1] = d[1];
                B.stage[0] = stage[slot];
                B.stage[1] = stage[slot] + bx;
                B.base8[0] = (uint32_t)off8(0, b0);
                B.base8[1] = (uint32_t)off8(1, b0);
                B.slot = slot;
                {
                    std::lock_guard<std::mutex> lk(mu);
                    batches.push_back(B);
                    ++published;
                    cv.notify_all();
                }
                b0 = b1;
            }
        }
        std::lock_guard<std::mutex> lk(mu);
        all_published = true;
        cv.notify_all();
    };

    const bool fused_unpack = codec_fused_default();
    auto decoder = [&]() {
        cudaSetDevice(dev);
        int64_t tmp[CODEC_BLOCK];
        size_t synced = (size_t)-1;  // batch whose landing this thread has already waited for
        for (;;) {
            CodecBatch B;
            int64_t g0, g1;
            size_t my;
            {
                std::unique_lock<std::mutex> lk(mu);
                for (;;) {
                    if (failed) return;
                    if (cur < batches.size()) {
                        const int64_t ngroups = (batches[cur].b1 - batches[cur].b0 + GROUP - 1) / GROUP;
                        if (next_group < ngroups) break;
                        // this batch is fully claimed: wait until it is finished, then move on
                        cv.wait(lk);
                        continue;
                    }
                    if (all_published) return;
                    cv.wait(lk);
                }
                my = cur;
                B = batches[cur];
                g0 = B.b0 + next_group * GROUP;
                g1 = g0 + GROUP < B.b1 ? g0 + GROUP : B.b1;
                ++next_group;
            }
            if (synced != my) {
                if (cudaEventSynchronize(landed[B.slot]) != cudaSuccess) {
                    std::lock_guard<std::mutex> lk(mu);
                    failed = true;
                    cv.notify_all();
                    return;
                }
                synced = my;
            }
            for (int64_t b = g0; b < g1; ++b) {
                const int64_t e0 = B.chunk_lo + b * CODEC_BLOCK;  // element index within the left slice
                const int len = (int)(B.chunk_len - b * CODEC_BLOCK < CODEC_BLOCK ? B.chunk_len - b * CODEC_BLOCK : CODEC_BLOCK);
                const int64_t e1m = e0 + len < mir.mirror_cnt ? e0 + len : mir.mirror_cnt;
                for (int a = 0; a < 2; ++a) {
                    const uint32_t ent = B.dir[a][b];
                    const uint8_t* p = B.stage[a] + (size_t)((ent >> 2) - B.base8[a]) * 8;
                    codec_unpack_block(p, (int)(ent & 3u), len, a == 0 ? xl : wl, a == 0 ? mir.xr : mir.wr, mir.cnt, e0,
                                       e1m, a == 0, fused_unpack, tmp);
                }
            }
            stream_fence();  // one fence per group of blocks, before the group is reported done
            {
                std::lock_guard<std::mutex> lk(mu);
                ++groups_done;
                const int64_t ngroups = (batches[my].b1 - batches[my].b0 + GROUP - 1) / GROUP;
                if (my == cur && groups_done == ngroups) {  // batch finished
                    ++cur;
                    ++batches_done;
                    next_group = 0;
                    groups_done = 0;
                }
                cv.notify_all();
            }
        }
    };
    pool.run(T + 1, [&](int t) {
        if (t == 0) feeder(); else decoder();
    });
    cudaError_t e1 = cudaStreamSynchronize(s_compute);
    cudaError_t e2 = cudaStreamSynchronize(s_copy);
    if (rc != FGQ_OK) {
        set_error("internal: transport coding failed");
        return rc;
    }
    FGQ_CUDA_TRY(err);
    FGQ_CUDA_TRY(e1);
    FGQ_CUDA_TRY(e2);
    if (failed) {
        set_error("device-to-host transfer failed");
        return FGQ_ECUDA;
    }
    return FGQ_OK;
}

// FGQ_HOST_CODEC=0 sends the plain values instead (comparison runs, tests)
static bool use_host_codec() {
    const char* e = getenv("FGQ_HOST_CODEC");
    return !(e && atoi(e) == 0);
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
    // half-indices per pipeline stage; FGQ_HOST_CHUNK_LOG2 is a tuning aid (tools/host_call_timing.py)
    static const int chunk_log2 = []() {
        const char* e = getenv("FGQ_HOST_CHUNK_LOG2");
        const int v = e ? atoi(e) : 0;
        return (v >= 21 && v <= 28) ? v : 24;
    }();
    for (double* p : {xl, wl, xr, wr}) hint_huge_pages(p, (size_t)cnt * 8);
    // The kernel of chunk j+1 runs while chunk j crosses PCIe and is mirrored.  A pinned destination
    // receives the DMA directly and the host threads only write the mirror image; a pageable one (a Julia
    // Vector, a numpy array) is filled through the pinned ring by the same threads, which also spreads
    // the first-touch page faults of a freshly allocated 16 GB result over all of them.
    // everything pinned: the kernel also leaves the +-x image in device memory, so that drain_to_host can
    // balance the mirror work between PCIe and the host threads
    if (use_host_codec() && cnt >= (1LL << 20)) {
        const MirrorSpec mirc{xr, wr, cnt, 0, (k_hi - 1 > mr) ? cnt - 1 : cnt};
        return codec_pair_pipeline(n, k_lo, cnt, xl, wl, mirc);
    }
    const bool all_pinned = is_pinned_host(xl) && is_pinned_host(wl) && is_pinned_host(xr) && is_pinned_host(wr);
    ChunkedProducer cp;
    FGQ_TRY(cp.init(cnt, 1LL << chunk_log2, all_pinned));
    cp.produce = [&](int64_t lo, int64_t hi, double* xd, double* wd, double* xrd, double* wrd, cudaStream_t st) {
        // left slice = output indices [k_lo-1+lo, k_lo-1+hi) = half-indices [k_lo+lo, k_lo+hi)
        if (xrd) return fgq_gausslegendre_device_pair(n, k_lo + lo, k_lo + hi, xd, wd, xrd, wrd, (void*)st);
        return fgq_gausslegendre_device(n, k_lo - 1 + lo, k_lo - 1 + hi, xd, wd, (void*)st);
    };
    // element e of the left slice (half-index k = k_lo + e) mirrors to right-slice position cnt-1-e;
    // the centre of an odd rule (k = m) has no mirror image
    const MirrorSpec mir{xr, wr, cnt, 0, (k_hi - 1 > mr) ? cnt - 1 : cnt};
    const int rc = drain_to_host(xl, wl, cnt, cp.s_copy,
                                 [&](int64_t lo, int64_t hi, PieceDev* pd) { return cp.piece(lo, hi, pd); }, &mir);
    return cp.finish(rc);
}

// counters of the threaded drains since the last call of this function: pieces handled by host threads,
// pieces whose mirror image went by DMA, pieces that had landed before a thread picked them up, and the
// last value of the adaptive DMA share (x 1000)
// bytes that crossed PCIe in coded form and the values they carried, since the last call of this function
extern "C" int fgq_debug_codec_stats(int64_t* bytes, int64_t* values) {
    if (bytes) *bytes = g_codec_bytes.exchange(0);
    if (values) *values = g_codec_values.exchange(0);
    return FGQ_OK;
}

// Round trip of the transport coding on the HOST (tests without a GPU): decode(encode(in)) -> out, stream size
// through *stream_bytes.  The product packs on the device; this is the reference encoder of delta_codec.h.
extern "C" int fgq_debug_codec_roundtrip_host(const double* in, int64_t cnt, double* out, int64_t* stream_bytes) {
    if (cnt < 0 || (cnt > 0 && (!in || !out))) return FGQ_EINVAL;
    const int64_t nb = (cnt + CODEC_BLOCK - 1) / CODEC_BLOCK;
    std::vector<uint8_t> payload((size_t)codec_worst_bytes(cnt) + 64);
    std::vector<uint32_t> dir((size_t)nb + 1);
    const int64_t bytes = codec_encode_host((const int64_t*)in, cnt, payload.data(), dir.data());
    for (int64_t b = 0; b < nb; ++b) {
        const int len = (int)(cnt - b * CODEC_BLOCK < CODEC_BLOCK ? cnt - b * CODEC_BLOCK : CODEC_BLOCK);
        const uint8_t* p = payload.data() + (size_t)(dir[(size_t)b] >> 2) * 8;
        const int cls = (int)(dir[(size_t)b] & 3u);
        codec_decode_block(p, cls, len, (int64_t*)out + b * CODEC_BLOCK);
        // the portable decoder must agree with the dispatched (AVX2) one
        int64_t ref[CODEC_BLOCK];
        codec_decode_block_scalar(p, cls, len, ref);
        if (memcmp(ref, (const int64_t*)out + b * CODEC_BLOCK, (size_t)len * 8) != 0) {
            set_error("internal: scalar and vector decoders disagree in block %lld", (long long)b);
            return FGQ_EINVAL;
        }
    }
    if (stream_bytes) *stream_bytes = bytes;
    return FGQ_OK;
}

// The host half of the coded transfer without a device (tests): `in` (cnt values, the left slice) is coded with the
// reference encoder and unpacked block by block through codec_unpack_block — the routine the pipeline's host threads
// run — into left[0 .. cnt) and the mirror image right[cnt - 1 - e] = (negate ? -in[e] : in[e]) for e < mirror_cnt.
extern "C" int fgq_debug_codec_unpack_host(const double* in, int64_t cnt, double* left, double* right, int64_t mirror_cnt,
                                           int negate, int fused) {
    if (cnt < 0 || mirror_cnt < 0 || mirror_cnt > cnt || (cnt > 0 && (!in || !left || !right))) return FGQ_EINVAL;
    const int64_t nb = (cnt + CODEC_BLOCK - 1) / CODEC_BLOCK;
    std::vector<uint8_t> payload((size_t)codec_worst_bytes(cnt) + 64);
    std::vector<uint32_t> dir((size_t)nb + 1);
    codec_encode_host((const int64_t*)in, cnt, payload.data(), dir.data());
    int64_t tmp[CODEC_BLOCK];
    for (int64_t b = 0; b < nb; ++b) {
        const int64_t e0 = b * CODEC_BLOCK;
        const int len = (int)(cnt - e0 < CODEC_BLOCK ? cnt - e0 : CODEC_BLOCK);
        const int64_t e1m = e0 + len < mirror_cnt ? e0 + len : mirror_cnt;
        codec_unpack_block(payload.data() + (size_t)(dir[(size_t)b] >> 2) * 8, (int)(dir[(size_t)b] & 3u), len, left,
                           right, cnt, e0, e1m, negate != 0, fused != 0, tmp);
    }
    stream_fence();
    return FGQ_OK;
}

extern "C" int fgq_debug_drain_stats(int64_t* pieces, int64_t* dma_pieces, int64_t* landed, int* share_milli) {
    if (pieces) *pieces = g_drain_pieces.exchange(0);
    if (dma_pieces) *dma_pieces = g_drain_dma.exchange(0);
    if (landed) *landed = g_drain_landed.exchange(0);
    if (share_milli) *share_milli = g_drain_share_milli.load();
    return FGQ_OK;
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) static void fill_stream_avx2(double* p, int64_t len, double value) {
    int64_t i = 0;
    while (i < len && ((uintptr_t)(p + i) & 31) != 0) p[i++] = value;
    const __m256d v = _mm256_set1_pd(value);
    for (; i + 4 <= len; i += 4) _mm256_stream_pd(p + i, v);
    for (; i < len; ++i) p[i] = value;
    _mm_sfence();
}
#endif

// Host-memory roofline of the host-returning calls: every result byte is written exactly once by a host thread
// (decoder output / mirror image) with non-temporal stores.  This measures that pattern alone — `threads` pool
// threads streaming constants into buf[0 .. count) — so that e2e can be stated as a fraction of what the host's
// memory system accepts (bench.py).  Needs no device.
extern "C" int fgq_measure_host_store_peak(double* buf, int64_t count, int threads, double* gbs) {
    clear_status();
    if (!buf || !gbs || count < (1 << 16)) {
        set_error("invalid host store measurement request");
        return FGQ_EINVAL;
    }
    WorkerPool& pool = WorkerPool::get();
    int T = threads > 0 ? threads : pool.max_threads();
    if (T > 64) T = 64;
    set_host_thread_budget(T);
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        const auto t0 = std::chrono::steady_clock::now();
        pool.run(T, [&](int t) {
            const int64_t a = (count * t / T) & ~(int64_t)3, b = t + 1 == T ? count : (count * (t + 1) / T) & ~(int64_t)3;
#if defined(__x86_64__)
            static const bool have_avx2 = __builtin_cpu_supports("avx2");
            if (have_avx2) {
                fill_stream_avx2(buf + a, b - a, 1.0 + rep);
                return;
            }
#endif
            for (int64_t i = a; i < b; ++i) buf[i] = 1.0 + rep;
        });
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const double g = (double)count * 8.0 / dt / 1e9;
        if (rep > 0 && g > best) best = g;
    }
    set_host_thread_budget(0);
    *gbs = best;
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

namespace fgq {
void host_parallel_for(int64_t n, const std::function<void(int64_t, int64_t)>& f) {
    unsigned hw = std::thread::hardware_concurrency();
    int64_t T = hw == 0 ? 4 : (hw > 16 ? 16 : hw);
    if (const char* e = getenv("FGQ_HOST_THREADS")) {
        const int v = atoi(e);
        if (v >= 1 && v <= 64) T = v;
    }
    if (n < (1 << 13) || T <= 1) {  // not worth ~50 us per thread
        f(0, n);
        return;
    }
    if (T > n / 2048) T = n / 2048;
    std::vector<std::thread> th;
    for (int64_t t = 1; t < T; ++t) th.emplace_back([&, t] { f(n * t / T, n * (t + 1) / T); });
    f(0, n / T);
    for (auto& x : th) x.join();
}

int small_batch_rules(int64_t nrules, const int32_t* n_i, const int64_t* offsets, int n_max, const char* what,
                      std::vector<SmallRule>* rules, int64_t* total) {
    *total = 0;
    if (nrules < 0 || (nrules > 0 && (!n_i || !offsets))) {
        set_error("invalid batch description");
        return FGQ_EINVAL;
    }
    if (nrules > 0x7fffffffLL) {
        set_error("batch too large");
        return FGQ_EINVAL;
    }
    if (rules) rules->resize((size_t)nrules);
    int64_t tot = 0;
    for (int64_t i = 0; i < nrules; ++i) {
        if (n_i[i] < 0) {
            set_error("DomainError(%d): %s rule %lld of the batch: n must be non-negative", n_i[i], what, (long long)i);
            return FGQ_EDOMAIN;
        }
        if (n_i[i] > n_max) {
            set_error("batched %s rules are limited to n <= %d (rule %lld has n=%d): use the single-rule call",
                      what, n_max, (long long)i, n_i[i]);
            return FGQ_EINVAL;
        }
        if (offsets[i] < 0) {
            set_error("negative offset");
            return FGQ_EINVAL;
        }
        const int64_t end = offsets[i] + n_i[i];
        if (end > tot) tot = end;
        if (rules) (*rules)[(size_t)i] = SmallRule{n_i[i], 0, offsets[i]};
    }
    *total = tot;
    return FGQ_OK;
}
}  // namespace fgq

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
