// common.cuh — shared plumbing of libfgq_cuda.so (error state, launch counter, scratch).
#ifndef FGQ_COMMON_CUH_
#define FGQ_COMMON_CUH_

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <functional>
#include <thread>
#include <vector>

#include "../../include/fgq.h"

namespace fgq {

// thread-local status (fgq_last_error / fgq_warnings)
void set_error(const char* fmt, ...);
void clear_status();
void add_warning(int bits);
void count_launch(int64_t n = 1);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#define FGQ_CUDA_TRY(expr)                                                              \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            fgq::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                           __FILE__, __LINE__);                                         \
            return FGQ_ECUDA + (int)_e;                                                 \
        }                                                                               \
    } while (0)

#define FGQ_TRY(expr)              \
    do {                           \
        int _rc = (expr);          \
        if (_rc != FGQ_OK) return _rc; \
    } while (0)

// Checks after a kernel launch (launch-configuration errors only; never synchronises).
#define FGQ_LAUNCH_CHECK()                 \
    do {                                   \
        fgq::count_launch();               \
        FGQ_CUDA_TRY(cudaGetLastError());  \
    } while (0)

// streaming (evict-first) stores: every output byte is written exactly once
__device__ __forceinline__ void st_stream(double* p, double v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream2(double* p, double a, double b) {
    __stcs(reinterpret_cast<double2*>(p), make_double2(a, b));
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Library-owned device scratch + pinned staging for the *_host entry points.
// Guarded by a per-device mutex (host entry points serialise per device).
struct HostPathLock {   // per current device
    HostPathLock();
    ~HostPathLock();
    int slot;
};
void set_host_thread_budget(int threads);  // thread-local cap on the host threads of this thread's *_host calls (0 = default)
int scratch_device(size_t bytes, int slot, void** out);   // grow-only, per (device, slot)
int scratch_stream(int which, cudaStream_t* out);          // two library streams
int scratch_event(int which, cudaEvent_t* out);

// Events / side streams for the duration of one call, taken from a per-device pool and handed back by the
// destructor on EVERY return path (the objects are reused, never destroyed: an event or stream may still be
// referenced by work the call enqueued).  Replaces per-call cudaEventCreate / Destroy pairs that leaked on early
// returns, and the library-wide side stream that concurrent device-API callers (or a graph capture) had to share.
struct ScopedEvents {
    ScopedEvents() {}
    ~ScopedEvents();
    int acquire(int count);          // timing disabled; ev[0 .. count)
    cudaEvent_t ev[64];
    int n = 0, dev = -1;
    ScopedEvents(const ScopedEvents&) = delete;
    ScopedEvents& operator=(const ScopedEvents&) = delete;
};
struct ScopedStream {
    ScopedStream() {}
    ~ScopedStream();
    // a non-blocking stream nobody else holds right now; high_priority: its blocks are scheduled ahead of those of
    // default-priority streams (a latency-bound chain beside a throughput kernel)
    int acquire(bool high_priority = false);
    cudaStream_t st = nullptr;
    int dev = -1;
    bool high = false;
    ScopedStream(const ScopedStream&) = delete;
    ScopedStream& operator=(const ScopedStream&) = delete;
};

// D2H of a finished result, ordered after `st`; blocks until the data is in place.  Pageable destinations
// go through a pinned ring drained by host threads (capi.cu).  Caller holds HostPathLock.
struct MirrorSpec {      // optional +-x mirroring done by the same host threads (Legendre pair shards)
    double *xr, *wr;     // right slice: xr[cnt-1-e] = -x_e, wr[cnt-1-e] = w_e
    int64_t cnt;         // half-indices of the whole left slice
    int64_t e0;          // element index (within the left slice) of this call's first entry
    int64_t mirror_cnt;  // elements e < mirror_cnt have a mirror image (the centre of an odd rule has none)
};
int drain_to_host(double* x, double* w, const double* x_dev, const double* w_dev, int64_t cnt, cudaStream_t st);
// general form: src(lo, hi, &x_dev, &w_dev) names (and may produce) the device data of piece [lo, hi)
struct PieceDev {
    const double *x, *w;    // device data of the piece
    const double *xr, *wr;  // optional: its +-x image (ascending, i.e. reversed and negated), or null
};
typedef std::function<int(int64_t, int64_t, PieceDev*)> PieceSource;
constexpr int64_t DRAIN_PIECE = 1LL << 18;         // entries per piece through the pinned ring (2 MiB)
constexpr int64_t DRAIN_PIECE_PINNED = 1LL << 20;  // entries per piece into a pinned destination (8 MiB)
int drain_to_host(double* x, double* w, int64_t cnt, cudaStream_t st, const PieceSource& src, const MirrorSpec* mirror);

// Chunked compute -> D2H pipeline shared by every *_host entry point: fn fills device
// buffers with output indices [lo, hi) of the rule; x/w receive hi-lo entries.
typedef int (*SliceFn)(void* ctx, int64_t lo, int64_t hi, double* x_dev, double* w_dev,
                       cudaStream_t st);
int host_pipeline(int64_t lo, int64_t hi, double* x, double* w, SliceFn fn, void* ctx);

// Batched small rules of the Jacobi / Hermite / Laguerre families.  The caller describes
// the batch with HOST arrays (n_i, parameters, CSR offsets): the per-rule constants need host libm
// (gamma functions, Bessel-zero guesses), exactly like the single-rule calls, and are uploaded into the
// workspace before the launch.
struct SmallRule {   // one rule of a batch: n nodes written at x[off .. off+n)
    int32_t n;
    int32_t kind;
    int64_t off;
};
// Validates (n_i, offsets): n < 0 -> FGQ_EDOMAIN, n > n_max or a negative offset -> FGQ_EINVAL.  Fills
// `rules` (may be null) and the output length `total` = max(off + n).
int small_batch_rules(int64_t nrules, const int32_t* n_i, const int64_t* offsets, int n_max, const char* what,
                      std::vector<SmallRule>* rules, int64_t* total);

// Transport coding of a chunk of (x, w) before it crosses PCIe (delta_codec.h, codec.cu).  Index 0 = x, 1 = w.
constexpr int64_t CODEC_MAX_BLOCKS = 1LL << 15;  // 1024-value blocks per chunk (2^25 values)
struct CodecArrays {
    const double* values[2];     // the chunk in device memory
    uint32_t* size8[2];          // per block: (stream size in 8-byte units << 2) | width class
    uint32_t* dir[2];            // per block: (stream offset in 8-byte units << 2) | width class
    uint8_t* payload[2];         // the streams (worst case: 8 B per value + 16 B per block)
    unsigned long long* total8;  // [2]: stream sizes in 8-byte units
    uint64_t capacity8;          // capacity of each payload buffer in 8-byte units
};
int codec_encode_launch(const CodecArrays& a, int64_t cnt, cudaStream_t st);

// f(lo, hi) over [0, n) cut into contiguous ranges, on up to 16 short-lived host threads when n is large
// (per-rule planning of a big batch: gamma functions, Bessel-zero guesses); f must not fail.
void host_parallel_for(int64_t n, const std::function<void(int64_t, int64_t)>& f);

// The per-rule constants of a large small-rule batch (gamma functions, Bessel-zero guesses: host libm, as for the
// single-rule calls) cost the host about as much as the batch costs the device (1e5 Laguerre rules: 6 ms on 16 threads
// against 8 ms of kernels).  The planner runs work(lo, hi) over the positions [0, n) on background threads IN ORDER,
// block by block, so that the first part of the launch order can be uploaded and launched after a fraction of the
// planning and the rest of it runs beside the kernels.  wait(hi) returns once every position below hi is done; the
// destructor joins.  work must not fail.
struct OrderedPlanner {
    static constexpr int64_t BLK = 1024;
    OrderedPlanner(int64_t n, std::function<void(int64_t, int64_t)> work_)
        : n_(n), nblk_((n + BLK - 1) / BLK), done_((size_t)((n + BLK - 1) / BLK)), work(std::move(work_)) {
        for (auto& d : done_) d.store(0, std::memory_order_relaxed);
        unsigned hw = std::thread::hardware_concurrency();
        int64_t T = hw == 0 ? 4 : (hw > 16 ? 16 : hw);
        if (const char* e = getenv("FGQ_HOST_THREADS")) {
            const int v = atoi(e);
            if (v >= 1 && v <= 64) T = v;
        }
        if (T > nblk_) T = nblk_;
        for (int64_t t = 0; t < T; ++t)
            threads_.emplace_back([this] {
                for (;;) {
                    const int64_t b = next_.fetch_add(1, std::memory_order_relaxed);
                    if (b >= nblk_) return;
                    work(b * BLK, (b + 1) * BLK < n_ ? (b + 1) * BLK : n_);
                    done_[(size_t)b].store(1, std::memory_order_release);
                }
            });
    }
    void wait(int64_t hi) {
        const int64_t last = (hi + BLK - 1) / BLK;
        for (; waited_ < last; ++waited_)
            while (!done_[(size_t)waited_].load(std::memory_order_acquire)) std::this_thread::yield();
    }
    ~OrderedPlanner() {
        for (auto& t : threads_) t.join();
    }
    OrderedPlanner(const OrderedPlanner&) = delete;
    OrderedPlanner& operator=(const OrderedPlanner&) = delete;

  private:
    int64_t n_, nblk_, waited_ = 0;
    std::atomic<int64_t> next_{0};
    std::vector<std::atomic<int>> done_;
    std::function<void(int64_t, int64_t)> work;
    std::vector<std::thread> threads_;
};

}  // namespace fgq
#endif
