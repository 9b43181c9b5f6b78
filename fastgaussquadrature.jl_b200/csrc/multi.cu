// multi.cu — the multi-GPU face of the C ABI.
//
// One process drives every visible B200: a Julia `ccall` client (the stated host of the north star) gets the
// 1/2/4/8-GPU device-resident rule — "hands back sharded device buffers" — and the optional replicated copy
// through plain C symbols, without torch, torchrun or CUDA.jl on its side.
//
//   fgq_init                               device list, one non-blocking stream + timing events per device, peer access
//   fgq_gausslegendre_device_sharded       allocates (first call) and fills the shards: device d owns the half-index
//                                          range of rank d (mirror-pair sharding: each half-index is evaluated once and
//                                          stored as +-x) — the same partition the one-process-per-GPU path uses, so the
//                                          results are bit-identical; no collective: every node is a function of (n, index)
//   fgq_timer_start / fgq_timer_stop       device-side timing of whatever was enqueued in between, max over devices
//   fgq_allgather                          optional replicated copy of the whole rule on every device, three transports:
//                                          FGQ_ALLGATHER=push (default) one hand-written kernel per device that reads
//                                          its shard once from local HBM and stores it into every peer's replica over
//                                          NVLink (peer-mapped pointers); =ce copy engines (cudaMemcpyPeerAsync);
//                                          =nccl grouped ncclSend/ncclRecv straight into place (libnccl.so.2 is
//                                          dlopen'ed, the library has no link-time NCCL dependency)
//   fgq_gausslegendre_device_replicated    the same replicated result WITHOUT any transfer: every device generates the
//                                          whole rule itself (2.3 ms per device at n = 1e9, against >= 15.6 ms for moving
//                                          14 GB into each GPU at 900 GB/s) — generation is cheaper than communication
//   fgq_free_shards / fgq_free_replicas / fgq_synchronize / fgq_finalize
//
// Host-side threading: device-resident calls are enqueued from the calling thread (cudaSetDevice + two launches per
// device, a few microseconds each; the GPUs run concurrently because nothing synchronises between the enqueues).
#include <dlfcn.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "common.cuh"

namespace fgq {

struct DevCtx {
    int dev = -1;
    cudaStream_t st = nullptr;
    cudaStream_t peer_st[16] = {};  // copy-engine all-gather: one stream per destination
    cudaEvent_t t0 = nullptr, t1 = nullptr, done = nullptr, pushed = nullptr;
    double* sums = nullptr;         // 2 doubles of device scratch for the moments
};

// One enqueue thread per device (created by fgq_init, parked on a condition variable): a steady-state sharded call
// hands each of them its device's two launches, so the devices are fed at the same moment instead of one after the
// other from the calling thread (8 devices x 2 launches of ~5 us would skew the last device's start by ~70 us).
// Deliberately leaked at process exit, like the host-path worker pool.
struct EnqueueWorker {
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, done = false;
    int rc = FGQ_OK;
    std::string err;
    int dev = -1;
    void loop() {
        cudaSetDevice(dev);
        for (;;) {
            std::function<int()> j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return has_job; });
                j = job;
                has_job = false;
            }
            const int r = j();
            std::string e = r != FGQ_OK ? std::string(fgq_last_error()) : std::string();
            std::lock_guard<std::mutex> lk(mu);
            rc = r;
            err = e;
            done = true;
            cv.notify_all();
        }
    }
    void submit(std::function<int()> j) {
        std::lock_guard<std::mutex> lk(mu);
        job = std::move(j);
        has_job = true;
        done = false;
        cv.notify_all();
    }
    int wait() {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return done; });
        return rc;
    }
};

static std::mutex g_multi_mu;
static std::vector<DevCtx> g_ctx;
static std::vector<EnqueueWorker*> g_workers;   // parallel to g_ctx
static bool g_peer_ok = false;

struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { cudaGetDevice(&prev); }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

static int init_locked(const int* device_ids, int ndev);  // caller holds g_multi_mu
static int ensure_init() {
    if (!g_ctx.empty()) return FGQ_OK;
    return init_locked(nullptr, 0);
}

// Rank r of `world` owns half-indices [k_lo, k_hi) (1-based): boundaries on even counts so that both segments keep
// 16-byte vector stores (identical to half_index_partition in the Python host mirror).
static void half_index_range(int64_t n, int world, int r, int64_t* k_lo, int64_t* k_hi) {
    const int64_t m = (n + 1) >> 1;
    auto bound = [&](int q) { return q >= world ? m : (((__int128)q * m) / world) & ~(int64_t)1; };
    *k_lo = (int64_t)bound(r) + 1;
    *k_hi = (int64_t)bound(r + 1) + 1;
}

// dst_j[i] = src[i] for every destination j: the all-gather as ONE read of the local shard and ndst remote (or
// local) 16-byte stores per element pair.  Remote stores travel over NVLink through the peer mapping.
struct PushArgs {
    const double* src[2];  // x, w of this segment
    double* dst[2][16];    // per destination replica: where element 0 of the segment goes
    int ndst;
    int64_t cnt;
};

__global__ void __launch_bounds__(512) allgather_push_kernel(const PushArgs a) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int arr = 0; arr < 2; ++arr) {
        const double* __restrict__ s = a.src[arr];
        // all segment starts share the parity of their 16-byte alignment by construction (even boundaries); the
        // host falls back to cnt2 = 0 when that is not the case
        const bool vec = ((reinterpret_cast<uintptr_t>(s) & 15u) == 0);
        bool vd = vec;
        for (int j = 0; j < a.ndst; ++j) vd = vd && ((reinterpret_cast<uintptr_t>(a.dst[arr][j]) & 15u) == 0);
        const int64_t cnt2 = vd ? a.cnt / 2 : 0;
        for (int64_t i = i0; i < cnt2; i += stride) {
            const double2 v = __ldcs(reinterpret_cast<const double2*>(s) + i);
            for (int j = 0; j < a.ndst; ++j) __stcs(reinterpret_cast<double2*>(a.dst[arr][j]) + i, v);
        }
        for (int64_t i = 2 * cnt2 + i0; i < a.cnt; i += stride) {
            const double v = __ldcs(s + i);
            for (int j = 0; j < a.ndst; ++j) __stcs(a.dst[arr][j] + i, v);
        }
    }
}

// The right half of a Gauss-Legendre rule from its left half (gausslegendre.jl:83-90: the reference mirrors too):
// x[n-1-i] = -x[i], w[n-1-i] = w[i] for i < n/2.  A replica therefore only needs the LEFT halves over NVLink; its
// owner writes the image locally (8 GB read + 8 GB written at n = 1e9, ~2.5 ms, against 7 GB less inbound traffic).
__global__ void __launch_bounds__(256) legendre_mirror_kernel(double* __restrict__ x, double* __restrict__ w, int64_t n) {
    const int64_t half = n >> 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((n & 1) == 0) {
        // pairs (i, i+1), i even -> destination pair (n-2-i, n-1-i), 16-byte aligned because n is even
        for (int64_t q = t0; 2 * q + 1 < half + (half & 1); q += stride) {
            const int64_t i = 2 * q;
            if (i + 1 < half) {
                const double2 vx = __ldcs(reinterpret_cast<const double2*>(x + i));
                const double2 vw = __ldcs(reinterpret_cast<const double2*>(w + i));
                __stcs(reinterpret_cast<double2*>(x + (n - 2 - i)), make_double2(-vx.y, -vx.x));
                __stcs(reinterpret_cast<double2*>(w + (n - 2 - i)), make_double2(vw.y, vw.x));
            } else if (i < half) {
                x[n - 1 - i] = -x[i];
                w[n - 1 - i] = w[i];
            }
        }
    } else {
        for (int64_t i = t0; i < half; i += stride) {
            x[n - 1 - i] = -__ldcs(x + i);
            w[n - 1 - i] = __ldcs(w + i);
        }
    }
}

// ---- NCCL through dlopen (no link-time dependency) ------------------------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::vector<ncclComm_t> comms;
};
static NcclApi g_nccl;
constexpr int NCCL_FLOAT64 = 8;  // ncclDouble (nccl.h)

static int nccl_load() {
    if (g_nccl.h) return FGQ_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        g_nccl.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.h) break;
    }
    if (!g_nccl.h) {
        set_error("FGQ_ALLGATHER=nccl: libnccl.so.2 not found (%s)", dlerror());
        return FGQ_ENOTIMPL;
    }
#define FGQ_NCCL_SYM(field, name)                                         \
    *(void**)(&g_nccl.field) = dlsym(g_nccl.h, name);                     \
    if (!g_nccl.field) {                                                  \
        set_error("libnccl: symbol %s missing", name);                    \
        return FGQ_ENOTIMPL;                                              \
    }
    FGQ_NCCL_SYM(CommInitAll, "ncclCommInitAll")
    FGQ_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    FGQ_NCCL_SYM(GroupStart, "ncclGroupStart")
    FGQ_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    FGQ_NCCL_SYM(Send, "ncclSend")
    FGQ_NCCL_SYM(Recv, "ncclRecv")
    FGQ_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef FGQ_NCCL_SYM
    return FGQ_OK;
}

#define FGQ_NCCL_TRY(expr)                                                                   \
    do {                                                                                     \
        ncclResult_t _r = (expr);                                                            \
        if (_r != 0) {                                                                       \
            set_error("%s failed: %s", #expr, g_nccl.GetErrorString(_r));                    \
            return FGQ_ECUDA;                                                                \
        }                                                                                    \
    } while (0)

static int nccl_comms() {
    FGQ_TRY(nccl_load());
    if (g_nccl.comms.size() == g_ctx.size()) return FGQ_OK;
    std::vector<int> devs;
    for (auto& c : g_ctx) devs.push_back(c.dev);
    g_nccl.comms.assign(g_ctx.size(), nullptr);
    FGQ_NCCL_TRY(g_nccl.CommInitAll(g_nccl.comms.data(), (int)devs.size(), devs.data()));
    return FGQ_OK;
}

// Layout of one device's allocation per array: [left segment | pad to even | right segment (cnt slots)].
struct ShardLayout {
    int64_t k_lo, k_hi, cnt, pad, skip;  // skip = 1 iff this shard holds the centre of an odd rule
};
static ShardLayout shard_layout(int64_t n, int world, int r) {
    ShardLayout s;
    half_index_range(n, world, r, &s.k_lo, &s.k_hi);
    s.cnt = s.k_hi - s.k_lo;
    s.pad = s.cnt + (s.cnt & 1);
    const int64_t m = (n + 1) >> 1;
    s.skip = ((n & 1) && s.cnt > 0 && s.k_hi - 1 == m) ? 1 : 0;
    return s;
}

}  // namespace fgq

using namespace fgq;

extern "C" int fgq_init(const int* device_ids, int ndev) {
    clear_status();
    std::lock_guard<std::mutex> guard(g_multi_mu);
    return init_locked(device_ids, ndev);
}

static int fgq::init_locked(const int* device_ids, int ndev) {
    int visible = 0;
    FGQ_CUDA_TRY(cudaGetDeviceCount(&visible));
    if (visible <= 0) {
        set_error("no CUDA device visible (this library has no CPU fallback)");
        return FGQ_ECUDA;
    }
    std::vector<int> ids;
    if (!device_ids || ndev <= 0) {
        const int want = ndev > 0 ? ndev : visible;
        if (want > visible) {
            set_error("fgq_init: %d devices requested, %d visible", want, visible);
            return FGQ_EINVAL;
        }
        for (int d = 0; d < want; ++d) ids.push_back(d);
    } else {
        for (int i = 0; i < ndev; ++i) {
            if (device_ids[i] < 0 || device_ids[i] >= visible) {
                set_error("fgq_init: device ordinal %d out of range (%d visible)", device_ids[i], visible);
                return FGQ_EINVAL;
            }
            for (int j = 0; j < i; ++j)
                if (device_ids[j] == device_ids[i]) {
                    set_error("fgq_init: device %d listed twice", device_ids[i]);
                    return FGQ_EINVAL;
                }
            ids.push_back(device_ids[i]);
        }
    }
    if (ids.size() > 16) {
        set_error("fgq_init: at most 16 devices");
        return FGQ_EINVAL;
    }
    // same list as before: nothing to do
    if (g_ctx.size() == ids.size()) {
        bool same = true;
        for (size_t i = 0; i < ids.size(); ++i) same = same && g_ctx[i].dev == ids[i];
        if (same) return FGQ_OK;
    }
    if (!g_ctx.empty()) {
        set_error("fgq_init: already initialised with a different device list; call fgq_finalize first");
        return FGQ_EINVAL;
    }
    DeviceGuard keep;
    std::vector<DevCtx> ctx(ids.size());
    for (size_t i = 0; i < ids.size(); ++i) {
        DevCtx& c = ctx[i];
        c.dev = ids[i];
        FGQ_CUDA_TRY(cudaSetDevice(c.dev));
        FGQ_CUDA_TRY(cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking));
        FGQ_CUDA_TRY(cudaEventCreate(&c.t0));
        FGQ_CUDA_TRY(cudaEventCreate(&c.t1));
        FGQ_CUDA_TRY(cudaEventCreateWithFlags(&c.done, cudaEventDisableTiming));
        FGQ_CUDA_TRY(cudaEventCreateWithFlags(&c.pushed, cudaEventDisableTiming));
        FGQ_CUDA_TRY(cudaMalloc(&c.sums, 4 * sizeof(double)));
    }
    // peer access between every pair (NVLink / NVSwitch on an HGX box); not fatal when unavailable — only the
    // push all-gather needs it
    bool peer = true;
    for (size_t i = 0; i < ids.size(); ++i) {
        FGQ_CUDA_TRY(cudaSetDevice(ids[i]));
        for (size_t j = 0; j < ids.size(); ++j) {
            if (i == j) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, ids[i], ids[j]);
            if (!can) {
                peer = false;
                continue;
            }
            cudaError_t e = cudaDeviceEnablePeerAccess(ids[j], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) {
                cudaGetLastError();
                peer = false;
            }
        }
    }
    g_peer_ok = peer;
    g_ctx.swap(ctx);
    g_workers.clear();
    for (auto& c : g_ctx) {
        EnqueueWorker* wk = new EnqueueWorker();
        wk->dev = c.dev;
        std::thread([wk]() { wk->loop(); }).detach();
        g_workers.push_back(wk);
    }
    return FGQ_OK;
}

extern "C" int fgq_debug_half_index_range(int64_t n, int world, int rank, int64_t* k_lo, int64_t* k_hi) {
    if (n < 0 || world < 1 || rank < 0 || rank >= world || !k_lo || !k_hi) return FGQ_EINVAL;
    half_index_range(n, world, rank, k_lo, k_hi);
    return FGQ_OK;
}

extern "C" int fgq_num_devices(void) {
    std::lock_guard<std::mutex> guard(g_multi_mu);
    return (int)g_ctx.size();
}

extern "C" int fgq_device_ordinal(int index) {
    std::lock_guard<std::mutex> guard(g_multi_mu);
    return (index >= 0 && index < (int)g_ctx.size()) ? g_ctx[(size_t)index].dev : -1;
}

extern "C" int fgq_finalize(void) {
    clear_status();
    std::lock_guard<std::mutex> guard(g_multi_mu);
    DeviceGuard keep;
    for (auto comm : g_nccl.comms)
        if (comm) g_nccl.CommDestroy(comm);
    g_nccl.comms.clear();
    for (auto& c : g_ctx) {
        cudaSetDevice(c.dev);
        cudaStreamSynchronize(c.st);
        for (auto& s : c.peer_st)
            if (s) cudaStreamDestroy(s);
        cudaStreamDestroy(c.st);
        cudaEventDestroy(c.t0);
        cudaEventDestroy(c.t1);
        cudaEventDestroy(c.done);
        cudaEventDestroy(c.pushed);
        cudaFree(c.sums);
    }
    g_ctx.clear();
    return FGQ_OK;
}

extern "C" int fgq_shard_count(int64_t n) {
    std::lock_guard<std::mutex> guard(g_multi_mu);
    if (ensure_init() != FGQ_OK) return 0;
    if (n <= 60) return n > 0 ? 1 : 0;
    return 2 * (int)g_ctx.size();
}

extern "C" int fgq_synchronize(void) {
    clear_status();
    std::lock_guard<std::mutex> guard(g_multi_mu);
    DeviceGuard keep;
    for (auto& c : g_ctx) {
        FGQ_CUDA_TRY(cudaSetDevice(c.dev));
        FGQ_CUDA_TRY(cudaStreamSynchronize(c.st));
        for (auto& s : c.peer_st)
            if (s) FGQ_CUDA_TRY(cudaStreamSynchronize(s));
    }
    return FGQ_OK;
}

extern "C" int fgq_timer_start(void) {
    clear_status();
    std::lock_guard<std::mutex> guard(g_multi_mu);
    FGQ_TRY(ensure_init());
    DeviceGuard keep;
    for (auto& c : g_ctx) {
        FGQ_CUDA_TRY(cudaSetDevice(c.dev));
        FGQ_CUDA_TRY(cudaStreamSynchronize(c.st));
    }
    for (auto& c : g_ctx) {
        FGQ_CUDA_TRY(cudaSetDevice(c.dev));
        FGQ_CUDA_TRY(cudaEventRecord(c.t0, c.st));
    }
    return FGQ_OK;
}

extern "C" int fgq_timer_stop(float* ms_max, float* ms_per_device) {
    clear_status();
    std::lock_guard<std::mutex> guard(g_multi_mu);
    DeviceGuard keep;
    for (auto& c : g_ctx) {
        FGQ_CUDA_TRY(cudaSetDevice(c.dev));
        FGQ_CUDA_TRY(cudaEventRecord(c.t1, c.st));
    }
    float worst = 0.f;
    for (size_t i = 0; i < g_ctx.size(); ++i) {
        DevCtx& c = g_ctx[i];
        FGQ_CUDA_TRY(cudaSetDevice(c.dev));
        FGQ_CUDA_TRY(cudaEventSynchronize(c.t1));
        float ms = 0.f;
        FGQ_CUDA_TRY(cudaEventElapsedTime(&ms, c.t0, c.t1));
        if (ms_per_device) ms_per_device[i] = ms;
        worst = ms > worst ? ms : worst;
    }
    if (ms_max) *ms_max = worst;
    return FGQ_OK;
}

extern "C" int fgq_gausslegendre_device_sharded(int64_t n, fgq_shard* shards, int max_shards, int* nshards,
                                                double* sums) {
    clear_status();
    if (n < 0) {
        set_error("DomainError(%lld): Input n must be a non-negative integer", (long long)n);
        return FGQ_EDOMAIN;
    }
    std::lock_guard<std::mutex> guard(g_multi_mu);
    FGQ_TRY(ensure_init());
    const int world = (int)g_ctx.size();
    const int need = n == 0 ? 0 : (n <= 60 ? 1 : 2 * world);
    if (nshards) *nshards = need;
    if (need == 0) {
        if (sums) sums[0] = sums[1] = 0.0;
        return FGQ_OK;
    }
    if (!shards || max_shards < need) {
        set_error("fgq_gausslegendre_device_sharded: %d shard descriptors needed, %d given", need, max_shards);
        return FGQ_EINVAL;
    }
    DeviceGuard keep;
    if (n <= 60) {
        // the recurrence path: one warp's worth of work, on the first device
        DevCtx& c = g_ctx[0];
        FGQ_CUDA_TRY(cudaSetDevice(c.dev));
        fgq_shard& s = shards[0];
        if (!s.x) {
            FGQ_CUDA_TRY(cudaMalloc(&s.x, 2 * 64 * sizeof(double)));
            s.w = s.x + 64;
            s.device = c.dev;
            s.lo = 0;
            s.hi = n;
            s.owner = 1;
        } else if (s.device != c.dev || s.lo != 0 || s.hi != n) {
            set_error("shard descriptors do not belong to gausslegendre(%lld) on this device list", (long long)n);
            return FGQ_EINVAL;
        }
        FGQ_TRY(fgq_gausslegendre_device(n, 0, n, s.x, s.w, c.st));
    } else {
        const int64_t m = (n + 1) >> 1;
        (void)m;
        for (int r = 0; r < world; ++r) {
            DevCtx& c = g_ctx[r];
            const ShardLayout L = shard_layout(n, world, r);
            fgq_shard& sl = shards[2 * r];
            fgq_shard& sr = shards[2 * r + 1];
            FGQ_CUDA_TRY(cudaSetDevice(c.dev));
            const int64_t llo = L.k_lo - 1, lhi = L.k_hi - 1;
            const int64_t rlo = n - L.k_hi + 1 + L.skip, rhi = n - L.k_lo + 1;
            if (!sl.x) {
                // one allocation per device: x = [left | pad | right], then w the same; 2 * pad entries each
                double* base = nullptr;
                FGQ_CUDA_TRY(cudaMalloc(&base, (size_t)(4 * L.pad > 0 ? 4 * L.pad : 2) * sizeof(double)));
                sl.device = sr.device = c.dev;
                sl.lo = llo; sl.hi = lhi;
                sr.lo = rlo; sr.hi = rhi;
                sl.x = base;
                sl.w = base + 2 * L.pad;
                sr.x = base + L.pad + L.skip;
                sr.w = base + 3 * L.pad + L.skip;
                sl.owner = 1;  // fgq_free_shards releases sl.x
                sr.owner = 0;  // a view into the same allocation
            } else if (sl.device != c.dev || sl.lo != llo || sl.hi != lhi || sr.lo != rlo || sr.hi != rhi ||
                       sr.x != sl.x + L.pad + L.skip) {
                set_error("shard descriptors do not belong to gausslegendre(%lld) on this device list "
                          "(zero-initialise them before the first call)", (long long)n);
                return FGQ_EINVAL;
            }
        }
        // every descriptor is valid: hand each device its launches (nothing above this line has been submitted, so an
        // early return cannot leave a worker with a job nobody waits for)
        for (int r = 0; r < world; ++r) {
            const ShardLayout L = shard_layout(n, world, r);
            const fgq_shard& sl = shards[2 * r];
            const fgq_shard& sr = shards[2 * r + 1];
            if (L.cnt > 0) {
                double *xl = sl.x, *wl = sl.w, *xr = sr.x - L.skip, *wr = sr.w - L.skip;
                cudaStream_t stq = g_ctx[r].st;
                const int64_t klo = L.k_lo, khi = L.k_hi;
                g_workers[r]->submit([=]() { return fgq_gausslegendre_device_pair(n, klo, khi, xl, wl, xr, wr, stq); });
            } else {
                g_workers[r]->submit([]() { return (int)FGQ_OK; });
            }
        }
        int rc_all = FGQ_OK;
        for (int r = 0; r < world; ++r) {
            const int rc = g_workers[r]->wait();
            if (rc != FGQ_OK && rc_all == FGQ_OK) {
                rc_all = rc;
                set_error("device %d: %s", g_ctx[r].dev, g_workers[r]->err.c_str());
            }
        }
        if (rc_all != FGQ_OK) return rc_all;
    }
    if (!sums) return FGQ_OK;  // stream-ordered: fgq_synchronize() / fgq_timer_stop() wait for the result
    // sum w, sum w x^2 per device, added on the host in device order (deterministic)
    for (int r = 0; r < (n <= 60 ? 1 : world); ++r) {
        DevCtx& c = g_ctx[r];
        FGQ_CUDA_TRY(cudaSetDevice(c.dev));
        FGQ_CUDA_TRY(cudaMemsetAsync(c.sums, 0, 2 * sizeof(double), c.st));
        for (int q = 0; q < (n <= 60 ? 1 : 2); ++q) {
            const fgq_shard& s = shards[(n <= 60 ? 0 : 2 * r) + q];
            if (s.hi > s.lo) FGQ_TRY(fgq_rule_moments_device(s.x, s.w, s.hi - s.lo, c.sums, c.st));
        }
    }
    sums[0] = sums[1] = 0.0;
    for (int r = 0; r < (n <= 60 ? 1 : world); ++r) {
        DevCtx& c = g_ctx[r];
        double h[2];
        FGQ_CUDA_TRY(cudaSetDevice(c.dev));
        FGQ_CUDA_TRY(cudaMemcpyAsync(h, c.sums, sizeof(h), cudaMemcpyDeviceToHost, c.st));
        FGQ_CUDA_TRY(cudaStreamSynchronize(c.st));
        sums[0] += h[0];
        sums[1] += h[1];
    }
    return FGQ_OK;
}

extern "C" int fgq_free_shards(fgq_shard* shards, int nshards) {
    clear_status();
    if (!shards) return FGQ_OK;
    DeviceGuard keep;
    for (int i = 0; i < nshards; ++i) {
        if (shards[i].owner && shards[i].x) {
            FGQ_CUDA_TRY(cudaSetDevice(shards[i].device));
            FGQ_CUDA_TRY(cudaFree(shards[i].x));
        }
        shards[i].x = shards[i].w = nullptr;
        shards[i].owner = 0;
    }
    return FGQ_OK;
}

static int alloc_replicas(int64_t n, double** x_repl, double** w_repl) {
    for (size_t d = 0; d < g_ctx.size(); ++d) {
        if (x_repl[d]) continue;
        FGQ_CUDA_TRY(cudaSetDevice(g_ctx[d].dev));
        double* base = nullptr;
        const int64_t pad = n + (n & 1);
        FGQ_CUDA_TRY(cudaMalloc(&base, (size_t)(2 * pad > 0 ? 2 * pad : 2) * sizeof(double)));
        x_repl[d] = base;
        w_repl[d] = base + pad;
    }
    return FGQ_OK;
}

extern "C" int fgq_free_replicas(double** x_repl, double** w_repl) {
    clear_status();
    std::lock_guard<std::mutex> guard(g_multi_mu);
    DeviceGuard keep;
    for (size_t d = 0; d < g_ctx.size(); ++d) {
        if (x_repl && x_repl[d]) {
            FGQ_CUDA_TRY(cudaSetDevice(g_ctx[d].dev));
            FGQ_CUDA_TRY(cudaFree(x_repl[d]));
            x_repl[d] = nullptr;
        }
        if (w_repl) w_repl[d] = nullptr;
    }
    return FGQ_OK;
}

extern "C" int fgq_gausslegendre_device_replicated(int64_t n, double** x_repl, double** w_repl) {
    clear_status();
    if (n < 0) {
        set_error("DomainError(%lld): Input n must be a non-negative integer", (long long)n);
        return FGQ_EDOMAIN;
    }
    if (!x_repl || !w_repl) {
        set_error("null replica table");
        return FGQ_EINVAL;
    }
    std::lock_guard<std::mutex> guard(g_multi_mu);
    FGQ_TRY(ensure_init());
    DeviceGuard keep;
    FGQ_TRY(alloc_replicas(n, x_repl, w_repl));
    if (n == 0) return FGQ_OK;
    for (size_t d = 0; d < g_ctx.size(); ++d) {
        DevCtx& c = g_ctx[d];
        FGQ_CUDA_TRY(cudaSetDevice(c.dev));
        if (n <= 60) {
            FGQ_TRY(fgq_gausslegendre_device(n, 0, n, x_repl[d], w_repl[d], c.st));
        } else {
            const int64_t m = (n + 1) >> 1;
            // the whole rule as one mirror pair: left = [0, m), right = [n - m, n) (its element 0 is the centre of
            // an odd rule and is left to the left slice)
            FGQ_TRY(fgq_gausslegendre_device_pair(n, 1, m + 1, x_repl[d], w_repl[d], x_repl[d] + (n - m),
                                                  w_repl[d] + (n - m), c.st));
        }
    }
    return FGQ_OK;
}

extern "C" int fgq_allgather(const fgq_shard* shards, int nshards, int64_t n, double** x_repl, double** w_repl) {
    clear_status();
    if (!shards || !x_repl || !w_repl || n <= 0) {
        set_error("invalid all-gather request");
        return FGQ_EINVAL;
    }
    std::lock_guard<std::mutex> guard(g_multi_mu);
    FGQ_TRY(ensure_init());
    const int world = (int)g_ctx.size();
    DeviceGuard keep;
    FGQ_TRY(alloc_replicas(n, x_repl, w_repl));
    auto ctx_of = [&](int dev) -> DevCtx* {
        for (auto& c : g_ctx)
            if (c.dev == dev) return &c;
        return nullptr;
    };
    for (int i = 0; i < nshards; ++i) {
        if (shards[i].hi > shards[i].lo && (!ctx_of(shards[i].device) || shards[i].lo < 0 || shards[i].hi > n)) {
            set_error("shard %d does not belong to this device list / rule", i);
            return FGQ_EINVAL;
        }
    }
    const char* mode_env = getenv("FGQ_ALLGATHER");
    std::string mode = mode_env ? mode_env : "push";
    if (mode == "push" && !g_peer_ok && world > 1) mode = "ce";
    // Mirror-pair shards: ship only the left segments (output indices below (n+1)/2) and let every replica's owner
    // write the +-x image locally — half the NVLink traffic (FGQ_ALLGATHER_HALF=0 ships both segments).
    bool half = n > 60;
    if (const char* e = getenv("FGQ_ALLGATHER_HALF")) half = half && atoi(e) != 0;
    const int64_t m_left = (n + 1) >> 1;
    std::vector<fgq_shard> list;
    for (int i = 0; i < nshards; ++i) {
        if (shards[i].hi <= shards[i].lo) continue;
        if (half && shards[i].lo >= m_left) continue;          // a right segment: regenerated by the mirror
        if (half && shards[i].hi > m_left) half = false;       // a segment straddling the centre: not pair shards
        list.push_back(shards[i]);
    }
    if (!half) {
        list.clear();
        for (int i = 0; i < nshards; ++i)
            if (shards[i].hi > shards[i].lo) list.push_back(shards[i]);
    }
    shards = list.data();
    nshards = (int)list.size();
    // The shards are produced on the library's per-device streams, and every transport below runs on them, so a
    // source's own data is ordered.  But a transport WRITES into the other devices' replicas: it must not overtake
    // what those devices still have queued on them (the mirror kernel of a previous all-gather, a previous
    // generation into the same replica): every stream first waits for every other stream's current tail.
    if (world > 1) {
        for (auto& c : g_ctx) {
            FGQ_CUDA_TRY(cudaSetDevice(c.dev));
            FGQ_CUDA_TRY(cudaEventRecord(c.done, c.st));
        }
        for (auto& c : g_ctx) {
            FGQ_CUDA_TRY(cudaSetDevice(c.dev));
            for (auto& o : g_ctx)
                if (o.dev != c.dev) FGQ_CUDA_TRY(cudaStreamWaitEvent(c.st, o.done, 0));
        }
    }
    if (mode == "push") {
        for (int i = 0; i < nshards; ++i) {
            const fgq_shard& s = shards[i];
            if (s.hi <= s.lo) continue;
            DevCtx& c = *ctx_of(s.device);
            PushArgs a;
            a.src[0] = s.x;
            a.src[1] = s.w;
            a.cnt = s.hi - s.lo;
            a.ndst = world;
            for (int d = 0; d < world; ++d) {
                a.dst[0][d] = x_repl[d] + s.lo;
                a.dst[1][d] = w_repl[d] + s.lo;
            }
            FGQ_CUDA_TRY(cudaSetDevice(c.dev));
            allgather_push_kernel<<<148 * 4, 512, 0, c.st>>>(a);
            FGQ_LAUNCH_CHECK();
        }
    } else if (mode == "ce") {
        for (int i = 0; i < nshards; ++i) {
            const fgq_shard& s = shards[i];
            if (s.hi <= s.lo) continue;
            DevCtx& c = *ctx_of(s.device);
            FGQ_CUDA_TRY(cudaSetDevice(c.dev));
            FGQ_CUDA_TRY(cudaEventRecord(c.done, c.st));
            const size_t bytes = (size_t)(s.hi - s.lo) * sizeof(double);
            for (int d = 0; d < world; ++d) {
                if (!c.peer_st[d]) FGQ_CUDA_TRY(cudaStreamCreateWithFlags(&c.peer_st[d], cudaStreamNonBlocking));
                FGQ_CUDA_TRY(cudaStreamWaitEvent(c.peer_st[d], c.done, 0));
                FGQ_CUDA_TRY(cudaMemcpyPeerAsync(x_repl[d] + s.lo, g_ctx[d].dev, s.x, c.dev, bytes, c.peer_st[d]));
                FGQ_CUDA_TRY(cudaMemcpyPeerAsync(w_repl[d] + s.lo, g_ctx[d].dev, s.w, c.dev, bytes, c.peer_st[d]));
            }
        }
        // join the copy streams back into the per-device streams so that fgq_timer_stop / fgq_synchronize see them
        for (auto& c : g_ctx) {
            FGQ_CUDA_TRY(cudaSetDevice(c.dev));
            for (int d = 0; d < world; ++d) {
                if (!c.peer_st[d]) continue;
                FGQ_CUDA_TRY(cudaEventRecord(c.done, c.peer_st[d]));
                FGQ_CUDA_TRY(cudaStreamWaitEvent(c.st, c.done, 0));
            }
        }
    } else if (mode == "nccl") {
        // An all-gather with unequal contributions delivered in place: every rank sends each of its segments to
        // every peer and receives each peer's segments straight into its replica, all inside one NCCL group (NCCL
        // fuses grouped point-to-point operations into one kernel per device).  ncclAllGather itself would need
        // equal counts and rank-ordered slots, which the mirrored right segments (descending in the rank) are not.
        FGQ_TRY(nccl_comms());
        auto rank_of = [&](int dev) {
            for (int r = 0; r < world; ++r)
                if (g_ctx[r].dev == dev) return r;
            return -1;
        };
        FGQ_NCCL_TRY(g_nccl.GroupStart());
        for (int i = 0; i < nshards; ++i) {
            const fgq_shard& s = shards[i];
            if (s.hi <= s.lo) continue;
            const int src = rank_of(s.device);
            const size_t cnt = (size_t)(s.hi - s.lo);
            for (int d = 0; d < world; ++d) {
                if (d == src) continue;
                FGQ_NCCL_TRY(g_nccl.Send(s.x, cnt, NCCL_FLOAT64, d, g_nccl.comms[src], g_ctx[src].st));
                FGQ_NCCL_TRY(g_nccl.Send(s.w, cnt, NCCL_FLOAT64, d, g_nccl.comms[src], g_ctx[src].st));
                FGQ_NCCL_TRY(g_nccl.Recv(x_repl[d] + s.lo, cnt, NCCL_FLOAT64, src, g_nccl.comms[d], g_ctx[d].st));
                FGQ_NCCL_TRY(g_nccl.Recv(w_repl[d] + s.lo, cnt, NCCL_FLOAT64, src, g_nccl.comms[d], g_ctx[d].st));
            }
        }
        FGQ_NCCL_TRY(g_nccl.GroupEnd());
        // the rank's own segments: a local copy
        for (int i = 0; i < nshards; ++i) {
            const fgq_shard& s = shards[i];
            if (s.hi <= s.lo) continue;
            const int src = rank_of(s.device);
            FGQ_CUDA_TRY(cudaSetDevice(s.device));
            const size_t bytes = (size_t)(s.hi - s.lo) * sizeof(double);
            FGQ_CUDA_TRY(cudaMemcpyAsync(x_repl[src] + s.lo, s.x, bytes, cudaMemcpyDeviceToDevice, g_ctx[src].st));
            FGQ_CUDA_TRY(cudaMemcpyAsync(w_repl[src] + s.lo, s.w, bytes, cudaMemcpyDeviceToDevice, g_ctx[src].st));
        }
    } else {
        set_error("FGQ_ALLGATHER must be push, ce or nccl (got %s)", mode.c_str());
        return FGQ_EINVAL;
    }
    if (half) {
        // every replica must be complete before its owner mirrors it: the pushes / copies into device d's replica
        // run on the SOURCE devices' streams (NCCL receives already run on d's own stream)
        if (mode != "nccl") {
            for (auto& c : g_ctx) {
                FGQ_CUDA_TRY(cudaSetDevice(c.dev));
                FGQ_CUDA_TRY(cudaEventRecord(c.pushed, c.st));
            }
            for (auto& c : g_ctx) {
                FGQ_CUDA_TRY(cudaSetDevice(c.dev));
                for (auto& o : g_ctx)
                    if (o.dev != c.dev) FGQ_CUDA_TRY(cudaStreamWaitEvent(c.st, o.pushed, 0));
            }
        }
        for (int d = 0; d < world; ++d) {
            DevCtx& c = g_ctx[d];
            FGQ_CUDA_TRY(cudaSetDevice(c.dev));
            legendre_mirror_kernel<<<148 * 8, 256, 0, c.st>>>(x_repl[d], w_repl[d], n);
            FGQ_LAUNCH_CHECK();
        }
    }
    return FGQ_OK;
}

// x[n-1-i] = -x[i], w[n-1-i] = w[i] for i < n/2 on the current device: the right half of a rule whose left half
// (incl. the centre of an odd rule) is in place.  Used by the one-process-per-GPU all-gather (distributed.py).
extern "C" int fgq_gausslegendre_mirror_device(int64_t n, double* x_dev, double* w_dev, void* stream) {
    clear_status();
    if (n < 0 || (n > 1 && (!x_dev || !w_dev))) {
        set_error("invalid mirror request");
        return FGQ_EINVAL;
    }
    if (n < 2) return FGQ_OK;
    legendre_mirror_kernel<<<148 * 8, 256, 0, as_stream(stream)>>>(x_dev, w_dev, n);
    FGQ_LAUNCH_CHECK();
    return FGQ_OK;
}

// Host-returning rule through every initialised device at once.  Each device runs the single-device mirror-pair
// pipeline (generate -> pack -> PCIe -> host threads unpack + mirror) on its own half-index range from its own
// host thread; per-device locks, staging and worker pools make the calls independent, and the host's cores are
// shared out between them.
extern "C" int fgq_gausslegendre_host_multi(int64_t n, double* x, double* w) {
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
    std::vector<int> devs;
    {
        std::lock_guard<std::mutex> guard(g_multi_mu);
        FGQ_TRY(ensure_init());
        for (auto& c : g_ctx) devs.push_back(c.dev);
    }
    const int world = (int)devs.size();
    DeviceGuard keep;
    if (world == 1 || n < (1LL << 23)) {  // small rules: one device, one pipeline
        FGQ_CUDA_TRY(cudaSetDevice(devs[0]));
        return fgq_gausslegendre_host(n, x, w);
    }
    unsigned hw = std::thread::hardware_concurrency();
    int total = hw == 0 ? 8 : (int)hw;
    if (const char* e = getenv("FGQ_HOST_THREADS")) {
        const int v = atoi(e);
        if (v >= 1) total = v;
    }
    const int per_dev = total / world > 1 ? total / world : 1;
    std::vector<int> rcs((size_t)world, FGQ_OK);
    std::vector<std::string> msgs((size_t)world);
    std::vector<std::thread> threads;
    for (int r = 0; r < world; ++r) {
        threads.emplace_back([&, r]() {
            if (cudaSetDevice(devs[r]) != cudaSuccess) {
                rcs[r] = FGQ_ECUDA;
                msgs[r] = "cudaSetDevice failed";
                return;
            }
            set_host_thread_budget(per_dev);
            int64_t k_lo, k_hi;
            half_index_range(n, world, r, &k_lo, &k_hi);
            if (k_hi > k_lo)
                rcs[r] = fgq_gausslegendre_host_pair(n, k_lo, k_hi, x + (k_lo - 1), w + (k_lo - 1), x + (n - k_hi + 1),
                                                     w + (n - k_hi + 1));
            if (rcs[r] != FGQ_OK) msgs[r] = fgq_last_error();
            set_host_thread_budget(0);
        });
    }
    for (auto& t : threads) t.join();
    for (int r = 0; r < world; ++r)
        if (rcs[r] != FGQ_OK) {
            set_error("device %d: %s", devs[r], msgs[r].c_str());
            return rcs[r];
        }
    return FGQ_OK;
}
