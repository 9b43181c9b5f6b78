// integrate.cu — dot(w, f.(x)) for a USER integrand (README.md:27-30).
//
// The caller passes f as an expression in x (CUDA C++: "exp(x) * cos(3 * x)", "1 / (1 + 25 * x * x)", ...).  The
// library compiles it at run time with NVRTC together with its own node-generation source (fgq_math.h,
// legendre_device.cuh, integrate_kernels.cuh are embedded as text at build time) for sm_100a, --fmad=false like every
// other kernel here, caches the module per (expression, series branch), and launches it:
//   fgq_gausslegendre_integrate_device   generate-and-reduce over a half-index range: x and w are never stored
//   fgq_rule_integrate_device            the same reduction over a stored rule of any family
// libnvrtc is dlopen'ed (no link-time dependency; a box without it gets FGQ_ENOTIMPL), the module is loaded through
// the CUDA runtime's library API (cudaLibraryLoadData), so libcuda is not linked either.
#include <dlfcn.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "legendre_device.cuh"

namespace fgq {

static const char* kRtcSource =
#include "_build/rtc_source.inc"
    ;

typedef struct _nvrtcProgram* nvrtcProgram;
struct NvrtcApi {
    void* h = nullptr;
    int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    int (*DestroyProgram)(nvrtcProgram*) = nullptr;
};
static NvrtcApi g_rtc;
static std::mutex g_rtc_mu;

struct RtcModule {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t k_integrate = nullptr, k_apply = nullptr;
};
static std::map<std::string, RtcModule> g_modules;   // key: device | NT | WT | HEAD | expression

static int nvrtc_load() {
    if (g_rtc.h) return FGQ_OK;
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"};
    for (const char* nm : names) {
        g_rtc.h = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (g_rtc.h) break;
    }
    if (!g_rtc.h) {
        set_error("user integrands need NVRTC (libnvrtc.so.12 not found: %s)", dlerror());
        return FGQ_ENOTIMPL;
    }
#define FGQ_RTC_SYM(field, name)                                  \
    *(void**)(&g_rtc.field) = dlsym(g_rtc.h, name);               \
    if (!g_rtc.field) {                                           \
        set_error("libnvrtc: symbol %s missing", name);           \
        return FGQ_ENOTIMPL;                                      \
    }
    FGQ_RTC_SYM(CreateProgram, "nvrtcCreateProgram")
    FGQ_RTC_SYM(CompileProgram, "nvrtcCompileProgram")
    FGQ_RTC_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    FGQ_RTC_SYM(GetCUBIN, "nvrtcGetCUBIN")
    FGQ_RTC_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    FGQ_RTC_SYM(GetProgramLog, "nvrtcGetProgramLog")
    FGQ_RTC_SYM(DestroyProgram, "nvrtcDestroyProgram")
#undef FGQ_RTC_SYM
    return FGQ_OK;
}

// The expression is spliced into a one-line function body: refuse what cannot be part of an expression, so that a
// malformed string fails here with a message instead of deep inside the compiler.
static int check_expression(const char* expr) {
    if (!expr || !*expr) {
        set_error("empty integrand expression");
        return FGQ_EINVAL;
    }
    const size_t len = strlen(expr);
    if (len > 4096) {
        set_error("integrand expression too long (%zu characters)", len);
        return FGQ_EINVAL;
    }
    int depth = 0;
    for (const char* c = expr; *c; ++c) {
        if (*c == '(') ++depth;
        if (*c == ')') --depth;
        if (depth < 0 || *c == ';' || *c == '{' || *c == '}' || *c == '#' || *c == '"' || *c == '\\' || *c == '\n') {
            set_error("integrand must be a single expression in x (offending character '%c')", *c);
            return FGQ_EINVAL;
        }
    }
    if (depth != 0) {
        set_error("integrand expression has unbalanced parentheses");
        return FGQ_EINVAL;
    }
    return FGQ_OK;
}

static int get_module(const char* expr, int nt, int wt, int head, RtcModule* out) {
    FGQ_TRY(check_expression(expr));
    int dev = 0;
    FGQ_CUDA_TRY(cudaGetDevice(&dev));
    char keyhead[64];
    snprintf(keyhead, sizeof keyhead, "%d|%d|%d|%d|", dev, nt, wt, head);
    const std::string key = std::string(keyhead) + expr;
    std::lock_guard<std::mutex> guard(g_rtc_mu);
    auto it = g_modules.find(key);
    if (it != g_modules.end()) {
        *out = it->second;
        return FGQ_OK;
    }
    FGQ_TRY(nvrtc_load());
    std::string src = kRtcSource;
    const std::string marker = "/*FGQ_USER_F*/";
    const size_t pos = src.find(marker);
    if (pos == std::string::npos) {
        set_error("internal: integrand marker missing from the embedded source");
        return FGQ_EINVAL;
    }
    src.replace(pos, marker.size(),
                std::string("__device__ __forceinline__ double fgq_user_f(double x) { return (") + expr + "); }\n");
    nvrtcProgram prog = nullptr;
    if (g_rtc.CreateProgram(&prog, src.c_str(), "fgq_user_integrand.cu", 0, nullptr, nullptr) != 0) {
        set_error("nvrtcCreateProgram failed");
        return FGQ_ECUDA;
    }
    char dnt[32], dwt[32], dhd[32];
    snprintf(dnt, sizeof dnt, "-DFGQ_NT=%d", nt);
    snprintf(dwt, sizeof dwt, "-DFGQ_WT=%d", wt);
    snprintf(dhd, sizeof dhd, "-DFGQ_HEAD=%d", head);
    const char* opts[] = {"--gpu-architecture=sm_100a", "--fmad=false", "--std=c++17", "-lineinfo", dnt, dwt, dhd};
    const int rc = g_rtc.CompileProgram(prog, (int)(sizeof opts / sizeof opts[0]), opts);
    if (rc != 0) {
        size_t ls = 0;
        g_rtc.GetProgramLogSize(prog, &ls);
        std::string log(ls + 1, '\0');
        if (ls) g_rtc.GetProgramLog(prog, &log[0]);
        // keep the first diagnostics that mention the user's line
        const size_t at = log.find("error");
        set_error("the integrand \"%s\" does not compile: %.300s", expr, at == std::string::npos ? log.c_str() : log.c_str() + at);
        g_rtc.DestroyProgram(&prog);
        return FGQ_EINVAL;
    }
    size_t bs = 0;
    g_rtc.GetCUBINSize(prog, &bs);
    std::vector<char> cubin(bs);
    g_rtc.GetCUBIN(prog, cubin.data());
    g_rtc.DestroyProgram(&prog);
    RtcModule m;
    FGQ_CUDA_TRY(cudaLibraryLoadData(&m.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    FGQ_CUDA_TRY(cudaLibraryGetKernel(&m.k_integrate, m.lib, "fgq_integrate_legendre_kernel"));
    FGQ_CUDA_TRY(cudaLibraryGetKernel(&m.k_apply, m.lib, "fgq_apply_rule_kernel"));
    g_modules[key] = m;
    *out = m;
    return FGQ_OK;
}

// the same launch descriptor as in integrate_kernels.cuh (host view)
struct FgqIntegrateLaunch {
    long long kmin, kmax, kq0_last, kext_first, kcentre;
    LegConsts c;
};

int legendre_branch(int64_t n, int* nt, int* wt);                           // legendre.cu
void legendre_bulk_split(int64_t n, int64_t* kq0_last, int64_t* kext_first);  // legendre.cu

}  // namespace fgq

using namespace fgq;

extern "C" int fgq_gausslegendre_integrate_device(int64_t n, int64_t k_lo, int64_t k_hi, const char* expr,
                                                  double* sums_dev, void* stream) {
    clear_status();
    if (n <= 60) {
        set_error("fused integration needs the asymptotic path (n > 60), got n=%lld; integrate a stored rule with "
                  "fgq_rule_integrate_device instead", (long long)n);
        return FGQ_EINVAL;
    }
    const int64_t m = (n + 1) >> 1;
    if (k_lo < 1 || k_hi < k_lo || k_hi > m + 1 || !sums_dev) {
        set_error("invalid half-index shard [%lld, %lld) of n=%lld", (long long)k_lo, (long long)k_hi, (long long)n);
        return FGQ_EINVAL;
    }
    FGQ_TRY(check_expression(expr));
    if (k_hi == k_lo) return FGQ_OK;
    int nt, wt;
    FGQ_TRY(legendre_branch(n, &nt, &wt));
    constexpr int64_t K_TAIL = 13192;
    for (int part = 0; part < 2; ++part) {
        const bool head = part == 0;
        FgqIntegrateLaunch p;
        p.kmin = head ? k_lo : (k_lo > K_TAIL ? k_lo : K_TAIL);
        p.kmax = head ? ((k_hi - 1 < K_TAIL - 1) ? k_hi - 1 : K_TAIL - 1) : k_hi - 1;
        if (p.kmin > p.kmax) continue;
        p.c = make_consts(n);
        p.kcentre = (n & 1) ? m : 0;
        p.kq0_last = 0;
        p.kext_first = 0;
        if (nt == 0 && !head) {
            int64_t a, b;
            legendre_bulk_split(n, &a, &b);
            p.kq0_last = a;
            p.kext_first = b;
        }
        RtcModule mod;
        FGQ_TRY(get_module(expr, nt, wt, head ? 1 : 0, &mod));
        const long long ntiles = (p.kmax - p.kmin) / 512 + 1;
        const unsigned blocks = (unsigned)(ntiles < 148 * 8 ? ntiles : 148 * 8);
        void* args[] = {&p, &sums_dev};
        FGQ_CUDA_TRY(cudaLaunchKernel((const void*)mod.k_integrate, dim3(blocks), dim3(256), args, 0, as_stream(stream)));
        FGQ_LAUNCH_CHECK();
    }
    return FGQ_OK;
}

extern "C" int fgq_rule_integrate_device(const double* x_dev, const double* w_dev, int64_t count, const char* expr,
                                         double* sums_dev, void* stream) {
    clear_status();
    FGQ_TRY(check_expression(expr));
    if (count < 0 || !sums_dev || (count > 0 && (!x_dev || !w_dev))) {
        set_error("invalid rule integration request");
        return FGQ_EINVAL;
    }
    if (count == 0) return FGQ_OK;
    RtcModule mod;
    FGQ_TRY(get_module(expr, 0, 0, 0, &mod));
    long long blocks = (count + 256 * 8 - 1) / (256 * 8);
    if (blocks > 148 * 8) blocks = 148 * 8;
    long long cnt = count;
    void* args[] = {&x_dev, &w_dev, &cnt, &sums_dev};
    FGQ_CUDA_TRY(cudaLaunchKernel((const void*)mod.k_apply, dim3((unsigned)blocks), dim3(256), args, 0, as_stream(stream)));
    FGQ_LAUNCH_CHECK();
    return FGQ_OK;
}
