/* fgq.h — C ABI of libfgq_cuda.so: B200-native Gauss quadrature node/weight generation.
 *
 * This is the drop-in boundary for the per-node generation path of
 * FastGaussQuadrature.jl v1.3.0.  The reference has no FFI layer of its own (it is pure
 * Julia); its boundary is the exported Julia functions (src/FastGaussQuadrature.jl:7-19).
 * Each entry point below states which reference function it replaces; the Julia
 * `ccall` stubs a maintainer would add are in julia/FGQCuda.jl and INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; Float64 everywhere (the reference's generic-T
 *     paths are out of scope: there is no CPU fallback in this library);
 *   - return 0 = ok, FGQ_EDOMAIN = the reference would throw DomainError,
 *     FGQ_ENOTIMPL = parameter region the reference serves with an eigen-solve or a
 *     non-Float64 type (not on the hot path), >= FGQ_ECUDA = CUDA failure;
 *     fgq_last_error() returns a thread-local message for the last non-zero return;
 *   - *_host variants fill caller-allocated HOST buffers (a Julia Vector{Float64}, a
 *     numpy array); the pointers are not retained after return;
 *   - *_device variants fill caller-allocated DEVICE buffers on the current CUDA
 *     device, asynchronously on `stream` (a cudaStream_t passed as void*; NULL = the
 *     default stream).  They never synchronise and never allocate, so they can be
 *     captured in a CUDA graph.  One process per GPU; shard by passing each rank its
 *     own index range;
 *   - nodes are ascending, indices are 0-based in this header (Julia's are 1-based).
 */
#ifndef FGQ_H_
#define FGQ_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FGQ_OK 0
#define FGQ_EDOMAIN 1
#define FGQ_ENOTIMPL 2
#define FGQ_EINVAL 3
#define FGQ_ECUDA 100

/* warning bits, cf. the reference's @warn sites (src/gausslaguerre.jl:118,257,424,570) */
#define FGQ_WARN_LARGE_ALPHA 1
#define FGQ_WARN_NEWTON_MAXITER 2
#define FGQ_WARN_NONFINITE 4

int fgq_version(void);
const char* fgq_last_error(void);
/* warning bits accumulated by this thread's last call */
int fgq_warnings(void);
/* number of visible CUDA devices (0 => every compute entry point fails with FGQ_ECUDA) */
int fgq_device_count(void);

/* Page-locked host memory for the destinations of the *_host calls.  The reference returns freshly allocated
 * arrays (`Vector{Float64}(undef, n)`, src/gausslegendre.jl:244-248): at n = 1e9 the first touch of such a 16 GB
 * result costs ~0.3 s of page faults, three times the whole transfer.  A client that reuses its result vectors
 * (or wraps these: `unsafe_wrap(Vector{Float64}, ptr, n)`) gets the DMA straight into place instead.  The memory
 * is usable from every device; fgq_host_free(NULL) is a no-op. */
int fgq_host_alloc(int64_t bytes, void** out);
int fgq_host_free(void* p);

/* ---- Gauss-Legendre: gausslegendre(n) — src/gausslegendre.jl:22-69 ---------------------- */

/* x[0..n), w[0..n) on the host.  n < 0 -> FGQ_EDOMAIN (gausslegendre.jl:26). */
int fgq_gausslegendre_host(int64_t n, double* x, double* w);
/* Same, output indices [lo, hi) only -> x[0 .. hi-lo), w[0 .. hi-lo): lets each process of a
 * multi-GPU job fetch its own part of a large rule over its own PCIe link. */
int fgq_gausslegendre_host_slice(int64_t n, int64_t lo, int64_t hi, double* x, double* w);

/* Host-returning mirror-pair shard (n > 60): half-indices [k_lo, k_hi) as in
 * fgq_gausslegendre_device_pair; only the left slice crosses PCIe, the mirrored slice
 * (x[n-1-i] = -x[i], w[n-1-i] = w[i]) is written by host threads while the next chunk is in flight.
 * xl/wl and xr/wr receive k_hi-k_lo entries each (element 0 of xr/wr is left untouched when the shard
 * holds the centre of an odd rule).  fgq_gausslegendre_host uses it for n >= 2^21.
 * Shards of >= 2^20 half-indices cross PCIe as a lossless second-difference coding of their bit patterns
 * (~1 byte per value at n = 1e9 instead of 8; csrc/delta_codec.h) that library-owned host threads unpack
 * straight into xl/wl and the mirror image; the result is bit-identical to the plain transfer
 * (environment FGQ_HOST_CODEC=0).  Destinations may be pinned or pageable. */
int fgq_gausslegendre_host_pair(int64_t n, int64_t k_lo, int64_t k_hi, double* xl, double* wl,
                                double* xr, double* wr);

/* Contiguous node-index shard: writes nodes/weights with output index in [lo, hi) to
 * x_dev[0 .. hi-lo), w_dev[0 .. hi-lo).  0 <= lo <= hi <= n. */
int fgq_gausslegendre_device(int64_t n, int64_t lo, int64_t hi, double* x_dev, double* w_dev,
                             void* stream);

/* Mirror-pair shard (asymptotic path, n > 60): one evaluation per half-index k serves
 * both +-x.  Half-indices k in [k_lo, k_hi) (1-based, 1 <= k_lo <= k_hi <= (n+1)/2 + 1):
 *   left  slice = output indices [k_lo-1, k_hi-1)      -> xl_dev/wl_dev[0 .. k_hi-k_lo)
 *   right slice = output indices [n-k_hi+1, n-k_lo+1)  -> xr_dev/wr_dev[0 .. k_hi-k_lo)
 * For odd n the centre node (k = (n+1)/2) belongs to the left slice only; the right slice
 * of a shard that contains it starts one element later (its element 0 is left untouched). */
int fgq_gausslegendre_device_pair(int64_t n, int64_t k_lo, int64_t k_hi, double* xl_dev,
                                  double* wl_dev, double* xr_dev, double* wr_dev, void* stream);

/* Batch of small rules gausslegendre(n_i), CSR layout: rule i occupies
 * [offsets[i], offsets[i]+n_i[i]) of x and w.  0 <= n_i <= 60 (closed forms for n <= 5,
 * gausslegendre.jl:27-60; Newton on the three-term recurrence, :233-309).  One warp per rule. */
int fgq_gausslegendre_batch_device(int64_t nrules, const int32_t* n_i_dev,
                                   const int64_t* offsets_dev, double* x_dev, double* w_dev,
                                   void* stream);
int fgq_gausslegendre_batch_host(int64_t nrules, const int32_t* n_i, const int64_t* offsets,
                                 double* x, double* w);

/* ---- Gauss-Jacobi: gaussjacobi(n, a, b) — src/gaussjacobi.jl:23-55,157-667 ------------------ */

/* Asymptotic branch jacobi_asy (n > 100, max(a,b) < 5: gaussjacobi.jl:48-49): Hale-Townsend
 * interior expansion + Newton on the device, 2 x 10 boundary nodes planned on the host.
 *   n < 0, min(a,b) <= -1          -> FGQ_EDOMAIN  (:26-27, :44-45)
 *   a == b == 0                    -> the Legendre path (:28-29)
 *   |a| == |b| == 1/2              -> the Chebyshev closed forms (:30-39)
 *   n <= 100                       -> jacobi_rec (:46-47), one block per rule; n = 1 closed form (:42-43)
 *   max(a,b) >= 5, n <= 4000       -> the Golub-Welsch rule (:50-51) computed per node: eigenvalue of the
 *                                     Jacobi matrix by Sturm bisection, weight from the orthonormal recurrence
 *   max(a,b) >= 5, n > 4000        -> FGQ_ENOTIMPL (the reference raises an ErrorException, :52-53) */
int fgq_gaussjacobi_host(int64_t n, double a, double b, double* x, double* w);
int64_t fgq_gaussjacobi_workspace_bytes(int64_t n);
/* Device-resident: the whole rule is solved on the current device (Newton stopping rule and series
 * truncation are global); output indices [lo, hi) go to x_dev/w_dev[0 .. hi-lo).  Enqueues a fixed
 * sequence of launches; no host synchronisation. */
int fgq_gaussjacobi_device(int64_t n, double a, double b, int64_t lo, int64_t hi, double* x_dev,
                           double* w_dev, void* workspace_dev, void* stream);
/* Newton sweeps (before the extra one) and series terms kept, for the right/left half. */
int fgq_gaussjacobi_stats(const void* workspace_dev, int* sweeps2, int* terms2);

/* gaussradau(n) — src/gaussradau.jl:31-49: gaussjacobi(n-1, 0, 1), w/(1+x), x[0] = -1, w[0] = 2/n^2.
 * gausslobatto(n) — src/gausslobatto.jl:31-52: gaussjacobi(n-2, 1, 1), w/(1-x^2), ends +-1.
 * n <= 0 (Radau) / n <= 1 (Lobatto) -> FGQ_EDOMAIN; inner rule with <= 100 nodes -> FGQ_ENOTIMPL.
 * workspace: fgq_gaussjacobi_workspace_bytes(n). */
int fgq_gaussradau_host(int64_t n, double* x, double* w);
int fgq_gausslobatto_host(int64_t n, double* x, double* w);
int fgq_gaussradau_device(int64_t n, double* x_dev, double* w_dev, void* workspace_dev, void* stream);
int fgq_gausslobatto_device(int64_t n, double* x_dev, double* w_dev, void* workspace_dev, void* stream);

/* gaussradau(n, a, b) — src/gaussradau.jl:52-69 (general Jacobi weight): eigenvalues of the modified
 * Jacobi matrix, computed per node as above; x[0] = -1 exactly (:67).  n <= 0 -> FGQ_EDOMAIN;
 * n > 4000 -> FGQ_ENOTIMPL.  workspace: fgq_gaussjacobi_workspace_bytes(n). */
int fgq_gaussradau_jacobi_host(int64_t n, double a, double b, double* x, double* w);
int fgq_gaussradau_jacobi_device(int64_t n, double a, double b, double* x_dev, double* w_dev,
                                 void* workspace_dev, void* stream);

/* approx_besselroots(nu, n) — src/besselroots.jl:23-67 (12-digit zeros of J_nu): one thread per zero
 * (McMahon's expansion, :74-116); the first 20 (nu = 0, table) / 6 (-1 <= nu <= 5, Piessens, :118-165) zeros
 * are parameter-only data planned on the host.  n < 0 -> FGQ_EDOMAIN (:41-43). */
int fgq_approx_besselroots(double nu, int64_t n, double* x);
int fgq_approx_besselroots_device(double nu, int64_t n, double* x_dev, void* stream);
/* J_nu(x), real order, moderate x: the host planner's stand-in for the reference's besselj. */
double fgq_besselj(double nu, double x);

/* ---- Gauss-Chebyshev: gausschebyshevt/u/v/w(n) — src/gausschebyshev.jl:22-113 ------------------ */

/* kind 1..4 = t, u, v, w (weights (1-x^2)^-1/2, (1-x^2)^1/2, sqrt((1+x)/(1-x)), sqrt((1-x)/(1+x)));
 * also what gaussjacobi(n, -+1/2, -+1/2) routes to (gaussjacobi.jl:30-39).  n < 0 -> FGQ_EDOMAIN. */
int fgq_gausschebyshev_host(int kind, int64_t n, double* x, double* w);
int fgq_gausschebyshev_device(int kind, int64_t n, int64_t lo, int64_t hi, double* x_dev, double* w_dev,
                              void* stream);

/* ---- Gauss-Hermite: gausshermite(n; normalize) — src/gausshermite.jl:28-89,171-323 ---------- */

/* n > 200: asymptotic (Airy) + Newton path (gausshermite.jl:49-50); 21..200: Newton on the regularised
 * recurrence (:46-48); 2..20: eigen-solve of the Jacobi matrix (:43-45); n = 1 closed form.
 * n < 0 -> FGQ_EDOMAIN (:38).  normalize != 0 applies x*sqrt(2), w/sqrt(pi) (:31). */
int fgq_gausshermite_host(int64_t n, int normalize, double* x, double* w);
/* Device-resident.  The whole half-range is always solved on the current device (the Newton
 * stopping rule and the normalisation sum are global); output indices [lo, hi) are written to
 * x_dev/w_dev[0 .. hi-lo).  workspace_dev: fgq_gausshermite_workspace_bytes(n) bytes, 16-byte
 * aligned.  Enqueues a fixed sequence of launches (no host synchronisation): the global
 * stopping rule is evaluated on the device. */
int64_t fgq_gausshermite_workspace_bytes(int64_t n);
int fgq_gausshermite_device(int64_t n, int normalize, int64_t lo, int64_t hi, double* x_dev,
                            double* w_dev, void* workspace_dev, void* stream);
/* Newton sweeps the last call on this workspace used (synchronises; for tests/reporting). */
int fgq_gausshermite_sweeps(const void* workspace_dev, int* sweeps);

/* ---- Gauss-Laguerre: gausslaguerre(n, alpha; reduced) — src/gausslaguerre.jl:59-261,273-529 --- */

/* Explicit asymptotic expansions (gausslaguerre_asy with T = -1, recompute = true: the call at
 * gausslaguerre.jl:91) for n >= 128 and alpha < 5; closed forms (n <= 2), Golub-Welsch (n < 15) and
 * the recurrence rule (n < 128, or alpha >= 5 up to n = 4096) on a single thread (:70-95).
 * alpha <= -1 or n < 0 -> FGQ_EDOMAIN (:60-65); alpha >= 5 with n > 4096 -> FGQ_ENOTIMPL.
 * reduced != 0: only the nodes before the first weight below 10*floatmin are returned (:163-169);
 * *n_out receives the number of entries written (n when reduced == 0).  Warning bits mirror the
 * reference's @warn sites (fgq_warnings()). */
int fgq_gausslaguerre_host(int64_t n, double alpha, int reduced, double* x, double* w, int64_t* n_out);
int64_t fgq_gausslaguerre_workspace_bytes(int64_t n);
/* Device-resident: x_dev/w_dev hold all n entries; the reduced length is available through
 * fgq_gausslaguerre_stats.  Enqueues a fixed sequence of launches; no host synchronisation.  workspace_dev:
 * fgq_gausslaguerre_workspace_bytes(n) bytes, 32-byte aligned (the error estimates, the list of nodes handed to the
 * recompute hook, and for n >= 128 the second candidate expansion of the ~0.1 n + 8 sqrt(n) nodes next to the two
 * hand-over indices).  The asymptotic rule forks onto library-owned side streams and joins them before it
 * returns to the caller's stream: stream-ordered and graph-capturable like the others. */
int fgq_gausslaguerre_device(int64_t n, double alpha, double* x_dev, double* w_dev, void* workspace_dev,
                             void* stream);
/* hand-over indices (1-based; n+1 = never), number of recompute candidates, reduced length. */
int fgq_gausslaguerre_stats(const void* workspace_dev, int64_t* k_b, int64_t* k_a, int64_t* n_flagged,
                            int64_t* n_reduced);

/* ---- truncated Gauss-Hermite rule ------------------------------------------ */

/* The analogue of gausslaguerre(n; reduced = true) (gausslaguerre.jl:163-169) for Gauss-Hermite: at
 * n = 1e6 only 2.4 % of the weights gausshermite returns (gausshermite.jl:28-33) are non-zero, because
 * exp(-x^2) (:30) is exactly 0.0 beyond |x| = 27.3.  These calls compute only the centre block of the rule:
 * the half-range nodes inside |x| <= 27.5 (bounded a priori through the Sturm spacing pi / sqrt(2n+1) of
 * the zeros), with the same kernels, Newton stopping rule (:79) and normalisation sum (:62) taken over
 * that block.  The entries agree with the full rule to the Newton tolerance (not bit for bit: the global
 * stopping rule sees fewer nodes); every omitted weight is 0.0 in the full rule.  n <= 200: the whole rule.
 *   *first  receives the 0-based index, within the full n-point rule, of the first entry written,
 *   *count / *n_out the number of entries.  fgq_gausshermite_reduced_bound(n) bounds it (size of x, w).
 * The host variant also drops the entries below 10 * floatmin at both ends (the reference's threshold for
 * the reduced Laguerre rule).  workspace_dev: fgq_gausshermite_workspace_bytes(n) bytes. */
int64_t fgq_gausshermite_reduced_bound(int64_t n);
int fgq_gausshermite_reduced_device(int64_t n, int normalize, double* x_dev, double* w_dev, void* workspace_dev,
                                    void* stream, int64_t* first, int64_t* count);
int fgq_gausshermite_reduced_host(int64_t n, int normalize, double* x, double* w, int64_t* first, int64_t* n_out);

/* ---- batched small rules of the other families ------------------------- */

/* Batches of small rules in the CSR layout of fgq_gausslegendre_batch_*: rule i occupies
 * [offsets[i], offsets[i] + n_i[i]) of x and w.  Each rule is computed by the small-n branch the
 * reference dispatches it to, with the same device code as the single-rule call, so a batch is
 * bit-identical to the concatenation of single calls:
 *   gaussjacobi   0 <= n_i <= 100   jacobi_rec (gaussjacobi.jl:61-155), one 128-thread block per rule; n = 1
 *                                   closed form (:42-43); (0,0) / (+-1/2,+-1/2) / max(a,b) >= 5 rules go
 *                                   through their Legendre / Chebyshev / Golub-Welsch routes (:28-39, :50-51)
 *   gausshermite  0 <= n_i <= 200   hermite_gw n <= 20 (gausshermite.jl:325-341), hermite_rec (:91-137); one block
 *                                   per rule incl. normalisation sum and fold (:55-62)
 *   gausslaguerre 0 <= n_i <= 127   closed forms (gausslaguerre.jl:70-78), gausslaguerre_GW n < 15 (:535-542),
 *                                   gausslaguerre_rec (:582-641); one thread per rule (the branch is sequential)
 * n_i, alpha_i, beta_i and offsets are HOST arrays in both variants: the per-rule constants (gamma
 * functions, Bessel-zero guesses) are planned on the host exactly as in the single-rule calls and uploaded
 * into workspace_dev (>= fgq_*_batch_workspace_bytes(nrules)) ahead of the launch.  The device variants
 * fill device buffers asynchronously on `stream` and never synchronise.
 *   n_i < 0 or an inadmissible parameter -> FGQ_EDOMAIN (as the single-rule call);
 *   n_i above the limit                  -> FGQ_EINVAL  (use the single-rule call). */
int64_t fgq_gaussjacobi_batch_workspace_bytes(int64_t nrules);
int fgq_gaussjacobi_batch_device(int64_t nrules, const int32_t* n_i, const double* alpha_i, const double* beta_i,
                                 const int64_t* offsets, double* x_dev, double* w_dev, void* workspace_dev,
                                 void* stream);
int fgq_gaussjacobi_batch_host(int64_t nrules, const int32_t* n_i, const double* alpha_i, const double* beta_i,
                               const int64_t* offsets, double* x, double* w);
int64_t fgq_gausshermite_batch_workspace_bytes(int64_t nrules);
int fgq_gausshermite_batch_device(int64_t nrules, const int32_t* n_i, const int64_t* offsets, int normalize,
                                  double* x_dev, double* w_dev, void* workspace_dev, void* stream);
int fgq_gausshermite_batch_host(int64_t nrules, const int32_t* n_i, const int64_t* offsets, int normalize,
                                double* x, double* w);
int64_t fgq_gausslaguerre_batch_workspace_bytes(int64_t nrules);
int fgq_gausslaguerre_batch_device(int64_t nrules, const int32_t* n_i, const double* alpha_i,
                                   const int64_t* offsets, double* x_dev, double* w_dev, void* workspace_dev,
                                   void* stream);
int fgq_gausslaguerre_batch_host(int64_t nrules, const int32_t* n_i, const double* alpha_i, const int64_t* offsets,
                                 double* x, double* w);

/* ---- multi-GPU, one process ------------------------------------------- */

/* A piece of a rule resident on one device: output indices [lo, hi) of the n-point rule live at
 * x[0 .. hi-lo), w[0 .. hi-lo) in the memory of CUDA device `device`.  The buffers are owned by the
 * library until fgq_free_shards.  This is the "sharded device buffers" handle of the device-resident
 * variant of gausslegendre(n) (src/gausslegendre.jl:22-69 returns two host Vectors; the replacement keeps
 * the result where it was generated). */
typedef struct fgq_shard {
    int32_t device;  /* CUDA device ordinal */
    int32_t owner;   /* 1: x is the base of an allocation (released by fgq_free_shards); 0: a view into one */
    int64_t lo, hi;  /* output indices [lo, hi), 0-based */
    double* x;       /* device pointers on `device` */
    double* w;
} fgq_shard;

/* Selects the devices (NULL / ndev <= 0: all visible ones, or the first ndev when only ndev > 0 is given),
 * creates one stream per device and enables peer access between them.  Optional: every call below
 * initialises with all visible devices on first use.  Idempotent for the same list. */
int fgq_init(const int* device_ids, int ndev);
int fgq_finalize(void);
int fgq_num_devices(void);
/* CUDA ordinal of the index-th initialised device (the order of x_repl / w_repl and of the shard pairs); -1 if out of range */
int fgq_device_ordinal(int index);
/* number of fgq_shard entries fgq_gausslegendre_device_sharded writes for this n (2 per device: the left
 * segment and its mirror image; 1 for n <= 60) */
int fgq_shard_count(int64_t n);

/* Device-resident gausslegendre(n) over all initialised devices: device d generates the half-index range of
 * rank d — output indices [k_lo-1, k_hi-1) and their mirror image [n-k_hi+1, n-k_lo+1), each half-index
 * evaluated once — into its own HBM.  No collective: every node is a function of (n, index).
 *   shards[2d], shards[2d+1]  left / right segment of device d, ascending in the output index.
 *   Zero-initialised descriptors are allocated by the call; descriptors filled by a previous call for the same
 *   n are reused (no allocation: steady-state calls only enqueue two launches per device).
 *   sums != NULL: sums[0] = sum w, sums[1] = sum w x^2 over the whole rule (per-device reductions added on the
 *   host in device order); the call then returns after the result is complete.  sums == NULL: the call only
 *   enqueues on the per-device streams; fgq_synchronize / fgq_timer_stop wait.
 * n < 0 -> FGQ_EDOMAIN (gausslegendre.jl:26). */
int fgq_gausslegendre_device_sharded(int64_t n, fgq_shard* shards, int max_shards, int* nshards, double* sums);
int fgq_free_shards(fgq_shard* shards, int nshards);
int fgq_synchronize(void);
/* Device-side timing of everything enqueued between the two calls: one event pair per device on the library's
 * per-device streams; *ms_max = the slowest device, ms_per_device[fgq_num_devices()] optional. */
int fgq_timer_start(void);
int fgq_timer_stop(float* ms_max, float* ms_per_device);

/* Optional replicated copy of the whole rule on every device (north star: "NCCL over NVLink is used only for
 * an optional allgather into a replicated array"): x_repl[d], w_repl[d] (d < fgq_num_devices()) are device
 * pointers on device d holding all n entries; NULL entries are allocated by the call (fgq_free_replicas).
 * Stream-ordered after the generation; fgq_synchronize waits.  Transport (environment FGQ_ALLGATHER):
 *   push (default)  one kernel per device: reads its shard once and stores it into every replica through the
 *                   peer mappings (NVLink / NVSwitch);
 *   ce              copy engines (cudaMemcpyPeerAsync), one stream per destination;
 *   nccl            grouped ncclSend / ncclRecv straight into place (libnccl.so.2, dlopen'ed). */
int fgq_allgather(const fgq_shard* shards, int nshards, int64_t n, double** x_repl, double** w_repl);
/* Mirror-pair shards travel as their left segments only (half the NVLink bytes): once a replica's left half is in
 * place its owner writes x[n-1-i] = -x[i], w[n-1-i] = w[i] (gausslegendre.jl:83-90 mirrors the same way) with
 * this kernel; FGQ_ALLGATHER_HALF=0 ships both segments instead.  Also callable on its own: fills output indices
 * [n - n/2, n) of x_dev, w_dev (n entries each, current device) from [0, n/2). */
int fgq_gausslegendre_mirror_device(int64_t n, double* x_dev, double* w_dev, void* stream);
/* The same replicated result without any transfer: every device generates the whole rule itself (generation at
 * ~7 TB/s of local HBM stores is cheaper than moving 14 of 16 GB into every GPU over NVLink). */
int fgq_gausslegendre_device_replicated(int64_t n, double** x_repl, double** w_repl);
int fgq_free_replicas(double** x_repl, double** w_repl);

/* Host-returning gausslegendre(n) that drains through ALL initialised devices at once: device d generates,
 * packs and ships the half-index range of rank d over its own PCIe link; the host threads (shared out between
 * the devices) unpack into x, w and write the mirror images.  Same result as fgq_gausslegendre_host, bit for bit. */
int fgq_gausslegendre_host_multi(int64_t n, double* x, double* w);

/* ---- consumers / checks ------------------------------------------------------------------ */

/* sums_dev[0] += sum w_i, sums_dev[1] += sum w_i x_i^2 over count entries (device buffers;
 * zero sums_dev first).  The README's only use of a rule is dot(w, f.(x)) (README.md:27-30);
 * this is that reduction for f = 1 and f = x^2, used as the sum-w / int x^2 check. */
int fgq_rule_moments_device(const double* x_dev, const double* w_dev, int64_t count,
                            double* sums_dev, void* stream);

/* Fused generate-and-reduce (no stores): sums_dev[0] += sum w, sums_dev[1] += sum w x^2 and, when
 * with_exp != 0, sums_dev[2] += sum w exp(x), over the nodes of half-indices [k_lo, k_hi) of the
 * n-point Gauss-Legendre rule AND their mirror images (n > 60).  sums_dev: 3 doubles, zeroed by
 * the caller; all-reduce across ranks for a sharded rule.  This is dot(w, f.(x)) of README.md:27-30
 * without ever materialising x and w. */
int fgq_gausslegendre_moments_fused_device(int64_t n, int64_t k_lo, int64_t k_hi, int with_exp,
                                           double* sums_dev, void* stream);

/* The same reduction for a USER integrand: `expr` is f as an expression in x in CUDA C++ syntax ("exp(x) * cos(3 * x)",
 * "1 / (1 + 25 * x * x)", "pow(fabs(x), 2.5)" ...).  It is compiled once per (expression, series branch) at run time
 * (NVRTC, sm_100a, --fmad=false) together with the library's node-generation source, cached, and launched:
 *   fgq_gausslegendre_integrate_device  generate-and-reduce over half-indices [k_lo, k_hi) of the n-point rule and
 *                                       their mirror images (n > 60): nodes and weights are never stored;
 *   fgq_rule_integrate_device           dot(w, f.(x)) over a stored rule of any family (count entries, device buffers).
 * sums_dev[0] += sum w f(x), sums_dev[1] += sum w (2 doubles, zeroed by the caller; all-reduce across ranks).
 * This is dot(w, f.(x)) of README.md:27-30.  A malformed expression -> FGQ_EINVAL with the compiler's diagnostic in
 * fgq_last_error(); no NVRTC on the box -> FGQ_ENOTIMPL. */
int fgq_gausslegendre_integrate_device(int64_t n, int64_t k_lo, int64_t k_hi, const char* expr, double* sums_dev,
                                       void* stream);
int fgq_rule_integrate_device(const double* x_dev, const double* w_dev, int64_t count, const char* expr,
                              double* sums_dev, void* stream);

/* ---- roofline denominators (measured on the current device) ------------------------------ */

/* DFMA issue-rate microbenchmark: independent FMA chains on every SM.  *tflops = 2*FMA/s/1e12. */
int fgq_measure_dfma_peak(double* tflops);
/* Store-only streaming bandwidth (16-byte st.global.cs over `bytes` of scratch), GB/s. */
int fgq_measure_store_peak(int64_t bytes, double* gbs);

/* Host-memory store bandwidth, GB/s: `threads` host threads (0 = the library's default) streaming non-temporal
 * stores into buf[0 .. count) — the pattern with which the host-returning calls write their result.  The e2e
 * figures of bench.py are stated against it.  Needs no device. */
int fgq_measure_host_store_peak(double* buf, int64_t count, int threads, double* gbs);

/* ---- test hooks ---------------------------------------------------------------------------- */

/* theta_k (the angle whose cosine is the node) for half-indices [k_lo, k_hi), asy path. */
int fgq_debug_legendre_theta_device(int64_t n, int64_t k_lo, int64_t k_hi, double* theta_dev,
                                    void* stream);
/* *mismatches_dev += number of half-indices in [k_lo, k_hi) (k_lo >= 13192, n >= 2^25) whose
 * node or weight differs between the tail kernel's slow-path-free reciprocal/quotient and the
 * IEEE operators.  Expected 0. */
int fgq_debug_fastdiv_check(int64_t n, int64_t k_lo, int64_t k_hi,
                            unsigned long long* mismatches_dev, void* stream);
/* Counters of the threaded device->host drains since the previous call of this function: pieces handled,
 * pieces whose mirror image travelled by DMA instead of being written by a host thread, pieces that had
 * already landed when a thread picked them up, and the last adaptive DMA share x 1000 (bench.py reports
 * them beside e2e).  Any pointer may be NULL. */
int fgq_debug_drain_stats(int64_t* pieces, int64_t* dma_pieces, int64_t* landed, int* share_milli);
/* Transport coding of large host-returning Legendre calls (csrc/delta_codec.h): bytes that crossed PCIe in
 * coded form and the number of values they carried, since the previous call of this function. */
int fgq_debug_codec_stats(int64_t* bytes, int64_t* values);
/* decode(encode(in)) on the host with the reference encoder of csrc/delta_codec.h (tests; no GPU needed):
 * out must equal in bit for bit for ANY content; *stream_bytes receives the coded size. */
int fgq_debug_codec_roundtrip_host(const double* in, int64_t cnt, double* out, int64_t* stream_bytes);
/* Angles beyond which the Jacobi truncation-norm passes skip a node (every candidate series term is provably below
 * a quarter of the reference's eps/100 threshold there), for the right half (a, b) and the mirrored left half (b, a);
 * +inf = no skipping.  Host-only; tests check the bound against the oracle's series terms. */
int fgq_debug_jacobi_skip_angles(int64_t n, double a, double b, double* t_right, double* t_left);
/* The tiles (128 nodes, counted within a half) the truncation-norm passes of gaussjacobi(n, a, b) walk — those that can
 * hold a node below the skip angle: out = {n_left (split index), tile_lo, tile_hi of the right half, of the left half}
 * (host only; tests). */
int fgq_debug_jacobi_norm_tiles(int64_t n, double a, double b, int64_t* out);
/* The host half of a coded transfer without a device (tests): codes `in` with the reference encoder and unpacks it
 * through the routine the pipeline's host threads run: left[e] = in[e], right[cnt-1-e] = +-in[e] for e < mirror_cnt.
 * fused != 0 selects the register-to-destination variant (FGQ_CODEC_FUSED=1 in the product; off by default). */
int fgq_debug_codec_unpack_host(const double* in, int64_t cnt, double* left, double* right, int64_t mirror_cnt,
                                int negate, int fused);
/* The half-index range [k_lo, k_hi) that device / rank `rank` of `world` owns in the sharded calls (host-only). */
int fgq_debug_half_index_range(int64_t n, int world, int rank, int64_t* k_lo, int64_t* k_hi);
/* Launch buckets of fgq_gausslaguerre_batch_* for a list of rule sizes (host only; tests): rules sorted by n,
 * longest first, are cut into groups of similar n; returns the number of buckets (first/count index the sorted list). */
int fgq_debug_laguerre_batch_buckets(int64_t nrules, const int32_t* n_i, int64_t* first, int64_t* count,
                                     int32_t* n_top, int max_buckets);
/* The in-order background planner of the batch calls (host only; tests): n positions are processed on the library's
 * host threads while the caller waits for growing prefixes (wait_step positions at a time).  Returns the number of
 * violations — a position below a waited-for bound not yet processed, or a position not processed exactly once. */
int64_t fgq_debug_ordered_planner(int64_t n, int64_t wait_step);
/* number of kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t fgq_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FGQ_H_ */
