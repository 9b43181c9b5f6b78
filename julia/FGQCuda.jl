# FGQCuda.jl — `ccall` binding of libfgq_cuda.so for FastGaussQuadrature.jl (NOT executed in this
# repository's image: Julia is not installed here).  Every function keeps the reference's name,
# signature, return type (a 2-tuple of caller-owned Vector{Float64}) and exceptions, so a
# maintainer can forward the Float64 methods of src/gauss*.jl to it:
#
#     gausslegendre(n::Integer) = FGQCuda.gausslegendre(n)        # src/gausslegendre.jl:69
#
module FGQCuda

const libfgq = get(ENV, "FGQ_CUDA_LIB", "libfgq_cuda.so")

const FGQ_OK, FGQ_EDOMAIN, FGQ_ENOTIMPL, FGQ_EINVAL = 0, 1, 2, 3

last_error() = unsafe_string(ccall((:fgq_last_error, libfgq), Cstring, ()))

function check(rc::Cint, arg)
    rc == FGQ_OK && return nothing
    msg = last_error()
    rc == FGQ_EDOMAIN && throw(DomainError(arg, msg))   # e.g. gausslegendre.jl:26
    rc == FGQ_ENOTIMPL && error(msg)                    # e.g. gaussjacobi.jl:53
    rc == FGQ_EINVAL && throw(ArgumentError(msg))
    error("libfgq_cuda: rc=$rc: $msg")                  # CUDA failure; there is no CPU fallback
end

# Page-locked result vectors (fgq_host_alloc): `x = pinned_vector(n)` is an ordinary Vector{Float64} whose memory
# the *_host calls fill by DMA, without staging and without the first-touch page faults of `Vector{Float64}(undef, n)`
# (0.3 s per 16 GB).  Meant to be reused: `gausslegendre!(x, w)`.
function pinned_vector(n::Integer)
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:fgq_host_alloc, libfgq), Cint, (Int64, Ptr{Ptr{Cvoid}}), 8 * n, p), n)
    v = unsafe_wrap(Vector{Float64}, Ptr{Float64}(p[]), n; own = false)
    n > 0 && finalizer(_ -> ccall((:fgq_host_free, libfgq), Cint, (Ptr{Cvoid},), p[]), v)
    return v
end

# in-place form: x, w are caller-owned (e.g. pinned_vector) and of the rule's length
function gausslegendre!(x::Vector{Float64}, w::Vector{Float64})
    n = length(x)
    length(w) == n || throw(ArgumentError("x and w must have the same length"))
    n == 0 && return x, w
    GC.@preserve x w check(ccall((:fgq_gausslegendre_host, libfgq), Cint,
        (Int64, Ptr{Float64}, Ptr{Float64}), n, x, w), n)
    return x, w
end

# gausslegendre(n) — src/gausslegendre.jl:22-69
function gausslegendre(n::Integer)
    n < 0 && throw(DomainError(n, "Input n must be a non-negative integer"))
    x = Vector{Float64}(undef, n); w = Vector{Float64}(undef, n)
    n == 0 && return x, w
    GC.@preserve x w check(ccall((:fgq_gausslegendre_host, libfgq), Cint,
        (Int64, Ptr{Float64}, Ptr{Float64}), n, x, w), n)
    return x, w
end
gausslegendre(::Type{Float64}, n::Integer) = gausslegendre(n)

# Batch of small rules, CSR layout (one warp per rule on the device)
function gausslegendre_batch(ns::Vector{Int32})
    offsets = Vector{Int64}(undef, length(ns) + 1); offsets[1] = 0
    cumsum!(view(offsets, 2:length(offsets)), Int64.(ns))
    x = Vector{Float64}(undef, offsets[end]); w = similar(x)
    GC.@preserve ns offsets x w check(ccall((:fgq_gausslegendre_batch_host, libfgq), Cint,
        (Int64, Ptr{Int32}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}),
        length(ns), ns, offsets, x, w), length(ns))
    return offsets, x, w
end

# Batches of small rules of the other families (same CSR layout; n_i <= 100 / 200 / 127)
function _offsets(ns::Vector{Int32})
    offsets = Vector{Int64}(undef, length(ns) + 1); offsets[1] = 0
    cumsum!(view(offsets, 2:length(offsets)), Int64.(ns))
    return offsets
end
function gaussjacobi_batch(ns::Vector{Int32}, α::Vector{Float64}, β::Vector{Float64})
    offsets = _offsets(ns); x = Vector{Float64}(undef, offsets[end]); w = similar(x)
    GC.@preserve ns α β offsets x w check(ccall((:fgq_gaussjacobi_batch_host, libfgq), Cint,
        (Int64, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}),
        length(ns), ns, α, β, offsets, x, w), length(ns))
    return offsets, x, w
end
function gausshermite_batch(ns::Vector{Int32}; normalize::Bool=false)
    offsets = _offsets(ns); x = Vector{Float64}(undef, offsets[end]); w = similar(x)
    GC.@preserve ns offsets x w check(ccall((:fgq_gausshermite_batch_host, libfgq), Cint,
        (Int64, Ptr{Int32}, Ptr{Int64}, Cint, Ptr{Float64}, Ptr{Float64}),
        length(ns), ns, offsets, normalize, x, w), length(ns))
    return offsets, x, w
end
function gausslaguerre_batch(ns::Vector{Int32}, α::Vector{Float64})
    offsets = _offsets(ns); x = Vector{Float64}(undef, offsets[end]); w = similar(x)
    GC.@preserve ns α offsets x w check(ccall((:fgq_gausslaguerre_batch_host, libfgq), Cint,
        (Int64, Ptr{Int32}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}),
        length(ns), ns, α, offsets, x, w), length(ns))
    return offsets, x, w
end

# Truncated Gauss-Hermite rule: (first, x, w) with x, w covering positions first+1 : first+length(x) (1-based) of
# gausshermite(n); every omitted weight is 0.0 in the full rule.  The entries agree with the full rule to the Newton
# tolerance (the stopping rule is taken over the computed block), bit for bit only when nothing is truncated.
function gausshermite_reduced(n::Integer; normalize::Bool=false)
    n < 0 && throw(DomainError(n, "Input n must be a non-negative integer"))
    bound = ccall((:fgq_gausshermite_reduced_bound, libfgq), Int64, (Int64,), n)
    x = Vector{Float64}(undef, bound); w = similar(x)
    first = Ref{Int64}(0); n_out = Ref{Int64}(0)
    n == 0 && return 0, x, w
    GC.@preserve x w check(ccall((:fgq_gausshermite_reduced_host, libfgq), Cint,
        (Int64, Cint, Ptr{Float64}, Ptr{Float64}, Ref{Int64}, Ref{Int64}), n, normalize, x, w, first, n_out), n)
    resize!(x, n_out[]); resize!(w, n_out[])
    return first[], x, w
end

# gaussjacobi(n, α, β) — src/gaussjacobi.jl:23-55 (asymptotic branch n > 100, max(α,β) < 5; (0,0) and
# (±1/2, ±1/2) are routed to the Legendre / Chebyshev kernels by the library, as the reference does)
function gaussjacobi(n::Integer, α::Real, β::Real)
    n < 0 && throw(DomainError(n, "gaussjacobi($n,$α,$β) not defined: n must be non-negative."))
    x = Vector{Float64}(undef, n); w = Vector{Float64}(undef, n)
    n == 0 && return x, w
    GC.@preserve x w check(ccall((:fgq_gaussjacobi_host, libfgq), Cint,
        (Int64, Float64, Float64, Ptr{Float64}, Ptr{Float64}), n, Float64(α), Float64(β), x, w), (α, β))
    return x, w
end

# gaussradau(n) — src/gaussradau.jl:31-49;  gausslobatto(n) — src/gausslobatto.jl:31-52
for (f, sym) in ((:gaussradau, :fgq_gaussradau_host), (:gausslobatto, :fgq_gausslobatto_host))
    @eval function $f(n::Integer)
        x = Vector{Float64}(undef, max(n, 0)); w = similar(x)
        GC.@preserve x w check(ccall(($(QuoteNode(sym)), libfgq), Cint,
            (Int64, Ptr{Float64}, Ptr{Float64}), n, x, w), n)
        return x, w
    end
end

# gaussradau(n, α, β) — src/gaussradau.jl:52-69
function gaussradau(n::Integer, α::Real, β::Real)
    n ≤ 0 && throw(DomainError(n, "Input N must be a positive integer"))
    x = Vector{Float64}(undef, n); w = similar(x)
    GC.@preserve x w check(ccall((:fgq_gaussradau_jacobi_host, libfgq), Cint,
        (Int64, Float64, Float64, Ptr{Float64}, Ptr{Float64}), n, Float64(α), Float64(β), x, w), n)
    return x, w
end

# gausschebyshevt/u/v/w(n) — src/gausschebyshev.jl:22-113
for (f, kind) in ((:gausschebyshevt, 1), (:gausschebyshevu, 2), (:gausschebyshevv, 3), (:gausschebyshevw, 4))
    @eval function $f(n::Integer)
        n < 0 && throw(DomainError(n, "Input n must be a non-negative integer"))
        x = Vector{Float64}(undef, n); w = similar(x)
        n == 0 && return x, w
        GC.@preserve x w check(ccall((:fgq_gausschebyshev_host, libfgq), Cint,
            (Cint, Int64, Ptr{Float64}, Ptr{Float64}), $kind, n, x, w), n)
        return x, w
    end
end

# deprecated dispatcher gausschebyshev(n, kind) — src/gausschebyshev.jl:117-144
function gausschebyshev(n::Integer, kind::Integer = 1)
    n < 0 && throw(DomainError(n, "Input n must be a non-negative integer"))
    1 <= kind <= 4 || throw(ArgumentError("Chebyshev kind should be 1, 2, 3, or 4"))
    f = (gausschebyshevt, gausschebyshevu, gausschebyshevv, gausschebyshevw)[kind]
    Base.depwarn("`gausschebyshev(n, $kind)` is deprecated and will be removed in the next breaking release. Please use `$(nameof(f))(n)` instead.", :gausschebyshev)
    return f(n)
end

# approx_besselroots(ν, n) — src/besselroots.jl:23-67; besselroots is its deprecated alias (:247)
function approx_besselroots(ν::Real, n::Integer)
    n < 0 && throw(DomainError(n, "Input n must be a non-negative integer"))
    x = Vector{Float64}(undef, n)
    n == 0 && return x
    GC.@preserve x check(ccall((:fgq_approx_besselroots, libfgq), Cint,
        (Float64, Int64, Ptr{Float64}), Float64(ν), n, x), n)
    return x
end
Base.@deprecate besselroots(ν::Real, n::Integer) approx_besselroots(ν, n)

# gausshermite(n; normalize) — src/gausshermite.jl:28-33 (asymptotic branch n > 200)
function gausshermite(n::Integer; normalize = false)
    n < 0 && throw(DomainError(n, "Input n must be a non-negative integer"))
    x = Vector{Float64}(undef, n); w = similar(x)
    n == 0 && return x, w
    GC.@preserve x w check(ccall((:fgq_gausshermite_host, libfgq), Cint,
        (Int64, Cint, Ptr{Float64}, Ptr{Float64}), n, normalize ? 1 : 0, x, w), n)
    return x, w
end

# gausslaguerre(n, α; reduced) — src/gausslaguerre.jl:59-96 (explicit asymptotics, n >= 128, α < 5)
function gausslaguerre(n::Integer, α::Real = 0.0; reduced = false)
    α ≤ -1 && throw(DomainError(α, "The Laguerre parameter α ≤ -1 corresponds to a nonintegrable weight function"))
    n < 0 && throw(DomainError(n, "gausslaguerre($n,$α) not defined: n must be positive."))
    x = Vector{Float64}(undef, n); w = similar(x)
    n == 0 && return x, w
    nout = Ref{Int64}(0)
    GC.@preserve x w check(ccall((:fgq_gausslaguerre_host, libfgq), Cint,
        (Int64, Float64, Cint, Ptr{Float64}, Ptr{Float64}, Ref{Int64}), n, Float64(α), reduced ? 1 : 0, x, w, nout), n)
    reduced && (resize!(x, nout[]); resize!(w, nout[]))
    return x, w
end

# ---- type-first methods of the reference (Float64 only: there is no CPU fallback and no generic-T kernel) ----
# gausshermite(::Type{T}, n; normalize) src/gausshermite.jl:28; gausslaguerre(::Type{T}, n) src/gausslaguerre.jl:1;
# gaussradau(::Type{T}, n) src/gaussradau.jl:31; gausslobatto(::Type{T}, n) src/gausslobatto.jl:31;
# gausschebyshev{t,u,v,w}(::Type{T}, n) src/gausschebyshev.jl:22,51,80,109
_only_f64(::Type{Float64}) = nothing
_only_f64(::Type{T}) where {T} = error("libfgq_cuda serves Float64 only (requested $T); use FastGaussQuadrature's generic method")
gausshermite(::Type{T}, n::Integer; normalize = false) where {T<:AbstractFloat} = (_only_f64(T); gausshermite(n; normalize = normalize))
gausslaguerre(::Type{T}, n::Integer) where {T<:AbstractFloat} = (_only_f64(T); gausslaguerre(n))
gaussradau(::Type{T}, n::Integer) where {T<:AbstractFloat} = (_only_f64(T); gaussradau(n))
gausslobatto(::Type{T}, n::Integer) where {T<:AbstractFloat} = (_only_f64(T); gausslobatto(n))
gausschebyshevt(::Type{T}, n::Integer) where {T<:AbstractFloat} = (_only_f64(T); gausschebyshevt(n))
gausschebyshevu(::Type{T}, n::Integer) where {T<:AbstractFloat} = (_only_f64(T); gausschebyshevu(n))
gausschebyshevv(::Type{T}, n::Integer) where {T<:AbstractFloat} = (_only_f64(T); gausschebyshevv(n))
gausschebyshevw(::Type{T}, n::Integer) where {T<:AbstractFloat} = (_only_f64(T); gausschebyshevw(n))

# ---- multi-GPU from ONE Julia process (include/fgq.h "multi-GPU, one process") ---------------------------------
# `fgq_shard` of include/fgq.h; x, w are device pointers on `device` (wrap them with CUDA.jl's unsafe_wrap to use them)
struct FGQShard
    device::Int32
    owner::Int32
    lo::Int64
    hi::Int64
    x::Ptr{Float64}
    w::Ptr{Float64}
end
FGQShard() = FGQShard(0, 0, 0, 0, C_NULL, C_NULL)

# fgq_init: devices = nothing -> all visible devices
function init(devices::Union{Nothing,Vector{Int32}} = nothing)
    rc = devices === nothing ?
        ccall((:fgq_init, libfgq), Cint, (Ptr{Cint}, Cint), C_NULL, 0) :
        ccall((:fgq_init, libfgq), Cint, (Ptr{Cint}, Cint), devices, length(devices))
    check(rc, devices)
    return Int(ccall((:fgq_num_devices, libfgq), Cint, ()))
end
finalize() = check(ccall((:fgq_finalize, libfgq), Cint, ()), nothing)
device_ordinal(i::Integer) = Int(ccall((:fgq_device_ordinal, libfgq), Cint, (Cint,), i))   # 0-based position -> CUDA ordinal
synchronize() = check(ccall((:fgq_synchronize, libfgq), Cint, ()), nothing)

# Device-resident gausslegendre(n) over every initialised GPU: returns (shards, (sum w, sum w x^2)).  Pass the
# shards of a previous call to refill them without allocation; free them with free_shards!.
function gausslegendre_sharded(n::Integer, shards::Vector{FGQShard} = FGQShard[]; sums::Bool = true)
    n < 0 && throw(DomainError(n, "Input n must be a non-negative integer"))
    cap = Int(ccall((:fgq_shard_count, libfgq), Cint, (Int64,), n))
    isempty(shards) && (shards = [FGQShard() for _ in 1:max(cap, 1)])
    cnt = Ref{Cint}(0); s = zeros(Float64, 2)
    GC.@preserve shards s check(ccall((:fgq_gausslegendre_device_sharded, libfgq), Cint,
        (Int64, Ptr{FGQShard}, Cint, Ref{Cint}, Ptr{Float64}), n, shards, length(shards), cnt, sums ? pointer(s) : C_NULL), n)
    resize!(shards, cnt[])
    return shards, (s[1], s[2])
end
free_shards!(shards::Vector{FGQShard}) =
    (GC.@preserve shards check(ccall((:fgq_free_shards, libfgq), Cint, (Ptr{FGQShard}, Cint), shards, length(shards)), nothing); empty!(shards))

# Optional replicated copy on every device: vectors of device pointers (one per device), n entries each
function allgather(shards::Vector{FGQShard}, n::Integer)
    nd = Int(ccall((:fgq_num_devices, libfgq), Cint, ()))
    xr = fill(Ptr{Float64}(C_NULL), nd); wr = fill(Ptr{Float64}(C_NULL), nd)
    GC.@preserve shards xr wr check(ccall((:fgq_allgather, libfgq), Cint,
        (Ptr{FGQShard}, Cint, Int64, Ptr{Ptr{Float64}}, Ptr{Ptr{Float64}}), shards, length(shards), n, xr, wr), n)
    return xr, wr
end
function gausslegendre_replicated(n::Integer)
    nd = Int(ccall((:fgq_num_devices, libfgq), Cint, ()))
    xr = fill(Ptr{Float64}(C_NULL), nd); wr = fill(Ptr{Float64}(C_NULL), nd)
    GC.@preserve xr wr check(ccall((:fgq_gausslegendre_device_replicated, libfgq), Cint,
        (Int64, Ptr{Ptr{Float64}}, Ptr{Ptr{Float64}}), n, xr, wr), n)
    return xr, wr
end
free_replicas!(xr::Vector{Ptr{Float64}}, wr::Vector{Ptr{Float64}}) =
    GC.@preserve xr wr check(ccall((:fgq_free_replicas, libfgq), Cint, (Ptr{Ptr{Float64}}, Ptr{Ptr{Float64}}), xr, wr), nothing)

# Host-returning gausslegendre(n) drained through all GPUs at once (same result as gausslegendre(n), bit for bit)
function gausslegendre_multi(n::Integer)
    n < 0 && throw(DomainError(n, "Input n must be a non-negative integer"))
    x = Vector{Float64}(undef, n); w = Vector{Float64}(undef, n)
    n == 0 && return x, w
    GC.@preserve x w check(ccall((:fgq_gausslegendre_host_multi, libfgq), Cint,
        (Int64, Ptr{Float64}, Ptr{Float64}), n, x, w), n)
    return x, w
end

# Device-side timing of what was enqueued between the two calls (slowest device, milliseconds)
timer_start() = check(ccall((:fgq_timer_start, libfgq), Cint, ()), nothing)
function timer_stop()
    ms = Ref{Cfloat}(0)
    check(ccall((:fgq_timer_stop, libfgq), Cint, (Ref{Cfloat}, Ptr{Cfloat}), ms, C_NULL), nothing)
    return ms[]
end

# Device-resident shard (with CUDA.jl): x_dev, w_dev are CuVector{Float64} of length hi-lo;
# asynchronous on `stream`; output indices are 0-based here, [lo, hi).
function gausslegendre_device!(x_dev, w_dev, n::Integer, lo::Integer, hi::Integer, stream::Ptr{Cvoid} = C_NULL)
    check(ccall((:fgq_gausslegendre_device, libfgq), Cint,
        (Int64, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
        n, lo, hi, pointer(x_dev), pointer(w_dev), stream), n)
    return x_dev, w_dev
end

end # module
