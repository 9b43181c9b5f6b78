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

# Truncated Gauss-Hermite rule: (first, x, w) with x, w = entries first+1 : first+length(x) (1-based) of
# gausshermite(n); every omitted weight is 0.0 in the full rule
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

# Device-resident shard (with CUDA.jl): x_dev, w_dev are CuVector{Float64} of length hi-lo;
# asynchronous on `stream`; output indices are 0-based here, [lo, hi).
function gausslegendre_device!(x_dev, w_dev, n::Integer, lo::Integer, hi::Integer, stream::Ptr{Cvoid} = C_NULL)
    check(ccall((:fgq_gausslegendre_device, libfgq), Cint,
        (Int64, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
        n, lo, hi, pointer(x_dev), pointer(w_dev), stream), n)
    return x_dev, w_dev
end

end # module
