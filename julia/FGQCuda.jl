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

# Device-resident shard (with CUDA.jl): x_dev, w_dev are CuVector{Float64} of length hi-lo;
# asynchronous on `stream`; output indices are 0-based here, [lo, hi).
function gausslegendre_device!(x_dev, w_dev, n::Integer, lo::Integer, hi::Integer, stream::Ptr{Cvoid} = C_NULL)
    check(ccall((:fgq_gausslegendre_device, libfgq), Cint,
        (Int64, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
        n, lo, hi, pointer(x_dev), pointer(w_dev), stream), n)
    return x_dev, w_dev
end

end # module
