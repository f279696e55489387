# SGCB200.jl — Julia shim: StaggeredGridCalculus.jl's hot-path methods on device arrays,
# forwarded to libsgc_b200.so through `ccall`.
#
# How it plugs in (see INTEGRATION.md): `include` this file at the end of
# src/StaggeredGridCalculus.jl (after line 17).  It adds METHODS to the reference's own
# generic functions — same names, same positional/keyword arguments — that are more specific
# than the reference's `AbsArrNumber` methods because the field arguments are `DeviceArray`s.
# The reference's wrappers (scalar-∆ `fill.`, `SVec` conversion; e.g.
# src/matrixfree/differential/curl.jl:12-49) keep working unchanged and dispatch lands here.
# `create_*` take no array argument to dispatch on; they are reached through
# `create_*(Val(:b200), args...; kwargs...)`.
#
# Dispatch: every method below repeats the argument types of the reference's CONCRETE method exactly
# (`SBool{K}`, `SNumber{K}`, `NTuple{K,AbsVecNumber}`, `SInt{K}` keywords) and narrows only the field arguments to
# `DeviceArray`, so it is strictly more specific than both the reference's concrete method and its
# `AbsVecBool` wrappers — no ambiguity (run `Test.detect_ambiguities(StaggeredGridCalculus)` once with a Julia at
# hand).  Shapes are validated here, like the reference's `@assert`s, because the C ABI carries no array lengths.
#
# NOTE: there is no Julia runtime in the build image, so this file has never been executed; it
# is kept deliberately thin — every decision (validation, promotion, kernel choice) lives in the
# C library, which is exercised through the identical `ctypes` binding by the test-suite.
#
# There is no CPU fallback: if the library or a GPU is missing every call throws.

module SGCB200

using SparseArrays
using AbbreviatedTypes   # SBool, SNumber, SInt, AbsVecNumber, … (the reference's own dependency)
import ..StaggeredGridCalculus: apply_∂!, apply_m̂!, apply_curl!, apply_divg!, apply_grad!, apply_mean!,
                                create_∂, create_m̂, create_mean, create_curl, create_divg, create_grad, create_πcmp

const LIB = get(ENV, "SGC_B200_LIB", joinpath(@__DIR__, "..", "csrc", "libsgc_b200.so"))

const SGC_F64, SGC_C128 = Cint(0), Cint(1)
const SGC_OP_SET, SGC_OP_ADD = Cint(0), Cint(1)

struct SGCError <: Exception
    code::Cint
    msg::String
end

function check(status::Cint)
    status == 0 && return nothing
    msg = unsafe_string(ccall((:sgc_last_error, LIB), Cstring, ()))
    status == 3 && throw(InexactError(:apply, Float64, msg))  # complex α/e/∆ with a Float64 field
    status == 1 && throw(ArgumentError(msg))                   # the reference's @assert / @error
    throw(SGCError(status, msg))
end

dtype_code(::Type{Float64}) = SGC_F64
dtype_code(::Type{ComplexF64}) = SGC_C128
opcode(::Val{:(=)}) = SGC_OP_SET
opcode(::Val{:(+=)}) = SGC_OP_ADD

# ---------------------------------------------------------------------------------------------
# DeviceArray: a column-major Julia array living in HBM
# ---------------------------------------------------------------------------------------------
mutable struct DeviceArray{T<:Union{Float64,ComplexF64},N} <: AbstractArray{T,N}
    ptr::Ptr{Cvoid}
    dims::NTuple{N,Int}
    owner::Any   # parent DeviceArray for views (keeps the allocation alive), `nothing` for owners
end

function DeviceArray{T}(::UndefInitializer, dims::Vararg{Int,N}) where {T,N}
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:sgc_malloc, LIB), Cint, (Ptr{Ptr{Cvoid}}, Csize_t), p, sizeof(T) * prod(dims)))
    a = DeviceArray{T,N}(p[], dims, nothing)
    finalizer(x -> ccall((:sgc_free, LIB), Cint, (Ptr{Cvoid},), x.ptr), a)
    return a
end

function DeviceArray(h::Array{T,N}) where {T,N}
    d = DeviceArray{T}(undef, size(h)...)
    copyto!(d, h)
    return d
end

Base.size(a::DeviceArray) = a.dims
Base.similar(a::DeviceArray{T}, dims::Dims) where {T} = DeviceArray{T}(undef, dims...)
Base.getindex(::DeviceArray, I...) = error("scalar indexing of a DeviceArray is not supported; copy to the host with Array(a)")

function Base.copyto!(d::DeviceArray{T}, h::Array{T}) where {T}
    @assert length(d) == length(h)
    check(ccall((:sgc_memcpy_h2d, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}), d.ptr, h, sizeof(h), C_NULL))
    check(ccall((:sgc_stream_sync, LIB), Cint, (Ptr{Cvoid},), C_NULL))
    return d
end

function Base.Array(d::DeviceArray{T,N}) where {T,N}
    h = Array{T,N}(undef, d.dims)
    check(ccall((:sgc_memcpy_d2h, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}), h, d.ptr, sizeof(h), C_NULL))
    check(ccall((:sgc_stream_sync, LIB), Cint, (Ptr{Cvoid},), C_NULL))
    return h
end

# `selectdim(F, K+1, w)` — the component slices the reference takes (curl.jl:68, :88) are
# contiguous because the component index is the last (slowest) dimension.
function Base.selectdim(a::DeviceArray{T,N}, d::Integer, w::Integer) where {T,N}
    d == N || error("only slices of the last dimension are contiguous")
    m = prod(a.dims[1:N-1])
    return DeviceArray{T,N-1}(a.ptr + sizeof(T) * m * (w - 1), a.dims[1:N-1], a)
end

# ---------------------------------------------------------------------------------------------
# argument marshalling
# ---------------------------------------------------------------------------------------------
pair(x::Number) = (z = ComplexF64(x); Cdouble[real(z), imag(z)])
pairs(xs) = (v = Cdouble[]; for x in xs; z = ComplexF64(x); push!(v, real(z), imag(z)); end; v)
cints(xs) = Cint[Cint(x) for x in xs]

# promote a tuple of per-axis vectors to one element type, as the concrete methods'
# promote_type does (matrix/differential/curl.jl:59)
function coefs(vs)
    T = any(v -> eltype(v) <: Complex, vs) ? ComplexF64 : Float64
    keep = [Vector{T}(v) for v in vs]
    return keep, Ptr{Cvoid}[pointer(k) for k in keep], Cint(T == ComplexF64)
end

# ---------------------------------------------------------------------------------------------
# matrix-free operators
# ---------------------------------------------------------------------------------------------
# apply_∂!  — src/matrixfree/differential/partial.jl:33 (3-D), :241 (2-D), :388 (1-D)
function apply_∂!(Gv::DeviceArray{T,K}, Fu::DeviceArray{T,K}, ::Val{OP}, nw::Integer, isfwd::Bool,
                  ∆w⁻¹::AbstractVector{<:Number}, isbloch::Bool, e⁻ⁱᵏᴸ::Number=1.0; α::Number=1.0) where {T,K,OP}
    size(Gv) == size(Fu) || throw(ArgumentError("size(Gv) != size(Fu)"))
    1 ≤ nw ≤ K || throw(ArgumentError("nw = $nw out of range 1:$K"))
    length(∆w⁻¹) == size(Fu, nw) || throw(ArgumentError("length(∆w⁻¹) = $(length(∆w⁻¹)) != size(Fu,$nw) = $(size(Fu, nw))"))
    keep, ptrs, cc = coefs((∆w⁻¹,))
    N = Int64[size(Fu)...]
    GC.@preserve keep check(ccall((:sgc_apply_partial, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Int64}, Cint, Cint, Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cvoid}),
        Gv.ptr, Fu.ptr, opcode(Val(OP)), dtype_code(T), K, N, nw, isfwd, ptrs[1], cc, isbloch, pair(e⁻ⁱᵏᴸ), pair(α), C_NULL))
    return nothing
end

# apply_m̂!  — src/matrixfree/mean.jl:102/295/432 (arithmetic), :493/705/856 (weighted)
function apply_m̂!(Gv::DeviceArray{T,K}, Fu::DeviceArray{T,K}, ::Val{OP}, nw::Integer, isfwd::Bool, isbloch::Bool,
                  e⁻ⁱᵏᴸ::Number=1.0; α::Number=1.0) where {T,K,OP}
    size(Gv) == size(Fu) || throw(ArgumentError("size(Gv) != size(Fu)"))
    1 ≤ nw ≤ K || throw(ArgumentError("nw = $nw out of range 1:$K"))
    N = Int64[size(Fu)...]
    check(ccall((:sgc_apply_mhat, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Int64}, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cvoid}),
        Gv.ptr, Fu.ptr, opcode(Val(OP)), dtype_code(T), K, N, nw, isfwd, C_NULL, C_NULL, 0, isbloch, pair(e⁻ⁱᵏᴸ), pair(α), C_NULL))
    return nothing
end

function apply_m̂!(Gv::DeviceArray{T,K}, Fu::DeviceArray{T,K}, ::Val{OP}, nw::Integer, isfwd::Bool,
                  ∆w::AbstractVector{<:Number}, ∆w′⁻¹::AbstractVector{<:Number}, isbloch::Bool, e⁻ⁱᵏᴸ::Number=1.0;
                  α::Number=1.0) where {T,K,OP}
    size(Gv) == size(Fu) || throw(ArgumentError("size(Gv) != size(Fu)"))
    1 ≤ nw ≤ K || throw(ArgumentError("nw = $nw out of range 1:$K"))
    (length(∆w) == size(Fu, nw) && length(∆w′⁻¹) == size(Fu, nw)) || throw(ArgumentError("∆w / ∆w′⁻¹ must have size(Fu,$nw) = $(size(Fu, nw)) entries"))
    keep, ptrs, cc = coefs((∆w, ∆w′⁻¹))
    N = Int64[size(Fu)...]
    GC.@preserve keep check(ccall((:sgc_apply_mhat, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Int64}, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cvoid}),
        Gv.ptr, Fu.ptr, opcode(Val(OP)), dtype_code(T), K, N, nw, isfwd, ptrs[1], ptrs[2], cc, isbloch, pair(e⁻ⁱᵏᴸ), pair(α), C_NULL))
    return nothing
end

# apply_curl!  — src/matrixfree/differential/curl.jl:52-63 (concrete method; the wrappers :12-49
# convert to SVec and land here).  ONE fused kernel for the full 3-D curl.
function apply_curl!(G::DeviceArray{T,K₊₁}, F::DeviceArray{T,K₊₁}, ::Val{OP}, isfwd::SBool{K}, ∆l⁻¹::NTuple{K,AbsVecNumber},
                     isbloch::SBool{K}, e⁻ⁱᵏᴸ::SNumber{K};
                     cmp_shp::SInt{K}=SVec(ntuple(identity, Val(K))),
                     cmp_out::SInt{Kout}=SVec(ntuple(identity, size(G, K₊₁))),
                     cmp_in::SInt{Kin}=SVec(ntuple(identity, size(G, K₊₁))),
                     α::Number=1.0) where {T,K,K₊₁,Kout,Kin,OP}
    K₊₁ == K + 1 || throw(ArgumentError("K₊₁ != K+1"))
    check_grid(size(G)[1:K], ∆l⁻¹)
    size(F)[1:K] == size(G)[1:K] || throw(ArgumentError("G and F live on different grids: $(size(G)) vs $(size(F))"))
    (size(G, K₊₁) == Kout && size(F, K₊₁) == Kin) || throw(ArgumentError("component dimensions $(size(G, K₊₁)), $(size(F, K₊₁)) != length(cmp_out), length(cmp_in) = $Kout, $Kin"))
    keep, ptrs, cc = coefs(∆l⁻¹)
    N = Int64[size(G)[1:K]...]
    GC.@preserve keep check(ccall((:sgc_apply_curl, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Int64}, Ptr{Cint}, Ptr{Ptr{Cvoid}}, Cint, Ptr{Cint}, Ptr{Cdouble},
         Ptr{Cint}, Cint, Ptr{Cint}, Cint, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cvoid}),
        G.ptr, F.ptr, opcode(Val(OP)), dtype_code(T), K, N, cints(isfwd), ptrs, cc, cints(isbloch), pairs(e⁻ⁱᵏᴸ),
        cints(cmp_shp), length(cmp_out), cints(cmp_out), length(cmp_in), cints(cmp_in), pair(α), C_NULL))
    return nothing
end

# apply_curlcurl!  — G OP ∇₂×(∇₁×F): the pair of apply_curl! calls either side of the intermediate field
# (src/matrixfree/differential/curl.jl:52-137 twice; matrix form `Cv*Cu`, test/curl.jl:123-153) in ONE pass over
# memory; bit-identical to the two calls.  New entry point (the reference has no fused form).
function apply_curlcurl!(G::DeviceArray{T,4}, F::DeviceArray{T,4}, ::Val{OP}, isfwd₁, ∆l₁⁻¹::NTuple{3,AbstractVector{<:Number}},
                         isfwd₂, ∆l₂⁻¹::NTuple{3,AbstractVector{<:Number}}, isbloch, e⁻ⁱᵏᴸ; α₁::Number=1.0, α₂::Number=1.0) where {T,OP}
    N = Int64[size(G)[1:3]...]
    ops = Ptr{Cvoid}[C_NULL, C_NULL]
    for (q, (isfwd, ∆l⁻¹, α)) in enumerate(((isfwd₁, ∆l₁⁻¹, α₁), (isfwd₂, ∆l₂⁻¹, α₂)))
        keep, ptrs, cc = coefs(∆l⁻¹)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve keep check(ccall((:sgc_op_create_curl, LIB), Cint,
            (Ptr{Ptr{Cvoid}}, Cint, Cint, Ptr{Int64}, Ptr{Cint}, Ptr{Ptr{Cvoid}}, Cint, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cint}, Cint,
             Ptr{Cint}, Cint, Ptr{Cint}, Ptr{Cdouble}),
            h, dtype_code(T), 3, N, cints(isfwd), ptrs, cc, cints(isbloch), pairs(e⁻ⁱᵏᴸ), cints(1:3), 3, cints(1:3), 3, cints(1:3), pair(α)))
        ops[q] = h[]
    end
    try
        check(ccall((:sgc_op_apply_curlcurl, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
                    ops[2], ops[1], G.ptr, F.ptr, opcode(Val(OP)), C_NULL))
        check(ccall((:sgc_stream_sync, LIB), Cint, (Ptr{Cvoid},), C_NULL))
    finally
        for h in ops
            h == C_NULL || ccall((:sgc_op_destroy, LIB), Cint, (Ptr{Cvoid},), h)
        end
    end
    return nothing
end

# apply_divg! / apply_grad!  — src/matrixfree/differential/divergence.jl:41-49, gradient.jl:41-49 (concrete methods)
function apply_divg!(g::DeviceArray{T,K}, F::DeviceArray{T,K₊₁}, ::Val{OP}, isfwd::SBool{K}, ∆l⁻¹::NTuple{K,AbsVecNumber},
                     isbloch::SBool{K}, e⁻ⁱᵏᴸ::SNumber{K}; α::Number=1.0) where {T,K,K₊₁,OP}
    K₊₁ == K + 1 || throw(ArgumentError("K₊₁ != K+1"))
    check_grid(size(g), ∆l⁻¹)
    size(F) == (size(g)..., K) || throw(ArgumentError("F must be a $K-component field on the grid of g: $(size(F)) vs $(size(g))"))
    divgrad_call(:sgc_apply_divg, g, F, Val(OP), size(g), isfwd, ∆l⁻¹, isbloch, e⁻ⁱᵏᴸ, α)
end
function apply_grad!(G::DeviceArray{T,K₊₁}, f::DeviceArray{T,K}, ::Val{OP}, isfwd::SBool{K}, ∆l⁻¹::NTuple{K,AbsVecNumber},
                     isbloch::SBool{K}, e⁻ⁱᵏᴸ::SNumber{K}; α::Number=1.0) where {T,K,K₊₁,OP}
    K₊₁ == K + 1 || throw(ArgumentError("K₊₁ != K+1"))
    check_grid(size(f), ∆l⁻¹)
    size(G) == (size(f)..., K) || throw(ArgumentError("G must be a $K-component field on the grid of f: $(size(G)) vs $(size(f))"))
    divgrad_call(:sgc_apply_grad, G, f, Val(OP), size(f), isfwd, ∆l⁻¹, isbloch, e⁻ⁱᵏᴸ, α)
end
for sym in (:sgc_apply_divg, :sgc_apply_grad)
    @eval function divgrad_call(::Val{$(QuoteNode(sym))}, A::DeviceArray{T}, B::DeviceArray{T}, ::Val{OP}, N, isfwd, ∆l⁻¹, isbloch, e⁻ⁱᵏᴸ, α) where {T,OP}
        K = length(N)
        keep, ptrs, cc = coefs(∆l⁻¹)
        GC.@preserve keep check(ccall(($(QuoteNode(sym)), LIB), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Int64}, Ptr{Cint}, Ptr{Ptr{Cvoid}}, Cint, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cvoid}),
            A.ptr, B.ptr, opcode(Val(OP)), dtype_code(T), K, Int64[N...], cints(isfwd), ptrs, cc, cints(isbloch), pairs(e⁻ⁱᵏᴸ), pair(α), C_NULL))
        return nothing
    end
end
divgrad_call(sym::Symbol, args...) = divgrad_call(Val(sym), args...)

# every per-axis vector must have as many entries as the grid has points along its axis (the reference's @assert)
function check_grid(N, vs::Tuple)
    length(vs) == length(N) || throw(ArgumentError("$(length(vs)) per-axis vectors for a $(length(N))-dimensional grid"))
    for d in 1:length(N)
        length(vs[d]) == N[d] || throw(ArgumentError("per-axis vector $d has $(length(vs[d])) entries, the grid $(N[d]) points"))
    end
end

# apply_mean!  — src/matrixfree/mean.jl:47-54 (arithmetic), :69-78 (weighted)
function apply_mean!(G::DeviceArray{T,K₊₁}, F::DeviceArray{T,K₊₁}, ::Val{OP}, isfwd::SBool{K}, isbloch::SBool{K}, e⁻ⁱᵏᴸ::SNumber{K};
                     α::Number=1.0) where {T,K,K₊₁,OP}
    K₊₁ == K + 1 || throw(ArgumentError("K₊₁ != K+1"))
    (size(G) == size(F) && size(G, K₊₁) == K) || throw(ArgumentError("G and F must be $K-component fields on one grid: $(size(G)) vs $(size(F))"))
    N = Int64[size(G)[1:K]...]
    check(ccall((:sgc_apply_mean, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Int64}, Ptr{Cint}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Cint, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cvoid}),
        G.ptr, F.ptr, opcode(Val(OP)), dtype_code(T), K, N, cints(isfwd), C_NULL, C_NULL, 0, cints(isbloch), pairs(e⁻ⁱᵏᴸ), pair(α), C_NULL))
    return nothing
end

function apply_mean!(G::DeviceArray{T,K₊₁}, F::DeviceArray{T,K₊₁}, ::Val{OP}, isfwd::SBool{K}, ∆l::NTuple{K,AbsVecNumber},
                     ∆l′⁻¹::NTuple{K,AbsVecNumber}, isbloch::SBool{K}, e⁻ⁱᵏᴸ::SNumber{K}; α::Number=1.0) where {T,K,K₊₁,OP}
    K₊₁ == K + 1 || throw(ArgumentError("K₊₁ != K+1"))
    (size(G) == size(F) && size(G, K₊₁) == K) || throw(ArgumentError("G and F must be $K-component fields on one grid: $(size(G)) vs $(size(F))"))
    check_grid(size(G)[1:K], ∆l)
    check_grid(size(G)[1:K], ∆l′⁻¹)
    keep, ptrs, cc = coefs((∆l..., ∆l′⁻¹...))
    N = Int64[size(G)[1:K]...]
    GC.@preserve keep check(ccall((:sgc_apply_mean, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Int64}, Ptr{Cint}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Cint, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cvoid}),
        G.ptr, F.ptr, opcode(Val(OP)), dtype_code(T), K, N, cints(isfwd), ptrs[1:K], ptrs[K+1:2K], cc, cints(isbloch), pairs(e⁻ⁱᵏᴸ), pair(α), C_NULL))
    return nothing
end

# ---------------------------------------------------------------------------------------------
# host `Array`s: the strict drop-in calls (sgc_op_apply_host / sgc_op_apply_curlcurl_host).  The library page-locks
# the caller's arrays in place the first time it sees them and keeps the registration; `pin!` ties that
# registration to the array's lifetime (the finalizer unregisters before the GC frees the memory).
# ---------------------------------------------------------------------------------------------
const PINNED = IdDict{Any,Bool}()
function pin!(A::Array)
    haskey(PINNED, A) && return A
    check(ccall((:sgc_host_register, LIB), Cint, (Ptr{Cvoid}, Csize_t), A, sizeof(A)))
    PINNED[A] = true
    finalizer(a -> (ccall((:sgc_host_unregister, LIB), Cint, (Ptr{Cvoid},), a); delete!(PINNED, a)), A)
    return A
end

# G OP ∇₂×(∇₁×F) on host Arrays in one library call (z-chunked full-duplex pipeline inside the library)
function apply_curlcurl!(G::Array{T,4}, F::Array{T,4}, ::Val{OP}, isfwd₁, ∆l₁⁻¹::NTuple{3,AbsVecNumber}, isfwd₂, ∆l₂⁻¹::NTuple{3,AbsVecNumber},
                         isbloch, e⁻ⁱᵏᴸ; α₁::Number=1.0, α₂::Number=1.0) where {T<:Union{Float64,ComplexF64},OP}
    size(G) == size(F) && size(G, 4) == 3 || throw(ArgumentError("G and F must be 3-component fields on one 3-D grid"))
    check_grid(size(G)[1:3], ∆l₁⁻¹); check_grid(size(G)[1:3], ∆l₂⁻¹)
    pin!(G); pin!(F)
    N = Int64[size(G)[1:3]...]
    ops = Ptr{Cvoid}[C_NULL, C_NULL]
    try
        for (q, (isfwd, ∆l⁻¹, α)) in enumerate(((isfwd₁, ∆l₁⁻¹, α₁), (isfwd₂, ∆l₂⁻¹, α₂)))
            keep, ptrs, cc = coefs(∆l⁻¹)
            h = Ref{Ptr{Cvoid}}(C_NULL)
            GC.@preserve keep check(ccall((:sgc_op_create_curl, LIB), Cint,
                (Ptr{Ptr{Cvoid}}, Cint, Cint, Ptr{Int64}, Ptr{Cint}, Ptr{Ptr{Cvoid}}, Cint, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cint}, Cint,
                 Ptr{Cint}, Cint, Ptr{Cint}, Ptr{Cdouble}),
                h, dtype_code(T), 3, N, cints(isfwd), ptrs, cc, cints(isbloch), pairs(e⁻ⁱᵏᴸ), cints(1:3), 3, cints(1:3), 3, cints(1:3), pair(α)))
            ops[q] = h[]
        end
        GC.@preserve G F check(ccall((:sgc_op_apply_curlcurl_host, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint),
                                     ops[2], ops[1], G, F, opcode(Val(OP)), 0))
    finally
        for h in ops
            h == C_NULL || ccall((:sgc_op_destroy, LIB), Cint, (Ptr{Cvoid},), h)
        end
    end
    return nothing
end

# ---------------------------------------------------------------------------------------------
# several GPUs in one Julia process (SURVEY.md §8e): z-slabs in peer-addressable memory
#   comms = comm_init_local(ngpus)          one context per device, peer access enabled
#   A = peer_array(comms[r], T, dims...)    a DeviceArray on device r-1 that every device's kernels may read
#   peer_barrier(comms[r])                  device-side barrier of all ranks (queue it on every device)
# The slab operators take the neighbours' planes as `ghost` pointers of sgc_op_apply_range; see INTEGRATION.md.
# ---------------------------------------------------------------------------------------------
function comm_init_local(ngpus::Integer)
    hs = Vector{Ptr{Cvoid}}(undef, ngpus)
    check(ccall((:sgc_comm_init_local, LIB), Cint, (Ptr{Ptr{Cvoid}}, Cint), hs, ngpus))
    return hs
end
function peer_array(comm::Ptr{Cvoid}, ::Type{T}, dims::Vararg{Int,N}) where {T,N}
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:sgc_peer_alloc, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}), comm, sizeof(T) * prod(dims), p, C_NULL))
    a = DeviceArray{T,N}(p[], dims, nothing)
    finalizer(x -> ccall((:sgc_peer_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), comm, x.ptr), a)
    return a
end
peer_barrier(comm::Ptr{Cvoid}) = check(ccall((:sgc_peer_barrier, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), comm, C_NULL))
# device-side waits on a neighbour that gave up (0 in a correct run; synchronises the default stream)
function comm_timeouts(comm::Ptr{Cvoid})
    n = Ref{Int64}(0)
    check(ccall((:sgc_comm_timeouts, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}, Ptr{Cvoid}), comm, n, C_NULL))
    return n[]
end

# ---------------------------------------------------------------------------------------------
# sparse assembly: handle → shape → colptr (count + scan) → fill → SparseMatrixCSC
# ---------------------------------------------------------------------------------------------
function finish_csc(h::Ptr{Cvoid})
    try
        m, n, vt = Ref{Int64}(0), Ref{Int64}(0), Ref{Cint}(0)
        check(ccall((:sgc_asm_shape, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Cint}), h, m, n, vt))
        Tv = vt[] == SGC_C128 ? ComplexF64 : Float64
        nnz = Ref{Int64}(0)
        check(ccall((:sgc_asm_nnz_host, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}), h, nnz))
        colptr, rowval, nzval = Vector{Int64}(undef, n[] + 1), Vector{Int64}(undef, nnz[]), Vector{Tv}(undef, nnz[])
        check(ccall((:sgc_asm_fill_host, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Cint, Ptr{Int64}, Ptr{Cvoid}),
                    h, colptr, 1, rowval, nzval))
        return SparseMatrixCSC{Tv,Int64}(m[], n[], colptr, rowval, nzval)  # 1-based arrays, used as they are
    finally
        ccall((:sgc_asm_destroy, LIB), Cint, (Ptr{Cvoid},), h)
    end
end

ecomplex(e) = Cint(eltype(e) <: Complex)

# create_∂  — src/matrix/differential/partial.jl:54-60
function create_∂(::Val{:b200}, nw::Integer, isfwd::Bool, N::AbstractVector{<:Integer}, ∆w⁻¹::AbstractVector{<:Number},
                  isbloch::Bool=true, e⁻ⁱᵏᴸ::Number=1.0)
    keep, ptrs, cc = coefs((∆w⁻¹,))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep check(ccall((:sgc_asm_create_partial, LIB), Cint,
        (Ptr{Ptr{Cvoid}}, Cint, Ptr{Int64}, Cint, Cint, Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Cint),
        h, length(N), Int64[N...], nw, isfwd, ptrs[1], cc, isbloch, pair(e⁻ⁱᵏᴸ), ecomplex(e⁻ⁱᵏᴸ)))
    return finish_csc(h[])
end

# create_m̂  — src/matrix/mean.jl:131-138
function create_m̂(::Val{:b200}, nw::Integer, isfwd::Bool, N::AbstractVector{<:Integer}, ∆w::AbstractVector{<:Number},
                  ∆w′⁻¹::AbstractVector{<:Number}, isbloch::Bool=true, e⁻ⁱᵏᴸ::Number=1.0)
    keep, ptrs, cc = coefs((∆w, ∆w′⁻¹))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep check(ccall((:sgc_asm_create_mhat, LIB), Cint,
        (Ptr{Ptr{Cvoid}}, Cint, Ptr{Int64}, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Cint),
        h, length(N), Int64[N...], nw, isfwd, ptrs[1], ptrs[2], cc, isbloch, pair(e⁻ⁱᵏᴸ), ecomplex(e⁻ⁱᵏᴸ)))
    return finish_csc(h[])
end

# create_curl  — src/matrix/differential/curl.jl:50-58
function create_curl(::Val{:b200}, isfwd, ∆l⁻¹::NTuple{K,AbstractVector{<:Number}}, isbloch, e⁻ⁱᵏᴸ;
                     cmp_shp=1:K, cmp_out=1:K, cmp_in=1:K, order_cmpfirst::Bool=true) where {K}
    keep, ptrs, cc = coefs(∆l⁻¹)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep check(ccall((:sgc_asm_create_curl, LIB), Cint,
        (Ptr{Ptr{Cvoid}}, Cint, Ptr{Int64}, Ptr{Cint}, Ptr{Ptr{Cvoid}}, Cint, Ptr{Cint}, Ptr{Cdouble}, Cint, Ptr{Cint}, Cint,
         Ptr{Cint}, Cint, Ptr{Cint}, Cint),
        h, K, Int64[length.(∆l⁻¹)...], cints(isfwd), ptrs, cc, cints(isbloch), pairs(e⁻ⁱᵏᴸ), ecomplex(e⁻ⁱᵏᴸ), cints(cmp_shp),
        length(cmp_out), cints(cmp_out), length(cmp_in), cints(cmp_in), order_cmpfirst))
    return finish_csc(h[])
end

# create_divg / create_grad  — src/matrix/differential/divergence.jl:37-42, gradient.jl:37-42
for (fn, sym) in ((:create_divg, :sgc_asm_create_divg), (:create_grad, :sgc_asm_create_grad))
    @eval function $fn(::Val{:b200}, isfwd, ∆l⁻¹::NTuple{K,AbstractVector{<:Number}}, isbloch, e⁻ⁱᵏᴸ;
                       order_cmpfirst::Bool=true) where {K}
        keep, ptrs, cc = coefs(∆l⁻¹)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve keep check(ccall(($(QuoteNode(sym)), LIB), Cint,
            (Ptr{Ptr{Cvoid}}, Cint, Ptr{Int64}, Ptr{Cint}, Ptr{Ptr{Cvoid}}, Cint, Ptr{Cint}, Ptr{Cdouble}, Cint, Cint),
            h, K, Int64[length.(∆l⁻¹)...], cints(isfwd), ptrs, cc, cints(isbloch), pairs(e⁻ⁱᵏᴸ), ecomplex(e⁻ⁱᵏᴸ), order_cmpfirst))
        return finish_csc(h[])
    end
end

# create_mean  — src/matrix/mean.jl:72-78
function create_mean(::Val{:b200}, isfwd, ∆l::NTuple{K,AbstractVector{<:Number}}, ∆l′⁻¹::NTuple{K,AbstractVector{<:Number}},
                     isbloch, e⁻ⁱᵏᴸ; order_cmpfirst::Bool=true) where {K}
    keep, ptrs, cc = coefs((∆l..., ∆l′⁻¹...))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep check(ccall((:sgc_asm_create_mean, LIB), Cint,
        (Ptr{Ptr{Cvoid}}, Cint, Ptr{Int64}, Ptr{Cint}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Cint, Ptr{Cint}, Ptr{Cdouble}, Cint, Cint),
        h, K, Int64[length.(∆l)...], cints(isfwd), ptrs[1:K], ptrs[K+1:2K], cc, cints(isbloch), pairs(e⁻ⁱᵏᴸ), ecomplex(e⁻ⁱᵏᴸ),
        order_cmpfirst))
    return finish_csc(h[])
end

# create_πcmp  — src/matrix/permutation.jl:15-19
function create_πcmp(::Val{:b200}, N::AbstractVector{<:Integer}, permute::AbstractVector{<:Integer};
                     scale::AbstractVector{<:Number}=ones(length(permute)), order_cmpfirst::Bool=true)
    Ts = eltype(scale) <: Complex ? ComplexF64 : Float64
    sc = Vector{Ts}(scale)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve sc check(ccall((:sgc_asm_create_picmp, LIB), Cint,
        (Ptr{Ptr{Cvoid}}, Cint, Ptr{Int64}, Cint, Ptr{Cint}, Ptr{Cvoid}, Cint, Cint),
        h, length(N), Int64[N...], length(permute), cints(permute), sc, Ts == ComplexF64, order_cmpfirst))
    return finish_csc(h[])
end

end # module SGCB200
