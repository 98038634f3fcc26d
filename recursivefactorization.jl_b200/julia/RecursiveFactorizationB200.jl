# RecursiveFactorizationB200.jl -- Julia host layer over librfb200.so (C ABI: include/rfb200.h).
#
# STATUS: shipped UN-EXECUTED.  No Julia runtime exists in the build image or on the GPU box, so
# this file has never been run; the executable twin used by the tests and benchmarks is the Python
# package next to it (same C calls through ctypes).  It is written against the reference's public
# surface (RecursiveFactorization.jl 0.2.30):
#
#   lu(A, pivot = Val(true), thread = Val(false); kwargs...)            src/lu.jl:19-21
#   lu!(A, pivot = Val(true), thread = Val(false); check, kwargs...)    src/lu.jl:67-83
#   lu!(A, ipiv, pivot, thread; check, blocksize, threshold)            src/lu.jl:97-130
#
# and returns the same LinearAlgebra.LU{T,Matrix{T},Vector{BlasInt}} (src/lu.jl:129): the caller's
# matrix mutated in place, the caller's pivot vector, info.  `thread`, `blocksize` and `threshold`
# are accepted and ignored (they tune the CPU kernels); `leaf_width` is the GPU analogue.
# pivot = Val(false) / NoPivot() (src/lu.jl:27-65) is supported: `NotIPIV` below is the reference's
# lazy identity pivot vector, a user ipiv is filled with 1:min(m,n) (:107-113), info is negative on
# a zero pivot (Julia >= 1.11, :24-25).  `🦋workspace` / `🦋solve!` (src/butterflylu.jl:20-55) and a
# batched `lu_batched!` are bound further down.
# Unsupported inputs (complex or generic eltypes, non-strided arrays) throw -- there is
# deliberately no CPU fallback.
module RecursiveFactorizationB200

using LinearAlgebra
using LinearAlgebra: BlasInt, LU, checknonsingular

const librfb200 = get(ENV, "RFB200_LIB", joinpath(@__DIR__, "..", "librfb200.so"))

# mirror of `struct rfb_opts` (16 x Int32)
struct RfbOpts
    mem_space::Int32
    leaf_width::Int32
    f32_mode::Int32
    trsm_block::Int32
    gemm_path::Int32
    laswp_path::Int32
    no_pivot::Int32
    reserved::NTuple{9, Int32}
end
RfbOpts(; mem_space = 0, leaf_width = 0, f32_mode = 0, trsm_block = 0, gemm_path = 0, laswp_path = 0, no_pivot = 0) =
    RfbOpts(mem_space, leaf_width, f32_mode, trsm_block, gemm_path, laswp_path, no_pivot, ntuple(_ -> Int32(0), 9))

# src/lu.jl:27-32: the lazy identity pivot vector of an unpivoted factorization
struct NotIPIV <: AbstractVector{BlasInt}
    len::Int
end
Base.size(A::NotIPIV) = (A.len,)
Base.getindex(::NotIPIV, i::Int) = i
Base.view(::NotIPIV, r::AbstractUnitRange) = NotIPIV(length(r))

mutable struct Context
    handle::Ptr{Cvoid}
    function Context(device::Integer = 0)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:rfb_create, librfb200), Cint, (Ptr{Ptr{Cvoid}}, Cint), ref, device)
        ctx = new(ref[])
        rc == 0 || (msg = last_error(ctx); destroy!(ctx); error("rfb_create failed ($rc): $msg"))
        finalizer(destroy!, ctx)
        return ctx
    end
end
last_error(ctx::Context) =
    ctx.handle == C_NULL ? "null context" :
    unsafe_string(ccall((:rfb_last_error, librfb200), Cstring, (Ptr{Cvoid},), ctx.handle))
function destroy!(ctx::Context)
    ctx.handle == C_NULL && return
    ccall((:rfb_destroy, librfb200), Cint, (Ptr{Cvoid},), ctx.handle)
    ctx.handle = C_NULL
    return
end

# one context per task (a context is not thread-safe, see rfb200.h)
default_context() = get!(() -> Context(parse(Int, get(ENV, "LOCAL_RANK", "0"))), task_local_storage(), :rfb200_ctx)::Context

# pivot normalisation, same accepted spellings as src/lu.jl:10-17
normalize_pivot(::Val{true}) = true
normalize_pivot(::Val{false}) = false
normalize_pivot(::LinearAlgebra.RowMaximum) = true
normalize_pivot(::LinearAlgebra.NoPivot) = false

_wants_check(check::Bool) = check
_wants_check(::Val{true}) = true
_wants_check(::Val{false}) = false

for (T, sym) in ((Float64, :rfb_lu_f64), (Float32, :rfb_lu_f32))
    @eval function _rfb_lu!(ctx::Context, A::StridedMatrix{$T}, ipiv::Union{Vector{BlasInt}, NotIPIV}, opts::RfbOpts)
        m, n = size(A)
        stride(A, 1) == 1 || throw(ArgumentError("rfb200 needs unit row stride (column-major storage)"))
        info = Ref{Int64}(0)
        o = Ref(opts)
        pp = ipiv isa NotIPIV ? Ptr{Int64}(C_NULL) : pointer(ipiv)          # NULL ipiv == NotIPIV (rfb200.h)
        rc = GC.@preserve A ipiv ccall(($(QuoteNode(sym)), librfb200), Cint,
            (Ptr{Cvoid}, Ptr{$T}, Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{RfbOpts}),
            ctx.handle, pointer(A), m, n, max(1, stride(A, 2)), pp, info, o)
        rc == 0 || error("librfb200 error $rc: $(last_error(ctx))")
        return info[]
    end
end

function lu!(A::StridedMatrix{T}, ipiv::AbstractVector{<:Integer}, pivot = Val(true), thread = Val(false);
        check::Union{Bool, Val{true}, Val{false}} = Val(true),
        blocksize::Integer = 0, threshold::Integer = 0,          # accepted, ignored (CPU knobs)
        leaf_width::Integer = 0, ctx::Context = default_context()) where {T <: Union{Float64, Float32}}
    piv_on = normalize_pivot(pivot)
    BlasInt === Int64 || error("rfb200 writes Int64 pivots; this Julia has BlasInt = $BlasInt")
    length(ipiv) == min(size(A)...) || throw(DimensionMismatch("ipiv must have length min(m, n)"))
    piv = (ipiv isa Vector{BlasInt} || ipiv isa NotIPIV) ? ipiv : Vector{BlasInt}(undef, length(ipiv))
    info = _rfb_lu!(ctx, A, piv, RfbOpts(leaf_width = leaf_width, no_pivot = piv_on ? 0 : 1))
    piv === ipiv || copyto!(ipiv, piv)
    _wants_check(check) && checknonsingular(info)                # SingularException / ZeroPivotException, src/lu.jl:128
    return LU(A, ipiv, BlasInt(info))                            # src/lu.jl:129
end

# init_pivot, src/lu.jl:33-40
init_pivot(piv_on::Bool, minmn) = piv_on ? Vector{BlasInt}(undef, minmn) : NotIPIV(minmn)
function lu!(A::StridedMatrix{T}, pivot = Val(true), thread = Val(false); kwargs...) where {T <: Union{Float64, Float32}}
    return lu!(A, init_pivot(normalize_pivot(pivot), min(size(A)...)), pivot, thread; kwargs...)
end

lu(A::AbstractMatrix, pivot = Val(true), thread = Val(false); kwargs...) =
    lu!(copy(A), pivot, thread; kwargs...)                       # src/lu.jl:19-21

# Adjoint / Transpose wrappers, same contract as src/lu.jl:85-87
for (f, W) in ((:adjoint, :Adjoint), (:transpose, :Transpose)), g in (:lu, :lu!)
    @eval $g(A::$W, args...; kwargs...) = $f($g(parent(A), args...; kwargs...))
end

# ldiv!(F, B) on the GPU for factorizations produced above (square, Float64): forward + back substitution
function ldiv_gpu!(F::LU{Float64, <:StridedMatrix{Float64}, <:Vector{BlasInt}}, B::StridedVecOrMat{Float64};
        ctx::Context = default_context())
    n = size(F.factors, 1)
    size(F.factors, 2) == n && size(B, 1) == n || throw(DimensionMismatch("square LU and n-row right-hand side expected"))
    o = Ref(RfbOpts())
    rc = GC.@preserve F B ccall((:rfb_solve_f64, librfb200), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Ptr{Int64}, Ptr{Float64}, Int64, Int64, Ptr{RfbOpts}),
        ctx.handle, pointer(F.factors), n, stride(F.factors, 2), pointer(F.ipiv), pointer(B), size(B, 2),
        B isa AbstractVector ? n : stride(B, 2), o)
    rc == 0 || error("librfb200 error $rc: $(last_error(ctx))")
    return B
end

# ldiv!(F::LU{T,<:StridedMatrix,<:NotIPIV}, B) (src/lu.jl:60-64): two triangular solves, no interchanges
function LinearAlgebra.ldiv!(F::LU{Float64, <:StridedMatrix{Float64}, <:NotIPIV}, B::StridedVecOrMat{Float64};
        ctx::Context = default_context())
    n = size(F.factors, 1)
    o = Ref(RfbOpts())
    rc = GC.@preserve F B ccall((:rfb_solve_f64, librfb200), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Ptr{Int64}, Ptr{Float64}, Int64, Int64, Ptr{RfbOpts}),
        ctx.handle, pointer(F.factors), n, stride(F.factors, 2), C_NULL, pointer(B), size(B, 2),
        B isa AbstractVector ? n : stride(B, 2), o)
    rc == 0 || error("librfb200 error $rc: $(last_error(ctx))")
    return B
end

# ---------------------------------------------------------------------------------------------
# 🦋 (src/butterflylu.jl:20-57).  The workspace keeps the caller's A and b; padding (pad!, :180-197),
# the transform (🦋mul!, :93-113), the unpivoted LU and both butterfly matrix-vector products run
# inside one library call.  `ws` holds the 4n butterfly values (generate_rand_butterfly_vals!, :7-13,
# drawn here from Julia's default RNG: the reference's VectorizedRNG stream is SIMD-width dependent).
# ---------------------------------------------------------------------------------------------
struct 🦋workspace{T}
    A::Matrix{T}
    b::Vector{T}
    ws::Vector{T}
    out::Vector{T}
    n::Int
    function 🦋workspace(A::Matrix{T}, b::Vector{T}) where {T <: Union{Float64, Float32}}
        n = size(A, 1)
        np = n % 4 == 0 ? n : n + (4 - n % 4)
        ws = T.(exp.(T(-0.05) .+ T(0.1) .* rand(T, 4np)) .* T(0.5))
        new{T}(A, b, ws, similar(b), n)
    end
end
const butterfly_workspace = 🦋workspace

function 🦋solve!(w::🦋workspace{Float64}, thread = Val(false); ctx::Context = default_context())
    copyto!(w.out, w.b)
    info = Ref{Int64}(0)
    o = Ref(RfbOpts())
    rc = GC.@preserve w ccall((:rfb_butterfly_solve_f64, librfb200), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Ptr{Float64}, Int64, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{RfbOpts}),
        ctx.handle, pointer(w.A), w.n, stride(w.A, 2), pointer(w.out), 1, w.n, pointer(w.ws), info, o)
    rc == 0 || error("librfb200 error $rc: $(last_error(ctx))")
    checknonsingular(info[])                                     # lu!(A, Val(false), thread) checks (:48)
    return w.out
end

# lu! over the slices A[:, :, b] of a 3-D array (one launch when n <= 64 and m <= 128)
function lu_batched!(A::Array{Float64, 3}; check = true, ctx::Context = default_context())
    m, n, batch = size(A)
    mn = min(m, n)
    ipiv = Matrix{BlasInt}(undef, mn, batch)
    info = Vector{Int64}(undef, batch)
    o = Ref(RfbOpts())
    rc = GC.@preserve A ipiv info ccall((:rfb_lu_batched_f64, librfb200), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{RfbOpts}),
        ctx.handle, pointer(A), m, n, max(1, m), m * n, batch, pointer(ipiv), pointer(info), o)
    rc == 0 || error("librfb200 error $rc: $(last_error(ctx))")
    check && foreach(checknonsingular, info)
    return [LU(view(A, :, :, b), view(ipiv, :, b), BlasInt(info[b])) for b in 1:batch]
end

# everything else is not on the GPU path: fail loudly instead of silently running on the CPU
lu!(A::AbstractMatrix, args...; kwargs...) =
    throw(ArgumentError("rfb200 supports strided Float64/Float32 matrices only, got $(typeof(A))"))

# ---------------------------------------------------------------------------------------------
# Julia-driven recursion over the kernel-level ABI: the restatement of reckernel! (src/lu.jl:189-263)
# a Julia maintainer would own if the recursion is to stay in Julia (north_star).  `dA`, `dipiv`,
# `dinfo` are device pointers (rfb_malloc / rfb_h2d); pivots are produced in global coordinates by
# passing the node's row offset as `ipiv_add`, so no P2 .+= n1 pass is needed.
# ---------------------------------------------------------------------------------------------
nsplit(::Type{T}, n) where {T} = (k = max(2, 128 ÷ sizeof(T)); n >= k ? ((n + k ÷ 2) ÷ k) * (k ÷ 2) : n ÷ 2)

function reckernel_device!(ctx::Context, dA::Ptr{Float64}, m, lda, c0, n, dipiv::Ptr{Int64}, dinfo::Ptr{Int64}; leaf = 64)
    at(r, c) = dA + 8 * (r + c * lda)
    A = at(c0, c0); mm = m - c0
    chk(rc) = rc == 0 || error("librfb200 error $rc: $(last_error(ctx))")
    if n <= leaf
        chk(ccall((:rfb_panel_getrf_f64, librfb200), Cint,
            (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64, Ptr{Int64}, Int64, Ptr{Int64}, Int64),
            ctx.handle, A, mm, n, lda, dipiv + 8 * c0, c0, dinfo, c0))
        return
    end
    n1 = nsplit(Float64, n); n2 = n - n1
    reckernel_device!(ctx, dA, m, lda, c0, n1, dipiv, dinfo; leaf)
    AR = at(c0, c0 + n1)
    chk(ccall((:rfb_laswp_f64, librfb200), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Ptr{Int64}, Int64, Int64),
        ctx.handle, AR, n2, lda, dipiv + 8 * c0, n1, c0))
    chk(ccall((:rfb_trsm_llnu_f64, librfb200), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64),
        ctx.handle, A, n1, AR, n2, lda))
    chk(ccall((:rfb_gemm_nn_sub_f64, librfb200), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Int64, Int64),
        ctx.handle, at(c0 + n1, c0 + n1), at(c0 + n1, c0), AR, mm - n1, n2, n1, lda))
    reckernel_device!(ctx, dA, m, lda, c0 + n1, n2, dipiv, dinfo; leaf)
    chk(ccall((:rfb_laswp_f64, librfb200), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Ptr{Int64}, Int64, Int64),
        ctx.handle, at(c0 + n1, c0), n1, lda, dipiv + 8 * (c0 + n1), n2, c0 + n1))
    return
end

end # module
