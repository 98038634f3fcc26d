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
# a zero pivot (Julia >= 1.11, :24-25).  `ldiv_gpu!` / the NotIPIV `ldiv!` overload (:60-64), `🦋workspace` /
# `🦋solve!` (src/butterflylu.jl:20-55), a batched `lu_batched!` and the multi-GPU handle (`MultiGpu`, `lu!(mg, A)`)
# are bound further down, each for Float64 AND Float32; the kwargs of `lu!` reach every field of `rfb_opts`.
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
    keep_factors::Int32
    reserved::NTuple{8, Int32}
end
RfbOpts(; mem_space = 0, leaf_width = 0, f32_mode = 0, trsm_block = 0, gemm_path = 0, laswp_path = 0, no_pivot = 0, keep_factors = 0) =
    RfbOpts(mem_space, leaf_width, f32_mode, trsm_block, gemm_path, laswp_path, no_pivot, keep_factors, ntuple(_ -> Int32(0), 8))

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

# How lu! on a PAGE-LOCKED matrix sends finished factors back while it is still factoring (rfb_set_early_download):
# 2 = finished tiles (default), 1 = row bands at the right spine of the recursion, 0 = off.  Results are identical.
function set_early_download!(ctx::Context, mode::Integer)
    rc = ccall((:rfb_set_early_download, librfb200), Cint, (Ptr{Cvoid}, Cint), ctx.handle, mode)
    rc == 0 || error("librfb200 error $rc: $(last_error(ctx))")
    return ctx
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
        blocksize::Integer = 0, threshold::Integer = 0,          # accepted, ignored (CPU knobs, src/lu.jl:101-102)
        leaf_width::Integer = 0,                                 # the GPU analogue: columns per panel launch (0 = 64)
        f32_mode::Integer = 0,                                   # RFB_F32_AUTO / TF32X3 / FP32 (include/rfb200.h)
        trsm_block::Integer = 0, gemm_path::Integer = 0, laswp_path::Integer = 0,
        keep::Bool = false,                                      # leave the factors on the device for ldiv_kept! (rfb_solve_kept_*)
        ctx::Context = default_context()) where {T <: Union{Float64, Float32}}
    piv_on = normalize_pivot(pivot)
    BlasInt === Int64 || error("rfb200 writes Int64 pivots; this Julia has BlasInt = $BlasInt")
    length(ipiv) == min(size(A)...) || throw(DimensionMismatch("ipiv must have length min(m, n)"))
    piv = (ipiv isa Vector{BlasInt} || ipiv isa NotIPIV) ? ipiv : Vector{BlasInt}(undef, length(ipiv))
    info = _rfb_lu!(ctx, A, piv, RfbOpts(leaf_width = leaf_width, f32_mode = f32_mode, trsm_block = trsm_block,
                                         gemm_path = gemm_path, laswp_path = laswp_path, no_pivot = piv_on ? 0 : 1,
                                         keep_factors = keep ? 1 : 0))
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
        size(A, 2) == n && length(b) == n || throw(DimensionMismatch("square A and length-n b expected"))
        np = n % 4 == 0 ? n : n + (4 - n % 4)
        ws = T.(exp.(T(-0.05) .+ T(0.1) .* rand(T, 4np)) .* T(0.5))
        new{T}(A, b, ws, similar(b), n)
    end
end
const butterfly_workspace = 🦋workspace

# ldiv!(F, B) on the GPU for factorizations produced above (square; Float64 and Float32 like the reference's
# `T <: BlasFloat` overload, src/lu.jl:60-64): row interchanges (unless NotIPIV), forward + back substitution
function _check_solve_dims(F::LU, B)
    n = size(F.factors, 1)
    size(F.factors, 2) == n || throw(DimensionMismatch("square LU expected, got $(size(F.factors))"))
    size(B, 1) == n || throw(DimensionMismatch("B has $(size(B, 1)) rows, the factorization has $n"))
    stride(F.factors, 1) == 1 && stride(B, 1) == 1 || throw(ArgumentError("rfb200 needs unit row stride"))
    return n
end

for (T, solve, bsolve, batched) in ((Float64, :rfb_solve_f64, :rfb_butterfly_solve_f64, :rfb_lu_batched_f64),
                                    (Float32, :rfb_solve_f32, :rfb_butterfly_solve_f32, :rfb_lu_batched_f32))
    @eval begin
        function _rfb_solve!(ctx::Context, F::LU{$T}, pp::Ptr{Int64}, B::StridedVecOrMat{$T})
            n = _check_solve_dims(F, B)
            o = Ref(RfbOpts())
            rc = GC.@preserve F B ccall(($(QuoteNode(solve)), librfb200), Cint,
                (Ptr{Cvoid}, Ptr{$T}, Int64, Int64, Ptr{Int64}, Ptr{$T}, Int64, Int64, Ptr{RfbOpts}),
                ctx.handle, pointer(F.factors), n, max(1, stride(F.factors, 2)), pp, pointer(B), size(B, 2),
                B isa AbstractVector ? max(1, n) : max(1, stride(B, 2)), o)
            rc == 0 || error("librfb200 error $rc: $(last_error(ctx))")
            return B
        end
        ldiv_gpu!(F::LU{$T, <:StridedMatrix{$T}, <:Vector{BlasInt}}, B::StridedVecOrMat{$T}; ctx::Context = default_context()) =
            GC.@preserve F _rfb_solve!(ctx, F, pointer(F.ipiv), B)
        # ldiv!(F::LU{T,<:StridedMatrix,<:NotIPIV}, B) (src/lu.jl:60-64): two triangular solves, no interchanges
        LinearAlgebra.ldiv!(F::LU{$T, <:StridedMatrix{$T}, <:NotIPIV}, B::StridedVecOrMat{$T}; ctx::Context = default_context()) =
            _rfb_solve!(ctx, F, Ptr{Int64}(C_NULL), B)

        function 🦋solve!(w::🦋workspace{$T}, thread = Val(false); ctx::Context = default_context())
            copyto!(w.out, w.b)
            info = Ref{Int64}(0)
            o = Ref(RfbOpts())
            rc = GC.@preserve w ccall(($(QuoteNode(bsolve)), librfb200), Cint,
                (Ptr{Cvoid}, Ptr{$T}, Int64, Int64, Ptr{$T}, Int64, Int64, Ptr{$T}, Ptr{Int64}, Ptr{RfbOpts}),
                ctx.handle, pointer(w.A), w.n, max(1, stride(w.A, 2)), pointer(w.out), 1, max(1, w.n), pointer(w.ws), info, o)
            rc == 0 || error("librfb200 error $rc: $(last_error(ctx))")
            checknonsingular(info[])                             # lu!(A, Val(false), thread) checks (src/butterflylu.jl:48)
            return w.out
        end

        # lu! over the slices A[:, :, b] of a 3-D array (one launch when n <= 64 and m <= 128)
        function lu_batched!(A::Array{$T, 3}; check = true, ctx::Context = default_context())
            m, n, batch = size(A)
            mn = min(m, n)
            ipiv = Matrix{BlasInt}(undef, mn, batch)
            info = Vector{Int64}(undef, batch)
            o = Ref(RfbOpts())
            rc = GC.@preserve A ipiv info ccall(($(QuoteNode(batched)), librfb200), Cint,
                (Ptr{Cvoid}, Ptr{$T}, Int64, Int64, Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{RfbOpts}),
                ctx.handle, pointer(A), m, n, max(1, m), m * n, batch, pointer(ipiv), pointer(info), o)
            rc == 0 || error("librfb200 error $rc: $(last_error(ctx))")
            check && foreach(checknonsingular, info)
            return [LU(view(A, :, :, b), view(ipiv, :, b), BlasInt(info[b])) for b in 1:batch]
        end
    end
end

# `ldiv!` with the factors that `lu!(A, ipiv; keep = true)` left on the device (only B travels).  `kept_id(ctx)` right after
# the factorization identifies them; a stale id (another host-mode call reused the staging buffer) is an error here -- fall
# back to ldiv_gpu!(F, B).
function kept_id(ctx::Context = default_context())
    id = Ref{Int64}(0)
    _chk_rc(ctx, ccall((:rfb_kept_id, librfb200), Cint, (Ptr{Cvoid}, Ptr{Int64}), ctx.handle, id))
    return id[]
end
_chk_rc(ctx, rc) = rc == 0 || error("librfb200 error $rc: $(last_error(ctx))")
for (T, sym) in ((Float64, :rfb_solve_kept_f64), (Float32, :rfb_solve_kept_f32))
    @eval function ldiv_kept!(id::Integer, B::StridedVecOrMat{$T}; ctx::Context = default_context())
        stride(B, 1) == 1 || throw(ArgumentError("rfb200 needs unit row stride"))
        _chk_rc(ctx, GC.@preserve B ccall(($(QuoteNode(sym)), librfb200), Cint, (Ptr{Cvoid}, Int64, Ptr{$T}, Int64, Int64),
            ctx.handle, id, pointer(B), size(B, 2), B isa AbstractVector ? max(1, length(B)) : max(1, stride(B, 2))))
        return B
    end
end

# ---------------------------------------------------------------------------------------------
# Multi-GPU (one process, G devices): the C++ driver behind rfb_mg_* (csrc/rfb_mg.cu) -- 1-D block-cyclic
# block columns, ncclBroadcast of every factored block column, replicated L.  `lu!(mg, A, ipiv)` is
# `lu!(A, ipiv)` (src/lu.jl:97-130) on all GPUs of the handle and returns the same LU object.
# ---------------------------------------------------------------------------------------------
mutable struct MultiGpu
    handle::Ptr{Cvoid}
    ngpus::Int
    function MultiGpu(ngpus::Integer, devices::Union{Nothing, Vector{Cint}} = nothing)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:rfb_mg_create_all, librfb200), Cint, (Ptr{Ptr{Cvoid}}, Cint, Ptr{Cint}), ref, ngpus,
                   devices === nothing ? Ptr{Cint}(C_NULL) : pointer(devices))
        mg = new(ref[], ngpus)
        rc == 0 || (msg = mg_last_error(mg); mg_destroy!(mg); error("rfb_mg_create_all failed ($rc): $msg"))
        finalizer(mg_destroy!, mg)
        return mg
    end
end
mg_last_error(mg::MultiGpu) = mg.handle == C_NULL ? "null handle" :
    unsafe_string(ccall((:rfb_mg_last_error, librfb200), Cstring, (Ptr{Cvoid},), mg.handle))
function mg_destroy!(mg::MultiGpu)
    mg.handle == C_NULL && return
    ccall((:rfb_mg_destroy, librfb200), Cint, (Ptr{Cvoid},), mg.handle)
    mg.handle = C_NULL
    return
end
for (T, sym) in ((Float64, :rfb_mg_lu_f64), (Float32, :rfb_mg_lu_f32))
    @eval function lu!(mg::MultiGpu, A::StridedMatrix{$T}, ipiv::Vector{BlasInt} = Vector{BlasInt}(undef, size(A, 1));
            check::Union{Bool, Val{true}, Val{false}} = Val(true), block::Integer = 0)
        n = size(A, 1)
        size(A, 2) == n || throw(DimensionMismatch("the multi-GPU path factors square matrices"))
        length(ipiv) == n || throw(DimensionMismatch("ipiv must have length n"))
        stride(A, 1) == 1 || throw(ArgumentError("rfb200 needs unit row stride (column-major storage)"))
        info = Ref{Int64}(0)
        rc = GC.@preserve A ipiv ccall(($(QuoteNode(sym)), librfb200), Cint,
            (Ptr{Cvoid}, Ptr{$T}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Int64),
            mg.handle, pointer(A), n, max(1, stride(A, 2)), pointer(ipiv), info, block)
        rc == 0 || error("librfb200 error $rc: $(mg_last_error(mg))")
        _wants_check(check) && checknonsingular(info[])
        return LU(A, ipiv, BlasInt(info[]))
    end
end

# everything else is not on the GPU path: fail loudly instead of silently running on the CPU
lu!(A::AbstractMatrix, args...; kwargs...) =
    throw(ArgumentError("rfb200 supports strided Float64/Float32 matrices only, got $(typeof(A))"))

# ---------------------------------------------------------------------------------------------
# Julia-driven recursion over the kernel-level ABI: the restatement of lu! / _recurse! / reckernel!
# (src/lu.jl:97-130, :145-156, :189-263) a Julia maintainer would own if the recursion is to stay in Julia
# (north_star).  Device pointers throughout.  It issues exactly the launches csrc/rfb_api.cu:lu_rec issues:
#   leaf (n <= leaf columns)  -> rfb_lu_range_*   one K1 launch; pivots and info in GLOBAL coordinates (so the
#                                                reference's `P2 .+= n1` / `info += n1`, :248-260, have nothing to do),
#                                                and the leaf's row-exchange list recorded for K2
#   apply_permutation! (:233, :246, :151)  -> rfb_laswp_range_* with use_lists = 1 (list-driven K2)
#   ldiv!(UnitLowerTriangular(A11), A12) (:235, :153) -> rfb_trsm_llnu_*
#   schur_complement! (:240)  -> rfb_gemm_nn_sub_*
# Its executable twin is recursivefactorization.jl_b200/host_recursion.py (same calls through ctypes), which
# tests/test_gpu_lu.py checks bit for bit against rfb_lu_*.
# ---------------------------------------------------------------------------------------------
nsplit(::Type{T}, n) where {T} = (k = max(2, 128 ÷ sizeof(T)); n >= k ? ((n + k ÷ 2) ÷ k) * (k ÷ 2) : n ÷ 2)

_chk(ctx, rc) = rc == 0 || error("librfb200 error $rc: $(last_error(ctx))")

for (T, suf) in ((Float64, "f64"), (Float32, "f32"))
    lu_range = QuoteNode(Symbol("rfb_lu_range_", suf)); laswp_range = QuoteNode(Symbol("rfb_laswp_range_", suf))
    trsm = QuoteNode(Symbol("rfb_trsm_llnu_", suf)); gemm = QuoteNode(Symbol("rfb_gemm_nn_sub_", suf))
    @eval begin
        function reckernel_device!(ctx::Context, dA::Ptr{$T}, m, lda, c0, n, dipiv::Ptr{Int64}, dinfo::Ptr{Int64},
                                   opts::Ref{RfbOpts}; leaf = 64)
            at(r, c) = dA + sizeof($T) * (r + c * lda)
            if n <= leaf                                                                     # :192-195
                _chk(ctx, ccall(($lu_range, librfb200), Cint,
                    (Ptr{Cvoid}, Ptr{$T}, Int64, Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{RfbOpts}),
                    ctx.handle, dA, m, lda, c0, n, dipiv, dinfo, opts))
                return
            end
            n1 = nsplit($T, n); n2 = n - n1                                                  # :196-198
            reckernel_device!(ctx, dA, m, lda, c0, n1, dipiv, dinfo, opts; leaf)             # :229
            _chk(ctx, ccall(($laswp_range, librfb200), Cint,                                 # :233  P1 -> AR
                (Ptr{Cvoid}, Ptr{$T}, Int64, Int64, Int64, Int64, Int64, Ptr{Int64}, Cint),
                ctx.handle, dA, lda, c0 + n1, n2, c0, c0 + n1, dipiv, 1))
            _chk(ctx, ccall(($trsm, librfb200), Cint, (Ptr{Cvoid}, Ptr{$T}, Int64, Ptr{$T}, Int64, Int64),
                ctx.handle, at(c0, c0), n1, at(c0, c0 + n1), n2, lda))                       # :235
            _chk(ctx, ccall(($gemm, librfb200), Cint,                                        # :240
                (Ptr{Cvoid}, Ptr{$T}, Ptr{$T}, Ptr{$T}, Int64, Int64, Int64, Int64),
                ctx.handle, at(c0 + n1, c0 + n1), at(c0 + n1, c0), at(c0, c0 + n1), m - c0 - n1, n2, n1, lda))
            reckernel_device!(ctx, dA, m, lda, c0 + n1, n2, dipiv, dinfo, opts; leaf)        # :244
            _chk(ctx, ccall(($laswp_range, librfb200), Cint,                                 # :246  P2 -> A21
                (Ptr{Cvoid}, Ptr{$T}, Int64, Int64, Int64, Int64, Int64, Ptr{Int64}, Cint),
                ctx.handle, dA, lda, c0, n1, c0 + n1, c0 + n, dipiv, 1))
            return
        end

        # lu!(A, ipiv; check = false) on a device matrix (src/lu.jl:97-130 + :145-156), pivoted.  dperm = three device
        # int32 arrays (2 cap, 2 cap, cap entries, cap >= min(m, n) + 64) for the row-exchange lists (rfb_perm_buffers).
        function lu_device!(ctx::Context, dA::Ptr{$T}, m, n, lda, dipiv::Ptr{Int64}, dinfo::Ptr{Int64},
                            dperm::NTuple{3, Ptr{Int32}}, cap; leaf = 64)
            mn = min(m, n)
            opts = Ref(RfbOpts(mem_space = 1, leaf_width = leaf))
            _chk(ctx, ccall((:rfb_perm_buffers, librfb200), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Int64),
                ctx.handle, dperm[1], dperm[2], dperm[3], cap))
            _chk(ctx, ccall((:rfb_memset, librfb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Csize_t), ctx.handle, dinfo, 0, 8))
            mn == 0 && return
            reckernel_device!(ctx, dA, m, lda, 0, mn, dipiv, dinfo, opts; leaf)              # :147
            if m < n                                                                         # fat tail, :148-154
                _chk(ctx, ccall(($laswp_range, librfb200), Cint,
                    (Ptr{Cvoid}, Ptr{$T}, Int64, Int64, Int64, Int64, Int64, Ptr{Int64}, Cint),
                    ctx.handle, dA, lda, m, n - m, 0, mn, dipiv, 1))
                _chk(ctx, ccall(($trsm, librfb200), Cint, (Ptr{Cvoid}, Ptr{$T}, Int64, Ptr{$T}, Int64, Int64),
                    ctx.handle, dA, m, dA + sizeof($T) * (m * lda), n - m, lda))
            end
            return
        end
    end
end

end # module
