# LBMB200.jl -- the Julia host side of the drop-in: keeps LazyBandedMatrices.jl's public API and
# layout dispatch, and routes `mul!` / `materialize` of GPU-resident banded operands to
# liblbm_b200.so through `ccall` (no CUDA.jl kernels, no multi-backend dispatch, no CPU fallback).
#
# NOT RUNNABLE IN THE BUILD IMAGE (there is no `julia`); it is kept mechanical -- pointer passing
# only -- and mirrors 1:1 the ctypes binding that IS tested (lbm_b200/_lib.py).  See INTEGRATION.md.
#
# How it hooks in.  The reference's types advertise their MemoryLayout from the layouts of their
# storage vectors (/root/reference/src/tridiag.jl:685-686, src/bidiag.jl:438), and BandedMatrices'
# `BandedMatrix{T,C}` advertises `BandedColumns{typeof(MemoryLayout(C))}`.  So one new array type
# `DevArray` with `MemoryLayout(::Type{<:DevArray}) = DeviceColumnMajor()` makes
#   BandedMatrix{T,<:DevMatrix}                      -> BandedColumns{DeviceColumnMajor}
#   LazyBandedMatrices.Tridiagonal{T,DevVec,...}      -> TridiagonalLayout{DeviceColumnMajor,...}
#   LazyBandedMatrices.SymTridiagonal / Bidiagonal    -> SymTridiagonalLayout / BidiagonalLayout
# without touching the package, and ArrayLayouts' `materialize!(::MulAdd{...})` methods below are the
# ONLY new dispatch.  LazyArrays' `ApplyMatrix`, `Mul` flattening, `bandwidths`, and this package's
# `BlockKron` / `BlockVcat` / `BlockHcat` structure code work unchanged: they only consult layouts
# and bandwidths (src/blockkron.jl:353-362, src/blockconcat.jl:364-395).
module LBMB200

using LinearAlgebra
import ArrayLayouts: MemoryLayout, AbstractColumnMajor, MulAdd, materialize!, MatMulVecAdd, MatMulMatAdd,
                     TridiagonalLayout, SymTridiagonalLayout, BidiagonalLayout,
                     diagonaldata, subdiagonaldata, supdiagonaldata
import BandedMatrices
import BandedMatrices: BandedMatrix, BandedColumns, bandeddata, bandwidths, bandwidth
import LazyBandedMatrices
import LazyBandedMatrices: bidiagonaluplo
import Base: size, getindex, similar, unsafe_convert, pointer, strides, copyto!

const LIB = get(ENV, "LBM_B200_LIB", "liblbm_b200.so")

# ---- handle -------------------------------------------------------------------------------------
mutable struct Handle
    ptr::Ptr{Cvoid}
end
function Handle(device::Integer = 0)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:lbm_create, LIB), Cint, (Ptr{Ptr{Cvoid}}, Cint), r, device)
    rc == 0 || error("lbm_create failed: ", unsafe_string(ccall((:lbm_last_error_string, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    h = Handle(r[])
    finalizer(h -> ccall((:lbm_destroy, LIB), Cint, (Ptr{Cvoid},), h.ptr), h)
    h
end
const HANDLE = Ref{Handle}()
handle() = (isassigned(HANDLE) || (HANDLE[] = Handle(0)); HANDLE[])
sync() = check(ccall((:lbm_sync, LIB), Cint, (Ptr{Cvoid},), handle().ptr), "lbm_sync")

# LAPACK-info style status -> Julia exceptions.  Shape errors never get here: they are thrown as
# DimensionMismatch BEFORE the ccall, as the reference does (src/bidiag.jl:363-377;
# test/test_tridiag.jl:272-279).
function check(rc::Integer, what)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:lbm_last_error_string, LIB), Cstring, (Ptr{Cvoid},), handle().ptr))
    rc < 0 ? throw(ArgumentError("$what: $msg")) : error("$what: $msg")
end

# ---- device array -------------------------------------------------------------------------------
struct DeviceColumnMajor <: AbstractColumnMajor end

mutable struct DevArray{T,N} <: AbstractArray{T,N}
    ptr::Ptr{T}
    dims::NTuple{N,Int}
    function DevArray{T,N}(::UndefInitializer, dims::NTuple{N,Int}) where {T,N}
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:lbm_malloc, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Csize_t), handle().ptr, r, sizeof(T) * prod(dims)), "lbm_malloc")
        a = new{T,N}(convert(Ptr{T}, r[]), dims)
        finalizer(a -> ccall((:lbm_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), handle().ptr, a.ptr), a)
        a
    end
end
const DevVector{T} = DevArray{T,1}
const DevMatrix{T} = DevArray{T,2}
DevArray{T}(u::UndefInitializer, dims::Int...) where T = DevArray{T,length(dims)}(u, dims)
size(a::DevArray) = a.dims
strides(a::DevArray) = Base.size_to_strides(1, a.dims...)
pointer(a::DevArray) = a.ptr
unsafe_convert(::Type{Ptr{T}}, a::DevArray{T}) where T = a.ptr
getindex(a::DevArray, i...) = error("scalar indexing of a device array is disabled; copy to the host with Array(a)")
similar(a::DevArray{T}, ::Type{S}, dims::Dims) where {T,S} = DevArray{S,length(dims)}(undef, dims)
MemoryLayout(::Type{<:DevArray}) = DeviceColumnMajor()

function DevArray(a::Array{T,N}) where {T,N}   # host -> device
    d = DevArray{T,N}(undef, size(a))
    check(ccall((:lbm_memcpy_h2d, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), handle().ptr, d.ptr, a, sizeof(a)), "lbm_memcpy_h2d")
    sync(); d
end
function Base.Array(d::DevArray{T,N}) where {T,N}  # device -> host
    a = Array{T,N}(undef, size(d))
    check(ccall((:lbm_memcpy_d2h, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), handle().ptr, a, d.ptr, sizeof(a)), "lbm_memcpy_d2h")
    sync(); a
end

# ---- size checks: same messages as check_A_mul_B!_sizes (src/bidiag.jl:363-377) --------------------
function check_mul_sizes(C, A, B)
    nA, mA = size(A, 1), size(A, 2)
    nB, mB = size(B, 1), size(B, 2)
    nC, mC = size(C, 1), size(C, 2)
    nA == nC || throw(DimensionMismatch("sizes size(A)=$(size(A)) and size(C) = $(size(C)) must match at first entry."))
    mA == nB || throw(DimensionMismatch("second entry of size(A)=$(size(A)) and first entry of size(B) = $(size(B)) must match."))
    mB == mC || throw(DimensionMismatch("sizes size(B)=$(size(B)) and size(C) = $(size(C)) must match at first second entry."))
end

for (T, p) in ((Float64, :d), (Float32, :s))
    gbmv = QuoteNode(Symbol(:lbm_, p, :gbmv)); tri = QuoteNode(Symbol(:lbm_, p, :tridiag_mv))
    bi = QuoteNode(Symbol(:lbm_, p, :bidiag_mv)); gbmm = QuoteNode(Symbol(:lbm_, p, :gbmm))
    @eval begin
        # BandedMatrix * vector: replaces BandedMatrices.gbmv! -> ccall(dgbmv_)   [SURVEY.md 3.1]
        function materialize!(M::MatMulVecAdd{<:BandedColumns{DeviceColumnMajor},DeviceColumnMajor,DeviceColumnMajor,$T})
            α, A, x, β, y = M.α, M.A, M.B, M.β, M.C
            check_mul_sizes(y, A, x)
            l, u = bandwidths(A); data = bandeddata(A)
            check(ccall(($gbmv, LIB), Cint,
                        (Ptr{Cvoid}, Cchar, Int64, Int64, Int64, Int64, $T, Ptr{$T}, Int64, Ptr{$T}, $T, Ptr{$T}),
                        handle().ptr, 'N', size(A, 1), size(A, 2), l, u, α, data, size(data, 1), x, β, y), "lbm_gbmv")
            y
        end
        # Tridiagonal / SymTridiagonal (layouts from src/tridiag.jl:685-686; accessors :705-712)
        function materialize!(M::MulAdd{<:Union{TridiagonalLayout{DeviceColumnMajor},SymTridiagonalLayout{DeviceColumnMajor}},
                                        DeviceColumnMajor,DeviceColumnMajor,$T})
            α, A, X, β, Y = M.α, M.A, M.B, M.β, M.C
            check_mul_sizes(Y, A, X)
            n = size(A, 1)
            check(ccall(($tri, LIB), Cint,
                        (Ptr{Cvoid}, Int64, Ptr{$T}, Ptr{$T}, Ptr{$T}, $T, Ptr{$T}, Int64, Int64, $T, Ptr{$T}, Int64),
                        handle().ptr, n, subdiagonaldata(A), diagonaldata(A), supdiagonaldata(A), α,
                        X, max(stride(X, 2), 1), size(X, 2), β, Y, max(stride(Y, 2), 1)), "lbm_tridiag_mv")
            Y                       # SymTridiagonal: sub == sup == ev (src/tridiag.jl:708-709) -> read once
        end
        # Bidiagonal (layout src/bidiag.jl:438, uplo :440, accessors :443-445)
        function materialize!(M::MulAdd{<:BidiagonalLayout{DeviceColumnMajor},DeviceColumnMajor,DeviceColumnMajor,$T})
            α, A, X, β, Y = M.α, M.A, M.B, M.β, M.C
            check_mul_sizes(Y, A, X)
            ev = bidiagonaluplo(A) == 'U' ? supdiagonaldata(A) : subdiagonaldata(A)
            check(ccall(($bi, LIB), Cint,
                        (Ptr{Cvoid}, Cchar, Int64, Ptr{$T}, Ptr{$T}, $T, Ptr{$T}, Int64, Int64, $T, Ptr{$T}, Int64),
                        handle().ptr, bidiagonaluplo(A), size(A, 1), diagonaldata(A), ev, α,
                        X, max(stride(X, 2), 1), size(X, 2), β, Y, max(stride(Y, 2), 1)), "lbm_bidiag_mv")
            Y
        end
        # banded * banded -> banded: materialises ApplyMatrix(*, A, B) (README.md:8-26); replaces gbmm!
        function materialize!(M::MatMulMatAdd{<:BandedColumns{DeviceColumnMajor},<:BandedColumns{DeviceColumnMajor},
                                              <:BandedColumns{DeviceColumnMajor},$T})
            α, A, B, β, C = M.α, M.A, M.B, M.β, M.C
            check_mul_sizes(C, A, B)
            (la, ua), (lb, ub), (lc, uc) = bandwidths(A), bandwidths(B), bandwidths(C)
            dA, dB, dC = bandeddata(A), bandeddata(B), bandeddata(C)
            check(ccall(($gbmm, LIB), Cint,
                        (Ptr{Cvoid}, Int64, Int64, Int64, $T, Int64, Int64, Ptr{$T}, Int64, Int64, Int64, Ptr{$T}, Int64,
                         $T, Int64, Int64, Ptr{$T}, Int64),
                        handle().ptr, size(A, 1), size(B, 2), size(A, 2), α, la, ua, dA, size(dA, 1), lb, ub, dB, size(dB, 1),
                        β, lc, uc, dC, size(dC, 1)), "lbm_gbmm")
            C
        end
    end
end

# `+`/`-` among structured kinds (src/special.jl:75-222, src/tridiag.jl:206-208,:570-572, src/bidiag.jl:330-348):
# the reference's methods are written on the diagonal vectors (`A.dv + B.diag`, `-B.ev`, `A.d .+ B.λ`), so giving
# DevVector these three operations makes every one of them run on the device unchanged; each is one launch of
# lbm_?tridiag_axpby on the main-diagonal slot.  `axpby3!` is the fused form (three diagonals, one launch) used by
# the Tridiagonal +/- Tridiagonal methods below.
for (T, p) in ((Float64, :d), (Float32, :s))
    ax = QuoteNode(Symbol(:lbm_, p, :tridiag_axpby)); pk = QuoteNode(Symbol(:lbm_, p, :tridiag_pack))
    @eval begin
        function axpby3!(odl, od, odu, α::$T, adl, ad, adu, β::$T, bdl, bd, bdu, γ::$T, n::Integer)
            P(v) = v === nothing ? Ptr{$T}(C_NULL) : pointer(v)
            check(ccall(($ax, LIB), Cint,
                        (Ptr{Cvoid}, Int64, $T, Ptr{$T}, Ptr{$T}, Ptr{$T}, $T, Ptr{$T}, Ptr{$T}, Ptr{$T}, $T, Ptr{$T}, Ptr{$T}, Ptr{$T}),
                        handle().ptr, n, α, P(adl), P(ad), P(adu), β, P(bdl), P(bd), P(bdu), γ, P(odl), P(od), P(odu)),
                  "lbm_tridiag_axpby")
        end
        function Base.:+(a::DevVector{$T}, b::DevVector{$T})
            length(a) == length(b) || throw(DimensionMismatch("dimensions must match: a has dims $(axes(a)), b has dims $(axes(b))"))
            o = similar(a); axpby3!(nothing, o, nothing, one($T), nothing, a, nothing, one($T), nothing, b, nothing, zero($T), length(a)); o
        end
        function Base.:-(a::DevVector{$T}, b::DevVector{$T})
            length(a) == length(b) || throw(DimensionMismatch("dimensions must match: a has dims $(axes(a)), b has dims $(axes(b))"))
            o = similar(a); axpby3!(nothing, o, nothing, one($T), nothing, a, nothing, -one($T), nothing, b, nothing, zero($T), length(a)); o
        end
        Base.:-(a::DevVector{$T}) = (o = similar(a); axpby3!(nothing, o, nothing, -one($T), nothing, a, nothing, zero($T), nothing, nothing, nothing, zero($T), length(a)); o)
        Base.:*(a::DevVector{$T}, c::Number) = (o = similar(a); axpby3!(nothing, o, nothing, $T(c), nothing, a, nothing, zero($T), nothing, nothing, nothing, zero($T), length(a)); o)
        Base.:*(c::Number, a::DevVector{$T}) = a * c
        # `A.d .+ B.λ` / `A.λ .- B.d` of the UniformScaling methods (src/special.jl:183-222)
        Base.broadcasted(::typeof(+), a::DevVector{$T}, λ::Number) = (o = similar(a); axpby3!(nothing, o, nothing, one($T), nothing, a, nothing, zero($T), nothing, nothing, nothing, $T(λ), length(a)); o)
        Base.broadcasted(::typeof(+), λ::Number, a::DevVector{$T}) = Base.broadcasted(+, a, λ)
        Base.broadcasted(::typeof(-), λ::Number, a::DevVector{$T}) = (o = similar(a); axpby3!(nothing, o, nothing, -one($T), nothing, a, nothing, zero($T), nothing, nothing, nothing, $T(λ), length(a)); o)
        # fused: all three diagonals in one launch (src/tridiag.jl:571-572)
        function Base.:+(A::LazyBandedMatrices.Tridiagonal{$T,<:DevVector}, B::LazyBandedMatrices.Tridiagonal{$T,<:DevVector})
            n = size(A, 1); dl, d, du = similar(A.dl), similar(A.d), similar(A.du)
            axpby3!(dl, d, du, one($T), A.dl, A.d, A.du, one($T), B.dl, B.d, B.du, zero($T), n); LazyBandedMatrices.Tridiagonal(dl, d, du)
        end
        function Base.:-(A::LazyBandedMatrices.Tridiagonal{$T,<:DevVector}, B::LazyBandedMatrices.Tridiagonal{$T,<:DevVector})
            n = size(A, 1); dl, d, du = similar(A.dl), similar(A.d), similar(A.du)
            axpby3!(dl, d, du, one($T), A.dl, A.d, A.du, -one($T), B.dl, B.d, B.du, zero($T), n); LazyBandedMatrices.Tridiagonal(dl, d, du)
        end
        # BandedMatrix(A, (l,u)) of a structured operand: the storage lbm_?gbmm consumes, so structured x structured
        # products (test/test_special.jl:154-160) reach the banded x banded kernel
        function BandedMatrices.BandedMatrix(A::Union{LazyBandedMatrices.Tridiagonal{$T,<:DevVector},LazyBandedMatrices.SymTridiagonal{$T,<:DevVector},LazyBandedMatrices.Bidiagonal{$T,<:DevVector}},
                                             (l, u)::NTuple{2,Integer} = bandwidths(A))
            n = size(A, 1); data = DevArray{$T}(undef, l + u + 1, n)
            bl, bu = bandwidths(A)
            check(ccall(($pk, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{$T}, Ptr{$T}, Ptr{$T}, Int64, Int64, Ptr{$T}, Int64),
                        handle().ptr, n, bl > 0 ? pointer(subdiagonaldata(A)) : Ptr{$T}(C_NULL), diagonaldata(A),
                        bu > 0 ? pointer(supdiagonaldata(A)) : Ptr{$T}(C_NULL), l, u, data, l + u + 1), "lbm_tridiag_pack")
            BandedMatrices._BandedMatrix(data, Base.OneTo(n), l, u)
        end
    end
end

# mul!(y, ApplyMatrix(*, F...), x, α, β) for all-banded device factors in ONE ccall (LazyArrays' flattening would issue k gbmv
# ccalls and allocate k-1 temporaries; the library ping-pongs them inside the handle's scratch).  A package extension forwards
# `materialize!(::MulAdd{<:ApplyLayout{typeof(*)}, DeviceColumnMajor, DeviceColumnMajor})` here.
function chain_mul!(y::DevVector{Float64}, factors::NTuple{N,BandedMatrix}, x::DevVector{Float64}, α = 1.0, β = 0.0) where N
    for f in 1:N-1
        size(factors[f], 2) == size(factors[f+1], 1) || throw(DimensionMismatch("second entry of size(A)=$(size(factors[f])) and first entry of size(B) = $(size(factors[f+1])) must match."))
    end
    size(factors[1], 1) == length(y) || throw(DimensionMismatch("sizes size(A)=$(size(factors[1])) and size(C) = $(size(y)) must match at first entry."))
    size(factors[N], 2) == length(x) || throw(DimensionMismatch("second entry of size(A)=$(size(factors[N])) and first entry of size(B) = $(size(x)) must match."))
    m = Int64[size(F, 1) for F in factors]; n = Int64[size(F, 2) for F in factors]
    kl = Int64[bandwidth(F, 1) for F in factors]; ku = Int64[bandwidth(F, 2) for F in factors]
    ptrs = Ptr{Float64}[pointer(bandeddata(F)) for F in factors]; ld = Int64[size(bandeddata(F), 1) for F in factors]
    check(ccall((:lbm_dgbmv_chain, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Ptr{Float64}}, Ptr{Int64}, Float64, Ptr{Float64}, Float64, Ptr{Float64}),
                handle().ptr, N, m, n, kl, ku, ptrs, ld, α, x, β, y), "lbm_dgbmv_chain")
    y
end

# 2-D Laplacian: Δ = BroadcastArray(+, BlockKron(A, Eye(n2)), BlockKron(Eye(n1), B)) applied without
# forming it (test/test_blockkron.jl:296-306): y = vec(X*A' + B*X).
function kronsum_mul!(y::DevVector{Float64}, A::BandedMatrix, B::BandedMatrix, x::DevVector{Float64}, α = 1.0, β = 0.0)
    n1, n2 = size(A, 1), size(B, 1)
    length(x) == n1 * n2 || throw(DimensionMismatch("second entry of size(A)=$((n1*n2, n1*n2)) and first entry of size(B) = $(size(x)) must match."))
    (la, ua), (lb, ub) = bandwidths(A), bandwidths(B)
    dA, dB = bandeddata(A), bandeddata(B)
    check(ccall((:lbm_dkronsum2d_mv, LIB), Cint,
                (Ptr{Cvoid}, Int64, Int64, Int64, Int64, Ptr{Float64}, Int64, Int64, Int64, Ptr{Float64}, Int64,
                 Float64, Ptr{Float64}, Float64, Ptr{Float64}),
                handle().ptr, n1, n2, la, ua, dA, size(dA, 1), lb, ub, dB, size(dB, 1), α, x, β, y), "lbm_dkronsum2d_mv")
    y
end

# brand on the device: every storage slot U[0,1) (the counter-based stream shared with the oracle)
function brand_dev(::Type{Float64}, m, n, l, u; seed = 0x1)
    data = DevArray{Float64}(undef, l + u + 1, n)
    check(ccall((:lbm_dfill_uniform, LIB), Cint, (Ptr{Cvoid}, UInt64, Int64, Int64, Ptr{Float64}),
                handle().ptr, seed, 0, length(data), data), "lbm_dfill_uniform")
    BandedMatrices._BandedMatrix(data, Base.OneTo(m), l, u)
end

end # module
