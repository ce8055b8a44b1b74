# MORB200.jl -- drop-in shim that reroutes the numeric hot path of ModelOrderReduction.jl to
# libmor_b200.so (hand-written sm_100a CUDA) through `ccall`.  `using MORB200` after
# `using ModelOrderReduction` adds MORE SPECIFIC methods for Float64 strided / vector-of-vector
# snapshots; every other element or container type keeps the reference's own Julia method
# (that is the reference running, not a fallback inside the library).
#
# Rerouted seams (reference file:line, relative to ModelOrderReduction.jl/src):
#   DataReduction/POD.jl:156-166  reduce!(pod, ::SVD)   -> mor_b200_pod_factor[_cols] + mor_b200_pod_basis
#   DataReduction/POD.jl:195-202  reduce!(pod, ::TSVD)  -> same, method = MOR_TSVD
#   DataReduction/POD.jl:231-238  reduce!(pod, ::RSVD)  -> same, method = MOR_RSVD
#   DataReduction/POD.jl:4-6      matricize             -> column pointers, no hcat copy
#   deim.jl:9-28                  deim_interpolation_indices(basis) -> mor_b200_deim_indices
# Untouched and still executed in Julia: POD constructors + errorhandle (ErrorHandle.jl:1-18),
# determine_truncation (POD.jl:121-132, quirks preserved), the symbolic deim(sys, ...) (deim.jl:207-265).
#
# Outside the library's limits the shim DEFERS to the reference's own method (`invoke`), it never errors where the
# reference would succeed: small dimension > 4096, more than 1024 DEIM basis vectors, backend keyword arguments the
# library does not understand (Types.jl:61-95 forwards `alg.kwargs` verbatim at POD.jl:157,196), non-convergence.
# `defers(...)` below is that rule; python/mor_b200/pod.py mirrors it (DeferToReference).
#
# Multi-GPU from ONE Julia session: `MORB200.init_multi(8)` or `MOR_B200_NGPUS=8` in the environment -- the library
# shards the rows of the snapshot matrix over the GPUs itself (include/mor_b200.h, mor_b200_init_multi).
#
# NOTE: Julia is not installed in the build image, so this file is exercised only through its
# 1:1 ctypes mirror (python/mor_b200/*.py, which the GPU parity tests drive).
module MORB200

using ModelOrderReduction
using ModelOrderReduction: POD, SVD, TSVD, RSVD, determine_truncation
import ModelOrderReduction: reduce!, deim_interpolation_indices
using LinearAlgebra: SingularException

const LIB = get(ENV, "MOR_B200_LIB", joinpath(@__DIR__, "..", "lib", "libmor_b200.so"))
const MOR_SVD, MOR_TSVD, MOR_RSVD = Cint(0), Cint(1), Cint(2)

const F64Mat = StridedMatrix{Float64}
const F64VoV = Vector{Vector{Float64}}

last_error() = unsafe_string(ccall((:mor_b200_last_error, LIB), Cstring, ()))

const MAX_SMALL_DIM = 4096      # include/mor_b200.h: the small side of the snapshot matrix
const MAX_DEIM_COLS = 1024

"`true` when this call must run the reference's own Julia method instead of the library."
function defers(m::Integer, n::Integer, kwargs)
    min(m, n) > MAX_SMALL_DIM && return true
    for key in keys(kwargs)
        key === :full && kwargs[key] == false && continue          # thin SVD is what the library computes
        key === :alg && continue                                    # LAPACK driver choice: same result
        key in (:initvec, :tolconv, :maxiter) && continue           # TSVD.jl's Lanczos knobs: exact top-k is returned
        return true
    end
    return false
end
snapsize(X::StridedMatrix) = size(X)
snapsize(V::Vector{<:Vector}) = (length(V[1]), length(V))

function check(st::Integer)
    st == 0 && return nothing
    msg = last_error()
    st == 5 && error("libmor_b200 status 5: $msg")               # caught by factor_or_nothing -> reference method
    (st == 1 || st == 2) && throw(ArgumentError(msg))          # bad argument / non-finite input
    st == 6 && throw(SingularException(0))                     # deim.jl:21 `\` on a singular P'U
    st == 4 && throw(OutOfMemoryError())
    error("libmor_b200 status $st: $msg")
end

"Bind this process to a GPU (one process per GPU); optional, device 0 is bound lazily."
init(device::Integer = parse(Int, get(ENV, "MOR_B200_DEVICE", "0"))) =
    check(ccall((:mor_b200_init, LIB), Cint, (Cint,), device))

"ONE Julia process drives `ngpus` GPUs: the library shards the rows of every host matrix (also: MOR_B200_NGPUS)."
init_multi(ngpus::Integer = parse(Int, get(ENV, "MOR_B200_NGPUS", "1"))) =
    check(ccall((:mor_b200_init_multi, LIB), Cint, (Cint,), ngpus))

"(margins, values) of the last `deim_interpolation_indices` call: relative gap to the runner-up and winning residual
of every greedy step (mor_b200_deim_last_margins)."
function deim_last_margins()
    n = Ref{Int64}(0)
    check(ccall((:mor_b200_deim_last_margins, LIB), Cint,
        (Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}), C_NULL, C_NULL, 0, n, C_NULL, C_NULL))
    margins, values = zeros(n[]), zeros(n[])
    check(ccall((:mor_b200_deim_last_margins, LIB), Cint,
        (Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}), margins, values, n[], C_NULL, C_NULL, C_NULL))
    return margins, values
end

"Multi-GPU: rank 0 calls `unique_id()`, the host runtime (Distributed / MPI) broadcasts the 128 bytes,
every rank calls `comm_init(rank, nranks, id)`; snapshot ROWS are then sharded over the ranks in rank order."
function unique_id()
    id = zeros(UInt8, 128)
    check(ccall((:mor_b200_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id))
    return id
end
comm_init(rank::Integer, nranks::Integer, id::Vector{UInt8}) =
    check(ccall((:mor_b200_comm_init, LIB), Cint, (Cint, Cint, Ptr{UInt8}), rank, nranks, id))

mutable struct Factor
    handle::Ptr{Cvoid}
    m::Int
    n::Int
    s::Vector{Float64}
    function Factor(handle, m, n, s)
        f = new(handle, m, n, s)
        finalizer(f -> (f.handle != C_NULL && ccall((:mor_b200_pod_free, LIB), Cint, (Ptr{Cvoid},), f.handle);
                        f.handle = C_NULL), f)
        return f
    end
end

omega_ptr(::Nothing) = Ptr{Float64}(C_NULL)
omega_ptr(Ω::Matrix{Float64}) = pointer(Ω)

function factor(X::F64Mat, method::Cint, k::Integer, p::Integer = 0; omega = nothing, seed::Integer = 0)
    stride(X, 1) == 1 || (X = Matrix(X))
    m, n = size(X)
    s = Vector{Float64}(undef, n)        # capacity min(M_global, n) <= n, also when this process holds a row shard
    nS = Ref{Int64}(0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve X omega begin
        check(ccall((:mor_b200_pod_factor, LIB), Cint,
            (Ptr{Float64}, Int64, Int64, Int64, Cint, Int64, Int64, Ptr{Float64}, UInt64, Ptr{Ptr{Cvoid}},
                Ptr{Float64}, Ptr{Int64}),
            X, m, n, max(stride(X, 2), 1), method, k, p, omega_ptr(omega), seed, h, s, nS))
    end
    return Factor(h[], m, n, resize!(s, nS[]))
end

function factor(V::F64VoV, method::Cint, k::Integer, p::Integer = 0; omega = nothing, seed::Integer = 0)
    n, m = length(V), length(V[1])
    all(v -> length(v) == m, V) || throw(DimensionMismatch("snapshots of unequal length"))
    cols = [pointer(v) for v in V]
    s = Vector{Float64}(undef, n)
    nS = Ref{Int64}(0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve V cols omega begin
        check(ccall((:mor_b200_pod_factor_cols, LIB), Cint,
            (Ptr{Ptr{Float64}}, Int64, Int64, Cint, Int64, Int64, Ptr{Float64}, UInt64, Ptr{Ptr{Cvoid}},
                Ptr{Float64}, Ptr{Int64}),
            cols, m, n, method, k, p, omega_ptr(omega), seed, h, s, nS))
    end
    return Factor(h[], m, n, resize!(s, nS[]))
end

function basis(f::Factor, k::Integer)
    U = Matrix{Float64}(undef, f.m, k)
    check(ccall((:mor_b200_pod_basis, LIB), Cint,
        (Ptr{Cvoid}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64),
        f.handle, k, U, max(f.m, 1), C_NULL, 0))
    return U
end

release!(f::Factor) = finalize(f)

const Snaps = Union{F64Mat, F64VoV}

function factor_or_nothing(args...; kw...)
    try
        return factor(args...; kw...)
    catch e
        # status 5 (CholeskyQR passes / Jacobi sweeps exhausted): the reference's LAPACK path takes over
        e isa ErrorException && occursin("status 5", e.msg) || rethrow()
        @warn "libmor_b200 did not converge; running the reference method" exception = e
        return nothing
    end
end

# POD.jl:156-166
function reduce!(pod::POD{S, Float64}, alg::SVD)::Nothing where {S <: Snaps}
    defers(snapsize(pod.snapshots)..., alg.kwargs) && return invoke(reduce!, Tuple{POD, SVD}, pod, alg)
    f = factor_or_nothing(pod.snapshots, MOR_SVD, pod.max_nmodes)
    f === nothing && return invoke(reduce!, Tuple{POD, SVD}, pod, alg)
    s = f.s                                                           # full spectrum, POD.jl:164
    pod.nmodes, pod.renergy = determine_truncation(s, pod.min_nmodes, pod.max_nmodes, pod.min_renergy)
    pod.rbasis = basis(f, pod.nmodes)                                 # POD.jl:163
    pod.spectrum = s
    release!(f)
    return nothing
end

# POD.jl:195-202 -- TSVD.jl's Lanczos knobs (initvec, tolconv, maxiter) are accepted and ignored:
# the library returns the exact top-k triplets the iteration converges to.
function reduce!(pod::POD{S, Float64}, alg::TSVD)::Nothing where {S <: Snaps}
    defers(snapsize(pod.snapshots)..., alg.kwargs) && return invoke(reduce!, Tuple{POD, TSVD}, pod, alg)
    f = factor_or_nothing(pod.snapshots, MOR_TSVD, pod.nmodes)
    f === nothing && return invoke(reduce!, Tuple{POD, TSVD}, pod, alg)
    s = f.s
    n_max = min(f.m, f.n)                                             # POD.jl:197
    pod.renergy = sum(s) / (sum(s) + (n_max - pod.nmodes) * s[end])   # POD.jl:198
    pod.rbasis = basis(f, pod.nmodes)
    pod.spectrum = s
    release!(f)
    return nothing
end

# POD.jl:231-238 -- Omega is drawn by the library from `MORB200.rsvd_seed[]` (the reference draws
# it from Julia's global RNG); pass `omega = randn(n, k + p)` through `rsvd_omega[]` to share one.
const rsvd_seed = Ref{UInt64}(0)
const rsvd_omega = Ref{Union{Nothing, Matrix{Float64}}}(nothing)
function reduce!(pod::POD{S, Float64}, alg::RSVD)::Nothing where {S <: Snaps}
    defers(snapsize(pod.snapshots)..., NamedTuple()) && return invoke(reduce!, Tuple{POD, RSVD}, pod, alg)
    f = factor_or_nothing(pod.snapshots, MOR_RSVD, pod.nmodes, alg.p; omega = rsvd_omega[], seed = rsvd_seed[])
    f === nothing && return invoke(reduce!, Tuple{POD, RSVD}, pod, alg)
    s = f.s
    n_max = min(f.m, f.n)                                             # POD.jl:233
    pod.renergy = sum(s) / (sum(s) + (n_max - pod.nmodes) * s[end])   # POD.jl:234
    pod.rbasis = basis(f, pod.nmodes)
    pod.spectrum = s
    release!(f)
    return nothing
end

# deim.jl:9-28
function deim_interpolation_indices(B::F64Mat)::Vector{Int}
    size(B, 2) > MAX_DEIM_COLS && return invoke(deim_interpolation_indices, Tuple{AbstractMatrix}, B)
    stride(B, 1) == 1 || (B = Matrix(B))
    m, k = size(B)
    idx = Vector{Int64}(undef, k)
    GC.@preserve B check(ccall((:mor_b200_deim_indices, LIB), Cint,
        (Ptr{Float64}, Int64, Int64, Int64, Ptr{Int64}), B, m, k, max(stride(B, 2), 1), idx))
    return idx
end

# deim.jl:150 -- `projector = ((@view U[indices, :])' \ (U' * V))'`.  The expression sits inside the
# numeric `deim(...)` method, so it cannot be rerouted by dispatch; a maintainer replaces that one
# line by `projector = MORB200.deim_projector(U, V, indices)` (see INTEGRATION.md).
function deim_projector(U::F64Mat, V::F64Mat, indices::Vector{Int})
    stride(U, 1) == 1 || (U = Matrix(U))
    stride(V, 1) == 1 || (V = Matrix(V))
    n, kd = size(U)
    kp = size(V, 2)
    P = Matrix{Float64}(undef, kp, kd)
    idx = Vector{Int64}(indices)
    GC.@preserve U V idx check(ccall((:mor_b200_deim_projector, LIB), Cint,
        (Ptr{Float64}, Int64, Int64, Int64, Ptr{Float64}, Int64, Int64, Ptr{Int64}, Ptr{Float64}, Int64),
        U, n, kd, max(stride(U, 2), 1), V, kp, max(stride(V, 2), 1), idx, P, max(kp, 1)))
    return P
end

end # module
