# RxGaussB200.jl -- the Julia-side shim a maintainer would add next to RxInfer to route the batched
# Gaussian hot path to librxgauss.so (include/rxgauss.h) through plain `ccall`.  No CUDA.jl codegen:
# the kernels are the hand-written sm_100a ones in csrc/.
#
# NOT RUNNABLE IN THE BUILD IMAGE (no Julia there); it is the reference-side binding that
# INTEGRATION.md describes, kept next to the C ABI it binds.  The Python package in this directory
# (`_lib.py`, `context.py`, `inference.py`, `rules.py`) is the same binding in the host language
# that *is* available, and is what the parity tests drive.
module RxGaussB200

using LinearAlgebra
# using RxInfer, ReactiveMP, ExponentialFamily, BayesBase    # in a real checkout

const LIB = get(ENV, "RXGAUSS_LIB", joinpath(@__DIR__, "..", "librxgauss.so"))

const RXG_PTR_DEVICE       = UInt32(1) << 0
const RXG_MODEL_PER_CHAIN  = UInt32(1) << 1
const RXG_ASYNC            = UInt32(1) << 2
const RXG_COV_SHARED_OUT   = UInt32(1) << 3
const RXG_PATH_PER_CHAIN   = UInt32(1) << 4
const RXG_TRANSITION_FIRST = UInt32(1) << 5
const RXG_COV_REPLICATE    = UInt32(1) << 6

struct RxGaussError <: Exception
    code::Cint
    msg::String
end

mutable struct Context
    handle::Ptr{Cvoid}
    function Context(device::Integer = 0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:rxg_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint, Cuint), h, device, 0)
        rc == 0 || throw(RxGaussError(rc, "rxg_create failed (no CUDA device? there is no CPU fallback)"))
        ctx = new(h[])
        finalizer(c -> ccall((:rxg_destroy, LIB), Cint, (Ptr{Cvoid},), c.handle), ctx)
        return ctx
    end
end

function check(ctx::Context, rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:rxg_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.handle))
    throw(RxGaussError(rc, msg))
end

f32(M) = Matrix{Float32}(M)      # shared model matrices are host constants, row-major on the C side
rowmajor(M::AbstractMatrix) = collect(transpose(f32(M)))   # Julia is column-major

"""
    smooth(ctx, y, A, B, P, Q, m0, S0; free_energy = false)

Batched replacement for `infer(model = linear_gaussian_ssm_smoothing(A = A, B = B, P = P, Q = Q),
data = (y = observations,))` (benchmarks/...Benchmark.ipynb:186-196) over `batch` independent series.
`y` is `Array{Float32,3}` of size `(batch, m, T)` -- i.e. the C layout `y[T][m][batch]` seen from
column-major Julia -- so no transposition is needed for the data.
Returns `(mean[batch, d, T], cov[batch, d, d, T], neg_log_evidence[batch])`.
"""
function smooth(ctx::Context, y::Array{Float32,3}, A, B, P, Q, m0, S0; free_energy::Bool = false)
    batch, m, T = size(y)
    d = size(A, 1)
    mean = Array{Float32}(undef, batch, d, T)
    cov  = Array{Float32}(undef, batch, d, d, T)
    nle  = free_energy ? Vector{Float32}(undef, batch) : Float32[]
    Ar, Br, Pr, Qr, S0r = rowmajor(A), rowmajor(B), rowmajor(P), rowmajor(Q), rowmajor(S0)
    m0r = Vector{Float32}(m0)
    GC.@preserve y mean cov nle Ar Br Pr Qr S0r m0r begin
        rc = ccall((:rxg_lgssm_smooth_f32, LIB), Cint,
            (Ptr{Cvoid}, Cint, Cint, Cint, Int64,
             Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32},
             Ptr{Float32}, Ptr{UInt8}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Int32}, Cuint),
            ctx.handle, d, m, T, batch, Ar, Br, Pr, Qr, m0r, S0r, C_NULL,     # u = NULL: no transition offset
            y, C_NULL, mean, cov, free_energy ? pointer(nle) : C_NULL, C_NULL, 0)   # host pointers: flags = 0
        check(ctx, rc)
    end
    return mean, cov, nle
end

"""
    filter_chunk!(ctx, y, A, B, P, Q, prev_mean, carry_cov; u = nothing)

One time-chunk of the streaming engine: replaces `Tc` ticks of the `RxInferenceEngine` executor
(src/inference/streaming.jl:344-430) with `@autoupdates x_min_t_mean, x_min_t_cov = mean_cov(q(x_t))`
(src/inference/autoupdates.jl:614-659; run as in benchmarks/...Benchmark.ipynb:199-216).  `y`, `prev_mean`,
and the outputs are DEVICE arrays here (`CuPtr` reinterpreted as `Ptr`): `y[batch, m, Tc]`, `prev_mean[batch, d]`
= means of q(x_{t0-1}); `carry_cov` is a host `Matrix{Float32}` (d x d, symmetric, so row/column major agree),
updated in place.  Returns the device pointers of `(filt_mean[batch, d, Tc], filt_cov[batch, d, d, Tc])`;
the next chunk's `prev_mean` is the last time slice of `filt_mean`.
"""
function filter_chunk!(ctx::Context, y::Ptr{Float32}, dims::NTuple{3,Int}, A, B, P, Q, prev_mean::Ptr{Float32},
                       carry_cov::Matrix{Float32}, filt_mean::Ptr{Float32}, filt_cov::Ptr{Float32}; u = nothing)
    batch, m, Tc = dims
    d = size(A, 1)
    Ar, Br, Pr, Qr = rowmajor(A), rowmajor(B), rowmajor(P), rowmajor(Q)
    ur = u === nothing ? Float32[] : Vector{Float32}(u)
    GC.@preserve Ar Br Pr Qr ur carry_cov begin
        rc = ccall((:rxg_lgssm_filter_chunk_f32, LIB), Cint,
            (Ptr{Cvoid}, Cint, Cint, Cint, Int64, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32},
             Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Cuint),
            ctx.handle, d, m, Tc, batch, Ar, Br, Pr, Qr, u === nothing ? C_NULL : pointer(ur),
            prev_mean, carry_cov, y, filt_mean, filt_cov, C_NULL, RXG_PTR_DEVICE)
        check(ctx, rc)
    end
    return filt_mean, filt_cov
end

"""
    stream_vmp_gamma!(ctx, y, dims, out; iterations = 4, w = 1f0, init = (0f0, 1f3, 1f0, 1f0), prev = C_NULL, fe = C_NULL)

Time-chunk of the streaming `test_model1` (test/inference/inference_tests.jl:752-775: one-step random walk observed
with unknown precision τ ~ Gamma, `MeanField()`, priors autoupdated from `q(x_t)`, `q(τ)`).  Device pointers:
`y[batch, Tc]`, `out[batch, 4, Tc]` = (m_x, v_x, shape, rate) per datum, optional `fe[batch, iterations, Tc]`;
`prev[batch, 4]` = last slice of the previous chunk's `out` (then `init` is ignored).
"""
function stream_vmp_gamma!(ctx::Context, y::Ptr{Float32}, dims::NTuple{2,Int}, out::Ptr{Float32}; iterations::Integer = 4,
                           w::Float32 = 1f0, init = (0f0, 1f3, 1f0, 1f0), prev::Ptr{Float32} = Ptr{Float32}(C_NULL),
                           fe::Ptr{Float32} = Ptr{Float32}(C_NULL))
    batch, Tc = dims
    ini = Float32[init...]
    GC.@preserve ini begin
        check(ctx, ccall((:rxg_stream_vmp_gamma_f32, LIB), Cint,
            (Ptr{Cvoid}, Cint, Int64, Cint, Cfloat, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Cuint),
            ctx.handle, Tc, batch, iterations, w, ini, prev, y, out, fe, RXG_PTR_DEVICE))
    end
    return out
end

# --- multi-GPU (one Julia process per GPU): id from rank 0 broadcast by the host (MPI.jl / Distributed), then
#     rxg_comm_init; after the sweep one all-gather of the posterior marginals.  With a shared model pass
#     RXG_COV_REPLICATE: only the means cross NVLink, the chain-independent covariances are filled locally.
function allgather_posteriors!(ctx::Context, d, T, batch_local, mean::Ptr{Float32}, cov::Ptr{Float32},
                               gmean::Ptr{Float32}, gcov::Ptr{Float32}; shared_model::Bool = true)
    flags = RXG_PTR_DEVICE | (shared_model ? RXG_COV_REPLICATE : UInt32(0))
    check(ctx, ccall((:rxg_allgather_posteriors, LIB), Cint,
        (Ptr{Cvoid}, Cint, Cint, Int64, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Cuint),
        ctx.handle, d, T, batch_local, mean, cov, gmean, gcov, flags))
end

# --- drop-in for the result object: user code only calls mean./cov./var. on posteriors[:x]
#     (test/models/statespace/mlgssm_test.jl:121-126), so unpack into MvNormalMeanCovariance:
#
# function infer_batched(; model, data, free_energy = false, kwargs...)
#     pattern = recognise(model)                       # GraphPPL graph -> (A, B, P, Q, prior) or nothing
#     pattern === nothing && return RxInfer.infer(; model, data, free_energy, kwargs...)   # stock path
#     isempty(intersect(keys(kwargs), (:callbacks, :constraints, :meta, :predictvars, :annotations))) ||
#         return RxInfer.infer(; model, data, free_energy, kwargs...)                       # never silently ignore
#     μ, Σ, F = smooth(CTX[], pack(data.y), pattern...; free_energy)
#     posteriors = Dict(:x => [MvNormalMeanCovariance(Float64.(μ[b, :, t]), Float64.(Σ[b, :, :, t]))
#                              for t in axes(μ, 3), b in axes(μ, 1)])
#     return InferenceResult(posteriors, nothing, free_energy ? F : nothing, model, nothing)
# end
#
# --- per-rule drop-in: ReactiveMP dispatches rules by message type, so a batched message type
#     makes `@rule` bodies one-liners over the C ABI:
#
# struct BatchedMvNormalMeanCovariance{P}; μ::P; Σ::P; n::Int; d::Int; end     # device pointers, SoA
# @rule typeof(*)(:out, Marginalisation) (m_A::PointMass{<:AbstractMatrix}, m_in::BatchedMvNormalMeanCovariance) = begin
#     out = similar(m_in)
#     check(CTX[], ccall((:rxg_rule_mul_out_f32, LIB), Cint,
#         (Ptr{Cvoid}, Int64, Cint, Cint, Ptr{Float32}, Cint, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Cuint),
#         CTX[].handle, m_in.n, m_in.d, m_in.d, rowmajor(mean(m_A)), 1, m_in.μ, m_in.Σ, out.μ, out.Σ, RXG_PTR_DEVICE))
#     return out
# end
# BayesBase.prod(::GenericProd, l::BatchedMvNormalWeightedMeanPrecision, r::BatchedMvNormalWeightedMeanPrecision) =
#     ...ccall((:rxg_prod_gaussian_f32, LIB), ...)

end # module
