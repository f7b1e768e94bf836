# DiffEqGPUDegkExt.jl -- CUDA-only package extension that routes DiffEqGPU.jl's EnsembleGPUKernel
# operator API through libdegk (include/degk.h).  SOURCE ONLY: Julia is not available in the build
# image, so this file has not been executed; it documents the binding a maintainer would add as
# `ext/DegkExt.jl` (replacing ext/CUDAExt.jl:9-10 for this path).
#
# Methods added:
#   DiffEqGPU.vectorized_solve(probs::DegkBatch, prob::ODEProblem, alg; ...)   lowerlevel_solve.jl:53-131
#   DiffEqGPU.vectorized_solve(probs::DegkBatch, prob::SDEProblem, alg; ...)   lowerlevel_solve.jl:134-199
#   DiffEqGPU.vectorized_asolve(probs::DegkBatch, prob::ODEProblem, alg; ...)  lowerlevel_solve.jl:253-346
# plus `Adapt.adapt(::DegkBackend, probs)` which turns the AoS Vector{ImmutableODEProblem} into the
# three strided device arrays the C ABI takes (u0, p, tspan).
module DiffEqGPUDegkExt

using DiffEqGPU, CUDA, SciMLBase, StaticArrays
import DiffEqGPU: vectorized_solve, vectorized_asolve, GPUTsit5, GPUVern7, GPUVern9,
                  GPURosenbrock23, GPURodas4, GPURodas5P, GPUEM, GPUSIEA

const libdegk = get(ENV, "DEGK_LIBRARY", "libdegk.so")

# ---- mirrors of the C structs (field order == include/degk.h) -------------------------------
struct ModelDesc
    builtin::Cstring; rhs_src::Cstring; jac_src::Cstring; tgrad_src::Cstring; noise_src::Cstring
    n_state::Int32; n_param::Int32; n_noise::Int32; noise_kind::Int32
    dtype::Int32; alg::Int32; fp_mode::Int32; force_jit::Int32
    events::Int32; n_callbacks::Int32                 # tstops / GPUDiscreteCallback lowering (degk.h)
    cb_condition_src::Ptr{Cstring}; cb_affect_src::Ptr{Cstring}
    jac_mode::Int32; reserved::Int32                  # 0 analytic/default, 1 finite differences, 2 ForwardDiff-style duals
    mass_src::Cstring                                 # constant mass matrix body (stiff solvers), C_NULL = identity
    n_ccallbacks::Int32; reserved3::Int32             # GPUContinuousCallback lowering
    cc_condition_src::Ptr{Cstring}; cc_affect_src::Ptr{Cstring}; cc_affect_neg_src::Ptr{Cstring}
    cc_rootfind::Ptr{Int32}; cc_abstol::Ptr{Float64}; cc_repeat_nudge::Ptr{Float64}; cc_dtrelax::Ptr{Float64}
end

struct SolveArgs
    n_traj::Int64; traj_offset::Int64
    u0::CuPtr{Cvoid}; u0_stride::Int64
    p::CuPtr{Cvoid}; p_stride::Int64
    tspan::CuPtr{Cvoid}; tspan_stride::Int64
    dt::Float64; adaptive::Int32
    abstol::Float64; reltol::Float64
    saveat::CuPtr{Cvoid}; n_saveat::Int32; save_everystep::Int32
    n_rows::Int64; us::CuPtr{Cvoid}; ts::CuPtr{Cvoid}
    out_layout::Int32; schedule::Int32
    retcode::CuPtr{Int32}; naccept::CuPtr{Int32}; nreject::CuPtr{Int32}
    seed::UInt64; reduce::CuPtr{Float64}; totals::CuPtr{UInt64}
    max_iters::Int64; engine::Int32; reserved::Int32
    tstops::CuPtr{Cvoid}; n_tstops::Int32; reserved2::Int32
    nsaved::CuPtr{Int32}
end

alg_id(::GPUTsit5) = 0; alg_id(::GPUVern7) = 1; alg_id(::GPUVern9) = 2
alg_id(::GPURosenbrock23) = 3; alg_id(::GPURodas4) = 4; alg_id(::GPURodas5P) = 5
alg_id(::GPUEM) = 6; alg_id(::GPUSIEA) = 7; alg_id(::GPUKvaerno3) = 8; alg_id(::GPUKvaerno5) = 9
dtype_id(::Type{Float32}) = 0; dtype_id(::Type{Float64}) = 1

check(ctx, st) = st == 0 || error(unsafe_string(ccall((:degk_last_error, libdegk), Cstring, (Ptr{Cvoid},), ctx)))

const CTX = Ref{Ptr{Cvoid}}(C_NULL)
function context()
    if CTX[] == C_NULL
        st = ccall((:degk_ctx_create, libdegk), Cint, (Cint, Ptr{Ptr{Cvoid}}), CUDA.deviceid(), CTX)
        check(C_NULL, st)
    end
    return CTX[]
end

"""
    DegkFunction(; builtin = nothing, rhs, jac = nothing, tgrad = nothing, noise = nothing)

CUDA C++ bodies for f / jac / tgrad / g.  With ModelingToolkit/Symbolics these come from
`build_function(rhs_exprs, u, p, t; target = Symbolics.CTarget())` (0-based `du[i] = ...`).
"""
Base.@kwdef struct DegkFunction
    builtin::Union{Nothing, String} = nothing
    rhs::Union{Nothing, String} = nothing
    jac::Union{Nothing, String} = nothing
    tgrad::Union{Nothing, String} = nothing
    noise::Union{Nothing, String} = nothing
end

# device-side batch: what `adapt(dev, probs)` produces for this backend
struct DegkBatch{T}
    u0::CuMatrix{T}      # n × N (column i = trajectory i  => stride n)
    p::CuMatrix{T}       # np × N
    tspan::CuMatrix{T}   # 2 × N or 2 × 1 (broadcast)
    f::DegkFunction
    seed::UInt64
end
Base.length(b::DegkBatch) = size(b.u0, 2)

function DegkBatch(probs::AbstractVector, f::DegkFunction)   # host loop #1 of src/solve.jl:187-202
    T = eltype(probs[1].u0)
    u0 = CuArray(reduce(hcat, [collect(T, pr.u0) for pr in probs]))
    p = CuArray(reduce(hcat, [collect(T, pr.p) for pr in probs]))
    same = all(pr -> pr.tspan == probs[1].tspan, probs)
    ts = same ? CuArray(reshape(collect(T, probs[1].tspan), 2, 1)) :
         CuArray(reduce(hcat, [collect(T, pr.tspan) for pr in probs]))
    seed = hasproperty(probs[1], :seed) ? UInt64(probs[1].seed) : UInt64(0)
    return DegkBatch{T}(u0, p, ts, f, seed)
end

const PROGRAMS = Dict{Any, Ptr{Cvoid}}()
function program(b::DegkBatch{T}, alg, fp_mode, noise_kind = 0, n_noise = 0) where {T}
    key = (b.f, typeof(alg), T, fp_mode)
    get!(PROGRAMS, key) do
        cs(x) = x === nothing ? Cstring(C_NULL) : Base.unsafe_convert(Cstring, x)
        GC.@preserve b begin
            desc = Ref(ModelDesc(cs(b.f.builtin), cs(b.f.rhs), cs(b.f.jac), cs(b.f.tgrad), cs(b.f.noise),
                                 size(b.u0, 1), size(b.p, 1), n_noise, noise_kind,
                                 dtype_id(T), alg_id(alg), fp_mode, 0))
            out = Ref{Ptr{Cvoid}}(C_NULL)
            check(context(), ccall((:degk_program_build, libdegk), Cint,
                                   (Ptr{Cvoid}, Ptr{ModelDesc}, Ptr{Ptr{Cvoid}}), context(), desc, out))
            out[]
        end
    end
end

function _launch(b::DegkBatch{T}, prob, alg; dt, adaptive, abstol, reltol, saveat, save_everystep,
                 fp_mode = 0) where {T}
    N, n = length(b), size(b.u0, 1)
    t0, tf = T.(prob.tspan)
    nsave = saveat === nothing ? 0 : length(saveat)
    len = ccall((:degk_output_rows, libdegk), Int64, (Cint, Cdouble, Cdouble, Cdouble, Cint, Cint, Cint),
                dtype_id(T), t0, tf, dt, adaptive, save_everystep, nsave)
    # same shapes as lowerlevel_solve.jl:81-83: (len × N), trajectory i = column i
    ts = CuMatrix{T}(undef, len, N)
    us = CuMatrix{SVector{n, T}}(undef, len, N)
    d_saveat = saveat === nothing ? nothing : CuArray(T.(collect(saveat)))
    args = Ref(SolveArgs(N, 0, pointer(b.u0), n, pointer(b.p), size(b.p, 1),
                         pointer(b.tspan), size(b.tspan, 2) == 1 ? 0 : 2,
                         Float64(T(dt)), adaptive, Float64(T(abstol)), Float64(T(reltol)),
                         d_saveat === nothing ? CU_NULL : pointer(d_saveat), nsave, save_everystep,
                         len, reinterpret(CuPtr{Cvoid}, pointer(us)), reinterpret(CuPtr{Cvoid}, pointer(ts)),
                         0, 2, CU_NULL, CU_NULL, CU_NULL, b.seed, CU_NULL, CU_NULL, 0, 0, 0))
    prog = program(b, alg, fp_mode)
    GC.@preserve b d_saveat check(context(), ccall((:degk_solve, libdegk), Cint,
        (Ptr{Cvoid}, Ptr{SolveArgs}, CUDA.CUstream), prog, args, CUDA.stream().handle))
    return ts, us                      # still on the device, asynchronous, like the reference
end

function vectorized_solve(probs::DegkBatch, prob::Union{ODEProblem, SDEProblem}, alg;
                          dt, saveat = nothing, save_everystep = true, debug = false, kwargs...)
    sv = saveat === nothing ? nothing : DiffEqGPU._convert_saveat(saveat, prob)   # :84-109
    _launch(probs, prob, alg; dt, adaptive = 0, abstol = 0, reltol = 0, saveat = sv,
            save_everystep = Int32(save_everystep))
end

function vectorized_asolve(probs::DegkBatch, prob::ODEProblem, alg;
                           dt = 0.1f0, saveat = nothing, save_everystep = false,
                           abstol = 1.0f-6, reltol = 1.0f-3, debug = false, kwargs...)
    sv = saveat === nothing ? nothing : DiffEqGPU._convert_saveat(saveat, prob)   # :271-306
    _launch(probs, prob, alg; dt, adaptive = 1, abstol, reltol, saveat = sv,
            save_everystep = Int32(save_everystep))
end

vectorized_asolve(probs::DegkBatch, prob::SDEProblem, alg; kwargs...) =
    error("Adaptive time-stepping is not supported yet with GPUEM.")            # :348-356

end # module
