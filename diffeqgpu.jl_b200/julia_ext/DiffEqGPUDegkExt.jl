# DiffEqGPUDegkExt.jl -- CUDA-only package extension that routes DiffEqGPU.jl's EnsembleGPUKernel
# operator API through libdegk (include/degk.h).  SOURCE ONLY: Julia is not available in the build
# image, so this file has not been executed; it documents the binding a maintainer would add as
# `ext/DegkExt.jl` (replacing ext/CUDAExt.jl:9-10 for this path).  tests/test_abi.py checks, against
# include/degk.h, the field order of the two struct mirrors AND that every constructor call below names only
# existing fields (the structs are keyword-constructed with defaults, so they cannot go stale by arity).
#
# Methods added:
#   DiffEqGPU.vectorized_solve(probs::DegkBatch, prob::ODEProblem, alg; ...)   lowerlevel_solve.jl:53-131
#   DiffEqGPU.vectorized_solve(probs::DegkBatch, prob::SDEProblem, alg; ...)   lowerlevel_solve.jl:134-199
#   DiffEqGPU.vectorized_asolve(probs::DegkBatch, prob::ODEProblem, alg; ...)  lowerlevel_solve.jl:253-346
# plus `DegkBatch(probs, f)` which turns the AoS Vector{ImmutableODEProblem} into the three strided device
# arrays the C ABI takes (u0, p, tspan), and `lower_symbolic` / `DegkFunction(sys)` which lower a
# Symbolics / ModelingToolkit right-hand side to the CUDA C++ bodies NVRTC inlines into the stepper kernels.
module DiffEqGPUDegkExt

using DiffEqGPU, CUDA, SciMLBase, StaticArrays
import DiffEqGPU: vectorized_solve, vectorized_asolve, GPUTsit5, GPUVern7, GPUVern9,
                  GPURosenbrock23, GPURodas4, GPURodas5P, GPUEM, GPUSIEA, GPUKvaerno3, GPUKvaerno5

const libdegk = get(ENV, "DEGK_LIBRARY", "libdegk.so")
const NULLSTR = Cstring(C_NULL)

# ---- mirrors of the C structs (field order == include/degk.h; every field has a default) ----------
Base.@kwdef struct ModelDesc
    builtin::Cstring = NULLSTR; rhs_src::Cstring = NULLSTR; jac_src::Cstring = NULLSTR; tgrad_src::Cstring = NULLSTR; noise_src::Cstring = NULLSTR
    n_state::Int32 = 0; n_param::Int32 = 0; n_noise::Int32 = 0; noise_kind::Int32 = 0
    dtype::Int32 = 0; alg::Int32 = 0; fp_mode::Int32 = 0; force_jit::Int32 = 0
    events::Int32 = 0; n_callbacks::Int32 = 0                 # tstops / GPUDiscreteCallback lowering (degk.h)
    cb_condition_src::Ptr{Cstring} = C_NULL; cb_affect_src::Ptr{Cstring} = C_NULL
    jac_mode::Int32 = 0; reserved::Int32 = 0                  # 0 analytic/default, 1 finite differences, 2 ForwardDiff-style duals
    mass_src::Cstring = NULLSTR                               # constant mass matrix body (stiff solvers), C_NULL = identity
    n_ccallbacks::Int32 = 0; reserved3::Int32 = 0             # GPUContinuousCallback lowering
    cc_condition_src::Ptr{Cstring} = C_NULL; cc_affect_src::Ptr{Cstring} = C_NULL; cc_affect_neg_src::Ptr{Cstring} = C_NULL
    cc_rootfind::Ptr{Int32} = C_NULL; cc_abstol::Ptr{Float64} = C_NULL; cc_repeat_nudge::Ptr{Float64} = C_NULL; cc_dtrelax::Ptr{Float64} = C_NULL
end

Base.@kwdef struct SolveArgs
    n_traj::Int64 = 0; traj_offset::Int64 = 0
    u0::CuPtr{Cvoid} = CU_NULL; u0_stride::Int64 = 0
    p::CuPtr{Cvoid} = CU_NULL; p_stride::Int64 = 0
    tspan::CuPtr{Cvoid} = CU_NULL; tspan_stride::Int64 = 0
    dt::Float64 = 0.0; adaptive::Int32 = 0
    abstol::Float64 = 0.0; reltol::Float64 = 0.0
    saveat::CuPtr{Cvoid} = CU_NULL; n_saveat::Int32 = 0; save_everystep::Int32 = 0
    n_rows::Int64 = 0; us::CuPtr{Cvoid} = CU_NULL; ts::CuPtr{Cvoid} = CU_NULL
    out_layout::Int32 = 0; schedule::Int32 = 2
    retcode::CuPtr{Int32} = CU_NULL; naccept::CuPtr{Int32} = CU_NULL; nreject::CuPtr{Int32} = CU_NULL
    seed::UInt64 = 0; reduce::CuPtr{Float64} = CU_NULL; totals::CuPtr{UInt64} = CU_NULL
    max_iters::Int64 = 0; engine::Int32 = 0; dae_init::Int32 = 0   # engine: degk_engine (0 auto, 1 per-thread kernels, 2 lock-step fixed dt)
    tstops::CuPtr{Cvoid} = CU_NULL; n_tstops::Int32 = 0; reserved2::Int32 = 0
    nsaved::CuPtr{Int32} = CU_NULL
    saveat_stride::Int64 = 0                                  # per-problem saveat grids (kernels.jl:15-17): elements between two grids
    order::CuPtr{Int32} = CU_NULL                             # optional start order of the trajectories (sorted by a parameter)
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
    DegkFunction(; builtin = nothing, rhs, jac = nothing, tgrad = nothing, noise = nothing, mass = nothing)

CUDA C++ bodies for f / jac / tgrad / g / the constant mass matrix (`du[i] = ...`, `J[i][j] = ...`, `dT[i] = ...`,
`g[i] = ...`, `Mm[i][j] = ...`; 0-based, scalar type `T`).  `lower_symbolic` produces them from symbolic
expressions, `DegkFunction(sys::ODESystem)` from a ModelingToolkit system.
"""
Base.@kwdef struct DegkFunction
    builtin::Union{Nothing, String} = nothing
    rhs::Union{Nothing, String} = nothing
    jac::Union{Nothing, String} = nothing
    tgrad::Union{Nothing, String} = nothing
    noise::Union{Nothing, String} = nothing
    mass::Union{Nothing, String} = nothing
end

"""
    DegkCallback(condition, affect)        # GPUDiscreteCallback:  `return <bool of u, p, t>;` / statements on u, p
    DegkContinuousCallback(condition, affect; affect_neg = affect, rootfind = 1)   # `return <value of u, p, t>;`

Callback bodies in CUDA C++ (the lowering target of the Julia closures, like the RHS); `terminate();` ends the
trajectory (callbacks.jl:1-36, integrator_utils.jl:69-150, 229-442).
"""
struct DegkCallback; condition::String; affect::String; end
Base.@kwdef struct DegkContinuousCallback
    condition::String; affect::Union{Nothing, String}; affect_neg::Union{Nothing, String} = affect; rootfind::Int32 = 1
end

# ---- Symbolics / ModelingToolkit lowering (docs/src/tutorials/modelingtoolkit.md) --------------------
# `Symbolics.build_function(exprs, u, p, t; target = Symbolics.CTarget())` prints
#     void diffeqf(double* du, const double* RHS1, const double* RHS2, const double RHS3) { du[0] = ...; ... }
# The body is taken as is, the three inputs are renamed to u / p / t, and floating literals get the kernel's scalar
# type (`1.5` -> `(T)(1.5)`, otherwise Float32 kernels would be promoted to double arithmetic).
function c_body(csrc::AbstractString, out::AbstractString)
    body = match(r"\{(.*)\}"s, csrc).captures[1]
    body = replace(body, "RHS1" => "u", "RHS2" => "p", "RHS3" => "t", "du[" => out * "[")
    body = replace(body, r"(?<![\w.\]])(\d+\.\d*(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)" => s"(T)(\1)")
    return replace(body, "pow(" => "pow_(")
end

"""
    lower_symbolic(rhs, u, p, t; jac = true) -> DegkFunction

`rhs`: vector of Symbolics expressions du_i(u, p, t).  With `jac = true` the analytic Jacobian and time gradient are
derived symbolically (what `ODEFunction(sys; jac = true, tgrad = true)` does) and lowered too, as flat `J[i][j]` entries.
"""
function lower_symbolic(rhs, u, p, t; jac::Bool = true, Symbolics = Main.Symbolics)
    f_c = Symbolics.build_function(rhs, u, p, t; target = Symbolics.CTarget())
    fn = DegkFunction(rhs = c_body(f_c, "du"))
    jac || return fn
    J = Symbolics.jacobian(rhs, u)
    n = length(rhs)
    J_c = Symbolics.build_function(vec(permutedims(J)), u, p, t; target = Symbolics.CTarget())   # row-major flat
    jbody = c_body(J_c, "Jflat")
    jsrc = "    T Jflat[$(n * n)];\n" * jbody * "\n    for (int i = 0; i < $n; ++i) for (int j = 0; j < $n; ++j) J[i][j] = Jflat[i * $n + j];\n"
    dT = Symbolics.derivative.(rhs, (t,))
    # an autonomous right-hand side: no time-gradient body at all -- a Jacobian body without one tells the library that
    # dT == 0, and the fast build drops the dT terms of the Rosenbrock stages (TGRAD_ZERO, degk_jit.cpp)
    all(iszero, dT) && return DegkFunction(rhs = fn.rhs, jac = jsrc)
    T_c = Symbolics.build_function(dT, u, p, t; target = Symbolics.CTarget())
    return DegkFunction(rhs = fn.rhs, jac = jsrc, tgrad = c_body(T_c, "dT"))
end

"DegkFunction(sys): lower a ModelingToolkit ODESystem (structurally simplified) like `ODEProblem{false}(sys, ...)` would use it"
function DegkFunction(sys; jac::Bool = true, MTK = Main.ModelingToolkit, Symbolics = Main.Symbolics)
    eqs = MTK.full_equations(sys)
    rhs = [eq.rhs for eq in eqs]
    return lower_symbolic(rhs, MTK.unknowns(sys), MTK.parameters(sys), MTK.get_iv(sys); jac, Symbolics)
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

cs(x) = x === nothing ? NULLSTR : Base.unsafe_convert(Cstring, x)

const PROGRAMS = Dict{Any, Ptr{Cvoid}}()
function program(b::DegkBatch{T}, alg, fp_mode; noise_kind = 0, n_noise = 0, callbacks = DegkCallback[],
                 ccallbacks = DegkContinuousCallback[], events = false) where {T}
    key = (b.f, typeof(alg), T, fp_mode, callbacks, ccallbacks, events)
    get!(PROGRAMS, key) do
        conds = [cs(c.condition) for c in callbacks]; affs = [cs(c.affect) for c in callbacks]
        cconds = [cs(c.condition) for c in ccallbacks]; caffs = [cs(c.affect) for c in ccallbacks]
        cnegs = [cs(c.affect_neg) for c in ccallbacks]; croot = Int32[c.rootfind for c in ccallbacks]
        GC.@preserve b callbacks ccallbacks conds affs cconds caffs cnegs croot begin
            desc = Ref(ModelDesc(builtin = cs(b.f.builtin), rhs_src = cs(b.f.rhs), jac_src = cs(b.f.jac), tgrad_src = cs(b.f.tgrad),
                                 noise_src = cs(b.f.noise), mass_src = cs(b.f.mass),
                                 n_state = size(b.u0, 1), n_param = size(b.p, 1), n_noise = n_noise, noise_kind = noise_kind,
                                 dtype = dtype_id(T), alg = alg_id(alg), fp_mode = fp_mode,
                                 events = Int32(events || !isempty(callbacks) || !isempty(ccallbacks)),
                                 n_callbacks = length(callbacks),
                                 cb_condition_src = isempty(conds) ? C_NULL : pointer(conds),
                                 cb_affect_src = isempty(affs) ? C_NULL : pointer(affs),
                                 n_ccallbacks = length(ccallbacks),
                                 cc_condition_src = isempty(cconds) ? C_NULL : pointer(cconds),
                                 cc_affect_src = isempty(caffs) ? C_NULL : pointer(caffs),
                                 cc_affect_neg_src = isempty(cnegs) ? C_NULL : pointer(cnegs),
                                 cc_rootfind = isempty(croot) ? C_NULL : pointer(croot)))
            out = Ref{Ptr{Cvoid}}(C_NULL)
            check(context(), ccall((:degk_program_build, libdegk), Cint,
                                   (Ptr{Cvoid}, Ptr{ModelDesc}, Ptr{Ptr{Cvoid}}), context(), desc, out))
            out[]
        end
    end
end

# saveat normalisation, restated from lowerlevel_solve.jl:84-109 (fixed dt, also the SDE method :159-178) and
# :271-306 (adaptive: the same plus the 100 000-point guard)
function convert_saveat(saveat, prob, adaptive::Bool)
    saveat === nothing && return nothing
    Tt = eltype(prob.tspan)
    if saveat isa AbstractRange
        return Tt.(collect(range(Tt(first(saveat)), Tt(last(saveat)), length = length(saveat))))
    elseif saveat isa AbstractVector
        return Tt.(collect(saveat))
    end
    t0, tf = Tt.(prob.tspan)
    Tt(saveat) == Tt(0.0) && return Tt.([t0, tf])
    num_points = Int(ceil(abs(tf - t0) / abs(Tt(saveat)))) + 1
    if adaptive && num_points > 100_000
        error("saveat would create too many save points ($num_points). Consider using a larger saveat value.")
    end
    return Tt.(collect(range(t0, tf, length = num_points)))
end

function _launch(b::DegkBatch{T}, prob, alg; dt, adaptive, abstol, reltol, saveat, save_everystep,
                 fp_mode = 0, tstops = nothing, callbacks = DegkCallback[], ccallbacks = DegkContinuousCallback[]) where {T}
    N, n = length(b), size(b.u0, 1)
    t0, tf = T.(prob.tspan)
    nsave = saveat === nothing ? 0 : length(saveat)
    len = ccall((:degk_output_rows, libdegk), Int64, (Cint, Cdouble, Cdouble, Cdouble, Cint, Cint, Cint),
                dtype_id(T), t0, tf, dt, adaptive, save_everystep, nsave)
    if tstops !== nothing && saveat === nothing && save_everystep != 0 && adaptive == 0
        len += length(tstops) - count(x -> x in tstops, t0:T(dt):tf)       # lowerlevel_solve.jl:74-77
    end
    # same shapes as lowerlevel_solve.jl:81-83: (len × N), trajectory i = column i
    ts = CuMatrix{T}(undef, len, N)
    us = CuMatrix{SVector{n, T}}(undef, len, N)
    d_saveat = saveat === nothing ? nothing : CuArray(T.(collect(saveat)))
    d_tstops = tstops === nothing ? nothing : CuArray(T.(collect(tstops)))
    args = Ref(SolveArgs(n_traj = N, u0 = reinterpret(CuPtr{Cvoid}, pointer(b.u0)), u0_stride = n,
                         p = reinterpret(CuPtr{Cvoid}, pointer(b.p)), p_stride = size(b.p, 1),
                         tspan = reinterpret(CuPtr{Cvoid}, pointer(b.tspan)), tspan_stride = size(b.tspan, 2) == 1 ? 0 : 2,
                         dt = Float64(T(dt)), adaptive = adaptive, abstol = Float64(T(abstol)), reltol = Float64(T(reltol)),
                         saveat = d_saveat === nothing ? CU_NULL : reinterpret(CuPtr{Cvoid}, pointer(d_saveat)), n_saveat = nsave,
                         save_everystep = save_everystep, n_rows = len,
                         us = reinterpret(CuPtr{Cvoid}, pointer(us)), ts = reinterpret(CuPtr{Cvoid}, pointer(ts)),
                         seed = b.seed,
                         tstops = d_tstops === nothing ? CU_NULL : reinterpret(CuPtr{Cvoid}, pointer(d_tstops)),
                         n_tstops = d_tstops === nothing ? 0 : length(d_tstops)))
    prog = program(b, alg, fp_mode; callbacks, ccallbacks, events = tstops !== nothing,
                   noise_kind = prob isa SDEProblem ? (SciMLBase.is_diagonal_noise(prob) ? 1 : 2) : 0,
                   n_noise = prob isa SDEProblem && !SciMLBase.is_diagonal_noise(prob) ? size(prob.noise_rate_prototype, 2) : 0)
    GC.@preserve b d_saveat d_tstops check(context(), ccall((:degk_solve, libdegk), Cint,
        (Ptr{Cvoid}, Ptr{SolveArgs}, CUDA.CUstream), prog, args, CUDA.stream().handle))
    return ts, us                      # still on the device, asynchronous, like the reference
end

function vectorized_solve(probs::DegkBatch, prob::Union{ODEProblem, SDEProblem}, alg;
                          dt, saveat = nothing, save_everystep = true, debug = false,
                          tstops = nothing, callbacks = DegkCallback[], ccallbacks = DegkContinuousCallback[], kwargs...)
    if prob isa SDEProblem && alg isa GPUSIEA && !SciMLBase.is_diagonal_noise(prob)
        error("The algorithm is not compatible with the chosen noise type. Please see the documentation on the solver methods")   # :185-186
    end
    sv = convert_saveat(saveat, prob, false)
    _launch(probs, prob, alg; dt, adaptive = Int32(0), abstol = 0, reltol = 0, saveat = sv,
            save_everystep = Int32(save_everystep), tstops, callbacks, ccallbacks)
end

function vectorized_asolve(probs::DegkBatch, prob::ODEProblem, alg;
                           dt = 0.1f0, saveat = nothing, save_everystep = false,
                           abstol = 1.0f-6, reltol = 1.0f-3, debug = false,
                           tstops = nothing, callbacks = DegkCallback[], ccallbacks = DegkContinuousCallback[], kwargs...)
    sv = convert_saveat(saveat, prob, true)
    _launch(probs, prob, alg; dt, adaptive = Int32(1), abstol, reltol, saveat = sv,
            save_everystep = Int32(save_everystep), tstops, callbacks, ccallbacks)
end

vectorized_asolve(probs::DegkBatch, prob::SDEProblem, alg; kwargs...) =
    error("Adaptive time-stepping is not supported yet with GPUEM.")            # :348-356

end # module
