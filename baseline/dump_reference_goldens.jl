# dump_reference_goldens.jl -- run the REAL reference (DiffEqGPU.jl's EnsembleGPUKernel path on its CPU backend) on
# the golden inputs of this repository and write its outputs in the layout of tests/golden/oracle_golden.npz.
#
#   python tests/golden/export_inputs.py                 # tests/golden/golden_inputs.npz  (committed)
#   julia --project=<env with DiffEqGPU, OrdinaryDiffEq, StaticArrays, NPZ> baseline/dump_reference_goldens.jl
#   python -m pytest tests/test_reference_goldens.py     # oracle == reference, case by case
#
# Julia is not available in the build image or on the GPU box, so this script has not been run there; it uses only
# the reference's documented lower-level API (test/lower_level_api.jl:42-73, docs/src/tutorials/lower_level_api.md):
#   DiffEqGPU.vectorized_solve(probs, prob, alg; dt, saveat, save_everystep)
#   DiffEqGPU.vectorized_asolve(probs, prob, alg; dt, saveat, save_everystep, abstol, reltol)
# with `probs` a plain Vector{ImmutableODEProblem} (=> KernelAbstractions CPU backend).
# Output arrays: ts (len x N), us (len x N) of SVector -> written as (N, len) and (N, len, n) like the oracle's.
using DiffEqGPU, OrdinaryDiffEq, StaticArrays, NPZ

lorenz(u, p, t) = SVector{3}(p[1] * (u[2] - u[1]), u[1] * (p[2] - u[3]) - u[2], u[1] * u[2] - p[3] * u[3])
function lorenz_jac(u, p, t)
    return SMatrix{3, 3}(-p[1], p[2] - u[3], u[2], p[1], -one(eltype(u)), u[1], zero(eltype(u)), -u[1], -p[3])
end
lorenz_tgrad(u, p, t) = zero(u)
henon_heiles(u, p, t) = SVector{4}(u[3], u[4], -u[1] - 2 * u[1] * u[2], -u[2] - (u[1]^2 - u[2]^2))
rober(u, p, t) = SVector{3}(-p[1] * u[1] + p[3] * u[2] * u[3], p[1] * u[1] - p[2] * (u[2] * u[2]) - p[3] * u[2] * u[3], p[2] * (u[2] * u[2]))
function rober_jac(u, p, t)
    T = eltype(u)
    return SMatrix{3, 3}(-p[1], p[1], zero(T), p[3] * u[3], T(-2) * p[2] * u[2] - p[3] * u[3], T(2) * p[2] * u[2],
                         p[3] * u[2], -(p[3] * u[2]), zero(T))
end
rober_tgrad(u, p, t) = zero(u)
decay(u, p, t) = SVector{1}(-p[1] * u[1])
decay_jac(u, p, t) = SMatrix{1, 1}(-p[1])
decay_tgrad(u, p, t) = zero(u)

const FUNCS = Dict(0 => ODEFunction(lorenz, jac = lorenz_jac, tgrad = lorenz_tgrad), 1 => ODEFunction(henon_heiles),
                   2 => ODEFunction(rober, jac = rober_jac, tgrad = rober_tgrad), 3 => ODEFunction(decay, jac = decay_jac, tgrad = decay_tgrad))
const ALGS = Dict(0 => GPUTsit5(), 1 => GPUVern7(), 2 => GPUVern9(), 3 => GPURosenbrock23(), 4 => GPURodas4(), 5 => GPURodas5P())

inp = npzread(joinpath(@__DIR__, "..", "tests", "golden", "golden_inputs.npz"))
out = Dict{String, Any}()
for name in inp["names"]
    ids = inp["$name/ids"]; tol = inp["$name/tol"]
    T = ids[5] == 1 ? Float64 : Float32
    U0 = T.(inp["$name/u0"]); P = T.(inp["$name/p"]); tspan = T.(inp["$name/tspan"]); sv = T.(inp["$name/saveat"])
    N, n = size(U0); np_ = size(P, 2)
    f = FUNCS[Int(ids[1])]; alg = ALGS[Int(ids[2])]
    mk(i) = ODEProblem{false}(f, SVector{n, T}(U0[i, :]), (tspan[1], tspan[2]), np_ == 0 ? SVector{0, T}() : SVector{np_, T}(P[i, :]))
    prob = mk(1)
    probs = [DiffEqGPU.make_prob_compatible(mk(i)) for i in 1:N]
    saveat = isempty(sv) ? nothing : sv
    if ids[3] == 1
        ts, us = DiffEqGPU.vectorized_asolve(probs, prob, alg; dt = T(tol[1]), saveat = saveat,
                                             save_everystep = false,     # adaptive integrators never save every step (nonstiff/types.jl:378)
                                             abstol = T(tol[2]), reltol = T(tol[3]))
    else
        ts, us = DiffEqGPU.vectorized_solve(probs, prob, alg; dt = T(tol[1]), saveat = saveat, save_everystep = ids[4] == 1)
    end
    ts = Array(ts); us = Array(us)
    len = size(ts, 1)
    U = Array{T}(undef, N, len, n)
    for i in 1:N, k in 1:len, c in 1:n
        U[i, k, c] = us[k, i][c]
    end
    out["$name/ts"] = permutedims(ts, (2, 1))
    out["$name/us"] = U
    println(name, "  ", size(U))
end
npzwrite(joinpath(@__DIR__, "..", "tests", "golden", "reference_golden.npz"), out)
println("wrote tests/golden/reference_golden.npz")
