# run_reference.jl -- the reference's own CPU arms for the headline workload (SURVEY §8d), to be run
# where Julia + DiffEqGPU/OrdinaryDiffEq are installed (they are NOT in the build image or on the GPU
# box: bench.py --impl reference times the C++ oracle port instead and says so, kind = "port").
#
#   julia -t auto baseline/run_reference.jl [trajectories]
#
# Prints one JSON line per arm: attempted trajectory-steps/s of
#   (1) EnsembleGPUKernel(CPU()) + GPUTsit5   -- the reference kernel path on its CPU backend
#   (2) EnsembleThreads() + OrdinaryDiffEq.Tsit5
# on C2: Lorenz, p = rand(3) .* (10, 28, 8/3), adaptive abstol = reltol = 1f-6, saveat 0:1:10, Float32.
using DiffEqGPU, OrdinaryDiffEq, StaticArrays, Random

function lorenz(u, p, t)
    σ, ρ, β = p
    return SVector{3}(σ * (u[2] - u[1]), u[1] * (ρ - u[3]) - u[2], u[1] * u[2] - β * u[3])
end

N = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 1_000_000
Random.seed!(1234)
u0 = @SVector [1.0f0, 0.0f0, 0.0f0]
p0 = @SVector [10.0f0, 28.0f0, 8.0f0 / 3.0f0]
prob = ODEProblem{false}(lorenz, u0, (0.0f0, 10.0f0), p0)
ps = [rand(SVector{3, Float32}) .* p0 for _ in 1:N]
monteprob = EnsembleProblem(prob, prob_func = (prob, i, repeat) -> remake(prob, p = ps[i]), safetycopy = false)
saveat = Float32.(0:1:10)

function arm(name, f)
    f()                                   # compile
    t = @elapsed sol = f()
    # the kernel path keeps no step counters: count attempts with OrdinaryDiffEq's stats where available,
    # else report trajectories/s and let the caller scale by the mean attempts per trajectory (173.0 on C2)
    steps = try sum(s.stats.naccept + s.stats.nreject for s in sol.u) catch; round(Int, 173.0 * N) end
    println("{\"impl\": \"reference\", \"arm\": \"$name\", \"trajectories\": $N, \"seconds\": $t, ",
            "\"value\": $(steps / t), \"unit\": \"trajectory-steps/s\", \"cores\": $(Threads.nthreads())}")
end

arm("EnsembleGPUKernel(CPU()) GPUTsit5", () -> solve(monteprob, GPUTsit5(), EnsembleGPUKernel(CPU()); trajectories = N,
    adaptive = true, dt = 0.1f0, abstol = 1.0f-6, reltol = 1.0f-6, saveat = saveat))
arm("EnsembleThreads() Tsit5", () -> solve(monteprob, Tsit5(), EnsembleThreads(); trajectories = N,
    adaptive = true, dt = 0.1f0, abstol = 1.0f-6, reltol = 1.0f-6, saveat = saveat))
