"""Small runs of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import diffeqgpu_b200 as dg  # noqa: E402
from cases import callback_sources, lorenz_sweep  # noqa: E402

f32 = np.float32
N = 3000
p = lorenz_sweep(N, seed=1)
prob = dg.ODEProblem(dg.models.lorenz, np.array([1, 0, 0], f32), (0.0, 3.0), np.array([10, 28, 8 / 3], f32))
probs = dg.ProblemBatch.from_arrays(prob, p=p, device="cuda:0")
sv = np.linspace(0, 3, 31).astype(f32)
for fp in ("fast", "strict"):
    for sched in ("queue", "static"):
        dg.vectorized_asolve(probs, prob, dg.GPUTsit5(), dt=f32(0.1), saveat=sv, abstol=f32(1e-6), reltol=f32(1e-6), fp_mode=fp,
                             schedule=sched, stats=True)
    dg.vectorized_asolve(probs, prob, dg.GPUTsit5(), dt=f32(0.1), save_everystep=False, fp_mode=fp)
    dg.vectorized_asolve(probs, prob, dg.GPURodas5P(), dt=f32(0.01), saveat=sv, fp_mode=fp)
    dg.vectorized_solve(probs, prob, dg.GPUTsit5(), dt=f32(0.01), fp_mode=fp)
    dg.vectorized_solve(probs, prob, dg.GPUVern7(), dt=f32(0.01), saveat=sv, fp_mode=fp)
# staged saves need a launch that fills the GPU
big = dg.ProblemBatch.from_arrays(prob, p=lorenz_sweep(148 * 256 * 3 + 77, seed=2), device="cuda:0")
prob_s = dg.ODEProblem(dg.models.lorenz, np.array([1, 0, 0], f32), (0.0, 0.5), np.array([10, 28, 8 / 3], f32))
dg.vectorized_solve(big, prob_s, dg.GPUTsit5(), dt=f32(0.01))
# lock-step fixed-dt kernel: both layouts, ragged sizes, short and long runs, Float64, end-point SDE loop
for fp in ("fast", "strict"):
    for n, tspan, dt in ((1, (0.0, 0.3), 0.1), (97, (0.0, 1.65), 0.1), (4099, (0.0, 10.0), 0.1), (641, (0.5, 3.0), 0.07)):
        pr = dg.ODEProblem(dg.models.lorenz, np.array([1, 0, 0], f32), tspan, np.array([10, 28, 8 / 3], f32))
        pb = dg.ProblemBatch.from_arrays(pr, p=lorenz_sweep(n, seed=3), device="cuda:0")
        for layout in ("ref", "soa"):
            dg.vectorized_solve(pb, pr, dg.GPUTsit5(), dt=f32(dt), fp_mode=fp, layout=layout, engine="lockstep")
        dg.vectorized_solve(pb, pr, dg.GPUVern9(), dt=f32(dt), fp_mode=fp, engine="lockstep")
    pr64 = dg.ODEProblem(dg.models.lorenz, np.array([1.0, 0, 0]), (0.0, 2.0), np.array([10, 28, 8 / 3]))
    pb64 = dg.ProblemBatch.from_arrays(pr64, p=lorenz_sweep(333, seed=4).astype(np.float64), device="cuda:0")
    dg.vectorized_solve(pb64, pr64, dg.GPUTsit5(), dt=0.05, fp_mode=fp, engine="lockstep")
    from cases import henon_heiles_u0  # noqa: E402
    for n_hh in (61, 1300):                  # 4-state Float32 (67.6 KB of staging per block) and its prepared-plan relaunch
        u0h = henon_heiles_u0(n_hh).astype(f32)
        prh = dg.ODEProblem(dg.models.henon_heiles, u0h[0], (0.0, 4.3), None)
        pbh = dg.ProblemBatch.from_arrays(prh, u0=u0h, device="cuda:0")
        plan = dg.vectorized_solve(pbh, prh, dg.GPUTsit5(), dt=f32(0.1), fp_mode=fp, prepare=True)
        plan(); plan()
    sp = dg.SDEProblem(dg.models.lorenz_additive, np.array([1, 0, 0], f32), (0.0, 0.05), np.array([10, 28, 8 / 3], f32), seed=5)
    dg.solve(dg.EnsembleProblem(sp, reduction=dg.EnsembleMoments()), dg.GPUEM(), dg.EnsembleGPUKernel(dev="cuda:0", fp_mode=fp),
             trajectories=1001, dt=f32(1e-3), save_everystep=False, adaptive=False)
cb = dg.DiscreteCallback(*callback_sources((("u_gt", 2, 30.0), ("u_scale", 2, 0.5))))
dg.vectorized_asolve(probs, prob, dg.GPUTsit5(), dt=f32(0.1), saveat=sv, callback=cb, tstops=[1.5])
dg.vectorized_asolve(probs, prob, dg.GPUKvaerno3(), dt=f32(0.01), save_everystep=False)
ts, us, st = dg.solve_host(prob, dg.GPUTsit5(), p=p, dt=f32(0.1), adaptive=True, abstol=1e-6, reltol=1e-6, saveat=sv, fp_mode="fast",
                           chunk_traj=700, stats=True)
torch.cuda.synchronize()
print("ok", float(us.sum()))
