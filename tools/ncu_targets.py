"""One launch of a BASELINE config for an ncu capture (development aid):  python tools/ncu_targets.py c1|c3|c4|c5 [fp]
Runs the solve twice (the first launch warms up / JIT-loads); capture the second with `ncu -k regex:<kernel> -s 1 -c 1`."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import diffeqgpu_b200 as dg  # noqa: E402

which = sys.argv[1]
fp = sys.argv[2] if len(sys.argv) > 2 else "fast"
dev = "cuda:0"
f32 = np.float32
P0 = np.array([10.0, 28.0, 8.0 / 3.0])
g = torch.Generator(device=dev).manual_seed(7)
for _ in range(2):
    if which == "c1":
        N = 1_000_000
        p = torch.rand((N, 3), generator=g, device=dev) * torch.tensor(P0, dtype=torch.float32, device=dev)
        prob = dg.ODEProblem(dg.models.lorenz, np.array([1, 0, 0], f32), (0.0, 10.0), P0.astype(f32))
        dg.vectorized_solve(dg.ProblemBatch.from_arrays(prob, p=p, device=dev), prob, dg.GPUTsit5(), dt=f32(0.1), fp_mode=fp)
    elif which == "c3":
        N = 1 << 18
        p = torch.rand((N, 3), generator=g, device=dev).double() * torch.tensor(P0, dtype=torch.float64, device=dev)
        prob = dg.ODEProblem(dg.models.lorenz, np.array([1.0, 0, 0]), (0.0, 10.0), P0)
        dg.vectorized_asolve(dg.ProblemBatch.from_arrays(prob, p=p, device=dev), prob, dg.GPUVern9(), dt=0.1, abstol=1e-10, reltol=1e-10,
                             save_everystep=False, fp_mode=fp)
    elif which == "c4":
        from cases import rober_sweep
        N = 1 << 20
        k = torch.as_tensor(rober_sweep(N), device=dev)
        prob = dg.ODEProblem(dg.models.rober, np.array([1, 0, 0], f32), (0.0, 1e5), np.array([0.04, 3e7, 1e4], f32))
        dg.vectorized_asolve(dg.ProblemBatch.from_arrays(prob, p=k, device=dev), prob, dg.GPURodas5P(), dt=f32(1e-4), abstol=f32(1e-8),
                             reltol=f32(1e-4), saveat=np.array([1.0, 10.0, 1e3, 1e5], f32), fp_mode=fp)
    elif which == "c5":
        N = 1_250_000
        prob = dg.SDEProblem(dg.models.lorenz_additive, np.array([1, 0, 0], f32), (0.0, 10.0), P0.astype(f32), seed=1234)
        red = torch.zeros((2, 3, 2), dtype=torch.float64, device=dev)
        dg.vectorized_solve(dg.ProblemBatch.from_arrays(prob, n_traj=N, device=dev), prob, dg.GPUEM(), dt=f32(1e-3), save_everystep=False,
                            fp_mode=fp, reduce=red)
    torch.cuda.synchronize()
