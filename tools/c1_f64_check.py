"""Float64 / Vern9 / Henon-Heiles spot check of the engine choice for fixed-dt runs (v1 against lock-step), graph-timed."""
import json, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import diffeqgpu_b200 as dg  # noqa: E402
from cases import henon_heiles_u0  # noqa: E402
dev = "cuda:0"
for name, dtype, alg in (("lorenz", np.float64, dg.GPUTsit5()), ("lorenz", np.float32, dg.GPUVern9()), ("henon", np.float32, dg.GPUTsit5()), ("lorenz", np.float64, dg.GPUVern9())):
    for N in ([int(float(x)) for x in sys.argv[1:]] or (1000, 10_000, 100_000, 1_000_000)):
        K = max(3, min(100, int(1e7 // N)))
        if name == "lorenz":
            P0 = np.array([10.0, 28.0, 8.0 / 3.0], dtype)
            p = (torch.rand((N, 3), device=dev, dtype=torch.float64) * torch.tensor(P0, device=dev)).to(torch.float64 if dtype == np.float64 else torch.float32)
            prob = dg.ODEProblem(dg.models.lorenz, np.array([1, 0, 0], dtype), (0.0, 10.0), P0)
            probs = dg.ProblemBatch.from_arrays(prob, p=p, device=dev)
        else:
            u0 = henon_heiles_u0(N).astype(dtype)
            prob = dg.ODEProblem(dg.models.henon_heiles, u0[0], (0.0, 10.0), None)
            probs = dg.ProblemBatch.from_arrays(prob, u0=u0, device=dev)
        for fp in ("strict", "fast"):
            r = {}
            for engine in ("v1", "lockstep"):
                plan = dg.vectorized_solve(probs, prob, alg, dt=dtype(0.1), fp_mode=fp, engine=engine, prepare=True)
                plan(); torch.cuda.synchronize(); plan.capture(K); plan.replay(); torch.cuda.synchronize()
                best = 1e9
                for _ in range(3):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); plan.replay(); e1.record(); torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1) / K)
                r[engine] = round(best * 1e3, 2)
            print(json.dumps(dict(model=name, dtype=np.dtype(dtype).name, alg=type(alg).__name__, N=N, fp=fp, us=r)), flush=True)
