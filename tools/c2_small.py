"""Adaptive Lorenz GPUTsit5 (tol 1e-6, saveat 0:1:10) over small ensemble sizes, graph-timed: strict (one trajectory per
thread) against fast (two per thread, packed); development aid."""
import json, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import diffeqgpu_b200 as dg  # noqa: E402
dev = "cuda:0"
P0 = np.array([10.0, 28.0, 8.0 / 3.0], np.float32)
sv = np.arange(0, 11, dtype=np.float32)
for N in ([int(float(x)) for x in sys.argv[1:]] or (1000, 10_000, 30_000, 100_000, 300_000, 1_000_000)):
    K = max(3, min(100, int(3e6 // N)))
    g = torch.Generator(device=dev).manual_seed(7)
    p = torch.rand((N, 3), generator=g, device=dev) * torch.tensor(P0, device=dev)
    prob = dg.ODEProblem(dg.models.lorenz, np.array([1, 0, 0], np.float32), (0.0, 10.0), P0)
    probs = dg.ProblemBatch.from_arrays(prob, p=p, device=dev)
    import os
    for fp, engine, w in (("strict", "auto", None), ("fast", "auto", "w1"), ("fast", "auto", "w2")):
        if True:
            os.environ["DEGK_ADAPTIVE_W1_BELOW"] = "0" if w == "w2" else str(1 << 40)
            plan = dg.vectorized_asolve(probs, prob, dg.GPUTsit5(), dt=np.float32(0.1), abstol=np.float32(1e-6), reltol=np.float32(1e-6),
                                        saveat=sv, fp_mode=fp, engine=engine, stats=True, prepare=True)
            plan(); torch.cuda.synchronize()
            att = int(plan.stats["totals"][0] + plan.stats["totals"][1])
            plan.capture(K); plan.replay(); torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); plan.replay(); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / K)
            print(json.dumps(dict(N=N, fp=fp, w=w or '1', us_per_solve=round(best * 1e3, 1), gsteps_per_s=round(att / best / 1e6, 2))), flush=True)
