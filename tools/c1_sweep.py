"""Fixed-dt Lorenz GPUTsit5 (dt = 0.1, tspan 0-10, every-step saves, reference layout) over ensemble sizes: the
one-thread-per-trajectory kernel against the lock-step kernel with one / two trajectories per thread, timed through a
CUDA graph of K launches (no host time).  Development aid for the dispatch thresholds in degk_api.cu; one JSON line each."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import diffeqgpu_b200 as dg  # noqa: E402

dev = "cuda:0"
P0 = np.array([10.0, 28.0, 8.0 / 3.0], np.float32)
LAYOUT = os.environ.get("C1_LAYOUT", "ref")
sizes = [int(float(x)) for x in sys.argv[1:]] or [100, 1000, 10_000, 30_000, 100_000, 200_000, 400_000, 1_000_000]
for N in sizes:
    K = max(3, min(200, int(2e7 // N)))
    g = torch.Generator(device=dev).manual_seed(7)
    p = torch.rand((N, 3), generator=g, device=dev) * torch.tensor(P0, device=dev)
    prob = dg.ODEProblem(dg.models.lorenz, np.array([1, 0, 0], np.float32), (0.0, 10.0), P0)
    probs = dg.ProblemBatch.from_arrays(prob, p=p, device=dev)
    for fp, variants in (("strict", (("v1", None), ("lockstep", None))), ("fast", (("v1", None), ("lockstep", "w1"), ("lockstep", "w2")))):
        for engine, w in variants:
            if w:
                os.environ["DEGK_LOCKSTEP_W1_BELOW"] = "0" if w == "w2" else str(1 << 40)
            plan = dg.vectorized_solve(probs, prob, dg.GPUTsit5(), dt=np.float32(0.1), fp_mode=fp, engine=engine, layout=LAYOUT, prepare=True)
            plan(); torch.cuda.synchronize()
            plan.capture(K)
            plan.replay(); torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); plan.replay(); e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / K)
            os.environ.pop("DEGK_LOCKSTEP_W1_BELOW", None)
            print(json.dumps(dict(N=N, layout=LAYOUT, fp=fp, engine=engine, w=w or "1", us_per_solve=round(best * 1e3, 2),
                                  gsteps_per_s=round(N * 100 / best / 1e6, 2))), flush=True)
    del plan, probs, p
