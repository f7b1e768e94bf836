"""BASELINE config 1 at its own size (10^4 trajectories, Lorenz GPUTsit5 dt = 0.1, every-step saves): time per solve
through the direct call (host-bound), a prepared plan, and a CUDA graph of K launches; one JSON line each."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import diffeqgpu_b200 as dg  # noqa: E402

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 200
dev = "cuda:0"
P0 = np.array([10.0, 28.0, 8.0 / 3.0], np.float32)
g = torch.Generator(device=dev).manual_seed(7)
p = torch.rand((N, 3), generator=g, device=dev) * torch.tensor(P0, device=dev)
prob = dg.ODEProblem(dg.models.lorenz, np.array([1, 0, 0], np.float32), (0.0, 10.0), P0)
probs = dg.ProblemBatch.from_arrays(prob, p=p, device=dev)


def timed(fn, k):
    fn(); torch.cuda.synchronize()
    best, host = 1e9, 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        fn()
        e1.record()
        host = min(host, (time.perf_counter() - t0) / k * 1e3)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / k)
    return best, host


for fp in ("strict", "fast"):
    for engine in ("auto", "lockstep"):
        kw = dict(dt=np.float32(0.1), fp_mode=fp, engine=engine)
        direct = lambda: [dg.vectorized_solve(probs, prob, dg.GPUTsit5(), **kw) for _ in range(K)]
        plan = dg.vectorized_solve(probs, prob, dg.GPUTsit5(), prepare=True, **kw)
        plan()
        planned = lambda: [plan() for _ in range(K)]
        plan.capture(K)
        for name, fn in (("direct", direct), ("plan", planned), ("graph", plan.replay)):
            ms, host = timed(fn, K)
            print(json.dumps(dict(N=N, fp=fp, engine=engine, path=name, ms_per_solve=round(ms, 4), host_ms_per_solve=round(host, 4),
                                  gsteps_per_s=round(N * 100 / ms / 1e6, 2))), flush=True)
