"""C1 probe (development aid): fixed-dt Lorenz Tsit5, every-step saves; the lock-step kernel against the one-thread-per-
trajectory kernel, REF vs SOA output layout.  K launches are enqueued back to back between two events so that the
host-side cost of a call (~0.1 ms of Python / ctypes / allocator) hides behind the GPU work of the previous ones."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import diffeqgpu_b200 as dg  # noqa: E402

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = "cuda:0"
P0 = np.array([10.0, 28.0, 8.0 / 3.0], np.float32)
g = torch.Generator(device=dev).manual_seed(7)
p = torch.rand((N, 3), generator=g, device=dev) * torch.tensor(P0, device=dev)
prob = dg.ODEProblem(dg.models.lorenz, np.array([1, 0, 0], np.float32), (0.0, 10.0), P0)
probs = dg.ProblemBatch.from_arrays(prob, p=p, device=dev)
for fp in ("strict", "fast"):
    for engine in ("v1", "lockstep"):
        for layout in ("ref", "soa"):
            fn = lambda: dg.vectorized_solve(probs, prob, dg.GPUTsit5(), dt=np.float32(0.1), fp_mode=fp, layout=layout, engine=engine)
            fn(); torch.cuda.synchronize()
            best, host = 1e9, 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0 = time.perf_counter()
                e0.record()
                for _k in range(K):
                    fn()
                e1.record()
                host = min(host, (time.perf_counter() - t0) / K * 1e3)
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / K)
            print(json.dumps(dict(N=N, fp=fp, engine=engine, layout=layout, ms=round(best, 4), host_ms_per_call=round(host, 4),
                                  gsteps_per_s=round(N * 100 / best / 1e6, 2))), flush=True)
