"""C1 probe (development aid): fixed-dt Lorenz Tsit5, every-step saves, REF vs SOA output layout."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import diffeqgpu_b200 as dg  # noqa: E402

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
dev = "cuda:0"
P0 = np.array([10.0, 28.0, 8.0 / 3.0], np.float32)
g = torch.Generator(device=dev).manual_seed(7)
p = torch.rand((N, 3), generator=g, device=dev) * torch.tensor(P0, device=dev)
prob = dg.ODEProblem(dg.models.lorenz, np.array([1, 0, 0], np.float32), (0.0, 10.0), P0)
probs = dg.ProblemBatch.from_arrays(prob, p=p, device=dev)
for fp in ("strict", "fast"):
    for layout in ("ref", "soa"):
        for save in (True, False):
            fn = lambda: dg.vectorized_solve(probs, prob, dg.GPUTsit5(), dt=np.float32(0.1), fp_mode=fp, layout=layout,
                                             save_everystep=save)
            fn(); torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            print(json.dumps(dict(N=N, fp=fp, layout=layout, save_everystep=save, ms=round(best, 3),
                                  gsteps_per_s=round(N * 100 / best / 1e6, 2))), flush=True)
