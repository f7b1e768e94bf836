"""End-to-end probe (development aid): degk_solve_host on the C2 workload with host buffers,
under different chunk sizes and with / without the host-side ts rebuild."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import diffeqgpu_b200 as dg  # noqa: E402

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
f32 = np.float32
P0 = np.array([10.0, 28.0, 8.0 / 3.0], f32)
U0 = np.array([1.0, 0.0, 0.0], f32)
SAVEAT = np.arange(0, 11, dtype=f32)
dev = torch.device("cuda:0")
p_host = torch.empty((N, 3), dtype=torch.float32, pin_memory=True)
p_host.copy_(torch.rand((N, 3)) * torch.tensor(P0))
us_h = torch.empty((N, 11, 3), dtype=torch.float32, pin_memory=True)
ts_h = torch.empty((N, 11), dtype=torch.float32, pin_memory=True)
prob = dg.ODEProblem(dg.models.lorenz, U0, (0.0, 10.0), P0)

# ceiling: D2H of the us bytes alone, in chunk-sized pieces on one stream
d = torch.empty((1 << 22, 11, 3), dtype=torch.float32, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
for c in range(0, N - (1 << 22) + 1, 1 << 22):
    us_h[c:c + (1 << 22)].copy_(d, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(json.dumps({"d2h_only_GBps": (N // (1 << 22)) * (1 << 22) * 132 / dt / 1e9}))
del d

for compact in (1, 0):
    if compact:
        os.environ.pop("DEGK_NO_COMPACT_TS", None)
    else:
        os.environ["DEGK_NO_COMPACT_TS"] = "1"
    for chunk in (1 << 21, 1 << 22, 1 << 23):
        hk = dict(p=p_host, dt=f32(0.1), adaptive=True, abstol=1e-6, reltol=1e-6, saveat=SAVEAT, fp_mode="fast",
                  out={"us": us_h, "ts": ts_h}, stats="totals", device=dev, chunk_traj=chunk)
        dg.solve_host(prob, dg.GPUTsit5(), **hk)
        t0 = time.perf_counter()
        _, _, st = dg.solve_host(prob, dg.GPUTsit5(), **hk)
        dt = time.perf_counter() - t0
        steps = int(st["totals"][0] + st["totals"][1])
        print(json.dumps({"compact_ts": compact, "chunk": chunk, "ms": round(dt * 1e3, 1), "gsteps_s": round(steps / dt / 1e9, 2),
                          "d2h_GBps": round(N * (136 if compact else 176) / dt / 1e9, 1)}), flush=True)
