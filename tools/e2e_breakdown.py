"""Where the end-to-end time of the C2 path goes on a multi-GPU host (run under torchrun; one JSON line per rank 0).

    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/e2e_breakdown.py [traj_per_gpu] [bind]

Per rank, all ranks at the same time (barrier before every leg), max over ranks:
  d2h / h2d   plain pinned copies in 2 M-trajectory pieces over three streams (the PCIe + host-memory ceiling)
  solve_host  degk_solve_host as bench.py's e2e leg calls it (ts rebuilt on the host from row counts)
  solve_host_ts_over_pcie   the same with DEGK_NO_COMPACT_TS=1 (ts transferred instead of rebuilt)
`bind` = 1 pins every rank to the CPUs local to its GPU before the pinned buffers are allocated
(parallel.bind_to_gpu_numa_node)."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import diffeqgpu_b200 as dg  # noqa: E402
from diffeqgpu_b200.parallel import bind_to_gpu_numa_node, init_from_env, max_over_ranks, sum_over_ranks  # noqa: E402

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
bind = len(sys.argv) > 2 and sys.argv[2] == "1"
rank, local, world = init_from_env("nccl")
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
cpus = bind_to_gpu_numa_node(local) if bind else set()
f32 = np.float32
P0 = np.array([10.0, 28.0, 8.0 / 3.0], f32)
t0 = time.perf_counter()
p_h = torch.empty((N, 3), dtype=torch.float32, pin_memory=True)
us_h = torch.empty((N, 11, 3), dtype=torch.float32, pin_memory=True)
ts_h = torch.empty((N, 11), dtype=torch.float32, pin_memory=True)
t_pin = time.perf_counter() - t0
p_h.copy_(torch.rand((N, 3)) * torch.tensor(P0))
us_d = torch.empty((N, 11, 3), dtype=torch.float32, device=dev)
streams = [torch.cuda.Stream(dev) for _ in range(3)]
CH = 1 << 21


def barrier():
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize(dev)


def timed(fn, reps=2):
    best = 1e30
    for _ in range(reps):
        barrier()
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize(dev)
        best = min(best, max_over_ranks(time.perf_counter() - t, dev))
    return best


def d2h():
    for i, c0 in enumerate(range(0, N, CH)):
        with torch.cuda.stream(streams[i % 3]):
            us_h[c0:c0 + CH].copy_(us_d[c0:c0 + CH], non_blocking=True)


def h2d():
    for i, c0 in enumerate(range(0, N, CH)):
        with torch.cuda.stream(streams[i % 3]):
            us_d[c0:c0 + CH].copy_(us_h[c0:c0 + CH], non_blocking=True)


prob = dg.ODEProblem(dg.models.lorenz, np.array([1, 0, 0], f32), (0.0, 10.0), P0)
hk = dict(p=p_h, dt=f32(0.1), adaptive=True, abstol=1e-6, reltol=1e-6, saveat=np.arange(0, 11, dtype=f32), fp_mode="fast",
          out={"us": us_h, "ts": ts_h}, stats="totals", device=dev, chunk_traj=CH)


def solve():
    dg.solve_host(prob, dg.GPUTsit5(), **hk)


solve()
t_d2h, t_h2d = timed(d2h), timed(h2d)
t_solve = timed(solve)
os.environ["DEGK_NO_COMPACT_TS"] = "1"
t_solve_ts = timed(solve)
del os.environ["DEGK_NO_COMPACT_TS"]
gb_us = N * 132 / 1e9
out = {"n_gpus": world, "traj_per_gpu": N, "bound_to_gpu_numa_node": bind, "cpus_per_rank": len(cpus) if cpus else len(os.sched_getaffinity(0)),
       "pin_alloc_s": round(max_over_ranks(t_pin, dev), 3),
       "d2h_GBps_per_gpu": round(gb_us / t_d2h, 1), "d2h_GBps_total": round(world * gb_us / t_d2h, 1),
       "h2d_GBps_per_gpu": round(gb_us / t_h2d, 1), "h2d_GBps_total": round(world * gb_us / t_h2d, 1),
       "solve_host_s": round(t_solve, 3), "solve_host_GBps_total": round(world * N * 148 / 1e9 / t_solve, 1),
       "solve_host_ts_over_pcie_s": round(t_solve_ts, 3),
       "d2h_only_floor_s": round(t_d2h * 136 / 132, 3)}
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
