"""C5 (BASELINE.json configs[4]): Lorenz + additive noise, GPUEM dt = 1e-3, tspan (0, 10), N trajectories
sharded by index range over the ranks of a torchrun job, ensemble mean / variance at tf by the
in-kernel reduction + ONE NCCL all-reduce of the moments (parallel.allreduce_moments).

    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/c5_multi.py [N_total]

Prints one JSON line on rank 0: steps/s over all ranks (device-timed, max over ranks), the mean, and a
shard-invariance check (the same ensemble on ONE rank gives the same moments, RNG streams are keyed by the
global trajectory index)."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import diffeqgpu_b200 as dg  # noqa: E402
from diffeqgpu_b200.parallel import allreduce_moments, init_from_env, max_over_ranks, shard_range  # noqa: E402

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
rank, local, world = init_from_env("nccl")
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
f32 = np.float32
prob = dg.SDEProblem(dg.models.lorenz_additive, np.array([1, 0, 0], f32), (0.0, 10.0), np.array([10, 28, 8 / 3], f32), seed=1234)
lo, hi = shard_range(N, rank, world)
probs = dg.ProblemBatch.from_arrays(prob, n_traj=hi - lo, device=dev)


def run(n_local, offset):
    red = torch.zeros((2, 3, 2), dtype=torch.float64, device=dev)
    b = probs if n_local == hi - lo else dg.ProblemBatch.from_arrays(prob, n_traj=n_local, device=dev)
    dg.vectorized_solve(b, prob, dg.GPUEM(), dt=f32(1e-3), save_everystep=False, fp_mode="fast", traj_offset=offset, reduce=red)
    return red


for _ in range(2):                       # warm-up (JIT / AOT lookup, clocks)
    run(hi - lo, lo)
torch.cuda.synchronize()
ms = 1e30
for _ in range(4):                       # best of 4, each = solve + in-kernel reduction + all-reduce
    if world > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    red = run(hi - lo, lo)
    mean, var, n = allreduce_moments(red, hi - lo)
    e1.record()
    torch.cuda.synchronize()
    ms = min(ms, max_over_ranks(e0.elapsed_time(e1), dev))
out = {"config": "C5 Lorenz + additive noise, GPUEM dt=1e-3, tspan (0,10), ensemble mean/variance at tf", "n_gpus": world,
       "trajectories": n, "ms": round(ms, 3), "gsteps_per_s": round(n * 10000 / ms / 1e6, 2),
       "gnormals_per_s": round(3 * n * 10000 / ms / 1e6, 2), "mean_tf": [float(x) for x in mean[1]],
       "var_tf": [float(x) for x in var[1]], "collective": "one NCCL all-reduce of 13 doubles" if world > 1 else "none"}
if rank == 0 and world > 1 and N <= 2_000_000:
    # shard invariance: the whole ensemble on this rank alone
    r1 = run(N, 0)
    m1 = (r1[..., 0] / N)[1]
    out["shard_invariant"] = bool(torch.allclose(m1, mean[1], rtol=1e-9, atol=1e-9))
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
