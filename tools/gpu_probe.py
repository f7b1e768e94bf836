"""Ad-hoc GPU probe (development aid): times the C2 hot path under the engine's knobs."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import diffeqgpu_b200 as dg  # noqa: E402


def lorenz_batch(N, dtype=np.float32, seed=0, dev="cuda:0"):
    g = torch.Generator(device=dev).manual_seed(seed)
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    r = torch.rand((N, 3), generator=g, device=dev, dtype=torch.float32).to(tdt)
    p = r * torch.tensor([10.0, 28.0, 8.0 / 3.0], device=dev, dtype=tdt)
    prob = dg.ODEProblem(dg.models.lorenz, np.array([1, 0, 0], dtype), (0.0, float(__import__('os').environ.get('DEGK_TF', '10'))),
                         np.array([10, 28, 8 / 3], dtype))
    return prob, dg.ProblemBatch.from_arrays(prob, p=p, device=dev)


def time_asolve(N, fp_mode, schedule, alg=None, reps=3, dtype=np.float32, tol=1e-6, engine="auto"):
    alg = alg or dg.GPUTsit5()
    prob, probs = lorenz_batch(N, dtype)
    saveat = np.linspace(0, float(__import__('os').environ.get('DEGK_TF', '10')), 11).astype(dtype)
    kw = dict(dt=dtype(0.1), saveat=saveat, abstol=dtype(tol), reltol=dtype(tol), fp_mode=fp_mode,
              schedule=schedule, stats=True, engine=engine)
    ts, us, st = dg.vectorized_asolve(probs, prob, alg, **kw)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ts, us, st = dg.vectorized_asolve(probs, prob, alg, **kw)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    tot = st["totals"].cpu().numpy()
    steps = int(tot[0] + tot[1])
    info = dg.get_program(prob, alg, fp_mode).info
    return dict(N=N, fp=fp_mode, sched=schedule, alg=type(alg).__name__, ms=round(best, 3),
                steps=steps, acc=int(tot[0]), rej=int(tot[1]), fail=int(tot[2]),
                gsteps_per_s=round(steps / best / 1e6, 3), engine=engine,
                regs=(info.regs_adaptive, info.regs_adaptive2), occ=(info.max_blocks_per_sm, info.max_blocks_per_sm2),
                slots=info.slots_per_thread2)


if __name__ == "__main__":
    out = []
    N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 22
    for eng in ("v1", "auto"):
        for fp in ("strict", "fast"):
            for sched in ("static", "queue"):
                r = time_asolve(N, fp, sched, engine=eng)
                print(json.dumps(r), flush=True)
                out.append(r)
    Path("gpurun_out").mkdir(exist_ok=True)
    Path("gpurun_out/probe.json").write_text(json.dumps(out, indent=1))
