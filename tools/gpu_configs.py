"""Time the other BASELINE.json configs (C1, C3, C4, C5) on one GPU; prints one JSON line each.
Development/reporting aid: numbers go to profiles/ and DESIGN.md, not to the bench contract."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import diffeqgpu_b200 as dg  # noqa: E402
from cases import henon_heiles_u0  # noqa: E402

dev = "cuda:0"
P0 = np.array([10.0, 28.0, 8.0 / 3.0])


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


def report(name, ms, st, flops, extra=None):
    tot = st["totals"].cpu().numpy()
    steps = int(tot[0] + tot[1])
    r = dict(config=name, ms=round(ms, 3), attempts=steps, accepted=int(tot[0]), failed=int(tot[2]),
             gsteps_per_s=round(steps / ms / 1e6, 3), tflops_alg=round(flops * steps / ms / 1e9, 3))
    if extra:
        r.update(extra)
    print(json.dumps(r), flush=True)
    return r


def main():
    out = []
    g = torch.Generator(device=dev).manual_seed(7)
    for fp in ("strict", "fast"):
        # C1: Lorenz Tsit5 fixed dt=0.1, save every step, N = 1e4 (and 1e6 for a meaningful time)
        for N in (10_000, 1_000_000):
            p = torch.rand((N, 3), generator=g, device=dev) * torch.tensor(P0, dtype=torch.float32, device=dev)
            prob = dg.ODEProblem(dg.models.lorenz, np.array([1, 0, 0], np.float32), (0.0, 10.0), P0.astype(np.float32))
            probs = dg.ProblemBatch.from_arrays(prob, p=p, device=dev)
            ms, o = timed(lambda: dg.vectorized_solve(probs, prob, dg.GPUTsit5(), dt=np.float32(0.1), fp_mode=fp, stats=True))
            out.append(report(f"C1 Lorenz Tsit5 fixed dt=0.1 f32 N={N} {fp}", ms, o[2], 190))
        # C2 at the lower end of its size range (the bench line is 10^8 per GPU)
        for N in (1_000_000, 10_000_000):
            p = torch.rand((N, 3), generator=g, device=dev) * torch.tensor(P0, dtype=torch.float32, device=dev)
            prob = dg.ODEProblem(dg.models.lorenz, np.array([1, 0, 0], np.float32), (0.0, 10.0), P0.astype(np.float32))
            probs = dg.ProblemBatch.from_arrays(prob, p=p, device=dev)
            sv = np.arange(0, 11, dtype=np.float32)
            ms, o = timed(lambda: dg.vectorized_asolve(probs, prob, dg.GPUTsit5(), dt=np.float32(0.1), abstol=np.float32(1e-6),
                                                       reltol=np.float32(1e-6), saveat=sv, fp_mode=fp, stats=True))
            out.append(report(f"C2 Lorenz Tsit5 adaptive tol 1e-6 saveat 0:1:10 f32 N={N} {fp}", ms, o[2], 263))
            del probs, p
        # C3: Vern9 adaptive f64 tol 1e-10
        N = 1_000_000
        p64 = (torch.rand((N, 3), generator=g, device=dev).double()) * torch.tensor(P0, dtype=torch.float64, device=dev)
        prob = dg.ODEProblem(dg.models.lorenz, np.array([1.0, 0, 0]), (0.0, 10.0), P0)
        probs = dg.ProblemBatch.from_arrays(prob, p=p64, device=dev)
        ms, o = timed(lambda: dg.vectorized_asolve(probs, prob, dg.GPUVern9(), dt=0.1, abstol=1e-10, reltol=1e-10,
                                                   fp_mode=fp, stats=True))
        out.append(report(f"C3 Lorenz Vern9 adaptive f64 tol 1e-10 N={N} {fp}", ms, o[2], 775))
        u0 = henon_heiles_u0(N)
        prob = dg.ODEProblem(dg.models.henon_heiles, u0[0], (0.0, 100.0), None)
        probs = dg.ProblemBatch.from_arrays(prob, u0=u0, device=dev)
        ms, o = timed(lambda: dg.vectorized_asolve(probs, prob, dg.GPUVern9(), dt=0.1, abstol=1e-10, reltol=1e-10,
                                                   fp_mode=fp, stats=True))
        out.append(report(f"C3 Henon-Heiles Vern9 adaptive f64 tol 1e-10 N={N} {fp}", ms, o[2], 1002))
        # C4: Robertson Rodas5P f32, N = 2^20
        N = 1 << 20
        k = (0.5 + torch.rand((N, 3), generator=g, device=dev)) * torch.tensor([0.04, 3e7, 1e4], dtype=torch.float32, device=dev)
        prob = dg.ODEProblem(dg.models.rober, np.array([1, 0, 0], np.float32), (0.0, 1e5), np.array([0.04, 3e7, 1e4], np.float32))
        probs = dg.ProblemBatch.from_arrays(prob, p=k, device=dev)
        sv = np.array([1.0, 10.0, 1e3, 1e5], np.float32)
        ms, o = timed(lambda: dg.vectorized_asolve(probs, prob, dg.GPURodas5P(), dt=np.float32(1e-4), abstol=np.float32(1e-8),
                                                   reltol=np.float32(1e-4), saveat=sv, fp_mode=fp, stats=True))
        out.append(report(f"C4 Robertson Rodas5P adaptive f32 N=2^20 {fp}", ms, o[2], 676))
        # C5: Lorenz + additive noise, EM dt=1e-3, 10,000 steps, endpoints + fused ensemble reduction
        N = 1_000_000
        sde = dg.SDEProblem(dg.models.lorenz_additive, np.array([1, 0, 0], np.float32), (0.0, 10.0), P0.astype(np.float32), seed=1234)
        probs = dg.ProblemBatch.from_arrays(sde, n_traj=N, device=dev, seed=1234)
        red = torch.zeros((2, 3, 2), dtype=torch.float64, device=dev)
        ms, o = timed(lambda: dg.vectorized_solve(probs, sde, dg.GPUEM(), dt=np.float32(1e-3), save_everystep=False,
                                                  fp_mode=fp, stats=True, reduce=red), reps=2)
        out.append(report(f"C5 Lorenz+noise EM dt=1e-3 f32 N={N} (1 GPU share of 1e7/8) {fp}", ms, o[2], 23,
                          dict(gnormals_per_s=round(3 * int(o[2]["totals"][0]) / ms / 1e6, 2))))
    Path("gpurun_out").mkdir(exist_ok=True)
    Path("gpurun_out/configs_r1.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
