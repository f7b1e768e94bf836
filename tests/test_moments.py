"""Ensemble-moment reduction (BASELINE config 5's path): solve(EnsembleProblem(prob; reduction = EnsembleMoments()), ...)
-> in-kernel sum(u) / sum(u^2) -> one all-reduce over the ranks.  Reference: the host `reduction` of
src/solve.jl:123-125, 145-146 and the ensemble mean of test/gpu_kernel_de/gpu_sde_regression.jl:36-42."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
f32, f64 = np.float32, np.float64


@pytest.fixture(scope="module")
def dg():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import diffeqgpu_b200 as dg
    return dg


@pytest.mark.gpu
def test_moments_reduction_equals_host_mean_and_is_batch_invariant(dg):
    import torch
    n = 20000
    prob = dg.SDEProblem(dg.models.lorenz_additive, np.array([1, 0, 0], f32), (0.0, 1.0), np.array([10, 28, 8 / 3], f32), seed=99)
    sv = np.array([0.0, 0.25, 0.5], f32)
    ens = dg.EnsembleProblem(prob, reduction=dg.EnsembleMoments())
    a = dg.solve(ens, dg.GPUEM(), dg.EnsembleGPUKernel(dev="cuda:0"), trajectories=n, dt=f32(1e-3), saveat=sv, adaptive=False).u
    b = dg.solve(ens, dg.GPUEM(), dg.EnsembleGPUKernel(dev="cuda:0"), trajectories=n, dt=f32(1e-3), saveat=sv, adaptive=False,
                 batch_size=3333).u
    assert a.n == b.n == n and a.n_ranks == 1
    # the same sample paths whatever the batching (streams keyed by the global trajectory index); only the
    # order of the floating-point sums differs
    assert np.allclose(a.mean, b.mean, rtol=1e-12, atol=1e-12) and np.allclose(a.var, b.var, rtol=1e-10, atol=1e-12)
    # against the paths themselves
    probs = dg.ProblemBatch.from_arrays(prob, n_traj=n, device="cuda:0", seed=99)
    ts, us = dg.vectorized_solve(probs, prob, dg.GPUEM(), dt=f32(1e-3), saveat=sv)
    torch.cuda.synchronize()
    x = us.cpu().numpy().astype(f64)
    assert np.allclose(a.mean, x.mean(0), rtol=1e-12, atol=1e-12) and np.allclose(a.var, x.var(0), rtol=1e-9, atol=1e-12)
    assert np.array_equal(a.t, sv)
    # an ODE ensemble reduces the same way (sums formed on the device from the saved states)
    oprob = dg.ODEProblem(dg.models.lorenz, np.array([1, 0, 0], f32), (0.0, 1.0), np.array([10, 28, 8 / 3], f32))
    r = np.random.default_rng(3).random((500, 3), dtype=f32) * np.array([10, 28, 8 / 3], f32)

    class PF:
        @staticmethod
        def batched(prob, ids):
            return dict(p=r[np.asarray(ids) - 1])
    oens = dg.EnsembleProblem(oprob, prob_func=PF(), reduction=dg.EnsembleMoments())
    m = dg.solve(oens, dg.GPUTsit5(), dg.EnsembleGPUKernel(dev="cuda:0"), trajectories=500, dt=f32(0.01), saveat=sv,
                 adaptive=True, abstol=f32(1e-6), reltol=f32(1e-6), batch_size=128).u
    pb = dg.ProblemBatch.from_arrays(oprob, p=r, device="cuda:0")
    ts, us = dg.vectorized_asolve(pb, oprob, dg.GPUTsit5(), dt=f32(0.01), saveat=sv, abstol=f32(1e-6), reltol=f32(1e-6))
    torch.cuda.synchronize()
    assert np.allclose(m.mean, us.cpu().numpy().astype(f64).mean(0), rtol=1e-12) and m.n == 500


WORKER = r"""
import os, sys
import numpy as np, torch
sys.path.insert(0, sys.argv[1])
import diffeqgpu_b200 as dg
from diffeqgpu_b200.parallel import init_from_env
rank, local, world = init_from_env("gloo")          # two ranks share the one GPU of the test box: gloo carries the all-reduce
f32 = np.float32
n = 30001
prob = dg.SDEProblem(dg.models.lorenz_additive, np.array([1, 0, 0], f32), (0.0, 0.5), np.array([10, 28, 8 / 3], f32), seed=5)
ens = dg.EnsembleProblem(prob, reduction=dg.EnsembleMoments())
sol = dg.solve(ens, dg.GPUEM(), dg.EnsembleGPUKernel(dev="cuda:0"), trajectories=n, dt=f32(1e-3), save_everystep=False, adaptive=False).u
assert sol.n == n and sol.n_ranks == world, (sol.n, sol.n_ranks)
np.save(sys.argv[2] + f"/mean_{world}_{rank}.npy", np.stack([sol.mean, sol.var]))
print("rank", rank, "of", world, "ok")
"""


@pytest.mark.gpu
def test_two_rank_moments_fed_by_the_kernel(dg, tmp_path):
    """world_size 2 (gloo): every rank solves its index-range shard on the GPU, the kernel's `reduce` output feeds the
    all-reduce, and the moments equal the single-rank run of the same ensemble."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER)

    def launch(world):
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world))
        procs = [subprocess.Popen([sys.executable, str(script), str(ROOT), str(tmp_path)], env=dict(env, RANK=str(r), LOCAL_RANK="0"),
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
        outs = [p.communicate(timeout=600)[0] for p in procs]
        for p, o in zip(procs, outs):
            assert p.returncode == 0, o
    launch(1)
    launch(2)
    one = np.load(tmp_path / "mean_1_0.npy")
    two0, two1 = np.load(tmp_path / "mean_2_0.npy"), np.load(tmp_path / "mean_2_1.npy")
    assert np.array_equal(two0, two1)
    assert np.allclose(one, two0, rtol=1e-11, atol=1e-12)
    assert np.isfinite(one).all() and one[1][1].min() > 0          # variance at tf (row 0 is the deterministic u0)
