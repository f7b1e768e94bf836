"""world_size-2 gloo test of the multi-GPU host logic (sharding + ensemble-moment all-reduce).
The per-rank `reduce` partials come from the oracle here (CPU only); on the GPU box the same chain is fed by the
kernel's reduce output with two ranks (tests/test_moments.py::test_two_rank_moments_fed_by_the_kernel)."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent

WORKER = r"""
import os, sys
import numpy as np, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import diffeqgpu_b200 as dg
from diffeqgpu_b200.parallel import init_from_env, shard_range, allreduce_moments, max_over_ranks
from oracle import oracle
rank, local, world = init_from_env("gloo")
N = 1001
lo, hi = shard_range(N, rank, world)
u0 = np.full((hi - lo, 1), 0.5, np.float32)
# shard-invariant RNG: the oracle keys streams by GLOBAL trajectory index, so emulate traj_offset
full = oracle.solve("scalar_sde", "em", np.full((N, 1), 0.5, np.float32), [1.0, 0.3], [0, 1], dt=1/32, save_everystep=False, seed=5)
mine = full["us"][lo:hi]
part = torch.tensor(np.stack([mine.sum(0), (mine.astype(np.float64) ** 2).sum(0)], -1), dtype=torch.float64)
mean, var, n = allreduce_moments(part, hi - lo)
ref_mean = full["us"].astype(np.float64).mean(0); ref_var = full["us"].astype(np.float64).var(0)
assert n == N, n
assert np.allclose(mean.numpy(), ref_mean, atol=1e-12) and np.allclose(var.numpy(), ref_var, atol=1e-10)
assert max_over_ranks(float(rank)) == world - 1
print("rank", rank, "ok", lo, hi)
"""


def test_two_rank_sharding_and_moment_allreduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    # build the oracle once here: two workers racing `make` on a stale library would both fail
    sys.path.insert(0, str(ROOT))
    from oracle import oracle
    oracle.build()
    import socket
    with socket.socket() as sk:                      # a free rendezvous port
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="2")
    procs = []
    for r in range(2):
        e = dict(env, RANK=str(r), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, str(script), str(ROOT)], env=e,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    assert "rank 0 ok 0 501" in outs[0] and "rank 1 ok 501 1001" in outs[1]
