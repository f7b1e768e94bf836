import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import diffeqgpu_b200 as dg
from cases import P0_LORENZ, U0_LORENZ, lorenz_sweep
f32 = np.float32
n = 20000
rng = np.random.default_rng(9)
p = lorenz_sweep(n, seed=12)
tspan = np.stack([rng.uniform(0.0, 0.4, n), rng.uniform(0.5, 6.0, n)], axis=1).astype(f32)
tspan[::7, 1] = tspan[::7, 0]
sv = np.array([0.25, 1.0, 2.5, 4.0, 5.5, 7.0], f32)
u0 = np.tile(U0_LORENZ.astype(f32), (n, 1))
prob = dg.ODEProblem(dg.models.lorenz, U0_LORENZ.astype(f32), (0.0, 6.0), P0_LORENZ.astype(f32))
def dev_run(sl=slice(None)):
    probs = dg.ProblemBatch.from_arrays(prob, u0=u0[sl], p=p[sl], tspan=tspan[sl], device="cuda:0")
    ts, us, st = dg.vectorized_asolve(probs, prob, dg.GPUTsit5(), dt=f32(0.1), abstol=f32(1e-5), reltol=f32(1e-5), saveat=sv, fp_mode="fast", stats=True)
    torch.cuda.synchronize()
    return ts.cpu().numpy(), us.cpu().numpy(), st["naccept"].cpu().numpy(), st["nreject"].cpu().numpy()
a = dev_run(); b = dev_run()
w = a[0] != tspan[:, :1]
print("dev vs dev: us equal", np.array_equal(a[1][w], b[1][w]), "nacc equal", np.array_equal(a[2], b[2]))
c = dev_run(slice(0, 3000))
w3 = w[:3000]
print("full vs first-3000 launch: us equal", np.array_equal(a[1][:3000][w3], c[1][w3]), "nacc equal", np.array_equal(a[2][:3000], c[2]))
d = np.abs(a[1][:3000] - c[1]); d[~w3] = 0
bad = np.argwhere(d.max(axis=2) > 0)
print("n differing rows", len(bad), "max abs", d.max())
for i, k in bad[:10]:
    print(i, k, tspan[i], a[2][i], c[2][i], a[3][i], c[3][i], a[1][i, k], c[1][i, k])
