"""SolvePlan (`vectorized_solve(..., prepare=True)`): a prepared launch gives what the direct call gives, re-launches into
the same arrays (picking up new parameter values), and replays from a CUDA graph."""
import numpy as np
import pytest

from cases import lorenz_sweep

pytestmark = pytest.mark.gpu
f32 = np.float32


@pytest.fixture(scope="module")
def dg():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import diffeqgpu_b200 as dg
    return dg


def _setup(dg, n, seed=5):
    import torch
    p = torch.as_tensor(lorenz_sweep(n, seed=seed)).to("cuda:0")
    prob = dg.ODEProblem(dg.models.lorenz, np.array([1, 0, 0], f32), (0.0, 10.0), np.array([10, 28, 8 / 3], f32))
    return prob, dg.ProblemBatch.from_arrays(prob, p=p, device="cuda:0"), p


@pytest.mark.parametrize("n", [1000, 10_000, 300_000])
def test_plan_fixed_dt_equals_direct_and_relaunches(dg, n):
    import torch
    prob, probs, p = _setup(dg, n)
    ts0, us0 = dg.vectorized_solve(probs, prob, dg.GPUTsit5(), dt=f32(0.1))
    plan = dg.vectorized_solve(probs, prob, dg.GPUTsit5(), dt=f32(0.1), prepare=True)
    assert isinstance(plan, dg.SolvePlan)
    ts1, us1 = plan()
    assert torch.equal(ts0, ts1) and torch.equal(us0, us1)
    # new parameter values in place: the same plan solves the new ensemble
    p2 = torch.as_tensor(lorenz_sweep(n, seed=6)).to("cuda:0")
    probs.p.copy_(p2)
    ts2, us2 = plan()
    assert us2.data_ptr() == us1.data_ptr()
    _, probs2, _ = _setup(dg, n, seed=6)
    ts3, us3 = dg.vectorized_solve(probs2, prob, dg.GPUTsit5(), dt=f32(0.1))
    assert torch.equal(us2, us3) and torch.equal(ts2, ts3)
    # CUDA graph of three launches
    plan.capture(3)
    us2.zero_()
    plan.replay()
    torch.cuda.synchronize()
    assert torch.equal(plan.us, us3)


def test_plan_adaptive_with_stats_does_not_accumulate(dg):
    import torch
    prob, probs, _ = _setup(dg, 20_000)
    sv = np.arange(0, 11, dtype=f32)
    kw = dict(dt=f32(0.1), saveat=sv, abstol=f32(1e-6), reltol=f32(1e-6), stats=True)
    ts0, us0, st0 = dg.vectorized_asolve(probs, prob, dg.GPUTsit5(), **kw)
    plan = dg.vectorized_asolve(probs, prob, dg.GPUTsit5(), prepare=True, **kw)
    for _ in range(3):
        ts1, us1, st1 = plan()
    assert torch.equal(us0, us1) and torch.equal(ts0, ts1)
    assert torch.equal(st0["totals"], st1["totals"]) and torch.equal(st0["naccept"], st1["naccept"])
    plan.capture(2)
    plan.replay()
    torch.cuda.synchronize()
    assert torch.equal(st0["totals"], plan.stats["totals"]) and torch.equal(us0, plan.us)
