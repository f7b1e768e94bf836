"""Host-side mirror of the reference interface: argument handling, sizing, error behaviour
(reference test/public_interface.jl, test/gpu_kernel_de/conversions.jl) -- no GPU needed."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import diffeqgpu_b200 as dg  # noqa: E402
from diffeqgpu_b200 import lowerlevel_solve as ll  # noqa: E402
from diffeqgpu_b200.parallel import shard_range  # noqa: E402

f32 = np.float32


def lorenz_prob(dtype=f32):
    return dg.ODEProblem(dg.models.lorenz, np.array([1, 0, 0], dtype), (0.0, 10.0), np.array([10, 28, 8 / 3], dtype))


def test_problem_types_and_remake():
    prob = lorenz_prob()
    assert prob.dtype == f32 and prob.tspan == (f32(0), f32(10)) and prob.u0.dtype == f32
    p2 = dg.remake(prob, p=np.array([1, 2, 3], f32))
    assert p2.p.tolist() == [1, 2, 3] and p2.u0 is not None and dg.make_prob_compatible(p2) is p2
    with pytest.raises(ValueError):
        dg.ODEProblem(dg.models.lorenz, np.zeros(2, f32), (0, 1), np.zeros(3, f32))
    assert dg.alg_order(dg.GPUTsit5()) == 5 and dg.alg_order(dg.GPUVern9()) == 9 and dg.alg_order(dg.GPURodas5P()) == 5
    assert dg.EnsembleGPUKernel("cuda").cpu_offload == 0.0
    with pytest.raises(ValueError):
        dg.EnsembleGPUKernel("cuda", 0.2)            # no CPU path in this engine


def test_saveat_conversions_like_reference():
    """test/gpu_kernel_de/conversions.jl:24-58: grids come out in the problem's Float32"""
    prob = dg.ODEProblem(dg.models.lorenz, np.array([1, 0, 0], f32), (1.0, 10.0), np.array([10, 28, 8 / 3], f32))
    a = ll._convert_saveat_adaptive(dg.Range(1, 10, step=1), prob)
    assert a.dtype == f32 and np.array_equal(a, np.arange(1, 11, dtype=f32))
    b = ll._convert_saveat_adaptive(dg.Range(1, 10, step=0.1), prob)
    assert len(b) == 91 and b.dtype == f32 and b[0] == 1 and b[-1] == 10
    assert np.allclose(b, np.arange(91) * 0.1 + 1, atol=1e-6)
    c = ll._convert_saveat_adaptive(1.0, prob)           # scalar step -> range(t0, tf, length=ceil(9/1)+1)
    assert np.array_equal(c, np.arange(1, 11, dtype=f32))
    d = ll._convert_saveat_adaptive([1.0, 5.0, 10.0], prob)
    assert d.dtype == f32 and d.tolist() == [1, 5, 10]
    e = ll._convert_saveat_adaptive(0.0, prob)           # saveat = 0 -> endpoints
    assert e.tolist() == [1, 10]
    with pytest.raises(ValueError, match="too many save points"):
        ll._convert_saveat_adaptive(1e-5, prob)
    # fixed-dt variant has no cap
    assert len(ll._convert_saveat_fixed(0.5, prob)) == 19


def test_error_behaviour_matches_reference():
    sde = dg.SDEProblem(dg.models.gbm, np.full(3, 0.1, f32), (0.0, 1.0), np.array([1.5, 0.01], f32))
    with pytest.raises(RuntimeError, match="Adaptive time-stepping is not supported yet with GPUEM"):
        dg.vectorized_asolve(None, sde, dg.GPUEM(), dt=0.1)
    nd = dg.SDEProblem(dg.models.gbm_nd, np.full(2, 0.1, f32), (0.0, 1.0), np.array([1.5, 0.01], f32))
    assert not nd.is_diagonal_noise() and sde.is_diagonal_noise()
    with pytest.raises(ValueError, match="not compatible with the chosen noise type"):
        dg.vectorized_solve(None, nd, dg.GPUSIEA(), dt=0.1)
    with pytest.raises(TypeError):
        dg.vectorized_solve(None, lorenz_prob(), dg.GPUEM(), dt=0.1)
    # tstops, discrete and continuous callbacks are lowered (tests/test_events.py); save_positions must be (false, false)
    with pytest.raises(ValueError, match="save_positions"):
        dg.ContinuousCallback("return u[0];", "u[0] = 0;", save_positions=(True, True))


def test_solving_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    prob = lorenz_prob()
    with pytest.raises(Exception):
        probs = dg.ProblemBatch.from_arrays(prob, p=np.ones((4, 3), f32), device="cpu")
        dg.vectorized_solve(probs, prob, dg.GPUTsit5(), dt=f32(0.1))


def test_problem_batch_from_problems_on_cpu_tensors():
    prob = lorenz_prob()
    probs = [dg.remake(prob, p=np.array([i, 2, 3], f32)) for i in range(5)]
    b = dg.ProblemBatch.from_problems(probs, device="cpu")
    assert len(b) == 5 and b.u0.shape == (5, 3) and b.p[:, 0].tolist() == [0, 1, 2, 3, 4] and b.tspan.ndim == 1
    probs[2] = dg.remake(probs[2], tspan=(f32(0), f32(5)))
    b = dg.ProblemBatch.from_problems(probs, device="cpu")
    assert b.tspan.shape == (5, 2)


@pytest.mark.parametrize("n,w", [(10, 1), (10, 3), (7, 8), (10 ** 8, 8), (0, 4)])
def test_shard_range_partitions(n, w):
    parts = [shard_range(n, r, w) for r in range(w)]
    assert parts[0][0] == 0 and parts[-1][1] == n
    assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
    sizes = [hi - lo for lo, hi in parts]
    assert max(sizes) - min(sizes) <= 1


def test_julia_range_elements_and_tstops_row_count():
    """range elements must be the ones Julia produces (twice-precision, rounded once): (0:0.1:1)[4] == 0.3.
    A naive t0 + k*dt in Float64 made tstops on grid points count as extra rows."""
    from diffeqgpu_b200.julia_ranges import lin_range, step_range
    from diffeqgpu_b200.lowerlevel_solve import _convert_saveat_adaptive, fixed_dt_rows
    r = step_range(0.0, 0.1, 11, np.float64)
    assert r[3] == 0.3 and r[6] == 0.6 and r[7] == 0.7 and r[10] == 1.0
    assert 0.1 * 3 != 0.3                                  # what the naive grid would have produced
    assert np.array_equal(lin_range(0.0, 1.0, 11, np.float64), r)
    assert np.array_equal(step_range(np.float32(0), np.float32(0.1), 101, np.float32)[[3, 100]], np.float32([0.3, 10.0]))
    assert fixed_dt_rows(np.float64, 0.0, 1.0, 0.1, [0.3]) == 11
    assert fixed_dt_rows(np.float64, 0.0, 1.0, 0.1, [0.3, 0.6, 0.7]) == 11
    assert fixed_dt_rows(np.float64, 0.0, 1.0, 0.1, [0.35]) == 12
    assert fixed_dt_rows(np.float32, 0.0, 1.0, 0.1, [0.3, 0.35]) == 12
    assert fixed_dt_rows(np.float64, 0.0, 1.0, 0.1) == 11
    prob = dg.ODEProblem(dg.models.lorenz, np.array([1.0, 0, 0]), (0.0, 1.0), np.array([10.0, 28.0, 8 / 3]))
    sv = _convert_saveat_adaptive(0.1, prob)               # saveat as a number: range(t0, tf, length = 11)
    assert sv.dtype == np.float64 and np.array_equal(sv, r)
    sv = _convert_saveat_adaptive(dg.Range(0.0, 1.0, step=0.1), prob)
    assert np.array_equal(sv, r)
