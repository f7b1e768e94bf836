"""The oracle (and the frozen goldens the GPU is compared with) against outputs of the REAL reference.

`tests/golden/reference_golden.npz` is produced by baseline/dump_reference_goldens.jl (DiffEqGPU.jl's own
EnsembleGPUKernel path on its CPU backend) from tests/golden/golden_inputs.npz.  Julia is not available where this
repository is built and graded, so the file is absent there and these tests skip -- the oracle's parity with the
reference stays "unpinned" (DESIGN.md section 2).  On a machine with Julia the three commands in
tests/golden/export_inputs.py turn it green or show exactly which rounding assumption (Float32 `^`, StaticArrays
`det`, MuladdMacro association) does not hold."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
GOLDEN = Path(__file__).resolve().parent / "golden"
REF = GOLDEN / "reference_golden.npz"
INPUTS = GOLDEN / "golden_inputs.npz"


def test_golden_inputs_match_the_cases():
    """the committed inputs are the ones the oracle goldens were frozen from"""
    from cases import golden_cases
    inp = np.load(INPUTS)
    names = set(inp["names"].tolist())
    cases = dict(golden_cases())
    assert names and names <= set(cases)
    for name in names:
        kw = cases[name]
        u0 = np.atleast_2d(np.asarray(kw["u0"], dtype=kw["dtype"]))
        assert np.array_equal(inp[f"{name}/u0"][0], u0[0]) and inp[f"{name}/u0"].dtype == np.dtype(kw["dtype"])
        if kw["p"] is not None:
            assert np.array_equal(inp[f"{name}/p"], np.atleast_2d(np.asarray(kw["p"], dtype=kw["dtype"]))) or inp[f"{name}/p"].shape[0] > 1
        assert np.array_equal(inp[f"{name}/tspan"], np.asarray(kw["tspan"], dtype=kw["dtype"]))


@pytest.mark.skipif(not REF.exists(), reason="tests/golden/reference_golden.npz not generated (needs Julia: baseline/dump_reference_goldens.jl)")
def test_oracle_goldens_equal_the_reference():
    ref = np.load(REF)
    gold = np.load(GOLDEN / "oracle_golden.npz")
    inp = np.load(INPUTS)
    bad = []
    for name in inp["names"].tolist():
        f64 = bool(inp[f"{name}/ids"][4])
        for k in ("ts", "us"):
            a, b = gold[f"{name}/{k}"], ref[f"{name}/{k}"]
            if a.shape != b.shape:
                bad.append((name, k, "shape", a.shape, b.shape))
                continue
            # rows the kernel never writes hold uninitialised memory in the reference's `us` (only `ts` is filled)
            written = np.ones(a.shape[:2], bool) if k == "ts" else (gold[f"{name}/ts"] != inp[f"{name}/tspan"][0]) | (np.arange(a.shape[1]) == 0)
            sel = written if k == "ts" else written[..., None] & np.ones(a.shape, bool)
            if f64:
                ok = np.allclose(a[sel], b[sel], rtol=1e-12, atol=1e-14)
            else:
                ok = np.array_equal(a[sel], b[sel])
            if not ok:
                bad.append((name, k, float(np.abs(a[sel].astype(np.float64) - b[sel].astype(np.float64)).max())))
    assert not bad, bad
