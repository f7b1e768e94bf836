"""Randomised differential runs of the strict build against the oracle (tools/fuzz_parity.py, tools/fuzz_events.py) with
fixed seeds: ensemble sizes, spans, saveat grids, tolerances, steppers, schedules, layouts, tstops and callbacks drawn at
random; every output must be bit-identical.  (The first run of fuzz_parity.py found the NaN-step difference that
test_non_finite_trajectory_ends_like_the_reference_loop pins.)"""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tool,cases,seed", [("fuzz_parity.py", 80, 20261018), ("fuzz_parity.py", 80, 99), ("fuzz_events.py", 40, 4242)])
def test_randomised_differential_run(tool, cases, seed):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    r = subprocess.run([sys.executable, str(ROOT / "tools" / tool), str(cases), str(seed)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-2000:])
