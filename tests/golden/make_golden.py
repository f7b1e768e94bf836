#!/usr/bin/env python3
"""Regenerate tests/golden/oracle_golden.npz from the CPU oracle.

The reference ships no golden vectors and cannot run here (no Julia), so -- as SURVEY §8c
prescribes -- the oracle's own outputs on the seeded cases of tests/cases.py are frozen as the
goldens: they pin the oracle against accidental change and travel to the GPU box, where the
CUDA engine is compared against them without needing the oracle's results to be recomputed.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from cases import golden_cases  # noqa: E402
from oracle import oracle  # noqa: E402

out = {}
for name, kw in golden_cases():
    kw = dict(kw)
    model, alg = kw.pop("model"), kw.pop("alg")
    r = oracle.solve(model, alg, kw.pop("u0"), kw.pop("p"), kw.pop("tspan"), **kw)
    for k in ("ts", "us", "naccept", "nreject", "retcode"):
        out[f"{name}/{k}"] = r[k]
    print(name, r["us"].shape, "acc", int(r["naccept"].sum()), "rej", int(r["nreject"].sum()))
np.savez_compressed(Path(__file__).resolve().parent / "oracle_golden.npz", **out)
