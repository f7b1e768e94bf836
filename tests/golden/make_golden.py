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
from cases import golden_cases, golden_cases_oracle_only  # noqa: E402
from oracle import oracle  # noqa: E402

out = {}
for name, kw in golden_cases() + golden_cases_oracle_only():
    kw = dict(kw)
    model, alg = kw.pop("model"), kw.pop("alg")
    r = oracle.solve(model, alg, kw.pop("u0"), kw.pop("p"), kw.pop("tspan"), **kw)
    for k in ("ts", "us", "naccept", "nreject", "retcode"):
        out[f"{name}/{k}"] = r[k]
    print(name, r["us"].shape, "acc", int(r["naccept"].sum()), "rej", int(r["nreject"].sum()))
path = Path(__file__).resolve().parent / "oracle_golden.npz"
if path.exists() and "--force" not in sys.argv:      # frozen entries may only be extended, not changed
    old = np.load(path)
    # --rekey-sde: the SDE noise stream was re-defined (round 2: seed and trajectory index moved into different
    # Philox words); only the stochastic cases may change then, every deterministic entry stays frozen
    rekey = "--rekey-sde" in sys.argv
    for k in old.files:
        if rekey and ("_em_" in k or "_siea_" in k):
            continue
        assert k in out and np.array_equal(old[k], out[k], equal_nan=True), f"{k} changed (pass --force to overwrite)"
np.savez_compressed(path, **out)
