#!/usr/bin/env python3
"""Write the inputs of the deterministic golden cases (tests/cases.py) to tests/golden/golden_inputs.npz so that the
REAL reference can be run on exactly these inputs wherever Julia is available:

    python tests/golden/export_inputs.py
    julia baseline/dump_reference_goldens.jl          # -> tests/golden/reference_golden.npz
    python -m pytest tests/test_reference_goldens.py  # oracle (and GPU goldens) against the reference itself

The oracle's parity is "unpinned" until reference_golden.npz exists (no Julia in the build image or on the GPU box);
this is the path that turns it green.  Only cases the reference's own EnsembleGPUKernel(CPU()) can run are exported:
the built-in ODE models without callbacks (the SDE noise stream is backend dependent in the reference)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "tests"))
from cases import golden_cases  # noqa: E402

MODELS = {"lorenz": 0, "henon_heiles": 1, "rober": 2, "decay": 3}
ALGS = {"tsit5": 0, "vern7": 1, "vern9": 2, "rosenbrock23": 3, "rodas4": 4, "rodas5p": 5}

out = {}
names = []
for name, kw in golden_cases():
    if kw["alg"] not in ALGS or kw["model"] not in MODELS:
        continue
    dt = np.dtype(kw["dtype"])
    u0 = np.atleast_2d(np.asarray(kw["u0"], dtype=dt))
    p = np.zeros((1, 0), dt) if kw["p"] is None else np.atleast_2d(np.asarray(kw["p"], dtype=dt))
    n = max(u0.shape[0], p.shape[0])
    u0 = np.broadcast_to(u0, (n, u0.shape[1])).copy()
    p = np.broadcast_to(p, (n, p.shape[1])).copy()
    sv = kw.get("saveat")
    out[f"{name}/u0"], out[f"{name}/p"] = u0, p
    out[f"{name}/tspan"] = np.asarray(kw["tspan"], dtype=dt)
    out[f"{name}/saveat"] = np.zeros(0, dt) if sv is None else np.asarray(sv, dtype=dt)
    # model, alg, adaptive, save_everystep, is_f64
    out[f"{name}/ids"] = np.array([MODELS[kw["model"]], ALGS[kw["alg"]], int(bool(kw.get("adaptive", False))),
                                   int(bool(kw.get("save_everystep", True))), int(dt == np.float64)], np.int64)
    out[f"{name}/tol"] = np.array([kw["dt"], kw.get("abstol", 0.0), kw.get("reltol", 0.0)], np.float64)
    names.append(name)
out["names"] = np.array(names)
path = Path(__file__).resolve().parent / "golden_inputs.npz"
np.savez_compressed(path, **out)
print("wrote", path, len(names), "cases")
