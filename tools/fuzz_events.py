"""Randomised differential run of the event-capable kernels (tstops, discrete callbacks, terminate!) of the strict build
against the CPU oracle; companion of tools/fuzz_parity.py.     python tools/fuzz_events.py [cases] [seed]"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import diffeqgpu_b200 as dg  # noqa: E402
from cases import lorenz_sweep  # noqa: E402
from oracle import oracle  # noqa: E402
import test_events as E  # noqa: E402
from test_gpu_parity import U0_LORENZ  # noqa: E402

f32 = np.float32
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 4242)
bad = 0
for it in range(cases):
    n = int(rng.choice([1, 31, 33, 65, 257, 1000]))
    alg = str(rng.choice(["tsit5", "tsit5", "vern7", "vern9", "rodas5p", "rosenbrock23"]))
    p = lorenz_sweep(n, seed=int(rng.integers(1 << 30)))
    if alg in ("rodas5p", "rosenbrock23"):
        p[:, 1] = 1.0 + p[:, 1] * (11.0 / 28.0)          # keep the stiff steppers away from the chaotic band
    tf = float(rng.choice([1.0, 2.5, 5.0]))
    tstops = sorted(float(f32(x)) for x in rng.uniform(0.05, tf * 0.98, int(rng.integers(0, 4))))
    if tstops and rng.random() < 0.4:
        tstops[0] = float(f32(round(tstops[0], 1)))       # on a round time
    pool = [(("u_gt", 2, float(rng.uniform(15, 35))), ("u_scale", 2, 0.5)),
            (("t_ge", 0, float(rng.uniform(0.2, tf))), ("p_set", 1, float(rng.uniform(5, 25)))),
            (("u_lt", 0, float(rng.uniform(-12, -2))), ("u_add", 0, 1.5)),
            (("u_gt", 1, float(rng.uniform(10, 25))), ("terminate", 0, 0.0))]
    if tstops:
        pool.append((("t_eq", 0, tstops[-1]), ("u_scale", 0, 0.25)))
    k = int(rng.integers(0, 3))
    cbs = [pool[i] for i in rng.choice(len(pool), size=k, replace=False)] if k else []
    if not cbs and not tstops:
        tstops = [float(f32(tf / 3))]
    mode = str(rng.choice(["adaptive_saveat", "adaptive_endpoints", "fixed", "fixed_saveat"]))
    kw = {}
    if "saveat" in mode:
        kw["saveat"] = np.sort(rng.uniform(0, tf, int(rng.integers(1, 9)))).astype(f32)
    if mode.startswith("adaptive"):
        kw.update(dt=float(rng.choice([0.01, 0.1])), adaptive=True, abstol=float(10 ** rng.uniform(-7, -4)), reltol=float(10 ** rng.uniform(-7, -4)))
        if mode == "adaptive_endpoints":
            kw["save_everystep"] = False
    else:
        kw["dt"] = float(rng.choice([0.01, 0.02, 0.05]))
    desc = dict(it=it, n=n, alg=alg, tf=tf, tstops=tstops, cbs=cbs, mode=mode,
                **{a: (b if not isinstance(b, np.ndarray) else b.tolist()) for a, b in kw.items()})
    try:
        g = E.gpu_events(dg, "lorenz", alg, U0_LORENZ, p, [0, tf], cbs, tstops=tstops or None, **kw)
        okw = dict(kw)
        if mode == "fixed":
            okw["length"] = g["us"].shape[1]
        r = oracle.solve("lorenz", alg, U0_LORENZ, p, [0, tf], callbacks=cbs, tstops=tstops or None, **okw)
        for key in ("ts", "naccept", "nreject", "retcode"):
            assert np.array_equal(g[key], r[key], equal_nan=True), f"{key} differs " + json.dumps(desc)
        w = np.ones(g["ts"].shape, bool)
        w[:, 1:] = g["ts"][:, 1:] != 0                     # rows never reached (terminated / failed) stay unwritten
        assert np.array_equal(g["us"][w], r["us"][w], equal_nan=True), "us differs (max |d| = %g) " % np.nanmax(
            np.abs(g["us"][w].astype(np.float64) - r["us"][w].astype(np.float64))) + json.dumps(desc)
    except AssertionError as e:
        bad += 1
        print("MISMATCH", str(e)[:600], flush=True)
    except Exception as e:
        bad += 1
        print("ERROR", type(e).__name__, str(e)[:300], json.dumps(desc), flush=True)
print(json.dumps(dict(cases=cases, mismatches=bad)))
sys.exit(1 if bad else 0)
