"""Randomised differential run of the strict build against the CPU oracle (development aid; the oracle is the checker,
exactly as in tests/): random ensemble sizes, time spans, saveat grids, tolerances, start steps, steppers and schedules;
every output array must be bit-identical.  Prints one line per mismatch and a summary; exit code 1 on any mismatch.

    python tools/fuzz_parity.py [cases] [seed]"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import diffeqgpu_b200 as dg  # noqa: E402
from cases import lorenz_sweep, rober_sweep  # noqa: E402
from oracle import oracle  # noqa: E402
import test_gpu_parity as T  # noqa: E402

T.ALGS.update(kvaerno3="GPUKvaerno3", kvaerno5="GPUKvaerno5")
import os
FAST_SMOKE = os.environ.get("FUZZ_FAST") == "1"
f32 = np.float32
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 20261018)
bad = 0
seen = {}
for it in range(cases):
    n = int(rng.choice([1, 2, 31, 32, 33, 63, 65, 127, 129, 255, 257, 1000, 2049, 4097]))
    stiff = rng.random() < 0.25
    if stiff:
        model, alg = "rober", str(rng.choice(["rodas5p", "rodas4", "rosenbrock23", "kvaerno3", "kvaerno5"]))
        u0, p = [1, 0, 0], rober_sweep(n, seed=int(rng.integers(1 << 30)))
        t0, tf = 0.0, float(10 ** rng.uniform(-1, 3))
        dt0 = float(10 ** rng.uniform(-5, -2))
        tols = (float(10 ** rng.uniform(-8, -5)), float(10 ** rng.uniform(-5, -2)))
    else:
        model, alg = "lorenz", str(rng.choice(["tsit5", "tsit5", "vern7", "vern9"]))
        u0, p = T.U0_LORENZ, lorenz_sweep(n, seed=int(rng.integers(1 << 30)))
        if rng.random() < 0.25:                  # four states, no parameters, per-trajectory initial values
            from cases import henon_heiles_u0
            model, u0, p = "henon_heiles", henon_heiles_u0(n, seed=int(rng.integers(1 << 30))).astype(f32), None
        t0 = float(rng.choice([0.0, 0.0, 0.37, 1.0]))
        tf = t0 + float(rng.choice([0.05, 0.5, 2.0, 5.0, 10.0, 17.3]))
        dt0 = float(rng.choice([1e-3, 0.01, 0.1, 0.5, 30.0]))
        tols = (float(10 ** rng.uniform(-8, -3)), float(10 ** rng.uniform(-8, -3)))
    mode = str(rng.choice(["saveat", "saveat", "saveat_dense", "endpoints", "fixed", "fixed_saveat"]))
    kw = dict(dt=dt0)
    if mode in ("saveat", "saveat_dense", "fixed_saveat"):
        m = int(rng.integers(1, 12)) if mode != "saveat_dense" else int(rng.integers(100, 700))
        sv = np.sort(rng.uniform(t0, tf, m)).astype(f32)
        if rng.random() < 0.5:
            sv[0] = f32(t0)
        if rng.random() < 0.5:
            sv[-1] = f32(tf)
        kw["saveat"] = sv
    if mode in ("fixed", "fixed_saveat"):
        kw["dt"] = float((tf - t0) / rng.integers(3, 150)) if not stiff else dt0 * 10
        if stiff:
            tf = t0 + kw["dt"] * int(rng.integers(3, 150))
    else:
        kw.update(adaptive=True, abstol=tols[0], reltol=tols[1])
        if mode == "endpoints":
            kw["save_everystep"] = False
    sched = str(rng.choice(["auto", "queue", "static"])) if kw.get("adaptive") else "auto"
    desc = dict(it=it, n=n, model=model, alg=alg, tspan=[t0, tf], mode=mode, sched=sched,
                **{k: (v if not isinstance(v, np.ndarray) else f"{len(v)} points") for k, v in kw.items()})
    layout = str(rng.choice(["ref", "ref", "soa"]))
    ragged = model == "lorenz" and mode != "fixed" and rng.random() < 0.3     # per-trajectory u0 / p / tspan arrays
    desc["layout"] = layout if not ragged else "ref"
    desc["ragged"] = bool(ragged)
    for key in (model, alg, mode, "ragged" if ragged else desc["layout"]):
        seen[key] = seen.get(key, 0) + 1
    try:
        if ragged:
            u0a = (rng.standard_normal((n, 3)) * 2).astype(f32)
            t0a = rng.choice([t0, t0 + 0.25], n)
            tspan_a = np.stack([t0a, t0a + rng.uniform(0.2, max(0.3, tf - t0), n)], 1).astype(f32)
            akw = {k: v for k, v in kw.items()}
            g = T.gpu_solve_arrays(dg, alg, u0a, p, tspan_a, **akw)
            r = oracle.solve(model, alg, u0a, p, tspan_a, **akw)
            t0col = tspan_a[:, :1]
            unwritten = (g["ts"] == t0col)
            unwritten[:, 0] &= not (mode == "endpoints" or ("saveat" in kw and False))
            if "saveat" in kw:
                unwritten[:, 0] = (g["ts"][:, 0] == t0col[:, 0]) & (kw["saveat"][0] != t0col[:, 0])
            for k in ("ts", "naccept", "nreject", "retcode"):
                assert np.array_equal(g[k], r[k], equal_nan=True), f"{k} differs " + json.dumps(desc)
            gu, ru = g["us"].copy(), r["us"].copy()
            gu[unwritten] = 0; ru[unwritten] = 0
            assert np.array_equal(gu, ru, equal_nan=True), "us differs (max |d| = %g) " % np.nanmax(np.abs(gu.astype(np.float64) - ru.astype(np.float64))) + json.dumps(desc)
            continue
        if FAST_SMOKE:      # the fast build on the same random cases: it must finish, save the same times and fail where the
            gf = T.gpu_solve(dg, model, alg, u0, p, [t0, tf], schedule=sched, layout=layout, fp_mode="fast", **kw)   # strict build fails
            gs = T.gpu_solve(dg, model, alg, u0, p, [t0, tf], schedule=sched, layout=layout, **kw)
            okc = gs["retcode"] == 1
            if okc.sum() >= 8:
                assert (gf["retcode"][okc] == 1).mean() >= 0.9, "fast build fails where strict succeeds (%d of %d) " % (
                    int((gf["retcode"][okc] != 1).sum()), int(okc.sum())) + json.dumps(desc)
                if layout != "soa":
                    same_rows = (gf["ts"] == gs["ts"]) | (np.isnan(gf["ts"]) & np.isnan(gs["ts"]))
                    assert same_rows.reshape(len(okc), -1)[okc].mean() >= 0.9, "saved times differ " + json.dumps(desc)
            continue
        g = T.gpu_solve(dg, model, alg, u0, p, [t0, tf], schedule=sched, layout=layout, **kw)
        if layout == "soa":                      # (rows, n, N) / (rows, N) -> the reference's (N, rows, n) / (N, rows)
            g["us"], g["ts"] = g["us"].transpose(2, 0, 1), g["ts"].T
        okw = dict(kw)
        if mode == "fixed":
            okw["length"] = g["us"].shape[1]
        r = oracle.solve(model, alg, u0, p, [t0, tf], **okw)
        # rows a trajectory never reaches (it failed, or the save point lies behind its last step) keep ts = t0 and are
        # never written: `us` is uninitialised memory there, in the reference (`allocate`) as here -- not compared
        unwritten = (g["ts"] == f32(t0))
        unwritten[:, 0] &= not (mode in ("fixed", "endpoints") or ("saveat" in kw and kw["saveat"][0] == f32(t0)))
        for k in ("ts", "naccept", "nreject", "retcode"):
            assert np.array_equal(g[k], r[k], equal_nan=True), f"{k} differs " + json.dumps(desc)
        gu, ru = g["us"].copy(), r["us"].copy()
        gu[unwritten] = 0; ru[unwritten] = 0
        assert np.array_equal(gu, ru, equal_nan=True), "us differs (max |d| = %g, failed trajectories %d) " % (
            np.nanmax(np.abs(gu.astype(np.float64) - ru.astype(np.float64))), int((r["retcode"] != 1).sum())) + json.dumps(desc)
    except AssertionError as e:
        bad += 1
        print("MISMATCH", str(e)[:400], flush=True)
    except Exception as e:                       # refused configurations must be refused by both sides
        print("ERROR", type(e).__name__, str(e)[:200], json.dumps(desc), flush=True)
        bad += 1
print(json.dumps(dict(cases=cases, mismatches=bad, seen=seen)))
sys.exit(1 if bad else 0)
