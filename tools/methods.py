"""Structured descriptions of the explicit Runge-Kutta methods on the hot path.

Built from tools/tableaus.json (reference literals) and tools/tsit5_coeffs.py. Consumed by
  tools/gen_oracle_tables.py  -> oracle/oracle_tables.inc        (runtime tables, CPU oracle)
  tools/gen_device_erk.py     -> diffeqgpu.jl_b200/csrc/device/gen_erk_*.cuh (unrolled CUDA)
Only *data* is shared between the two; the stepping code is written twice, independently.

Conventions: stages are 1-based. A row is a list of (j, literal) in the order the reference
writes the summands (ascending j) -- the order matters for bit-parity:
  Tsit5  perform_step/gpu_tsit5_perform_step.jl:36-49,104-120
  Vern7  perform_step/gpu_vern7_perform_step.jl:117-150  + interpolants.jl:28-190
  Vern9  perform_step/gpu_vern9_perform_step.jl:171-262  + interpolants.jl:192-370
"""
import json
import re
from pathlib import Path

from tsit5_coeffs import AS, BTILDES, CS, RS

_TAB = json.loads((Path(__file__).resolve().parent / "tableaus.json").read_text())


def _rows_from_names(block, pat, width_i, width_j):
    """Collect a<i><j> coefficients of a block into {i: [(j, lit), ...]} sorted by j."""
    rows = {}
    for name, lit in block.items():
        m = re.fullmatch(pat, name)
        if not m:
            continue
        digits = m.group(1)
        if len(digits) != width_i + width_j:
            continue
        i, j = int(digits[:width_i]), int(digits[width_i:])
        rows.setdefault(i, []).append((j, lit))
    for i in rows:
        rows[i].sort()
    return rows


def tsit5():
    a = {}
    idx = 0
    for i in range(2, 8):
        a[i] = []
        for j in range(1, i):
            a[i].append((j, AS[idx]))
            idx += 1
    c = {i + 2: CS[i] for i in range(6)}  # c for stages 2..7 (c1..c6 in the reference)
    # stage 7 row doubles as the solution weights (FSAL): u = uprev + dt*(a71 k1 + ... + a76 k6)
    interp = {1: ["0", RS[0], RS[1], RS[2], RS[3]]}
    for i in range(2, 8):
        base = 4 + 3 * (i - 2)
        interp[i] = ["0", "0", RS[base], RS[base + 1], RS[base + 2]]
    return dict(
        name="tsit5", order=5, stages=7, fsal=True,
        a=a, c=c,
        # which stage rows are evaluated as RHS arguments, in order; stage 7 argument is u itself
        b=a[7],
        btilde=[(i + 1, BTILDES[i]) for i in range(7)],
        interp_kind="poly_b",  # u(θ) = y0 + dt * Σ b_i(θ) k_i   (@muladd, Horner)
        interp=interp, interp_stages=[1, 2, 3, 4, 5, 6, 7],
        kept=[1, 2, 3, 4, 5, 6, 7], extra=[],
        # threshold literals: Tsit5 uses T(1.0e-14) for both
        dtmin_lit="1.0e-14", dtmin_via_f32=False, land_lit="1.0e-14", land_via_f32=False,
    )


def vern7():
    t = _TAB["Vern7Tableau"]
    a = _rows_from_names(t, r"a(\d{3})", 2, 1)
    assert sorted(a) == list(range(2, 11)), sorted(a)
    c = {i: t[f"c{i}"] for i in range(2, 9)}
    c[9] = "1"   # k9 = f(g9, p, t + dt)
    c[10] = "1"
    b = [(int(k[1:]), v) for k, v in t.items() if re.fullmatch(r"b\d+", k)]
    b.sort()
    bt = [(int(k[6:]), v) for k, v in t.items() if re.fullmatch(r"btilde\d+", k)]
    bt.sort()
    ex = _TAB["Vern7ExtraStages"]
    ea = _rows_from_names(ex, r"a(\d{4})", 2, 2)
    extra = [dict(stage=i, c=ex[f"c{i}"], row=ea[i]) for i in sorted(ea)]
    ip = _TAB["Vern7InterpolationCoefficients"]
    interp = {}
    for name, lit in ip.items():
        m = re.fullmatch(r"r(\d\d)(\d)", name)
        interp.setdefault(int(m.group(1)), {})[int(m.group(2))] = lit
    # r01x has powers 1..7, others 2..7
    return dict(
        name="vern7", order=7, stages=10, fsal=False,
        a=a, c=c, b=b, btilde=bt,
        interp_kind="vern7_q",
        interp=interp, interp_stages=[1, 4, 5, 6, 7, 8, 9, 11, 12, 13, 14, 15, 16],
        kept=list(range(1, 11)), extra=extra,
        dtmin_lit="1.0e-14", dtmin_via_f32=True, land_lit="1.0e-14", land_via_f32=True,
    )


def vern9():
    t = _TAB["Vern9Tableau"]
    a = _rows_from_names(t, r"a(\d{4})", 2, 2)
    assert sorted(a) == list(range(2, 17)), sorted(a)
    # reference names c1..c13 belong to stages 2..14; stages 15,16 are evaluated at t+dt
    c = {i + 1: t[f"c{i}"] for i in range(1, 14)}
    c[15] = "1"
    c[16] = "1"
    b = [(int(k[1:]), v) for k, v in t.items() if re.fullmatch(r"b\d+", k)]
    b.sort()
    bt = [(int(k[6:]), v) for k, v in t.items() if re.fullmatch(r"btilde\d+", k)]
    bt.sort()
    ex = _TAB["Vern9ExtraStages"]
    ea = _rows_from_names(ex, r"a(\d{4})", 2, 2)
    extra = [dict(stage=i, c=ex[f"c{i}"], row=ea[i]) for i in sorted(ea)]
    ip = _TAB["Vern9InterpolationCoefficients"]
    interp = {}
    for name, lit in ip.items():
        m = re.fullmatch(r"r(\d\d)(\d)", name)
        interp.setdefault(int(m.group(1)), {})[int(m.group(2))] = lit
    return dict(
        name="vern9", order=9, stages=16, fsal=False,
        a=a, c=c, b=b, btilde=bt,
        interp_kind="poly_b",
        interp={i: (["0"] if i == 1 else ["0", "0"]) + [v[p] for p in sorted(v)]
                for i, v in interp.items()},
        interp_stages=[1, 8, 9, 10, 11, 12, 13, 14, 15, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26],
        # only these true stages survive the step (slots k1..k10, gpu_vern9_perform_step.jl:281-292)
        kept=[1, 8, 9, 10, 11, 12, 13, 14, 15, 16], extra=extra,
        # dtmin convert(T,1.0f-14) but land-on-tf convert(T,1.0e-14)
        dtmin_lit="1.0e-14", dtmin_via_f32=True, land_lit="1.0e-14", land_via_f32=False,
    )


def all_methods():
    return [tsit5(), vern7(), vern9()]


if __name__ == "__main__":
    for m in all_methods():
        print(m["name"], "stages", m["stages"], "terms", sum(len(r) for r in m["a"].values()),
              "extra", len(m["extra"]), "interp stages", len(m["interp"]))
        for i in sorted(m["a"]):
            print("  ", i, [j for j, _ in m["a"][i]])
        print("   b", [j for j, _ in m["b"]], "bt", [j for j, _ in m["btilde"]])
        for e in m["extra"]:
            print("   extra", e["stage"], [j for j, _ in e["row"]])
        for i in sorted(m["interp"]):
            print("   interp", i, len(m["interp"][i]))
