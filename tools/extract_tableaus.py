#!/usr/bin/env python3
"""Extract Runge-Kutta / Rosenbrock coefficient literals from the reference tree.

Reads (read-only, in the build container only):
  /root/reference/src/ensemblegpukernel/tableaus/verner_tableaus.jl
  /root/reference/src/ensemblegpukernel/tableaus/rodas_tableaus.jl
and writes tools/tableaus.json : {block_name: {coef_name: "decimal literal"}}.

The literals are kept as *strings* exactly as the reference spells them (incl. the
truncated Vern7 `a1211 = -0.0160443457`, SURVEY Q10, and the sign-outside-convert
form `-convert(T, x)` in the Rodas4 block), so that both the oracle and the CUDA
code round the same decimal to float/double as Julia's `convert(T, literal)` does.

Tsit5 is not in the reference tree (it lives in SimpleDiffEq, un-vendored); its
coefficients are stated in tools/tsit5_coeffs.py and verified by order conditions in
tests/test_tableaus.py.
"""
import json
import re
import sys
from pathlib import Path

REF = Path("/root/reference/src/ensemblegpukernel/tableaus")
OUT = Path(__file__).resolve().parent / "tableaus.json"

FUNC_RE = re.compile(r"^function\s+(\w+)\(")
ASSIGN_RE = re.compile(
    r"^\s*([A-Za-zγ]\w*)\s*=\s*(-?)\s*convert\(\s*(T2?)\s*,\s*(-?[0-9.]+(?:[eE][-+]?\d+)?)\s*\)\s*$"
)


def parse(path: Path):
    blocks = {}
    cur = None
    for line in path.read_text().splitlines():
        m = FUNC_RE.match(line)
        if m:
            cur = m.group(1)
            blocks[cur] = {}
            continue
        if line.startswith("end"):
            cur = None
            continue
        if cur is None:
            continue
        m = ASSIGN_RE.match(line)
        if m:
            name, sign, _ty, lit = m.groups()
            name = name.replace("γ", "gamma")
            if sign == "-":
                lit = lit[1:] if lit.startswith("-") else "-" + lit
            blocks[cur][name] = lit
    return blocks


def main():
    if not REF.exists():
        print("reference tree not present; keeping committed tools/tableaus.json", file=sys.stderr)
        return 0
    out = {}
    for fn in ("verner_tableaus.jl", "rodas_tableaus.jl"):
        out.update(parse(REF / fn))
    OUT.write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    for k, v in out.items():
        print(f"{k}: {len(v)} coefficients")
    return 0


if __name__ == "__main__":
    sys.exit(main())
