"""The C-ABI boundary without a GPU: the library loads, exports exactly what include/degk.h
declares, the host helpers agree with the reference's sizing rules, NVRTC compiles a user model,
and there is no CPU fallback (context creation fails loudly without a device)."""
import ctypes
import re
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import diffeqgpu_b200 as dg  # noqa: E402
from diffeqgpu_b200 import _lib  # noqa: E402


def header_symbols():
    text = (ROOT / "include" / "degk.h").read_text()
    return sorted(set(re.findall(r"DEGK_API\s+[\w\s\*]+?\b(degk_\w+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 14
    assert sorted(syms) == sorted(_lib.API_SYMBOLS)
    for s in syms:
        assert hasattr(lib, s), s
    assert lib.degk_version() == 100


def test_struct_layouts_match_the_c_compiler(tmp_path):
    """ctypes mirrors vs sizeof/offsetof from gcc on include/degk.h"""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('''#include <stdio.h>
#include <stddef.h>
#include "degk.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(degk_model_desc), sizeof(degk_program_info),
         sizeof(degk_solve_args), offsetof(degk_solve_args, abstol), offsetof(degk_solve_args, n_rows),
         offsetof(degk_solve_args, seed), offsetof(degk_solve_args, max_iters));
  return 0; }''')
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    S = _lib.SolveArgs
    assert got == [ctypes.sizeof(_lib.ModelDesc), ctypes.sizeof(_lib.ProgramInfo), ctypes.sizeof(S),
                   S.abstol.offset, S.n_rows.offset, S.seed.offset, S.max_iters.offset]


def test_builtin_table_lists_the_hot_path_kernels():
    lib = _lib.lib()
    names = {lib.degk_builtin_name(i).decode() for i in range(lib.degk_builtin_count())}
    for need in ("lorenz/alg0/f32/fixed", "lorenz/alg0/f32/adaptive", "lorenz/alg2/f64/adaptive",
                 "henon_heiles/alg2/f64/adaptive", "rober/alg5/f32/adaptive", "lorenz/alg6/f32/fixed",
                 "gbm/alg7/f32/fixed"):
        assert need in names, need


@pytest.mark.parametrize("args,expect", [
    ((0, 0.0, 10.0, 0.1, 0, 1, 0), 101),      # length(0f0:0.1f0:10f0)
    ((0, 0.0, 10.0, 0.01, 0, 1, 0), 1001),
    ((1, 0.0, 10.0, 0.1, 0, 1, 0), 101),
    ((0, 0.0, 1.0, 0.3, 0, 1, 0), 4),         # 0:0.3:1 -> 0,0.3,0.6,0.9
    ((1, 0.0, 1.0, 1.0 / 64, 0, 1, 0), 65),
    ((0, 0.0, 10.0, 0.1, 0, 0, 0), 2),        # endpoints only
    ((0, 0.0, 10.0, 0.1, 1, 0, 0), 2),
    ((0, 0.0, 10.0, 0.1, 1, 1, 0), 101),      # adaptive save_everystep: ceil((tf-t0)/dt)+1
    ((0, 0.0, 10.0, 0.1, 1, 1, 7), 7),        # saveat wins
])
def test_output_rows(args, expect):
    assert _lib.lib().degk_output_rows(*args) == expect


def test_no_cpu_fallback_context_fails_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.DegkError) as e:
        _lib.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_nvrtc_compiles_user_model_without_gpu():
    for alg in (0, 2, 5):
        d = _lib.make_desc(rhs_src=dg.models.LORENZ_RHS, jac_src=dg.models.LORENZ_JAC, n_state=3,
                           n_param=3, alg=alg, dtype=_lib.F32, fp_mode=_lib.FP_STRICT)
        st, nbytes, log = _lib.jit_compile_check(d)
        assert st == _lib.OK and nbytes > 10000, log
    d = _lib.make_desc(rhs_src=dg.models.gbm_src.f.rhs, noise_src=dg.models.gbm_src.g, n_state=3, n_param=2,
                       noise_kind=_lib.NOISE_DIAGONAL, alg=6, dtype=_lib.F64, fp_mode=_lib.FP_FAST)
    st, nbytes, log = _lib.jit_compile_check(d)
    assert st == _lib.OK and nbytes > 5000, log


def test_nvrtc_reports_errors_not_crashes():
    d = _lib.make_desc(rhs_src="du[0] = undefined_symbol;", n_state=1, alg=0)
    st, nbytes, log = _lib.jit_compile_check(d)
    assert st == _lib.ERR_NVRTC and "undefined_symbol" in log
    d = _lib.make_desc(rhs_src=dg.models.LORENZ_RHS, n_state=3, n_param=3, alg=5)   # stiff without jac:
    st, nb, log = _lib.jit_compile_check(d)                                            # forward-mode duals
    assert st == _lib.OK and nb > 0, log
    st, _, log = _lib.jit_compile_check(_lib.make_desc(rhs_src=dg.models.LORENZ_RHS, n_state=3, n_param=3, alg=5, jac_mode=7))
    assert st == _lib.ERR_INVALID and "jac_mode" in log


def test_product_does_not_import_oracle():
    pkg = ROOT / "diffeqgpu.jl_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.h")):
        txt = f.read_text()
        assert "oracle" not in txt.replace("the oracle", "").replace("oracle/", "").replace("oracle's", "").replace("oracle (", "").lower() \
            or "import oracle" not in txt and "from oracle" not in txt, f
        assert "import oracle" not in txt and "from oracle" not in txt and "degk_oracle" not in txt, f


def _c_struct_fields(name):
    """field names of `typedef struct { ... } name;` in include/degk.h, in declaration order"""
    import re
    h = (ROOT / "include" / "degk.h").read_text()
    m = re.search(r"typedef struct\s*\{((?:(?!typedef struct).)*?)\}\s*%s\s*;" % name, h, re.S)
    assert m, name
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    out = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        parts = stmt.split(",")
        out.append(parts[0].split()[-1].strip("* "))
        out.extend(p.strip().strip("* ") for p in parts[1:])
    return out


def test_bindings_mirror_the_header_field_for_field():
    """the ctypes mirrors and the Julia `ext/` structs list the fields of include/degk.h in order"""
    import re
    jl = (ROOT / "diffeqgpu.jl_b200" / "julia_ext" / "DiffEqGPUDegkExt.jl").read_text()
    for cname, jname, ctype in (("degk_model_desc", "ModelDesc", _lib.ModelDesc), ("degk_solve_args", "SolveArgs", _lib.SolveArgs)):
        want = _c_struct_fields(cname)
        assert [f[0] for f in ctype._fields_] == want, cname
        body = re.search(r"struct %s\n(.*?)\nend" % jname, jl, re.S).group(1)
        body = re.sub(r"#.*", "", body)
        got = re.findall(r"(\w+)::", body)
        assert got == want, (jname, got, want)
        # the mirrors are keyword-constructed (Base.@kwdef, a default for every field): a call can only go stale by
        # naming a field that does not exist -- check every call site
        assert re.search(r"Base\.@kwdef struct %s\n" % jname, jl), jname
        assert len(re.findall(r"::[^;=\n]+=", body)) == len(want), (jname, "every field needs a default")
        calls = re.findall(r"%s\((.*?)\)\)" % jname, jl, re.S)
        assert calls, jname
        for call in calls:
            call = re.sub(r"#.*", "", call)
            depth, cur, parts = 0, "", []
            for ch in call:                                   # split the argument list at top-level commas
                if ch in "([{":
                    depth += 1
                if ch in ")]}":
                    depth -= 1
                if ch == "," and depth == 0:
                    parts.append(cur); cur = ""
                else:
                    cur += ch
            parts.append(cur)
            for part in parts:
                m = re.match(r"\s*(\w+)\s*=(?!=)", part)
                assert m, (jname, "positional argument in a constructor call", part.strip()[:60])
                assert m.group(1) in want, (jname, "unknown field", m.group(1))
    assert "_convert_saveat" not in jl                        # (round 1 called a function the reference does not have)
