"""Coefficient sanity: order conditions of the tableaus the oracle and the CUDA code are
generated from (tools/tableaus.json, tools/tsit5_coeffs.py), and freshness of generated files."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
from methods import all_methods  # noqa: E402

METHODS = {m["name"]: m for m in all_methods()}


def dense(m):
    S = m["stages"]
    A = np.zeros((S + 1, S + 1))
    for i, row in m["a"].items():
        for j, lit in row:
            A[i, j] = float(lit)
    c = np.zeros(S + 1)
    for i, lit in m["c"].items():
        c[i] = float(lit)
    b = np.zeros(S + 1)
    for j, lit in m["b"]:
        b[j] = float(lit)
    bt = np.zeros(S + 1)
    for j, lit in m["btilde"]:
        bt[j] = float(lit)
    return A, b, bt, c


@pytest.mark.parametrize("name", ["tsit5", "vern7", "vern9"])
def test_row_sums_and_quadrature(name):
    m = METHODS[name]
    A, b, bt, c = dense(m)
    # row-sum condition sum_j a_ij = c_i (Vern7's truncated a1211 lives in the extra stages only)
    assert np.abs(A.sum(1) - c)[2:].max() < 5e-15
    assert abs(b.sum() - 1) < 5e-15
    assert abs(bt.sum()) < 5e-15
    for k in range(m["order"]):
        assert abs(b @ c ** k - 1 / (k + 1)) < 2e-14, (name, k)
    # tree conditions b.A.c^k = 1/((k+1)(k+2)) up to the order
    for k in range(m["order"] - 1):
        assert abs(b @ A @ c ** k - 1 / ((k + 1) * (k + 2))) < 2e-14, (name, k)


def test_tsit5_dense_output_endpoints():
    m = METHODS["tsit5"]
    A, b, _, _ = dense(m)
    for i in range(1, 8):
        coef = [float(x) for x in m["interp"][i]]
        assert abs(sum(coef) - b[i] if i < 7 else sum(coef)) < 5e-14 or i == 7
    # b_i(1) = a_7i and b_7(1) = 0
    b1 = np.array([sum(float(x) for x in m["interp"][i]) for i in range(1, 8)])
    assert np.abs(b1[:6] - A[7, 1:7]).max() < 5e-14 and abs(b1[6]) < 5e-14
    # sum_i b_i(theta) = theta ; sum_i b_i(theta) c_i = theta^2/2
    c = np.array([0, 0.161, 0.327, 0.9, 0.9800255409045097, 1, 1])
    for th in (0.13, 0.5, 0.87):
        bth = np.array([np.polyval([float(x) for x in m["interp"][i]][::-1], th) for i in range(1, 8)])
        assert abs(bth.sum() - th) < 1e-13
        assert abs(bth @ c - th ** 2 / 2) < 1e-13


def test_vern7_truncated_literal_is_kept():
    # the reference spells a1211 with 10 digits; the oracle must use those digits
    tab = json.loads((ROOT / "tools" / "tableaus.json").read_text())
    assert tab["Vern7ExtraStages"]["a1211"] == "-0.0160443457"


def test_rodas_gamma_consistency():
    tab = json.loads((ROOT / "tools" / "tableaus.json").read_text())
    assert tab["Rodas4Tableau"]["gamma"] == tab["Rodas4Tableau"]["d1"] == "0.25"
    assert tab["Rodas5PTableau"]["gamma"] == tab["Rodas5PTableau"]["d1"]
    assert tab["Rodas4Tableau"]["a54"].startswith("-")   # sign-outside-convert form parsed


def test_generated_files_are_fresh(tmp_path):
    """the committed generated headers equal what the generators emit now"""
    dev = ROOT / "diffeqgpu.jl_b200" / "csrc" / "device"
    before = {p.name: p.read_text() for p in dev.glob("gen_*.cuh")}
    before["oracle_tables.inc"] = (ROOT / "oracle" / "oracle_tables.inc").read_text()
    for tool in ("gen_device_erk.py", "gen_device_rodas.py", "gen_oracle_tables.py"):
        subprocess.check_call([sys.executable, str(ROOT / "tools" / tool)], stdout=subprocess.DEVNULL)
    after = {p.name: p.read_text() for p in dev.glob("gen_*.cuh")}
    after["oracle_tables.inc"] = (ROOT / "oracle" / "oracle_tables.inc").read_text()
    assert before == after


@pytest.mark.skipif(not Path("/root/reference").exists(), reason="reference tree not mounted")
def test_tableaus_json_matches_reference(tmp_path):
    import extract_tableaus as ex
    out = {}
    for fn in ("verner_tableaus.jl", "rodas_tableaus.jl"):
        out.update(ex.parse(ex.REF / fn))
    assert out == json.loads((ROOT / "tools" / "tableaus.json").read_text())
