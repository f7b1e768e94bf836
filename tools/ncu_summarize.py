"""Summarise an ncu report for profiles/: key metrics of the first kernel (raw page), the
share of instructions / stall samples per code region (source page), and the launch list of a
`--metrics gpu__time_duration.sum` CSV.

    python tools/ncu_summarize.py rep  <file.ncu-rep> <name> "<workload line>"   -> profiles/<name>.txt (+ r1_ncu_summary.json)
    python tools/ncu_summarize.py list <launches.csv> <out.txt> "<command line>"
"""
import csv
import io
import json
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active"]


def ncu_csv(rep, page):
    out = subprocess.check_output(["ncu", "-i", str(rep), "--page", page, "--csv"], text=True, stderr=subprocess.DEVNULL)
    return list(csv.reader(io.StringIO(out)))


def summarize_rep(rep, name, workload):
    rows = ncu_csv(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    d, lines = {}, [f"# ncu --set full --clock-control none --import-source on  ({Path(rep).name})", f"# workload: {workload}"]
    for i, h in enumerate(hdr):
        stall = "issue_stalled" in h and h.endswith("per_issue_active.ratio")
        if h in KEYS or (stall and vals[i] and float(vals[i]) > 0.05):
            d[h] = vals[i]
            lines.append(f"{h} [{units[i]}] = {vals[i]}")
    # source page: instructions and stall samples by execution-frequency class
    src = ncu_csv(rep, "source")
    h2 = src[1]
    ix = {h: i for i, h in enumerate(h2)}
    data = [r for r in src[2:] if len(r) > ix["Instructions Executed"]]
    mx = max(int(r[ix["Instructions Executed"]]) for r in data)
    tot_i = sum(int(r[ix["Instructions Executed"]]) for r in data)
    tot_s = sum(int(r[ix["# Samples"]]) for r in data)
    cls = defaultdict(lambda: [0, 0, 0])
    ops = defaultdict(int)
    for r in data:
        e, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
        c = "every iteration" if e > 0.5 * mx else ("sometimes (service / deferred saves)" if e > 0.002 * mx else "cold")
        cls[c][0] += 1; cls[c][1] += e; cls[c][2] += s
        if e > 0.5 * mx:
            t = r[ix["Source"]].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            ops[op] += 1
    lines.append("# code regions by execution frequency (SASS instructions, share of executed warp-instructions, share of stall samples)")
    for c, (n, e, s) in cls.items():
        lines.append(f"#   {c:40s} n={n:5d}  inst={100 * e / tot_i:5.1f}%  samples={100 * s / tot_s:5.1f}%")
    lines.append("# opcodes of the every-iteration path: " + ", ".join(f"{k} {v}" for k, v in sorted(ops.items(), key=lambda x: -x[1])))
    (ROOT / "profiles" / f"{name}.txt").write_text("\n".join(lines) + "\n")
    js = ROOT / "profiles" / ("r2_ncu_summary.json" if name.startswith("r2") else "r1_ncu_summary.json")
    allj = json.loads(js.read_text()) if js.exists() else {}
    allj[name] = d
    js.write_text(json.dumps(allj, indent=1))
    print("\n".join(lines))


def summarize_list(csvfile, out, command):
    rows = list(csv.reader(l for l in open(csvfile) if l.startswith('"')))
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if len(r) <= iv or r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
            continue
        unit = r[hdr.index("Metric Unit")]
        v = float(r[iv].replace(",", ""))
        ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
        agg[r[ik]][0] += 1; agg[r[ik]][1] += ms
    tot = sum(v[1] for v in agg.values())
    lines = [f"# {command}", "# every kernel launched by the bench process (cold-cache, serialised: compare SHARES, not absolutes)",
             f"# {sum(v[0] for v in agg.values())} launches, {tot:.1f} ms total device time", "# count   total_ms   share   kernel"]
    for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append(f"{n:6d} {ms:10.3f} {100 * ms / tot:6.2f}%  {k[:100]}")
    Path(out).write_text("\n".join(lines) + "\n")
    print("\n".join(lines[:12]))


if __name__ == "__main__":
    if sys.argv[1] == "rep":
        summarize_rep(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        summarize_list(sys.argv[2], sys.argv[3], sys.argv[4])
