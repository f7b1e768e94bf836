import re, sys, collections, subprocess
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
import sass_hotloop as sh
obj = sys.argv[1]
fun = subprocess.run("cuobjdump -sass %s | grep 'Function :' | sed 's/.*Function : //'" % obj, shell=True, capture_output=True, text=True).stdout.split()[0]
ins = sh.disasm(obj, fun)
addr = {a: i for i, (a, t) in enumerate(ins)}
start = next(i for i, (a, t) in enumerate(ins) if "0.16099999845" in t)
# walk forward: straight line; conditional forward branches are assumed not taken except the one whose target
# is beyond the last packed op region and leads to a backward unconditional BRA (the tail)
i = start; path = []
last_packed = max(i for i, (a, t) in enumerate(ins) if re.search(r"F(FMA|MUL|ADD)2", t))
while True:
    a, t = ins[i]; path.append((a, t))
    m = re.search(r"BRA\S*\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
    if m and "BRA" in t:
        tgt = int(m.group(1), 16)
        cond = t.startswith("@")
        if i > last_packed and tgt > a and cond and (tgt - a) > 0x400:
            i = addr[tgt]; continue          # jump to the tail
        if not cond and tgt < a:
            break                            # back edge
        if not cond:
            i = addr[tgt]; continue
    i += 1
hist = collections.Counter(sh.opcode(t) for a, t in path)
total = sum(hist.values()); packed = hist["FFMA2"] + hist["FMUL2"] + hist["FADD2"]
cycles = sum(sh.COST[k] * v for k, v in hist.items())
print(f"{total} instr (+2 loop head), packed {packed}, cycles {cycles + 3}: " + ", ".join(f"{k} {v}" for k, v in hist.most_common()))
