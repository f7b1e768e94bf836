#!/usr/bin/env python
"""Static view of a kernel's every-iteration path: dump the SASS of one function from an object file
(cuobjdump), find the densest packed-FP32 region and the blocks chained to it by unconditional / loop
branches, and print the opcode histogram with the measured issue-cost model of profiles/r1_issue_probe.json
(FFMA2/FMUL2/FADD2 2 cycles on fmaheavy, IMAD 2 on fmaheavy, ALU-pipe ops 2, MUFU 4 per quarter-rate op).

    python tools/sass_hotloop.py diffeqgpu.jl_b200/csrc/build/aot_fast_0.o \
        _ZN4degk13k_ode_asolve2ILi0EfNS_6LorenzENS_8ErkTsit5ELi2EEEvNS_5KArgsE [lo hi]

`lo hi` (hex addresses) override the automatic region.  No GPU needed.
"""
import collections
import re
import subprocess
import sys

FMA_PIPE = {"FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "IMAD", "HFMA2"}
ALU_PIPE = {"FSETP", "FSEL", "FMNMX", "ISETP", "SEL", "LOP3", "PLOP3", "IADD3", "VIADD", "MOV", "SHF", "LEA", "P2R", "VOTE",
            "POPC", "F2F", "VIADDMNMX", "UIADD3", "ULOP3"}
COST = collections.defaultdict(lambda: 2, {"FFMA": 1, "FMUL": 1, "FADD": 1, "MUFU": 4, "BRA": 1, "BSSY": 1, "BSYNC": 1,
                                            "LDCU": 1, "LDC": 1, "LDS": 1, "STS": 1, "LD": 1, "NOP": 1})


def disasm(obj, fun):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True, check=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def opcode(text):
    return re.sub(r"^@!?U?P\d\s+", "", text).split()[0].split(".")[0]


def main():
    obj, fun = sys.argv[1], sys.argv[2]
    ins = disasm(obj, fun)
    if len(sys.argv) >= 5:
        lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
    else:
        dens = collections.Counter()
        for a, t in ins:
            if re.search(r"\bF(FMA|MUL|ADD)2\b", t):
                dens[a // 0x400] += 1
        hot = [k for k, v in dens.items() if v >= 8]
        lo, hi = min(hot) * 0x400, (max(hot) + 1) * 0x400
        # extend to the enclosing branch-free stretch boundaries
        addrs = [a for a, t in ins if lo <= a <= hi]
        lo, hi = min(addrs), max(addrs)
    hist = collections.Counter(opcode(t) for a, t in ins if lo <= a <= hi)
    total = sum(hist.values())
    fma = sum(v for k, v in hist.items() if k in FMA_PIPE)
    alu = sum(v for k, v in hist.items() if k in ALU_PIPE)
    packed = hist["FFMA2"] + hist["FMUL2"] + hist["FADD2"]
    cycles = sum(COST[k] * v for k, v in hist.items())
    print(f"function {fun}\nregion 0x{lo:x}..0x{hi:x}: {total} instructions of {len(ins)}")
    print("opcodes:", ", ".join(f"{k} {v}" for k, v in hist.most_common()))
    print(f"packed FP32: {packed} ({100 * packed / total:.0f} %), fma-pipe instructions: {fma}, alu-pipe: {alu}, MUFU: {hist['MUFU']}")
    print(f"issue-model cycles per pass: {cycles}  (packed-FP floor {2 * packed}, i.e. {100 * 2 * packed / cycles:.0f} % of the modelled issue time)")


if __name__ == "__main__":
    main()
