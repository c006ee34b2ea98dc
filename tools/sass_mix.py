#!/usr/bin/env python
"""Per-kernel SASS instruction mix of a built library (no GPU needed).

usage: tools/sass_mix.py direct-nbody_b200/lib/libnbk.so [name-substring]
For every kernel: FP64-pipe ops (DFMA/DMUL/DADD/DSETP), MUFU, LDS, total; and the same for the
hottest loop body (the backward-branch span holding the most DFMA), which is what the FP64
roofline of DESIGN.md is computed from.
"""
import re
import subprocess
import sys
from collections import Counter


def kernels(lib):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, out = None, {}
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            out[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return out


def opname(ins):
    ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
    return ins.split()[0].split(".")[0] if ins else ""


def mix(instrs):
    c = Counter(opname(i) for _, i in instrs)
    return c


def hot_loop(instrs):
    addr_index = {a: k for k, (a, _) in enumerate(instrs)}
    best, best_span = -1, None
    for k, (a, ins) in enumerate(instrs):
        m = re.search(r"BRA\S*\s+.*?0x([0-9a-f]+)", ins)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_index:
                span = instrs[addr_index[tgt]:k + 1]
                n = sum(1 for _, i in span if opname(i) == "DFMA")
                if n > best:
                    best, best_span = n, span
    return best_span


def show(c, label):
    fp64 = c["DFMA"] + c["DMUL"] + c["DADD"] + c["DSETP"]
    tot = sum(c.values())
    print(f"  {label:10s} total {tot:5d} | fp64 {fp64:4d} (DFMA {c['DFMA']}, DMUL {c['DMUL']}, "
          f"DADD {c['DADD']}, DSETP {c['DSETP']}) | MUFU {c['MUFU']} | LDS {c['LDS']} | "
          f"other {tot - fp64 - c['MUFU'] - c['LDS']}")


if __name__ == "__main__":
    lib = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 else ""
    for name, ins in kernels(lib).items():
        if pat not in name:
            continue
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        print(dem[:150])
        show(mix(ins), "kernel")
        hl = hot_loop(ins)
        if hl:
            show(mix(hl), "hot loop")
            others = Counter(opname(i) for _, i in hl)
            rest = {k: v for k, v in others.items() if k not in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU", "LDS")}
            print("   other ops:", dict(sorted(rest.items(), key=lambda kv: -kv[1])))
