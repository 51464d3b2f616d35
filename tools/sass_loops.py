#!/usr/bin/env python
"""List the loops (backward branches) of one kernel in a cuobjdump -sass listing with their instruction mix.

    cuobjdump -sass -fun <mangled> lib.so > k.sass ; python tools/sass_loops.py k.sass [min_fp64]
"""
import re
import sys
from collections import Counter

lines = open(sys.argv[1]).read().split("\n")
min_fp64 = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ins = []
for l in lines:
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
addr = {a: i for i, (a, _) in enumerate(ins)}


def opcode(x):
    t = x.split()
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]


for i, (a, t) in enumerate(ins):
    if "BRA" in t:
        m2 = re.search(r"0x([0-9a-f]+)", t)
        if m2:
            tgt = int(m2.group(1), 16)
            if tgt < a and tgt in addr:
                body = ins[addr[tgt]:i + 1]
                c = Counter(opcode(x) for _, x in body)
                nd = c["DFMA"] + c["DADD"] + c["DMUL"]
                if nd >= min_fp64:
                    print(f"loop {tgt:#x}..{a:#x}: {len(body)} instr, fp64 {nd}")
                    print("   ", sorted(c.items(), key=lambda kv: -kv[1]))
