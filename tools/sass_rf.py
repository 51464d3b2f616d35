#!/usr/bin/env python
"""Register-file read estimate of a SASS loop: python tools/sass_rf.py k.sass 0xSTART 0xEND
Counts, per instruction, the vector-register source operands that are not served by the reuse cache (.reuse on the
previous use of the same slot is not modelled: a flagged operand is counted once where it is flagged)."""
import re, sys
from collections import Counter
lines = open(sys.argv[1]).read().split("\n")
lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
tot = Counter(); n = Counter(); hist = Counter()
for l in lines:
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if not m: continue
    a = int(m.group(1), 16)
    if a < lo or a > hi: continue
    t = m.group(2).split()
    if t[0].startswith("@"): t = t[1:]
    op = t[0].split(".")[0]
    ops = " ".join(t[1:]).split(",")
    srcs = ops[1:] if op not in ("STS", "STG", "BRA", "ISETP", "DSETP") else ops
    regs = [s for s in srcs if re.search(r"\bR\d+", s)]
    k = len(regs)
    wide = op in ("DFMA", "DADD", "DMUL")
    tot[op] += k; n[op] += 1
    if wide: hist[k] += 1
print("fp64 instr by #vector-register source operands:", dict(hist))
print("fp64 cycles at max(2, operands):", sum(max(2, k) * v for k, v in hist.items()))
print({o: (n[o], tot[o]) for o in n})
