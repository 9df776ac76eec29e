#!/usr/bin/env python
"""Opcode histogram of the hottest loop (largest backward branch span) of one kernel in a cuobjdump -sass dump.

    cuobjdump -sass build/obj/cloudsc2_nl_kernel.o > /tmp/nl.sass
    python tools/sass_loop_hist.py /tmp/nl.sass 'k_cloudsc2_nlILb0ELi2ELi128ELi128ELb0ELb0ELi0E' [--dump]
"""
import collections
import re
import sys

txt = open(sys.argv[1]).read().split("Function : ")
pat = sys.argv[2]
body = next(b for b in txt if pat in b.split("\n", 1)[0])
ins = []
for line in body.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
best = None
for addr, t in ins:
    m = re.search(r"\bBRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < addr and (best is None or addr - tgt > best[1] - best[0]):
            best = (tgt, addr)
lo, hi = best
loop = [t for a, t in ins if lo <= a <= hi]
print(f"loop 0x{lo:x}..0x{hi:x}: {len(loop)} instructions")
if "--dump" in sys.argv:
    print("\n".join(loop))
    sys.exit(0)
c = collections.Counter()
for t in loop:
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    c[t.split()[0]] += 1
fp64 = sum(v for k, v in c.items() if k.startswith(("DFMA", "DMUL", "DADD", "DSETP", "MUFU")))
print(f"FP64-pipe (+MUFU): {fp64}, other: {len(loop) - fp64}")
for k, v in c.most_common():
    print(f"{v:5d}  {k}")
