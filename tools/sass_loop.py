#!/usr/bin/env python3
"""Histogram of SASS opcodes inside the innermost backward-branch loops of each kernel in a binary/cubin/.so."""
import re, subprocess, sys, collections
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
pat = sys.argv[2] if len(sys.argv) > 2 else ""
fn = None; ins = {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: fn = m.group(1); ins[fn] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m and fn: ins[fn].append((int(m.group(1), 16), m.group(2).strip()))
for fn, lst in ins.items():
    if pat not in fn: continue
    loops = []
    for addr, text in lst:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(?:`\(\S+\)|0x([0-9a-f]+))", text)
        if m and m.group(1):
            tgt = int(m.group(1), 16)
            if tgt < addr: loops.append((tgt, addr))
    print("==", fn, "total", len(lst), "loops", [(hex(a), hex(b), (b - a) // 16) for a, b in loops])
    for a, b in loops:
        c = collections.Counter()
        for addr, text in lst:
            if a <= addr <= b:
                t = text.split()
                op = t[1] if t[0].startswith("@") else t[0]
                c[op] += 1
        print("   loop", hex(a), dict(c.most_common()))
