#!/usr/bin/env python3
"""Static SASS opcode histogram per kernel of a built library (development aid):
    python tools/sass_count.py plonky2.5_b200/libgl_commit.so ntt_pass_kernelILi8ELi10
Straight-line kernels (the NTT passes are fully unrolled) execute about what they contain."""
import collections, re, subprocess, sys

lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, hist = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_.]+)?)", line)
    if m and cur and pat in cur:
        hist.setdefault(cur, collections.Counter())[m.group(1).split(".")[0] + (".WIDE" if ".WIDE" in m.group(1) else "")] += 1
for fn, h in hist.items():
    print(fn, "total", sum(h.values()))
    print("   ", ", ".join(f"{k} {v}" for k, v in h.most_common(16)))
