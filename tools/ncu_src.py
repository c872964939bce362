#!/usr/bin/env python3
"""Aggregate the per-instruction stall samples of `ncu --page source --csv` by opcode and by address range (loops).
usage: ncu_src.py file.csv [lo_hex hi_hex]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = rows[2:]
base = int(data[0][ix["Address"]], 16)
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 60
by_op = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
n_inst = collections.Counter()
for r in data:
    off = int(r[ix["Address"]], 16) - base
    if not (lo <= off <= hi):
        continue
    t = r[ix["Source"]].split()
    op = t[1] if t[0].startswith("@") else t[0]
    op = op.rstrip(";")
    ex = int(r[ix["Instructions Executed"]])
    n_inst[op] += ex
    for s in stalls:
        v = int(r[ix[s]])
        by_op[op][s] += v
        tot[s] += v
        by_op[op]["all"] += v
        tot["all"] += v
print("range", hex(lo), hex(hi), "total samples", tot["all"], "inst executed", sum(n_inst.values()))
print("by reason:", {k[6:]: round(v / tot["all"], 3) for k, v in tot.most_common() if k != "all" and v / tot["all"] > 0.005})
print("%-18s %10s %7s %8s  top reasons" % ("opcode", "executed", "smp%", "smp/inst"))
tot_ex = sum(n_inst.values())
for op, c in sorted(by_op.items(), key=lambda kv: -kv[1]["all"])[:16]:
    top = ", ".join("%s %.0f%%" % (k[6:], 100 * v / c["all"]) for k, v in c.most_common(5) if k != "all")
    print("%-18s %10d %6.1f%% %8.2f  %s" % (op, n_inst[op], 100 * c["all"] / tot["all"], c["all"] / max(n_inst[op], 1) * tot_ex / tot["all"], top))
