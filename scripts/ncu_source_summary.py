#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: executed warp-instructions and stall samples by
opcode, stall reasons, and the hottest SASS lines.  Usage: ncu_source_summary.py src.csv [top]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = next(r for r in rows if "Address" in r and "Source" in r)
iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [r for r in rows if len(r) > iE and r[0].startswith("0x")]
ops, samp = collections.Counter(), collections.Counter()
tot = tots = 0
for r in data:
    s = re.sub(r"^@!?U?P\d+\s+", "", r[iS].strip())
    op = (s.split()[0] if s else "?").split(".")[0]
    e, n = int(r[iE] or 0), int(r[iN] or 0)
    ops[op] += e; samp[op] += n; tot += e; tots += n
print("kernel:", rows[0][1] if len(rows[0]) > 1 else "?")
print("SASS lines %d, warp instructions executed %d, stall samples %d" % (len(data), tot, tots))
for op, c in ops.most_common(top):
    print("  %-12s %14d %5.1f%%   samples %5.1f%%" % (op, c, 100.0 * c / tot, 100.0 * samp[op] / max(1, tots)))
print("stall reasons (all samples):")
for name in hdr:
    if name.startswith("stall_") and "Not Issued" not in name:
        i = hdr.index(name)
        v = sum(int(r[i] or 0) for r in data)
        if v:
            print("  %-24s %8d %5.1f%%" % (name, v, 100.0 * v / max(1, tots)))
print("hottest SASS lines by samples:")
for r in sorted(data, key=lambda r: -int(r[iN] or 0))[:top]:
    print("  %6s %10s  %s" % (r[iN], r[iE], r[iS].strip()[:100]))
