#!/usr/bin/env python
"""Block-level view of `ncu --page source --csv`: consecutive SASS instructions with the same
execution count are one block (loop body / phase); prints each block's share of the stall samples,
its stall-reason mix and opcode mix.   usage: ncu_blocks.py src.csv [top]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
hdr = next(r for r in rows if "Address" in r and "Source" in r)
iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
st = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
seen, data = set(), []
for r in rows:
    if len(r) > iE and r[0].startswith('0x') and r[0] not in seen:
        seen.add(r[0]); data.append(r)
tot = sum(int(r[iN] or 0) for r in data)
blocks, cur = [], None
for r in data:
    e, n = int(r[iE] or 0), int(r[iN] or 0)
    if cur is None or abs(e - cur['e']) > 0.02 * max(e, cur['e'], 1):
        cur = {'e': e, 'n': 0, 'cnt': 0, 'st': collections.Counter(), 'ops': collections.Counter()}
        blocks.append(cur)
    cur['n'] += n; cur['cnt'] += 1
    cur['ops'][re.sub(r"^@!?U?P\d+\s+", "", r[iS].strip()).split()[0].split('.')[0]] += 1
    for i in st:
        v = int(r[i] or 0)
        if v: cur['st'][hdr[i][6:]] += v
print("kernel:", rows[0][1] if len(rows[0]) > 1 else "?", " total samples", tot)
for b in sorted(blocks, key=lambda b: -b['n'])[:top]:
    print("%5.1f%% samples %4d SASS exec=%9d | %s | %s" % (100.0 * b['n'] / max(1, tot), b['cnt'], b['e'],
          " ".join("%s:%d" % (k, 100 * v // max(1, b['n'])) for k, v in b['st'].most_common(4)),
          " ".join("%s:%d" % kv for kv in b['ops'].most_common(6))))
