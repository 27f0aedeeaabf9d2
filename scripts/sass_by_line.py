#!/usr/bin/env python
"""Static view: SASS opcode counts per CUDA source line of one kernel (nvdisasm -g on a cubin/object).
   usage: sass_by_line.py <cubin or .o> <mangled kernel name> [opcode filter regex]"""
import collections, re, subprocess, sys
obj, kname = sys.argv[1:3]
filt = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
dis = subprocess.run(["nvdisasm", "-g", obj], capture_output=True, text=True).stdout
sect = dis.split(".text." + kname + ":")[1].split("\n.text.")[0]
agg = collections.defaultdict(collections.Counter)
cur = None
for ln in sect.split("\n"):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
    if m and cur:
        op = re.sub(r"^@!?U?P\d+\s+", "", m.group(1).strip()).split()[0].split(".")[0]
        agg[cur][op] += 1
tot = collections.Counter()
for k, c in agg.items():
    tot.update(c)
print("total", sum(tot.values()), tot.most_common(12))
for k, c in sorted(agg.items(), key=lambda kv: -sum(v for o, v in kv[1].items() if not filt or filt.search(o)))[:25]:
    n = sum(v for o, v in c.items() if not filt or filt.search(o))
    if n:
        print("%-18s:%-4d %5d  %s" % (k[0], k[1], n, " ".join("%s:%d" % x for x in c.most_common(5))))
