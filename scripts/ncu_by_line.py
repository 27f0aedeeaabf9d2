#!/usr/bin/env python
"""Join ncu's SASS-level source page with nvdisasm line info: stall samples and executed warp
instructions per CUDA source line.
   usage: ncu_by_line.py <src.csv from `ncu --page source --csv`> <cubin> <mangled kernel name> [top]
The library must be the SAME build the profile was taken with."""
import collections, csv, re, subprocess, sys
src_csv, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
sect = dis.split(".text." + kname + ":")[1]
sect = sect.split("\n.text.")[0].split("\n\t.section")[0]
line_of = {}           # instruction offset -> (file, line)
cur = None
for ln in sect.split("\n"):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = next(r for r in rows if "Address" in r and "Source" in r)
iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data, seen = [], set()
for r in rows:
    if len(r) > iE and r[0].startswith("0x") and r[0] not in seen:
        seen.add(r[0]); data.append(r)
base = min(int(r[0], 16) for r in data)
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot_s = tot_e = 0
for r in data:
    off = int(r[0], 16) - base
    key = line_of.get(off, ("?", 0))
    e, n = int(r[iE] or 0), int(r[iN] or 0)
    a = agg[key]; a[0] += n; a[1] += e
    a[2][re.sub(r"^@!?U?P\d+\s+", "", r[iS].strip()).split()[0].split(".")[0]] += e
    tot_s += n; tot_e += e
print("total: %d samples, %d warp instructions" % (tot_s, tot_e))
srcs = {}
for (f, l), (n, e, ops) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in srcs:
        try: srcs[f] = open("pylda_b200/csrc/" + f).read().split("\n")
        except Exception: srcs[f] = []
    text = srcs[f][l - 1].strip()[:70] if 0 < l <= len(srcs[f]) else ""
    print("%5.1f%% samples %5.1f%% instr  %-16s:%-4d %-70s %s" % (100.0 * n / tot_s, 100.0 * e / tot_e, f, l, text,
          " ".join("%s:%d" % (o, c * 1000 // max(1, tot_e)) for o, c in ops.most_common(3))))
