#!/usr/bin/env python
"""profiles/traffic.json from an ncu metrics pass over ONE E-step of the bench workload: DRAM bytes
(dram__bytes_read.sum + dram__bytes_write.sum) per per-document kernel launch and their sum.  Input: the CSV log of
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file ...` (one row per
launch and metric) or an .ncu-rep of the same pass.
   usage: ncu_traffic.py log.csv|report.ncu-rep docs config build_digest > profiles/traffic.json"""
import collections, csv, io, json, subprocess, sys
src, docs, config, digest = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
if src.endswith(".ncu-rep"):
    text = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
else:
    text = open(src).read()
lines = [ln for ln in text.split("\n") if ln.startswith('"')]
rows = list(csv.reader(io.StringIO("\n".join(lines))))
hdr = rows[0]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}


def clean(name):
    for a in ("void pylda::", "(pylda::EParams)", "(pylda::NParams)", "(pylda::LParams)"):
        name = name.replace(a, "")
    return name


kernels = []
if "Metric Name" in hdr:                      # long format: one row per (launch, metric)
    iid, ik, im, iu, iv = (hdr.index(x) for x in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    per = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        e = per.setdefault(r[iid], {"kernel": clean(r[ik])})
        val = float(r[iv].replace(",", ""))
        if r[im].startswith("dram__bytes_read"):
            e["dram_read_bytes"] = val * scale[r[iu]]
        elif r[im].startswith("dram__bytes_write"):
            e["dram_write_bytes"] = val * scale[r[iu]]
        elif r[im].startswith("gpu__time_duration"):
            e["ms_under_ncu"] = val * tscale[r[iu]]
    kernels = list(per.values())
else:                                         # raw page of a report: one row per launch
    units = rows[1]
    ik, ir, iw, it = (hdr.index(x) for x in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
    for r in rows[2:]:
        kernels.append({"kernel": clean(r[ik]), "dram_read_bytes": float(r[ir].replace(",", "")) * scale[units[ir]],
                        "dram_write_bytes": float(r[iw].replace(",", "")) * scale[units[iw]],
                        "ms_under_ncu": float(r[it].replace(",", "")) * tscale[units[it]]})
total = sum(k["dram_read_bytes"] + k["dram_write_bytes"] for k in kernels)
json.dump({"what": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, one cold E-step of the bench "
                   "workload (scripts/tune.py); DRAM bytes of every per-document kernel launch of that E-step",
           "config": config, "docs": docs, "build_digest": digest, "dram_bytes_per_estep": total, "kernels": kernels},
          sys.stdout, indent=1)
