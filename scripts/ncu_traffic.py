#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` capture of ONE E-step of the bench workload: DRAM bytes
(dram__bytes_read.sum + dram__bytes_write.sum) per per-document kernel launch and their sum.
   usage: ncu_traffic.py report.ncu-rep docs config build_digest > profiles/traffic.json"""
import csv, io, json, subprocess, sys
rep, docs, config, digest = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ik, ir, iw, it = (hdr.index(x) for x in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
kernels, total = [], 0.0
for r in rows[2:]:
    name = r[ik].replace("void pylda::", "").replace("(pylda::EParams)", "").replace("(pylda::NParams)", "")
    rd = float(r[ir].replace(",", "")) * scale[units[ir]]
    wr = float(r[iw].replace(",", "")) * scale[units[iw]]
    ms = float(r[it].replace(",", "")) * tscale[units[it]]
    kernels.append({"kernel": name, "dram_read_bytes": rd, "dram_write_bytes": wr, "ms_under_ncu": ms})
    total += rd + wr
json.dump({"what": "ncu --set full --clock-control none, one cold E-step of the bench workload (scripts/tune.py); "
                   "DRAM bytes of every per-document kernel launch of that E-step",
           "config": config, "docs": docs, "build_digest": digest, "dram_bytes_per_estep": total, "kernels": kernels},
          sys.stdout, indent=1)
